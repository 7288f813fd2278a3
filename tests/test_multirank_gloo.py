"""world_size-2 gloo test of the N>1 host logic on CPU: round-robin ownership
(F:1162), the unique-id broadcast used to bootstrap NCCL, and "sum of the
rank partials, then fold" == the single-process result.  The per-rank compute
is the CPU oracle here (no GPU in this container); on the GPU box the same
host helpers drive the CUDA library (tests/test_gpu_multi.py, bench.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pyoracle as O
from tests import util as U


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import mrg_b200 as mrg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1) unique-id plumbing (a fake id stands in for ncclGetUniqueId)
        uid = mrg.broadcast_unique_id(rank, make_id=lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # 2) ownership + moment sum
        p = U.make_parm(8, 6, 8)
        sp, _ = U.load_species(p, 6)
        f12 = U.smooth_fields(p, seed=2)
        a6 = O.field_prep(p, f12)
        npr = len(sp[2][0])
        ipar = rank + 1
        sl = mrg.owned_slice(ipar, world)
        mine = [np.ascontiguousarray(a[sl]) for a in sp[2]]
        assert len(mine[0]) == mrg.owned_count(npr, ipar, world)
        r = O.fulmov(p, a6, *mine, -1.0, 1.0, 1, nranks=1, want_raw=True)
        buf = torch.from_numpy(np.concatenate(r["raw"] + [np.array([r["wkix"], r["wkih"]])]))
        dist.all_reduce(buf)                              # stands in for ncclAllReduce(ncclDouble, ncclSum)
        n = O.mxyzA(p)
        tot = [buf[c * n:(c + 1) * n].numpy().copy() for c in range(4)]
        O.vmesh3(p, tot[0], tot[1], tot[2])
        O.vmesh1(p, tot[3])
        ref = O.fulmov(p, a6, *[a.copy() for a in sp[2]], -1.0, 1.0, 1, nranks=world)
        err = max(U.rel_l2(tot[c], ref["mom"][c]) for c in range(4))
        wk_err = abs(float(buf[4 * n]) - ref["wkix"]) / abs(ref["wkix"])
        q.put((rank, err, wk_err))
    finally:
        dist.destroy_process_group()


def test_two_rank_moment_sum_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, err, wk_err in res:
        assert err < 1e-12 and wk_err < 1e-12, (rank, err, wk_err)


def test_ownership_helpers():
    import mrg_b200 as mrg
    for npr in (0, 1, 7, 8, 9, 1000):
        for size in (1, 2, 3, 8):
            counts = [mrg.owned_count(npr, ipar, size) for ipar in range(1, size + 1)]
            assert sum(counts) == npr
            idx = np.arange(npr)
            got = np.sort(np.concatenate([idx[mrg.owned_slice(ipar, size)] for ipar in range(1, size + 1)]))
            assert np.array_equal(got, idx)
            for ipar in range(1, size + 1):
                assert len(idx[mrg.owned_slice(ipar, size)]) == counts[ipar - 1]
