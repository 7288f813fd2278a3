"""2-GPU test of the NCCL moment sum (needs two devices: run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).  Each
rank is one process with one context; the library's own ncclAllReduce replaces
mpi_allreduce (F:2379-2384, 2533, 1312-1315)."""
import os
import socket

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import util as U

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import mrg_b200 as mrg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = mrg.broadcast_unique_id(rank)
        p = U.make_parm(16, 12, 16)
        sp, ranfb = U.load_species(p, 10)
        f12 = U.smooth_fields(p, seed=2)
        a6 = O.field_prep(p, f12)
        c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
        for name, arr in zip(mrg.host.FIELD_NAMES, f12):
            getattr(c, name)[:] = arr
        c.ranfb = ranfb
        fm = mrg.Fulmov(c, ipar=rank + 1, size=world, device=rank, uid=uid)
        npr = len(sp[1][0])
        host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
        errs = []
        st = np.full(world, ranfb, dtype=np.int32)
        ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
        for k in (1, 2):
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 1, k)
            r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=world)
            got = (c.qix, c.qiy, c.qiz, c.qi) if k == 1 else (c.qex, c.qey, c.qez, c.qe)
            errs.append(max(U.rel_l2(got[m], r["mom"][m]) for m in range(4)))
            errs.append(abs(c.wkix - r["wkix"]) / abs(r["wkix"]))
        for k in (1, 2):
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 0, k)
            r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=world, ranfb=st)
            errs.append(abs(c.wkix - r["wkix"]) / abs(r["wkix"]))
            fm.pull(k, *host[k], npr)
            sl = mrg.owned_slice(rank + 1, world)
            errs.append(U.particle_err([a[sl] for a in host[k]], [a[sl] for a in ref[k]], p.hx, U.vth(k)) * 1e2)
        ok_rng = int(c.ranfb == int(st[rank]))
        q.put((rank, max(errs), ok_rng))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_nccl_moment_sum():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, err, ok_rng in res:
        assert err < 1e-10 and ok_rng == 1, (rank, err, ok_rng)
