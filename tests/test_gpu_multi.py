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


def _worker_slab(rank, world, port, q):
    """The configuration bench.py runs for N > 1: device loader with z-slab ownership, restricted field
    preparation (planes = -1 -> on), deferred moment sums on the second stream, NCCL allreduce."""
    import torch
    import torch.distributed as dist
    import mrg_b200 as mrg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = mrg.broadcast_unique_id(rank)
        p = U.make_parm(12, 8, 48, Ez00=0.0)       # Ez00 = 0: the kick draws its random numbers but changes nothing,
        ppc = 8                                    # so the particles do not depend on which rank owns them (Q4)
        sp, ranfb = U.load_species(p, ppc)
        npr = len(sp[1][0])
        ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, rank=rank, nranks=world, device=rank)
        ctx.comm_init(uid)
        ctx.set_option("shard", 1)
        ctx.set_option("defer", 1)
        own = {}
        for k in (1, 2):
            ctx.loadpt(k, ppc, U.vth(k), 0.0, U.VBEAM[k])
            zc = (sp[k][2] + 0.5 * p.hz) / p.zmax * world
            own[k] = np.nonzero(np.clip(zc.astype(np.int64), 0, world - 1) == rank)[0]
            assert ctx.num_local(k) == len(own[k])
            ctx.sort(k, p.hdt)
        par = mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)
        ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
        errs = []
        st_gpu = ranfb
        for step in range(2):
            f12 = U.smooth_fields(p, seed=70 + step)
            a6 = O.field_prep(p, f12)
            ctx.set_fields(f12)
            wk = {k: ctx.fulmov_deferred(k, U.QSPEC[k], U.WSPEC[k], par) for k in (1, 2)}
            for k in (1, 2):
                r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=1)
                mom = ctx.moments(k)
                errs.append(max(U.rel_l2(mom[m], r["mom"][m]) for m in range(4)))
                errs.append(abs(wk[k][0].value - r["wkix"]) / abs(r["wkix"]))
            f12 = U.smooth_fields(p, seed=80 + step)
            a6 = O.field_prep(p, f12)
            ctx.set_fields(f12)
            for k in (1, 2):
                O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=1)
                _, _, st_gpu = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 0, par, st_gpu)
                ctx.sort(k, p.hdt)
        for k in (1, 2):
            got = ctx.download(k, len(own[k]))
            errs.append(U.particle_err(got, [a[own[k]] for a in ref[k]], p.hx, U.vth(k)) * 1e2 / 2)
        stats = ctx.prep_stats()
        # every preparation restricted to the slab; after the first step (where the ranks vote) the moments of
        # both species are summed slab-wise (halo strips + in-place allgather) instead of by a whole-grid allreduce
        ok = int(stats["restricted"] == stats["preps"] == 4 and stats["compact_sums"] == 2)
        if not ok:
            print("rank", rank, stats)
        ctx.close()
        q.put((rank, max(errs), ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_slab_deferred_restricted():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_slab, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, err, ok in res:
        assert err < 1e-10 and ok == 1, (rank, err, ok)
