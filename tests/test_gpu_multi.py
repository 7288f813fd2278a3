"""Multi-GPU tests of the rank sum of the moments (2, 4 and 8 ranks; each case needs that many devices: run with
`gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py -m gpu`).  Each rank is one process with one context;
the library's own ncclAllReduce replaces mpi_allreduce (F:2379-2384, 2533, 1312-1315)."""
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


def _spawn(worker, world, *extra):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q) + extra) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 4, 8])
def test_round_robin_ranks_nccl_moment_sum(world):
    """the reference's own partition (l = rank+1 mod N) with the reference's serial per-rank kick streams: moments,
    wkix/wkih, every owned particle and every rank's ranfp state against the oracle run as `world` ranks"""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    for rank, err, ok_rng in _spawn(_worker, world):
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
        p = U.make_parm(12, 8, 32 * world, Ez00=0.0)   # Ez00 = 0: the kick draws its random numbers but changes nothing,
        ppc = 8                                    # so the particles do not depend on which rank owns them (Q4)
        sp, ranfb = U.load_species(p, ppc)
        npr = len(sp[1][0])
        ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, rank=rank, nranks=world, device=rank)
        ctx.comm_init(uid)
        ctx.set_option("shard", 1)
        ctx.set_option("defer", 1)
        ctx.map_peers(2)                            # NVLink peer memory: the exchange ends with the fused add + push kernel
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
        # ... and the electrons' predictor of step 2 (last species, fresh order) ran as a split launch whose first part's
        # planes were pushed to the peers while the second part was still depositing
        ok = int(stats["restricted"] == stats["preps"] == 4 and stats["compact_sums"] == 2 and ctx.peer_pushes() == 2 and ctx.split_pushes() == 1)
        if not ok:
            print("rank", rank, stats, ctx.peer_pushes(), ctx.split_pushes())
        ctx.close()
        q.put((rank, max(errs), ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_ranks_deferred_restricted(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    for rank, err, ok in _spawn(_worker_slab, world):
        assert err < 1e-10 and ok == 1, (rank, err, ok)


def _worker_slab_kick(rank, world, port, q):
    """z-slab ownership WITH the drive (Ez00 != 0): the kick draws by particle index (no reference stream exists for this
    ownership), so the kicked set is not the oracle's -- but everything else must be: the summed moments before any kick,
    every coordinate except vy of every particle, vy of the particles that were not kicked, the size of a kick
    (Ez00 / bxa at the nearest node, F:1354) and the kick probability 0.001 per slab particle (F:1353)."""
    import torch
    import torch.distributed as dist
    import mrg_b200 as mrg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = mrg.broadcast_unique_id(rank)
        p = U.make_parm(16, 40, 16 * world)
        p0 = U.make_parm(16, 40, 16 * world, Ez00=0.0)
        ppc = 24
        sp, ranfb = U.load_species(p, ppc)
        ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, rank=rank, nranks=world, device=rank)
        ctx.comm_init(uid)
        ctx.set_option("shard", 1)
        own = {}
        for k in (1, 2):
            ctx.loadpt(k, ppc, U.vth(k), 0.0, U.VBEAM[k])
            zc = (sp[k][2] + 0.5 * p.hz) / p.zmax * world
            own[k] = np.nonzero(np.clip(zc.astype(np.int64), 0, world - 1) == rank)[0]
            ctx.sort(k, p.hdt)
        par = mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)
        f12 = U.smooth_fields(p, seed=90)
        a6 = O.field_prep(p, f12)
        ctx.set_fields(f12)
        errs, st = [], ranfb
        ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
        for k in (1, 2):
            r = O.fulmov(p, a6, *[a.copy() for a in sp[k]], U.QSPEC[k], U.WSPEC[k], 1, nranks=1)
            ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 1, par, st)
            mom = ctx.moments(k)
            errs.append(max(U.rel_l2(mom[m], r["mom"][m]) for m in range(4)))
        kicked = in_slab = 0
        for k in (1, 2):
            O.fulmov(p0, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=1)        # the same push without the kick
            _, _, st = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 0, par, st)
            got = ctx.download(k, len(own[k]))
            want = [a[own[k]] for a in ref[k]]
            for c in (0, 1, 2, 3, 5):
                fl = p.hx if c < 3 else U.vth(k)
                errs.append(float(np.max(np.abs(got[c] - want[c]) / np.maximum(np.abs(want[c]), fl))) * 1e2)
            dvy = got[4] - want[4]
            hit = np.abs(dvy) > 1e-12 * U.vth(k)
            y, z = want[1], want[2]
            slab = (np.abs(z - p.zcent) < 0.15 * p.zmax) & ((np.abs(y - p.ycent2) < 0.025 * p.ymax) | (np.abs(y - p.ycent1) < 0.025 * p.ymax))
            assert not np.any(hit & ~slab)                                          # only slab particles are kicked
            near2 = np.abs(y[hit] - p.ycent2) < 0.05 * p.ymax                        # vy -= vy0 near ycent2, += near ycent1
            bx = np.where(near2, -1.0, 1.0) * p.Ez00 / dvy[hit]                       # = bxa at the nearest node
            assert np.all((bx > 0.05) & (bx < 0.5)), bx                              # bxc = 0.2 +- the smooth modes
            kicked += int(hit.sum())
            in_slab += int(slab.sum())
        ctx.close()
        q.put((rank, max(errs), kicked, in_slab))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_ranks_drive_kick_statistics(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    res = _spawn(_worker_slab_kick, world)
    kicked = sum(r[2] for r in res)
    in_slab = sum(r[3] for r in res)
    for rank, err, _, _ in res:
        assert err < 1e-10, (rank, err)
    mean = 0.001 * in_slab
    assert in_slab > 20000 and abs(kicked - mean) < 5.0 * np.sqrt(mean) + 1, (kicked, in_slab)
