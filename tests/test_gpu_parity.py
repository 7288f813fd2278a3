"""GPU parity tests (cases C1-C10): the CUDA path, called
through the C ABI (ctypes over include/mrg_fulmov.h), against the CPU oracle
on identical seeded inputs.

Tolerances (BASELINE.json north_star): per-particle position/velocity
<= 1e-12 relative after one step (positions floored at one cell, velocities
at the thermal speed); deposited moments <= 1e-10 relative L2 (atomic
summation order differs); prepared fields and the synthetic loader are
bit-exact.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import util as U

pytestmark = pytest.mark.gpu

PTOL = 1e-12
MTOL = 1e-10


@pytest.fixture(scope="module")
def mrg():
    import mrg_b200
    mrg_b200.capi.load()
    return mrg_b200


def params_of(mrg, p, drive_on=True, ifil=(1, 1, 1)):
    return mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, ifil[0], ifil[1], ifil[2],
                          1 if drive_on else 0, p.Ez00, p.zcent, p.ycent1, p.ycent2)


def new_ctx(mrg, p, **kw):
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, **kw)
    if os.environ.get("MRG_TEST_TILE"):      # run the whole suite on another kernel family
        ctx.set_option("tile", int(os.environ["MRG_TEST_TILE"]))
    return ctx


@pytest.fixture(scope="module")
def case(mrg):
    """config-1-like load shrunk so the oracle finishes in seconds"""
    p = U.make_parm(16, 12, 16)
    sp, ranfb = U.load_species(p, 20)
    f12 = U.smooth_fields(p, seed=7)
    a6 = O.field_prep(p, f12)
    return p, sp, ranfb, f12, a6


# ---- a1, a12, a13: field preparation ---------------------------------------
@pytest.mark.parametrize("ifil", [(1, 1, 1), (0, 0, 0), (2, 1, 3)])
def test_prepared_fields_bit_exact(mrg, case, ifil):
    p, sp, ranfb, f12, _ = case
    po = U.make_parm(p.mx, p.my, p.mz)
    po.ifilx, po.ifily, po.ifilz = ifil
    a6 = O.field_prep(po, f12)
    ctx = new_ctx(mrg, p)
    ctx.set_fields(f12)
    got = ctx.prepared_fields(params_of(mrg, p, ifil=ifil))
    for c in range(6):
        np.testing.assert_array_equal(got[c], a6[c])
    ctx.close()


# ---- C1: corrector ------------------------------------------------------------
@pytest.mark.parametrize("ksp", [1, 2])
@pytest.mark.parametrize("sort,tile", [(False, 1), (True, 1), (True, 0), ("adt", 1)])
def test_corrector_particles(mrg, case, ksp, sort, tile):
    p, sp, ranfb, f12, a6 = case
    q, w = U.QSPEC[ksp], U.WSPEC[ksp]
    ref = [a.copy() for a in sp[ksp]]
    st = np.array([ranfb], dtype=np.int32)
    r = O.fulmov(p, a6, *ref, q, w, 0, nranks=1, ranfb=st)
    ctx = new_ctx(mrg, p)
    ctx.set_option("tile", tile)
    ctx.set_fields(f12)
    ctx.upload(ksp, *sp[ksp])
    if sort:
        ctx.sort(ksp, p.adt * 3 if sort == "adt" else p.hdt)
    wkix, wkih, st_gpu = ctx.fulmov(ksp, q, w, 0, params_of(mrg, p), ranfb)
    got = ctx.download(ksp, len(ref[0]))
    assert U.particle_err(got, ref, p.hx, U.vth(ksp)) < PTOL
    assert abs(wkix - r["wkix"]) < MTOL * abs(r["wkix"])
    assert abs(wkih - r["wkih"]) < MTOL * abs(r["wkih"])
    assert st_gpu == int(r["ranfb"][0])          # drive-kick RNG stream consumed identically
    ctx.close()


# ---- C2: predictor + deposit, every deposit mode, sorted and unsorted -----------
@pytest.mark.parametrize("ksp", [1, 2])
@pytest.mark.parametrize("deposit,iters,sort,tile", [(0, 8, False, 0), (1, 8, False, 0), (1, 8, True, 0), (2, 4, True, 0),
                                                     (2, 8, True, 0), (2, 8, False, 0), (2, 32, True, 0),
                                                     (2, 8, True, 1), (2, 8, "adt", 1), (2, 8, "stale", 1)])
def test_predictor_moments(mrg, case, ksp, deposit, iters, sort, tile):
    """sort: False = load order; True = sorted by the gather cell (x + hdt*v); "adt" = sorted by another
    key (many particles gather outside their tile); "stale" = sorted, then moved by a corrector step
    without re-sorting.  tile=1 runs the TMA-staged shared-memory kernels."""
    p, sp, ranfb, f12, a6 = case
    q, w = U.QSPEC[ksp], U.WSPEC[ksp]
    orig = [a.copy() for a in sp[ksp]]
    r = O.fulmov(p, a6, *[a.copy() for a in orig], q, w, 1, nranks=1, want_raw=True)
    ctx = new_ctx(mrg, p)
    ctx.set_option("deposit", deposit)
    ctx.set_option("iters", iters)
    ctx.set_option("tile", tile)
    ctx.set_fields(f12)
    if sort == "stale":
        # a sort index that no longer matches the positions: sort, push one step, restore order-independent truth
        ctx.upload(ksp, *orig)
        ctx.sort(ksp, p.hdt)
        ctx.fulmov(ksp, q, w, 0, params_of(mrg, p, drive_on=False))
        orig = ctx.download(ksp, len(orig[0]))
        r = O.fulmov(p, a6, *[a.copy() for a in orig], q, w, 1, nranks=1, want_raw=True)
    else:
        ctx.upload(ksp, *orig)
        if sort:
            ctx.sort(ksp, p.adt * 3 if sort == "adt" else p.hdt)
    wkix, wkih, _ = ctx.fulmov(ksp, q, w, 1, params_of(mrg, p))
    raw = ctx.moments(ksp, folded=False)
    mom = ctx.moments(ksp, folded=True)
    for c in range(4):
        assert U.rel_l2(raw[c], r["raw"][c]) < MTOL, (c, "raw")
        assert U.rel_l2(mom[c], r["mom"][c]) < MTOL, (c, "folded")
    n = len(orig[0])
    assert abs(raw[3].sum() - q * n) < 1e-9 * n          # invariant 1
    assert abs(wkix - r["wkix"]) < MTOL * abs(r["wkix"])
    assert abs(wkih - r["wkih"]) < MTOL * abs(r["wkih"])
    got = ctx.download(ksp, n)                              # predictor leaves x,v untouched
    for c in range(6):
        np.testing.assert_array_equal(got[c], orig[c])
    ctx.close()


# ---- C3: the call sequence of trans (F:664-706, 749-807) -------------------------
def test_step_sequence(mrg, case):
    p, sp, ranfb, f12_a, _ = case
    nsteps = 3
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.array([ranfb], dtype=np.int32)
    ctx = new_ctx(mrg, p)
    for k in (1, 2):
        ctx.upload(k, *sp[k])
    st_gpu = ranfb
    n = len(sp[1][0])

    def both(ipc, pp, f12, sort_after):
        nonlocal st_gpu
        a6 = O.field_prep(pp, f12)
        ctx.set_fields(f12)
        for k in (1, 2):
            q, w = U.QSPEC[k], U.WSPEC[k]
            r = O.fulmov(pp, a6, *ref[k], q, w, ipc, nranks=1, ranfb=st)
            wkix, wkih, st_gpu = ctx.fulmov(k, q, w, ipc, params_of(mrg, pp), st_gpu)
            assert abs(wkix - r["wkix"]) <= MTOL * abs(r["wkix"]), (ipc, k)
            if ipc >= 1:
                mom = ctx.moments(k)
                for c in range(4):
                    assert U.rel_l2(mom[c], r["mom"][c]) < MTOL, (ipc, k, c)
            else:
                assert st_gpu == int(st[0])
                if sort_after:
                    ctx.sort(k, pp.hdt)                  # uses the keys the tiled corrector emitted

    # it = 0: dt = adt = hdt = 0, moments only
    p0 = U.make_parm(p.mx, p.my, p.mz, dt=0.0)
    p0.adt = 0.0
    p0.hdt = 0.0
    both(1, p0, f12_a, False)
    for step in range(nsteps):
        f_pred = U.smooth_fields(p, seed=100 + step)
        both(1, p, f_pred, False)
        f_corr = U.smooth_fields(p, seed=200 + step)    # "emfild" changed the fields
        both(0, p, f_corr, step % 2 == 0)
        # accumulated drift of the two implementations stays at rounding level
        for k in (1, 2):
            got = ctx.download(k, n)
            assert U.particle_err(got, ref[k], p.hx, U.vth(k)) < 10 * PTOL * (step + 1), (step, k)
    ctx.close()


# ---- C4: seams, walls, cell boundaries ----------------------------------------------
@pytest.mark.parametrize("tile", [None, 1])
def test_edge_particles(mrg, case, tile):
    """tile=None: load order (any-order kernels); 1: sorted, tiled kernels"""
    p, sp, ranfb, f12, a6 = case
    rng = np.random.default_rng(42)
    n = 4096
    x = rng.uniform(-p.hx / 2, p.xmax - p.hx / 2, n)
    y = rng.uniform(0, p.ymax, n)
    z = rng.uniform(-p.hz / 2, p.zmax - p.hz / 2, n)
    v = [rng.normal(scale=0.3, size=n) for _ in range(3)]
    # exact seam / wall / cell-boundary positions and their neighbours by one ulp
    specials_x = [np.nextafter(-p.hx / 2, 1), np.nextafter(p.xmax - p.hx / 2, 0), 0.5 * p.hx, np.nextafter(0.5 * p.hx, 0),
                  np.nextafter(0.5 * p.hx, 1), 0.0, 1.5 * p.hx]
    for q, val in enumerate(specials_x):
        x[q] = val
        z[q + 16] = val * p.hz / p.hx
    y[32:40] = [np.nextafter(0, 1), np.nextafter(p.ymax, 0), p.hy, np.nextafter(p.hy, 0), np.nextafter(p.hy, 1),
                p.ymax - 1e-9, 1e-9, (p.my - 1) * p.hy]
    v[1][32] = -0.5; v[1][33] = 0.5; v[1][37] = 0.5; v[1][38] = -0.5   # reflect off the walls
    v[0][0] = -0.5; v[0][1] = 0.5; v[2][16] = -0.5; v[2][17] = 0.5      # wrap through the seams
    arrs = [x, y, z] + v
    pz = U.make_parm(p.mx, p.my, p.mz, Ez00=0.0)
    for ksp in (1, 2):
        q, w = U.QSPEC[ksp], U.WSPEC[ksp]
        ref = [a.copy() for a in arrs]
        O.fulmov(pz, a6, *ref, q, w, 0, nranks=1)
        r1 = O.fulmov(p, a6, *[a.copy() for a in arrs], q, w, 1, nranks=1, want_raw=True)
        ctx = new_ctx(mrg, p)
        ctx.set_fields(f12)
        ctx.upload(ksp, *arrs)
        if tile is not None:
            ctx.set_option("tile", tile)
            ctx.sort(ksp, p.hdt)
        ctx.fulmov(ksp, q, w, 1, params_of(mrg, p))
        raw = ctx.moments(ksp, folded=False)
        for c in range(4):
            assert U.rel_l2(raw[c], r1["raw"][c]) < MTOL
        ctx.fulmov(ksp, q, w, 0, params_of(mrg, pz, drive_on=False))
        got = ctx.download(ksp, n)
        assert U.particle_err(got, ref, p.hx, 0.3) < PTOL
        # wall reflections flipped vy exactly like partbc
        assert np.array_equal(np.sign(got[4][32:40]), np.sign(ref[4][32:40]))
        ctx.close()


# ---- tiled kernels on grids whose x size is not a multiple of the 32-cell tile, thin and fat cells ---------
@pytest.mark.parametrize("tile", [1])
@pytest.mark.parametrize("mx,ppc", [(40, 7), (17, 33), (8, 70), (70, 3)])
def test_tiled_kernels_ragged_tiles(mrg, mx, ppc, tile):
    p = U.make_parm(mx, 6, 8)
    sp, ranfb = U.load_species(p, ppc)
    f12 = U.smooth_fields(p, seed=11)
    a6 = O.field_prep(p, f12)
    for ksp in (1, 2):
        q, w = U.QSPEC[ksp], U.WSPEC[ksp]
        n = len(sp[ksp][0])
        ctx = new_ctx(mrg, p)
        ctx.set_option("tile", tile)
        ctx.set_fields(f12)
        ctx.upload(ksp, *sp[ksp])
        ctx.sort(ksp, p.hdt)
        r = O.fulmov(p, a6, *[a.copy() for a in sp[ksp]], q, w, 1, nranks=1, want_raw=True)
        wkix, wkih, _ = ctx.fulmov(ksp, q, w, 1, params_of(mrg, p))
        raw = ctx.moments(ksp, folded=False)
        for c in range(4):
            assert U.rel_l2(raw[c], r["raw"][c]) < MTOL, c
        assert abs(raw[3].sum() - q * n) < 1e-9 * n
        assert abs(wkix - r["wkix"]) < MTOL * abs(r["wkix"])
        ref = [a.copy() for a in sp[ksp]]
        st = np.array([ranfb], dtype=np.int32)
        r0 = O.fulmov(p, a6, *ref, q, w, 0, nranks=1, ranfb=st)
        _, _, st_gpu = ctx.fulmov(ksp, q, w, 0, params_of(mrg, p), ranfb)
        got = ctx.download(ksp, n)
        assert U.particle_err(got, ref, p.hx, U.vth(ksp)) < PTOL
        assert st_gpu == int(r0["ranfb"][0])
        # the keys the corrector emitted drive the next sort; a second step stays on the oracle
        ctx.sort(ksp, p.hdt)
        r = O.fulmov(p, a6, *[a.copy() for a in ref], q, w, 1, nranks=1, want_raw=True)
        ctx.fulmov(ksp, q, w, 1, params_of(mrg, p))
        raw = ctx.moments(ksp, folded=False)
        for c in range(4):
            assert U.rel_l2(raw[c], r["raw"][c]) < MTOL, (c, "step 2")
        ctx.close()


# ---- C5: pure rotation, up to the large-Dt regime -------------------------------------
@pytest.mark.parametrize("bmag", [0.1, 1.0, 10.0, 100.0])
def test_pure_rotation(mrg, bmag):
    p = U.make_parm(8, 6, 8, wce=0.0)
    n = O.mxyzA(p)
    f12 = [np.zeros(n) for _ in range(12)]
    for c, b in zip((3, 4, 5), (0.6 * bmag, -0.48 * bmag, 0.64 * bmag)):
        f12[c][:] = b
        f12[c + 6][:] = b
    rng = np.random.default_rng(5)
    N = 5000
    x = rng.uniform(0.2 * p.xmax, 0.8 * p.xmax, N)
    y = rng.uniform(2.1 * p.hy, 3.9 * p.hy, N)   # rows 2..4: untouched by the wall mirror of filt3e
    z = rng.uniform(0.2 * p.zmax, 0.8 * p.zmax, N)
    v = [rng.normal(scale=0.01, size=N) for _ in range(3)]
    v0 = np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    ctx = new_ctx(mrg, p)
    ctx.set_fields(f12)
    ctx.upload(1, x, y, z, *v)
    ctx.fulmov(1, 1.0, 1.0, 0, params_of(mrg, p, drive_on=False))
    got = ctx.download(1, N)
    v1 = np.sqrt(got[3] ** 2 + got[4] ** 2 + got[5] ** 2)
    assert np.max(np.abs(v1 - v0) / v0) < 1e-13
    # rotation angle 2*atan(ht*|B|) about B (closed form of F:1278-1280)
    ht = 0.5 * p.dt
    bhat = np.array([0.6, -0.48, 0.64])
    vpar0 = v[0] * bhat[0] + v[1] * bhat[1] + v[2] * bhat[2]
    vpar1 = got[3] * bhat[0] + got[4] * bhat[1] + got[5] * bhat[2]
    assert np.max(np.abs(vpar1 - vpar0)) < 1e-13 * 0.01 * 100
    perp0 = np.stack(v) - np.outer(bhat, vpar0)
    perp1 = np.stack(got[3:6]) - np.outer(bhat, vpar1)
    cosang = (perp0 * perp1).sum(0) / ((perp0 ** 2).sum(0))
    t = ht * bmag
    assert np.max(np.abs(cosang - (1 - t * t) / (1 + t * t))) < 1e-9
    ctx.close()


# ---- C6: the it=0 call -------------------------------------------------------------------
def test_zero_dt_identity(mrg, case):
    p, sp, ranfb, f12, a6 = case
    p0 = U.make_parm(p.mx, p.my, p.mz, dt=0.0)
    p0.adt = 0.0
    p0.hdt = 0.0
    ksp = 2
    ctx = new_ctx(mrg, p)
    ctx.set_fields(f12)
    ctx.upload(ksp, *sp[ksp])
    ctx.fulmov(ksp, -1.0, 1.0, 0, params_of(mrg, p0, drive_on=False))
    got = ctx.download(ksp, len(sp[ksp][0]))
    for c in range(6):
        np.testing.assert_array_equal(got[c], sp[ksp][c])       # bit-identical
    r = O.fulmov(p0, a6, *[a.copy() for a in sp[ksp]], -1.0, 1.0, 1, nranks=1)
    ctx.fulmov(ksp, -1.0, 1.0, 1, params_of(mrg, p0))
    mom = ctx.moments(ksp)
    for c in range(4):
        assert U.rel_l2(mom[c], r["mom"][c]) < MTOL
    ctx.close()


# ---- C8: drive kick with 1/2/4/8 simulated ranks -------------------------------------------
@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_drive_kick_rank_streams(mrg, case, nranks):
    p, sp, ranfb, f12, a6 = case
    ksp = 2
    q, w = U.QSPEC[ksp], U.WSPEC[ksp]
    pk = U.make_parm(p.mx, p.my, p.mz, Ez00=0.25)      # a kick large enough to see
    ref = [a.copy() for a in sp[ksp]]
    st = np.full(nranks, ranfb, dtype=np.int32)
    O.fulmov(pk, a6, *ref, q, w, 0, nranks=nranks, ranfb=st)
    npr = len(ref[0])
    got = [np.zeros(npr) for _ in range(6)]
    nkicked = 0
    for r in range(nranks):                             # each rank = its own context (owned subset)
        ctx = new_ctx(mrg, p)
        ctx.set_fields(f12)
        ctx.upload(ksp, *sp[ksp], first=r + 1, stride=nranks)
        if r % 2 == 1:
            ctx.sort(ksp, pk.adt)                       # device order must not matter
        _, _, st_gpu = ctx.fulmov(ksp, q, w, 0, params_of(mrg, pk), ranfb)
        assert st_gpu == int(st[r])
        ctx.download(ksp, npr, first=r + 1, stride=nranks, out=got)
        ctx.close()
    assert U.particle_err(got, ref, p.hx, U.vth(ksp)) < PTOL
    # the kicked set is identical: a kick changes vy by Ez00/bxa ~ 1, far above PTOL
    p_nokick = U.make_parm(p.mx, p.my, p.mz, Ez00=0.0)
    base = [a.copy() for a in sp[ksp]]
    O.fulmov(p_nokick, a6, *base, q, w, 0, nranks=1)
    kicked_ref = np.abs(ref[4] - base[4]) > 1e-3
    kicked_gpu = np.abs(got[4] - base[4]) > 1e-3
    assert kicked_ref.sum() > 0 and np.array_equal(kicked_ref, kicked_gpu)


# ---- C10 on one GPU: moments are ownership-independent ---------------------------------------
@pytest.mark.parametrize("nranks", [2, 4])
def test_sharded_moments_sum(mrg, case, nranks):
    p, sp, ranfb, f12, a6 = case
    ksp = 1
    q, w = U.QSPEC[ksp], U.WSPEC[ksp]
    r = O.fulmov(p, a6, *[a.copy() for a in sp[ksp]], q, w, 1, nranks=nranks, want_raw=True)
    raw_sum = [np.zeros(O.mxyzA(p)) for _ in range(4)]
    wk = 0.0
    for rank in range(nranks):
        ctx = new_ctx(mrg, p)
        ctx.set_fields(f12)
        ctx.upload(ksp, *sp[ksp], first=rank + 1, stride=nranks)
        ctx.sort(ksp, p.adt)
        wkix, _, _ = ctx.fulmov(ksp, q, w, 1, params_of(mrg, p))
        wk += wkix
        part = ctx.moments(ksp, folded=False)
        for c in range(4):
            raw_sum[c] += part[c]
        ctx.close()
    for c in range(4):
        assert U.rel_l2(raw_sum[c], r["raw"][c]) < MTOL
    assert abs(wk - r["wkix"]) < MTOL * abs(r["wkix"])


# ---- maintenance: sort keeps the particle set and restores the order on download ---------------
def test_sort_and_download_order(mrg, case):
    p, sp, ranfb, f12, a6 = case
    ksp = 2
    ctx = new_ctx(mrg, p)
    ctx.upload(ksp, *sp[ksp])
    n = len(sp[ksp][0])
    for la in (0.0, p.adt, 0.0):
        ctx.sort(ksp, la)                                   # repeated sorts compose
        got = ctx.download(ksp, n)
        for c in range(6):
            np.testing.assert_array_equal(got[c], sp[ksp][c])
    ctx.close()


# ---- synthetic loader: bit-exact with loadpt, for any ownership ----------------------------------
@pytest.mark.parametrize("nranks", [1, 3])
def test_device_loadpt_bit_exact(mrg, nranks):
    p = U.make_parm(12, 10, 8)
    ppc = 9
    npr = p.mx * p.my * p.mz * ppc
    for ksp in (1, 2):
        arrs, a, b = O.loadpt(p, ppc, U.vth(ksp), 0.0, U.VBEAM[ksp])
        got = [np.zeros(npr) for _ in range(6)]
        for rank in range(nranks):
            ctx = new_ctx(mrg, p, rank=rank, nranks=nranks)
            ga, gb = ctx.loadpt(ksp, ppc, U.vth(ksp), 0.0, U.VBEAM[ksp])
            assert (ga, gb) == (a, b)
            ctx.download(ksp, npr, first=rank + 1, stride=nranks, out=got)
            ctx.close()
        for c in range(6):
            np.testing.assert_array_equal(got[c], arrs[c])


# ---- z-slab ownership of the device loader: same particle set, same summed moments --------------
@pytest.mark.parametrize("nranks", [2, 3])
def test_device_loadpt_slab_ownership(mrg, nranks):
    p = U.make_parm(12, 10, 16)
    ppc = 9
    npr = p.mx * p.my * p.mz * ppc
    f12 = U.smooth_fields(p, seed=5)
    a6 = O.field_prep(p, f12)
    ksp = 2
    q, w = U.QSPEC[ksp], U.WSPEC[ksp]
    arrs, a, b = O.loadpt(p, ppc, U.vth(ksp), 0.0, U.VBEAM[ksp])
    r = O.fulmov(p, a6, *[x.copy() for x in arrs], q, w, 1, nranks=1, want_raw=True)
    raw_sum = [np.zeros(O.mxyzA(p)) for _ in range(4)]
    total, zs = 0, []
    for rank in range(nranks):
        ctx = new_ctx(mrg, p, rank=rank, nranks=nranks)
        ctx.set_option("shard", 1)
        ga, gb = ctx.loadpt(ksp, ppc, U.vth(ksp), 0.0, U.VBEAM[ksp])
        assert (ga, gb) == (a, b)
        n = ctx.num_local(ksp)
        total += n
        got = ctx.download(ksp, n)                  # local order = increasing l
        lo, hi = -p.hz / 2 + p.zmax * rank / nranks, -p.hz / 2 + p.zmax * (rank + 1) / nranks
        assert np.all(got[2] >= lo - 1e-9) and np.all(got[2] < hi + 1e-9)
        zs.append(got[2])
        ctx.close()
        one = new_ctx(mrg, p)                       # the rank's share pushed without a communicator
        one.set_fields(f12)
        one.upload(ksp, *got)
        one.sort(ksp, p.hdt)
        one.fulmov(ksp, q, w, 1, params_of(mrg, p))
        part = one.moments(ksp, folded=False)
        for c in range(4):
            raw_sum[c] += part[c]
        one.close()
    assert total == npr
    np.testing.assert_array_equal(np.sort(np.concatenate(zs)), np.sort(arrs[2]))
    for c in range(4):
        assert U.rel_l2(raw_sum[c], r["raw"][c]) < MTOL


# ---- error behaviour of the C ABI --------------------------------------------------------------------
def test_abi_errors(mrg, case):
    p, sp, ranfb, f12, a6 = case
    ctx = new_ctx(mrg, p)
    with pytest.raises(mrg.MrgError):
        ctx.fulmov(1, 1.0, 100.0, 1, params_of(mrg, p))       # fields not set
    ctx.set_fields(f12)
    with pytest.raises(mrg.MrgError):
        ctx.fulmov(3, 1.0, 100.0, 1, params_of(mrg, p))       # species out of range
    with pytest.raises(mrg.MrgError):
        ctx.moments(1)                                         # nothing deposited yet
    with pytest.raises(mrg.MrgError):
        ctx.set_option("no_such_option", 1)
    # an empty species is legal: zero moments
    ctx.fulmov(1, 1.0, 100.0, 1, params_of(mrg, p))
    assert all(np.all(m == 0) for m in ctx.moments(1))
    ctx.close()


# ---- BASELINE configs[4]: three species, dt * wce > 10 ------------------------------------------------
def test_three_species_large_dt(mrg):
    """a heavy positive third species (q = +1, m = 1600; qspec(4), wspec(4) exist in the reference, F:1100, although trans
    only calls fulmov for two) next to ions and electrons, with wce/wpe = 9 so that dt * wce = 10.8 (ht * |B| ~ 5 for the
    electrons): two full steps through the C ABI and through the Python mirror with nspecies = 3, against the oracle (which
    is pinned to the reference in this regime by tests/test_ref_pin.py::test_large_dt_heavy_species_bit_identical)"""
    p = O.make_parm(24, 10, 12, U.HX * 24, U.HY * 10, U.HZ * 12, 1.2, 0.6, 9.0, 0.25e-2)
    Q = {1: 1.0, 2: -1.0, 3: 1.0}
    W = {1: 100.0, 2: 1.0, 3: 1600.0}
    sp, ranfb = U.load_species(p, 12)
    arrs3, _, _ = O.loadpt(p, 12, U.VETH / np.sqrt(W[3]), 0.0, 0.0)
    sp[3] = arrs3
    n = len(sp[1][0])
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2, 3)}
    st = np.array([ranfb], dtype=np.int32)
    c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, dt=p.dt, aimpl=p.aimpl, wce_by_wpe=p.bxc, Ez00=p.Ez00)
    c.ranfb, c.it = ranfb, 1
    fm = mrg.Fulmov(c, ipar=1, size=1, nspecies=3)
    with pytest.raises(ValueError):
        mrg.Fulmov(c, ipar=1, size=1, ctx=fm.ctx)(*sp[3], Q[3], W[3], n, 1, 3)      # the default mirror keeps the reference's ksp = 1|2
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2, 3)}
    FN = mrg.host.FIELD_NAMES
    for step in range(2):
        f12 = U.smooth_fields(p, seed=600 + 2 * step, amp_b=0.5)
        for name, a in zip(FN, f12):
            getattr(c, name)[:] = a
        fm.fields_changed()
        a6 = O.field_prep(p, f12)
        for k in (1, 2, 3):
            r = O.fulmov(p, a6, *ref[k], Q[k], W[k], 1, nranks=1, ranfb=st)
            fm(*host[k], Q[k], W[k], n, 1, k)
            got = fm._moment_arrays(k)
            for ci in range(4):
                assert U.rel_l2(got[ci], r["mom"][ci]) < MTOL, (step, k, ci)
            assert abs(c.wkix - r["wkix"]) < MTOL * abs(r["wkix"])
        f12 = U.smooth_fields(p, seed=601 + 2 * step, amp_b=0.5)
        for name, a in zip(FN, f12):
            getattr(c, name)[:] = a
        fm.fields_changed()
        a6 = O.field_prep(p, f12)
        for k in (1, 2, 3):
            O.fulmov(p, a6, *ref[k], Q[k], W[k], 0, nranks=1, ranfb=st)
            fm(*host[k], Q[k], W[k], n, 0, k)
        assert c.ranfb == int(st[0])
    for k in (1, 2, 3):
        fm.pull(k, *host[k], n)
        vfl = U.VETH / np.sqrt(W[k]) if Q[k] > 0 else U.VETH
        assert U.particle_err(host[k], ref[k], p.hx, vfl) < 2 * PTOL, k
    fm.ctx.close()
