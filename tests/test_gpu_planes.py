"""GPU parity of the multi-GPU plumbing that changes WHERE work runs but must
not change any result: restricted field preparation (option "planes": only the
z planes near the rank's particles are blended/filtered/ghost-filled), the
deferred mode (option "defer": moment sum + fold + D2H on a second stream),
zero-copy field binding and the host's field-change hints.  All against the
CPU oracle on the same seeded inputs, tolerances as in test_gpu_parity.py
(particles 1e-12 relative, moments 1e-10 relative L2)."""
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


def params_of(mrg, p):
    return mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)


def slab_subset(p, sp, zlo, zhi):
    """the particles of both species whose z lies in [zlo, zhi) cells (wrapping when zlo > zhi): what a rank
    that owns a z slab holds"""
    out = {}
    for k in (1, 2):
        zc = (sp[k][2] + 0.5 * p.hz) / p.hz
        m = ((zc >= zlo) & (zc < zhi)) if zlo <= zhi else ((zc >= zlo) | (zc < zhi))
        out[k] = [np.ascontiguousarray(a[m]) for a in sp[k]]
    return out


@pytest.mark.parametrize("zlo,zhi", [(6.0, 11.0), (20.5, 3.5), (0.0, 5.0)])
def test_restricted_preparation_matches_oracle(mrg, zlo, zhi):
    """A slab of particles (interior, across the periodic seam, at the seam) through three full steps with new
    fields in every phase: same particles and moments as the oracle, which prepares every plane."""
    p = U.make_parm(12, 10, 24)
    sp_all, ranfb = U.load_species(p, 12)
    sp = slab_subset(p, sp_all, zlo, zhi)
    n = {k: len(sp[k][0]) for k in (1, 2)}
    assert min(n.values()) > 500
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.array([ranfb], dtype=np.int32)
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    ctx.set_option("planes", 1)
    for k in (1, 2):
        ctx.upload(k, *sp[k])
        ctx.sort(k, p.hdt)
    st_gpu = ranfb
    ctx.prep_stats(reset=True)
    for step in range(3):
        for ipc in (1, 0):
            f12 = U.smooth_fields(p, seed=300 + 10 * step + ipc)
            a6 = O.field_prep(p, f12)
            ctx.set_fields(f12)
            for k in (1, 2):
                q, w = U.QSPEC[k], U.WSPEC[k]
                r = O.fulmov(p, a6, *ref[k], q, w, ipc, nranks=1, ranfb=st)
                wkix, wkih, st_gpu = ctx.fulmov(k, q, w, ipc, params_of(mrg, p), st_gpu)
                assert np.isfinite(wkix) and abs(wkix - r["wkix"]) <= MTOL * abs(r["wkix"]), (step, ipc, k)
                if ipc >= 1:
                    mom = ctx.moments(k)
                    for c in range(4):
                        assert U.rel_l2(mom[c], r["mom"][c]) < MTOL, (step, k, c)
                else:
                    ctx.sort(k, p.hdt)
        assert st_gpu == int(st[0])
        for k in (1, 2):
            got = ctx.download(k, n[k])
            assert U.particle_err(got, ref[k], p.hx, U.vth(k)) < 10 * PTOL * (step + 1), (step, k)
    stats = ctx.prep_stats()
    assert stats["preps"] == 6 and stats["restricted"] == 6, stats
    assert stats["planes"] < 6 * (p.mz + 4) * 0.75, stats         # well under the full grid
    # the full preparation is still what the host sees
    got = ctx.prepared_fields(params_of(mrg, p))
    for c in range(6):
        np.testing.assert_array_equal(got[c], a6[c])
    ctx.close()


def test_restricted_preparation_follows_new_particles(mrg):
    """Particles uploaded after a restricted preparation lie on other planes: the cached preparation must not be
    reused for them."""
    p = U.make_parm(12, 10, 24)
    sp_all, ranfb = U.load_species(p, 8)
    a = slab_subset(p, sp_all, 2.0, 6.0)
    b = slab_subset(p, sp_all, 14.0, 19.0)
    f12 = U.smooth_fields(p, seed=5)
    a6 = O.field_prep(p, f12)
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    ctx.set_option("planes", 1)
    ctx.set_fields(f12)
    ctx.upload(1, *a[1]); ctx.sort(1, p.hdt)
    ctx.fulmov(1, U.QSPEC[1], U.WSPEC[1], 1, params_of(mrg, p))
    ctx.upload(2, *b[2]); ctx.sort(2, p.hdt)          # same field version, other planes
    ref = [x.copy() for x in b[2]]
    r = O.fulmov(p, a6, *ref, U.QSPEC[2], U.WSPEC[2], 1, nranks=1)
    ctx.fulmov(2, U.QSPEC[2], U.WSPEC[2], 1, params_of(mrg, p))
    mom = ctx.moments(2)
    for c in range(4):
        assert U.rel_l2(mom[c], r["mom"][c]) < MTOL, c
    # unsorted upload: no plane information, full preparation
    ctx.upload(1, *b[1])
    ctx.set_fields(f12)
    ctx.prep_stats(reset=True)
    ctx.fulmov(1, U.QSPEC[1], U.WSPEC[1], 1, params_of(mrg, p))
    assert ctx.prep_stats()["restricted"] == 0
    ctx.close()


@pytest.mark.parametrize("planes", [0, 1])
def test_deferred_mode_and_field_hints(mrg, planes):
    """The Fulmov mirror with defer=True (moment sum, fold and D2H on the second stream) and hints=True (only
    the members of COMMON /fields/ the host marked are uploaded; ex0 <- ex repeated on the device) over three
    steps of the trans protocol: prefld changes bx..bz, emfild changes ex..bz, then the renewal."""
    p = U.make_parm(12, 10, 16)
    sp, ranfb = U.load_species(p, 10)
    npr = len(sp[1][0])
    c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, dt=p.dt, aimpl=p.aimpl, wce_by_wpe=p.bxc, Ez00=p.Ez00)
    c.ranfb = ranfb
    c.it, c.nha, c.ldec = 5, 5, 2
    f0 = U.smooth_fields(p, seed=40, ghost_nan=False)
    for name, arr in zip(mrg.host.FIELD_NAMES, f0):
        getattr(c, name)[:] = arr
    fm = mrg.Fulmov(c, ipar=1, size=1, hints=True, defer=True)
    fm.ctx.set_option("planes", planes)
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.array([ranfb], dtype=np.int32)
    FN = mrg.host.FIELD_NAMES
    for step in range(3):
        newb = U.smooth_fields(p, seed=50 + step, ghost_nan=False)
        for i in (3, 4, 5):                                  # prefld: bx, by, bz
            getattr(c, FN[i])[:] = newb[i]
        fm.fields_changed(fm.MASK_B)
        a6 = O.field_prep(p, c.fields())
        rr = {}
        for k in (1, 2):
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 1, k)
            rr[k] = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=1, ranfb=st)
        fm.finish_moments()                                  # before "emfild" reads /srimp7/
        for k, got in ((1, (c.qix, c.qiy, c.qiz, c.qi)), (2, (c.qex, c.qey, c.qez, c.qe))):
            for m in range(4):
                assert U.rel_l2(got[m], rr[k]["mom"][m]) < MTOL, (step, k, m)
        assert abs(c.wkix - rr[2]["wkix"]) <= MTOL * abs(rr[2]["wkix"])
        assert c.edec[4, c.ldec - 1] != 0.0 and abs(c.edec[4, c.ldec - 1] - rr[1]["wkix"]) <= MTOL * abs(rr[1]["wkix"])
        newf = U.smooth_fields(p, seed=60 + step, ghost_nan=False)
        for i in range(6):                                   # emfild: ex..bz
            getattr(c, FN[i])[:] = newf[i]
        fm.fields_changed(fm.MASK_NEW)
        a6 = O.field_prep(p, c.fields())
        for k in (1, 2):
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 0, k)
            O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=1, ranfb=st)
        for i in range(6):                                   # renewal, F:796-807
            getattr(c, FN[i + 6])[:] = getattr(c, FN[i])
        fm.fields_renewed()
        assert c.ranfb == int(st[0])
    for k in (1, 2):
        fm.pull(k, *host[k], npr)
        assert U.particle_err(host[k], ref[k], p.hx, U.vth(k)) < 30 * PTOL, k
    cnt = fm.ctx.counters()
    # all 12 arrays once, 6 for the first corrector, then 3 + 6 per step: never the whole of /fields/ again
    assert cnt["h2d_bytes"] - 2 * 6 * npr * 8 == (12 + 6 + 2 * 9) * fm.ctx.n_grid * 8, cnt
    fm.ctx.close()


def test_bind_fields_device_is_zero_copy_and_identical(mrg):
    import torch
    p = U.make_parm(12, 10, 16)
    sp, ranfb = U.load_species(p, 10)
    f12 = U.smooth_fields(p, seed=9, ghost_nan=False)
    a6 = O.field_prep(p, f12)
    dev = [torch.tensor(a, dtype=torch.float64, device="cuda:0") for a in f12]
    torch.cuda.synchronize()
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    ctx.bind_fields_device([t.data_ptr() for t in dev])
    got = ctx.prepared_fields(params_of(mrg, p))
    for c in range(6):
        np.testing.assert_array_equal(got[c], a6[c])
    # the bound arrays are read in place: changing one and re-binding changes the result, no copy was kept
    dev[0].mul_(2.0)
    torch.cuda.synchronize()
    ctx.bind_fields_device([t.data_ptr() for t in dev], mask=0x1)
    f12b = [a.copy() for a in f12]
    f12b[0] *= 2.0
    a6b = O.field_prep(p, f12b)
    got = ctx.prepared_fields(params_of(mrg, p))
    np.testing.assert_array_equal(got[0], a6b[0])
    # renewal on bound arrays copies them into the context's own ex0..bz0
    ctx.renew_fields()
    f12c = f12b[:6] + [a.copy() for a in f12b[:6]]
    a6c = O.field_prep(p, f12c)
    got = ctx.prepared_fields(params_of(mrg, p))
    for c in range(6):
        np.testing.assert_array_equal(got[c], a6c[c])
    assert ctx.counters()["h2d_bytes"] == 0
    ctx.close()


def test_inline_drive_kick_by_particle_index(mrg):
    """Option "kick" = 1 (what z-slab ownership uses): particle id draws ranfp number id+1 of the stream at *ranfb and
    is kicked inside the tiled corrector.  Checked against the oracle's corrector without kick (Ez00 = 0) plus the
    kick of F:1342-1364 applied here with that draw rule; *ranfb advances by the number of particles."""
    p = U.make_parm(16, 12, 16, Ez00=0.25)                 # a large Ez00 makes a missed or spurious kick obvious
    p0 = U.make_parm(16, 12, 16, Ez00=0.0)
    sp, ranfb = U.load_species(p, 40)
    f12 = U.smooth_fields(p, seed=21)
    a6 = O.field_prep(p, f12)
    bxa = a6[3]
    ksp = 2
    q, w = U.QSPEC[ksp], U.WSPEC[ksp]
    ref = [a.copy() for a in sp[ksp]]
    O.fulmov(p0, a6, *ref, q, w, 0, nranks=1)
    n = len(ref[0])
    x, y, z, vy = ref[0], ref[1], ref[2], ref[4]
    in_slab = (np.abs(z - p.zcent) < 0.15 * p.zmax) & ((np.abs(y - p.ycent2) < 0.025 * p.ymax) | (np.abs(y - p.ycent1) < 0.025 * p.ymax))
    nk = 0
    for l in np.nonzero(in_slab)[0]:
        u = O.lcg_skip(ranfb, int(l) + 1) / 2147483648.0
        if u > 0.999:
            ip, jp, kp = int(p.hxi * x[l] + 0.500000001), int(p.hyi * y[l] + 0.000000001), int(p.hzi * z[l] + 0.500000001)
            vy0 = p.Ez00 / bxa[U.idx(p, ip, jp, kp)]
            if abs(y[l] - p.ycent2) < 0.05 * p.ymax:
                vy[l] -= vy0
            elif abs(y[l] - p.ycent1) < 0.05 * p.ymax:
                vy[l] += vy0
            nk += 1
    assert nk >= 2, nk
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    ctx.set_option("kick", 1)
    ctx.set_fields(f12)
    ctx.upload(ksp, *sp[ksp])
    ctx.sort(ksp, p.hdt)
    _, _, st = ctx.fulmov(ksp, q, w, 0, params_of(mrg, p), ranfb)
    got = ctx.download(ksp, n)
    assert U.particle_err(got, ref, p.hx, U.vth(ksp)) < PTOL
    assert st == O.lcg_skip(ranfb, n)
    ctx.close()


@pytest.mark.parametrize("planes", [0, 1])
def test_lazy_host_fields_follow_the_trans_protocol(mrg, planes):
    """mrg_set_fields_lazy + mrg_renew_fields_host through the Python mirror of fulmov, with the three field marks of trans
    (prefld, emfild, renewal): a slab of particles steps three times; every result matches the oracle, which sees whole
    arrays, and with restricted preparation the rank uploads well under the replicated arrays."""
    p = U.make_parm(12, 10, 24)
    sp_all, ranfb = U.load_species(p, 10)
    sp = slab_subset(p, sp_all, 7.0, 12.0)
    n = {k: len(sp[k][0]) for k in (1, 2)}
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.array([ranfb], dtype=np.int32)
    c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, dt=p.dt, aimpl=p.aimpl, wce_by_wpe=p.bxc, Ez00=p.Ez00)
    c.ranfb, c.it = ranfb, 1
    fm = mrg.Fulmov(c, ipar=1, size=1, hints=True, lazy=True)
    fm.ctx.set_option("planes", planes)
    FN = mrg.host.FIELD_NAMES
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    f_old = U.smooth_fields(p, seed=400, ghost_nan=False)
    for name, a in zip(FN, f_old):
        setattr(c, name, a.copy())
    fm.ctx.counters(reset=True)
    for step in range(3):
        # prefld: new bx,by,bz (F:759)
        f_b = U.smooth_fields(p, seed=410 + step, ghost_nan=False)
        for i in (3, 4, 5):
            setattr(c, FN[i], f_b[i].copy())
        fm.fields_changed(fm.MASK_B)
        a6 = O.field_prep(p, c.fields())
        mom = {}
        for k in (1, 2):
            r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=1, ranfb=st)
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], n[k], 1, k)
            got = (c.qix, c.qiy, c.qiz, c.qi) if k == 1 else (c.qex, c.qey, c.qez, c.qe)
            for ci in range(4):
                assert U.rel_l2(got[ci], r["mom"][ci]) < MTOL, (step, k, ci)
        # emfild: new ex..bz (F:771)
        f_n = U.smooth_fields(p, seed=420 + step, ghost_nan=False)
        for i in range(6):
            setattr(c, FN[i], f_n[i].copy())
        fm.fields_changed(fm.MASK_NEW)
        a6 = O.field_prep(p, c.fields())
        for k in (1, 2):
            O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=1, ranfb=st)
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], n[k], 0, k)
        assert c.ranfb == int(st[0])
        # renewal ex0 <- ex (F:796-807): the host copies, the mirror repeats it on the device
        for i in range(6):
            setattr(c, FN[i + 6], getattr(c, FN[i]).copy())
        fm.fields_renewed()
        c.it += 1
    for k in (1, 2):
        fm.pull(k, *host[k], n[k])
        assert U.particle_err(host[k], ref[k], p.hx, U.vth(k)) < 30 * PTOL, k
    h2d = fm.ctx.counters()["h2d_bytes"] - 48 * (n[1] + n[2])      # field bytes only (the first calls uploaded the particles)
    full = 8 * O.mxyzA(p) * (12 + 3 * 9)              # eager protocol: 12 arrays once, then 9 per step
    if planes:
        assert h2d < 0.85 * full, (h2d, full)
    fm.ctx.close()


def test_fast_particles_disable_the_plane_record(mrg):
    """|vz| dt >= hz breaks the +-1 plane bound of the recorded gather planes (the precondition was only
    documented).  The tiled corrector now reports it, the host ignores the record and prepares every plane: results still
    match the oracle, and no restricted preparation runs."""
    p = U.make_parm(12, 10, 24)
    sp_all, ranfb = U.load_species(p, 10)
    sp = slab_subset(p, sp_all, 6.0, 11.0)
    k = 2
    sp[k][5][::7] = 1.3 * p.hz / p.dt * np.sign(sp[k][5][::7] + 1e-30)      # every 7th electron crosses more than a plane per step
    n = len(sp[k][0])
    ref = [a.copy() for a in sp[k]]
    st = np.array([ranfb], dtype=np.int32)
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    ctx.set_option("planes", 1)
    ctx.upload(k, *sp[k])
    ctx.sort(k, p.hdt)
    st_gpu = ranfb
    restricted = []
    for step in range(3):
        for ipc in (1, 0):
            f12 = U.smooth_fields(p, seed=500 + 10 * step + ipc)
            a6 = O.field_prep(p, f12)
            ctx.set_fields(f12)
            ctx.prep_stats(reset=True)
            r = O.fulmov(p, a6, *ref, U.QSPEC[k], U.WSPEC[k], ipc, nranks=1, ranfb=st)
            _, _, st_gpu = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], ipc, params_of(mrg, p), st_gpu)
            restricted.append(ctx.prep_stats()["restricted"])
            if ipc >= 1:
                mom = ctx.moments(k)
                for c in range(4):
                    assert U.rel_l2(mom[c], r["mom"][c]) < MTOL, (step, c)
            else:
                ctx.sort(k, p.hdt)
    got = ctx.download(k, n)
    assert U.particle_err(got, ref, p.hx, U.vth(k)) < 30 * PTOL
    assert st_gpu == int(st[0])
    assert sum(restricted[2:]) == 0, restricted        # after the first corrector reported the violation: full preparations only
    ctx.close()


@pytest.mark.parametrize("planes", [0, 1])
def test_lazy_fields_with_the_b_updates_on_the_device(mrg, planes):
    """The same protocol with the two marks that keep bx,by,bz off PCIe (Fulmov.prefld_done / emfild_done): with lazily held
    fields every preparation computes exactly the planes of bx,by,bz it reads, after fetching the planes of ex..ez, ex0..ez0
    around them.  The host's arrays are what its own prefld / emfild would hold (the oracle's orc_update_b, pinned to the
    reference); results match the oracle on whole arrays, and on the steps without smoothing no plane of bx,by,bz is uploaded."""
    p = U.make_parm(12, 10, 24)
    sp_all, ranfb = U.load_species(p, 10)
    sp = slab_subset(p, sp_all, 7.0, 12.0)
    n = {k: len(sp[k][0]) for k in (1, 2)}
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.array([ranfb], dtype=np.int32)
    c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, dt=p.dt, aimpl=p.aimpl, wce_by_wpe=p.bxc, Ez00=p.Ez00)
    c.ranfb, c.it = ranfb, 4                                   # steps it = 4, 5, 6: the last one smooths (mod(it,5) = 1)
    fm = mrg.Fulmov(c, ipar=1, size=1, hints=True, lazy=True)
    fm.ctx.set_option("planes", planes)
    FN = mrg.host.FIELD_NAMES
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    for name, a in zip(FN, U.smooth_fields(p, seed=500, ghost_nan=False)):
        setattr(c, name, a.copy())
    per_step = []
    for step in range(3):
        fm.ctx.counters(reset=True)
        hb = O.update_b(p, [np.ascontiguousarray(a).copy() for a in c.fields()], 0)           # the host's prefld
        for i in (3, 4, 5):
            setattr(c, FN[i], hb[i])
        fm.prefld_done()
        a6 = O.field_prep(p, c.fields())
        for k in (1, 2):
            r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=1, ranfb=st)
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], n[k], 1, k)
            got = (c.qix, c.qiy, c.qiz, c.qi) if k == 1 else (c.qex, c.qey, c.qez, c.qe)
            for ci in range(4):
                assert U.rel_l2(got[ci], r["mom"][ci]) < MTOL, (step, k, ci)
        f_n = U.smooth_fields(p, seed=520 + step, ghost_nan=False)                             # the host's emfild: new E ...
        for i in range(3):
            setattr(c, FN[i], f_n[i].copy())
        hb = O.update_b(p, [np.ascontiguousarray(a).copy() for a in c.fields()], c.it % 5 == 1)   # ... and the B it leaves
        for i in (3, 4, 5):
            setattr(c, FN[i], hb[i])
        fm.emfild_done()
        a6 = O.field_prep(p, c.fields())
        for k in (1, 2):
            O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=1, ranfb=st)
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], n[k], 0, k)
        assert c.ranfb == int(st[0])
        for i in range(6):
            setattr(c, FN[i + 6], getattr(c, FN[i]).copy())
        fm.fields_renewed()
        per_step.append(fm.ctx.counters()["h2d_bytes"] - (48 * (n[1] + n[2]) if step == 0 else 0))
        c.it += 1
    for k in (1, 2):
        fm.pull(k, *host[k], n[k])
        assert U.particle_err(host[k], ref[k], p.hx, U.vth(k)) < 30 * PTOL, k
    # step 2 (it = 5): nothing but planes of ex,ey,ez (+ what the widened halo of the B update asks of the old arrays)
    assert per_step[1] <= 8 * O.mxyzA(p) * 3 * 1.35, per_step
    assert per_step[2] > per_step[1]                           # the smoothing step uploads bx,by,bz as well
    fm.ctx.close()
