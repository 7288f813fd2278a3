"""The host-side mirrors of `subroutine fulmov` (F:1044): the Python `Fulmov`
and the C++ `fulmov` of csrc/mrg_host.cpp (what the ISO_C_BINDING shim calls)
reproduce the subroutine's side effects on COMMON-like storage."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import util as U

pytestmark = pytest.mark.gpu
MTOL, PTOL = 1e-10, 1e-12


def _oracle_step(p, f_pred, f_corr, sp, ranfb):
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.array([ranfb], dtype=np.int32)
    mom, wk = {}, {}
    a6 = O.field_prep(p, f_pred)
    for k in (1, 2):
        r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=1, ranfb=st)
        mom[k] = r["mom"]
    a6 = O.field_prep(p, f_corr)
    for k in (1, 2):
        r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=1, ranfb=st)
        wk[k] = (r["wkix"], r["wkih"])
    return ref, mom, wk, int(st[0])


@pytest.fixture(scope="module")
def case():
    p = U.make_parm(12, 10, 12)
    sp, ranfb = U.load_species(p, 12)
    return p, sp, ranfb, U.smooth_fields(p, seed=21), U.smooth_fields(p, seed=22)


def test_python_fulmov_mirror(case):
    import mrg_b200 as mrg
    p, sp, ranfb, f_pred, f_corr = case
    ref, mom, wk, st_ref = _oracle_step(p, f_pred, f_corr, sp, ranfb)
    c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, dt=p.dt, aimpl=p.aimpl, wce_by_wpe=p.bxc, Ez00=p.Ez00)
    c.ranfb = ranfb
    c.it, c.nha, c.ldec = 5, 5, 2                      # mod(it,nha)==0 -> edec row ldec is written
    fm = mrg.Fulmov(c, ipar=1, size=1)
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    npr = len(sp[1][0])
    for name, arr in zip(mrg.host.FIELD_NAMES, f_pred):
        getattr(c, name)[:] = arr
    for k in (1, 2):
        fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 1, k)
    for cidx, (a, b) in enumerate(zip((c.qix, c.qiy, c.qiz, c.qi), mom[1])):
        assert U.rel_l2(a, b) < MTOL, cidx
    for cidx, (a, b) in enumerate(zip((c.qex, c.qey, c.qez, c.qe), mom[2])):
        assert U.rel_l2(a, b) < MTOL, cidx
    for name, arr in zip(mrg.host.FIELD_NAMES, f_corr):
        getattr(c, name)[:] = arr
    fm.fields_changed()
    for k in (1, 2):
        fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 0, k)
        assert abs(c.wkix - wk[k][0]) < MTOL * abs(wk[k][0])
        col = 5 if k == 1 else 7
        assert c.edec[col - 1, c.ldec - 1] == c.wkix and c.edec[col, c.ldec - 1] == c.wkih   # F:1320-1328
    assert c.ranfb == st_ref
    for k in (1, 2):
        fm.pull(k, *host[k], npr)
        assert U.particle_err(host[k], ref[k], p.hx, U.vth(k)) < PTOL


def test_hints_survive_the_it0_sequence(case):
    """at it = 0 trans calls the pair with dt = adt = hdt = 0, emfld0 then rewrites ALL of COMMON /fields/ and the
    renewal loop runs (F:664-706) -- with only the three optional marks a mirror in hints mode would keep stale ex..bz and
    copy them into ex0..bz0.  The mirrors treat the first call after the it = 0 pair as "everything changed"."""
    import mrg_b200 as mrg
    p, sp, ranfb, f_a, f_b = case
    c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, dt=p.dt, aimpl=p.aimpl, wce_by_wpe=p.bxc, Ez00=p.Ez00)
    c.ranfb, c.it = ranfb, 0
    fm = mrg.Fulmov(c, ipar=1, size=1, hints=True)
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    npr = len(sp[1][0])
    FN = mrg.host.FIELD_NAMES
    for name, arr in zip(FN, f_a):
        getattr(c, name)[:] = arr
    dt, adt, hdt = c.dt, c.adt, c.hdt
    c.dt = c.adt = c.hdt = 0.0                          # F:678-680
    for k in (1, 2):
        fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 1, k)
    c.dt, c.adt, c.hdt = dt, adt, hdt
    for name, arr in zip(FN[:6], f_b[:6]):              # emfld0 (F:691): every field is new, no mark exists for it
        getattr(c, name)[:] = arr
    for i in range(6):                                  # renewal loop + its mark
        getattr(c, FN[i + 6])[:] = getattr(c, FN[i])
    fm.fields_renewed()
    c.it = 1
    f_c = U.smooth_fields(p, seed=23)
    for i in (3, 4, 5):                                 # prefld + its mark
        getattr(c, FN[i])[:] = f_c[i]
    fm.fields_changed(fm.MASK_B)
    a6 = O.field_prep(p, c.fields())
    for k in (1, 2):
        r = O.fulmov(p, a6, *[a.copy() for a in sp[k]], U.QSPEC[k], U.WSPEC[k], 1, nranks=1)
        fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 1, k)
        got = (c.qix, c.qiy, c.qiz, c.qi) if k == 1 else (c.qex, c.qey, c.qez, c.qe)
        for ci in range(4):
            assert U.rel_l2(got[ci], r["mom"][ci]) < MTOL, (k, ci)
    fm.ctx.close()


_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class View(C.Structure):
    """mirror of mrg_common_view (csrc/mrg_host.h)"""
    _fields_ = ([("mx", C.c_int32), ("my", C.c_int32), ("mz", C.c_int32)]
                + [(n, _dp) for n in ("ex", "ey", "ez", "bx", "by", "bz", "ex0", "ey0", "ez0", "bx0", "by0", "bz0")]
                + [(n, _dp) for n in ("qix", "qiy", "qiz", "qex", "qey", "qez", "qi", "qe")]
                + [(n, _ip) for n in ("it", "ldec", "ifilx", "ifily", "ifilz", "nha")]
                + [(n, _dp) for n in ("xmax", "ymax", "zmax", "dt", "aimpl", "adt", "hdt", "bxc", "byc", "bzc", "edec")]
                + [(n, _dp) for n in ("wkix", "wkih", "zcent", "ycent1", "ycent2", "Ez00")]
                + [("ranfb", _ip), ("io_pe", _ip)])


def test_cpp_fulmov_mirror(case):
    """C++ `fulmov` with the Fortran calling convention (everything by reference)."""
    import mrg_b200 as mrg
    p, sp, ranfb, f_pred, f_corr = case
    ref, mom, wk, st_ref = _oracle_step(p, f_pred, f_corr, sp, ranfb)
    mrg.build.build_host()
    lib = C.CDLL(mrg.build.HOSTLIB)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.mrg_host_fulmov.argtypes = [dp] * 8 + [ip] * 5
    lib.mrg_host_fulmov.restype = None
    lib.mrg_host_pull_particles.argtypes = [C.c_int32] + [dp] * 6 + [C.c_int32] * 3
    n = O.mxyzA(p)
    store = {}

    def darr(name, val):
        store[name] = np.array(val, dtype=np.float64).reshape(-1).copy() if not isinstance(val, np.ndarray) else val
        return store[name].ctypes.data_as(dp)

    def iarr(name, val):
        store[name] = np.array([val], dtype=np.int32)
        return store[name].ctypes.data_as(ip)

    v = View()
    v.mx, v.my, v.mz = p.mx, p.my, p.mz
    fnames = ("ex", "ey", "ez", "bx", "by", "bz", "ex0", "ey0", "ez0", "bx0", "by0", "bz0")
    for name, arr in zip(fnames, f_pred):
        setattr(v, name, darr(name, arr.copy()))
    for name in ("qix", "qiy", "qiz", "qex", "qey", "qez", "qi", "qe"):
        setattr(v, name, darr(name, np.zeros(n)))
    for name, val in (("it", 5), ("ldec", 2), ("ifilx", 1), ("ifily", 1), ("ifilz", 1), ("nha", 5), ("ranfb", ranfb), ("io_pe", 1)):
        setattr(v, name, iarr(name, val))
    for name, val in (("xmax", p.xmax), ("ymax", p.ymax), ("zmax", p.zmax), ("dt", p.dt), ("aimpl", p.aimpl),
                      ("adt", p.adt), ("hdt", p.hdt), ("bxc", p.bxc), ("byc", p.byc), ("bzc", p.bzc),
                      ("wkix", 0.0), ("wkih", 0.0), ("zcent", p.zcent), ("ycent1", p.ycent1), ("ycent2", p.ycent2),
                      ("Ez00", p.Ez00)):
        setattr(v, name, darr(name, [val]))
    v.edec = darr("edec", np.zeros(3000 * 12))
    assert lib.mrg_host_bind(C.byref(v), 0) == 0
    lib.mrg_host_set_exit_on_error(0)
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    npr = C.c_int32(len(sp[1][0]))
    one, size = C.c_int32(1), C.c_int32(1)

    def call(k, ipc):
        q, w = C.c_double(U.QSPEC[k]), C.c_double(U.WSPEC[k])
        lib.mrg_host_fulmov(*[a.ctypes.data_as(dp) for a in host[k]], C.byref(q), C.byref(w), C.byref(npr),
                   C.byref(C.c_int32(ipc)), C.byref(C.c_int32(k)), C.byref(one), C.byref(size))
        assert lib.mrg_host_status() == 0

    call(1, 1); call(2, 1)
    for cidx, name in enumerate(("qix", "qiy", "qiz", "qi")):
        assert U.rel_l2(store[name], mom[1][cidx]) < MTOL, name
    for cidx, name in enumerate(("qex", "qey", "qez", "qe")):
        assert U.rel_l2(store[name], mom[2][cidx]) < MTOL, name
    for name, arr in zip(fnames, f_corr):                 # "emfild" wrote new fields into COMMON /fields/
        store[name][:] = arr
    lib.mrg_host_fields_changed()
    call(1, 0)
    assert abs(store["wkix"][0] - wk[1][0]) < MTOL * abs(wk[1][0])
    assert store["edec"][1 + 3000 * 4] == store["wkix"][0] and store["edec"][1 + 3000 * 5] == store["wkih"][0]
    call(2, 0)
    assert store["edec"][1 + 3000 * 6] == store["wkix"][0]
    assert int(store["ranfb"][0]) == st_ref
    for k in (1, 2):
        assert lib.mrg_host_pull_particles(k, *[a.ctypes.data_as(dp) for a in host[k]], npr.value, 1, 1) == 0
        assert U.particle_err(host[k], ref[k], p.hx, U.vth(k)) < PTOL
    lib.mrg_host_unbind()


def test_device_prefld_bit_exact(case):
    """mrg_prefld = entry prefld of emfild (F:3820-3873) on the device copies of COMMON /fields/: every interior node of bx, by,
    bz bit-identical to the oracle's restatement (which tests/test_ref_pin.py holds bit for bit to the reference's own);
    nothing else is written."""
    import mrg_b200 as mrg
    p = case[0]
    rng = np.random.default_rng(11)
    f12 = [rng.normal(scale=0.01, size=O.mxyzA(p)) for _ in range(12)]
    want = O.prefld(p, [a.copy() for a in f12])
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    ctx.set_fields(f12)
    ctx.prefld(p.dt, p.aimpl)
    got = ctx.get_fields()
    ctx.close()
    sh = (p.mz + 4, p.my + 3, p.mx + 4)
    inner = (slice(2, p.mz + 2), slice(1, p.my + 2), slice(2, p.mx + 2))
    for c in range(12):
        if 3 <= c <= 5:
            np.testing.assert_array_equal(got[c].reshape(sh)[inner], want[c].reshape(sh)[inner])
        else:
            np.testing.assert_array_equal(got[c], f12[c])


def test_prefld_mark_replaces_the_upload_of_b(case):
    """hints mode: after the host's prefld the mirror repeats the entry on the device (prefld_done) instead of uploading
    bx, by, bz -- the prepared fields the predictor gathers from are bit-identical either way, 3 grid arrays less H2D"""
    import mrg_b200 as mrg
    p, sp, ranfb, f_a, f_b = case
    npr = len(sp[1][0])
    FN = mrg.host.FIELD_NAMES
    out = {}
    for mode in ("upload", "device"):
        c = mrg.Common(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax, dt=p.dt, aimpl=p.aimpl, wce_by_wpe=p.bxc, Ez00=p.Ez00)
        c.ranfb, c.it = ranfb, 1
        fm = mrg.Fulmov(c, ipar=1, size=1, hints=True)
        host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
        for name, arr in zip(FN, f_a):
            getattr(c, name)[:] = arr
        for k in (1, 2):                                   # a first pair of calls: everything is uploaded
            fm(*host[k], U.QSPEC[k], U.WSPEC[k], npr, 1, k)
        fm.finish_moments()
        fm.ctx.counters(reset=True)
        hostf = O.prefld(p, [np.ascontiguousarray(a, dtype=np.float64).copy() for a in c.fields()])      # the host's own prefld
        for i in (3, 4, 5):
            getattr(c, FN[i])[:] = hostf[i]
        if mode == "upload":
            fm.fields_changed(fm.MASK_B)
        else:
            fm.prefld_done()
        fm(*host[1], U.QSPEC[1], U.WSPEC[1], npr, 1, 1)
        fm.finish_moments()
        cnt = fm.ctx.counters()
        par = c.step_params()
        out[mode] = (fm.ctx.prepared_fields(par), cnt["h2d_bytes"], [a.copy() for a in (c.qix, c.qiy, c.qiz, c.qi)])
        fm.ctx.close()
    for a, b in zip(out["upload"][0], out["device"][0]):
        np.testing.assert_array_equal(a, b)
    assert out["upload"][1] - out["device"][1] == 3 * 8 * O.mxyzA(p)
    for a, b in zip(out["upload"][2], out["device"][2]):
        assert U.rel_l2(a, b) < 1e-13


@pytest.mark.parametrize("smooth", [0, 1])
def test_device_b_update_after_emfild_bit_exact(case, smooth):
    """mrg_update_b: bx,by,bz as emfild leaves them behind its solve (F:4238-4302) -- prefld's update from the new E, and on
    the smoothing steps (mod(it,5) = 1) outmesh3 + filt3e -- against the oracle's orc_update_b, which tests/test_ref_pin.py
    holds bit for bit to the reference's emfild"""
    import mrg_b200 as mrg
    p = case[0]
    rng = np.random.default_rng(12)
    f12 = [rng.normal(scale=0.01, size=O.mxyzA(p)) for _ in range(12)]
    want = O.update_b(p, [a.copy() for a in f12], smooth)
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    ctx.set_fields(f12)
    ctx.update_b(p.dt, p.aimpl, smooth)
    got = ctx.get_fields(0x038)
    # the scratch of the sweeps is the preparation's: a preparation afterwards must still be right
    par = mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)
    a6 = ctx.prepared_fields(par)
    ctx.close()
    sh = (p.mz + 4, p.my + 3, p.mx + 4)
    inner = (slice(2, p.mz + 2), slice(1, p.my + 2), slice(2, p.mx + 2))
    for c in (3, 4, 5):
        np.testing.assert_array_equal(got[c].reshape(sh)[inner], want[c].reshape(sh)[inner])
    full = [a.copy() for a in f12]
    for c in (3, 4, 5):
        full[c].reshape(sh)[inner] = want[c].reshape(sh)[inner]
    for a, b in zip(a6, O.field_prep(p, full)):
        np.testing.assert_array_equal(a, b)


def test_cpp_mirror_marks_keep_b_on_the_device(case):
    """The C++ mirror in hints mode with the marks a Fortran host would place (mrg_host_prefld_done, mrg_host_emfild_done,
    mrg_host_fields_renewed): bx,by,bz are never uploaded after the first call, yet two steps (the second one a smoothing
    step of emfild, it = 6) match the oracle run on the host's whole arrays."""
    import mrg_b200 as mrg
    p, sp, ranfb, f_a, _ = case
    mrg.build.build_host()
    lib = C.CDLL(mrg.build.HOSTLIB)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.mrg_host_fulmov.argtypes = [dp] * 8 + [ip] * 5
    lib.mrg_host_fulmov.restype = None
    lib.mrg_host_pull_particles.argtypes = [C.c_int32] + [dp] * 6 + [C.c_int32] * 3
    lib.mrg_host_context.restype = C.c_void_p
    n = O.mxyzA(p)
    store = {}
    v = View()
    v.mx, v.my, v.mz = p.mx, p.my, p.mz
    fnames = ("ex", "ey", "ez", "bx", "by", "bz", "ex0", "ey0", "ez0", "bx0", "by0", "bz0")
    f_start = U.smooth_fields(p, seed=31, ghost_nan=False)
    for name, arr in zip(fnames, f_start):
        store[name] = arr.copy()
        setattr(v, name, store[name].ctypes.data_as(dp))
    for name in ("qix", "qiy", "qiz", "qex", "qey", "qez", "qi", "qe"):
        store[name] = np.zeros(n)
        setattr(v, name, store[name].ctypes.data_as(dp))
    for name, val in (("it", 5), ("ldec", 2), ("ifilx", 1), ("ifily", 1), ("ifilz", 1), ("nha", 5), ("ranfb", ranfb), ("io_pe", 0)):
        store[name] = np.array([val], dtype=np.int32)
        setattr(v, name, store[name].ctypes.data_as(ip))
    for name, val in (("xmax", p.xmax), ("ymax", p.ymax), ("zmax", p.zmax), ("dt", p.dt), ("aimpl", p.aimpl),
                      ("adt", p.adt), ("hdt", p.hdt), ("bxc", p.bxc), ("byc", p.byc), ("bzc", p.bzc),
                      ("wkix", 0.0), ("wkih", 0.0), ("zcent", p.zcent), ("ycent1", p.ycent1), ("ycent2", p.ycent2), ("Ez00", p.Ez00)):
        store[name] = np.array([val])
        setattr(v, name, store[name].ctypes.data_as(dp))
    store["edec"] = np.zeros(3000 * 12)
    v.edec = store["edec"].ctypes.data_as(dp)
    assert lib.mrg_host_bind(C.byref(v), 0) == 0
    lib.mrg_host_set_exit_on_error(0)
    lib.mrg_host_set_auto_fields(0)
    host = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    ref = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.array([ranfb], dtype=np.int32)
    npr = C.c_int32(len(sp[1][0]))
    one, size = C.c_int32(1), C.c_int32(1)

    def call(k, ipc):
        q, w = C.c_double(U.QSPEC[k]), C.c_double(U.WSPEC[k])
        lib.mrg_host_fulmov(*[a.ctypes.data_as(dp) for a in host[k]], C.byref(q), C.byref(w), C.byref(npr),
                            C.byref(C.c_int32(ipc)), C.byref(C.c_int32(k)), C.byref(one), C.byref(size))
        assert lib.mrg_host_status() == 0

    def fields():
        return [store[nm] for nm in fnames]

    for it in (5, 6):
        store["it"][0] = it
        hb = O.update_b(p, [a.copy() for a in fields()], 0)                    # the host's prefld
        for i in (3, 4, 5):
            store[fnames[i]][:] = hb[i]
        lib.mrg_host_prefld_done()
        a6 = O.field_prep(p, fields())
        for k in (1, 2):
            r = O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=1, ranfb=st)
            call(k, 1)
            names = ("qix", "qiy", "qiz", "qi") if k == 1 else ("qex", "qey", "qez", "qe")
            for cidx, name in enumerate(names):
                assert U.rel_l2(store[name], r["mom"][cidx]) < MTOL, (it, name)
        f_n = U.smooth_fields(p, seed=40 + it, ghost_nan=False)               # the host's emfild: new E and the B it leaves
        for i in range(3):
            store[fnames[i]][:] = f_n[i]
        hb = O.update_b(p, [a.copy() for a in fields()], it % 5 == 1)
        for i in (3, 4, 5):
            store[fnames[i]][:] = hb[i]
        lib.mrg_host_emfild_done()
        a6 = O.field_prep(p, fields())
        for k in (1, 2):
            O.fulmov(p, a6, *ref[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=1, ranfb=st)
            call(k, 0)
        assert int(store["ranfb"][0]) == int(st[0])
        for i in range(6):                                                     # renewal + its mark
            store[fnames[i + 6]][:] = store[fnames[i]]
        lib.mrg_host_fields_renewed()
    for k in (1, 2):
        assert lib.mrg_host_pull_particles(k, *[a.ctypes.data_as(dp) for a in host[k]], npr.value, 1, 1) == 0
        assert U.particle_err(host[k], ref[k], p.hx, U.vth(k)) < 4 * PTOL
    lib.mrg_host_unbind()
