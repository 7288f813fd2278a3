"""ctypes binding of the CPU oracle (oracle/fulmov_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libfulmov_oracle.so")

dp = C.POINTER(C.c_double)


class Parm(C.Structure):
    """Mirror of orc_parm (COMMON /parm1/,/parm2/,/ptable/,/profl/ subset)."""

    _fields_ = [
        ("mx", C.c_int32), ("my", C.c_int32), ("mz", C.c_int32),
        ("ifilx", C.c_int32), ("ifily", C.c_int32), ("ifilz", C.c_int32),
        ("xmax", C.c_double), ("ymax", C.c_double), ("zmax", C.c_double),
        ("hx", C.c_double), ("hy", C.c_double), ("hz", C.c_double),
        ("hxi", C.c_double), ("hyi", C.c_double), ("hzi", C.c_double),
        ("xmaxe", C.c_double), ("ymaxe", C.c_double), ("zmaxe", C.c_double),
        ("dt", C.c_double), ("aimpl", C.c_double), ("adt", C.c_double), ("hdt", C.c_double),
        ("bxc", C.c_double), ("byc", C.c_double), ("bzc", C.c_double),
        ("Ez00", C.c_double), ("zcent", C.c_double), ("ycent1", C.c_double), ("ycent2", C.c_double),
    ]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    src = [os.path.join(_HERE, f) for f in ("fulmov_oracle.c", "fulmov_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src)):
        return _LIB
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        P = C.POINTER(Parm)
        i64, i32p = C.c_int64, C.POINTER(C.c_int32)
        L.orc_mxyzA.restype = i64
        L.orc_mxyzA.argtypes = [P]
        L.orc_parm_init.argtypes = [P, C.c_int, C.c_int, C.c_int] + [C.c_double] * 7
        L.orc_ranf.restype = C.c_double
        L.orc_ranf.argtypes = [i32p]
        L.orc_ranfp.restype = C.c_double
        L.orc_ranfp.argtypes = [i32p]
        L.orc_lcg_skip.restype = C.c_int32
        L.orc_lcg_skip.argtypes = [C.c_int32, C.c_uint64]
        L.orc_outmesh3.argtypes = [P, dp, dp, dp]
        L.orc_vmesh3.argtypes = [P, dp, dp, dp]
        L.orc_vmesh1.argtypes = [P, dp]
        L.orc_filt3e.argtypes = [P, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                 C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_field_prep.argtypes = [P, C.POINTER(dp), C.POINTER(dp)]
        L.orc_partbc.argtypes = [P, dp, dp, dp, dp, i64, i64, i64]
        L.orc_partbcEST.argtypes = [P, dp, dp, dp, i64, i64, i64]
        L.orc_srimp1_scatter.argtypes = [P] + [dp] * 6 + [C.c_double] + [dp] * 3 + [i64] * 3
        L.orc_srimp2_scatter.argtypes = [P] + [dp] * 3 + [C.c_double, dp] + [i64] * 3
        L.orc_fulmov.argtypes = [P, C.POINTER(dp)] + [dp] * 6 + [C.c_double, C.c_double, i64,
                                 C.c_int, C.c_int, i32p, C.POINTER(dp), C.POINTER(dp), dp,
                                 C.POINTER(dp)]
        L.orc_loadpt.restype = i64
        L.orc_loadpt.argtypes = [P, C.c_int, C.c_double, C.c_double, C.c_double] + [dp] * 6 + [i32p, i32p]
        L.orc_loadpt_fv2.argtypes = [C.c_double, C.c_double, dp, dp, dp]
        L.orc_num_threads.restype = C.c_int
        L.orc_time_step.restype = C.c_double
        L.orc_time_step.argtypes = [P, C.POINTER(dp), C.POINTER(dp)] + [dp] * 6 + [C.c_double, C.c_double, i64, C.c_int, dp, dp]
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(dp)


def _parr(arrs):
    return (dp * len(arrs))(*[_p(a) if a is not None else dp() for a in arrs])


def make_parm(mx, my, mz, xmax, ymax, zmax, dt=1.2, aimpl=0.6, wce_by_wpe=0.2, Ez00=0.25e-2):
    p = Parm()
    lib().orc_parm_init(C.byref(p), mx, my, mz, xmax, ymax, zmax, dt, aimpl, wce_by_wpe, Ez00)
    return p


def mxyzA(p):
    return (p.mx + 4) * (p.my + 3) * (p.mz + 4)


def grid_shape(p):
    """numpy shape (k, j, i) of an extended array; element [k+2, j+1, i+2]."""
    return (p.mz + 4, p.my + 3, p.mx + 4)


def field_prep(p, f12):
    a6 = [np.empty(mxyzA(p)) for _ in range(6)]
    lib().orc_field_prep(C.byref(p), _parr(f12), _parr(a6))
    return a6


def prefld(p, f12):
    """entry prefld of emfild (F:3820-3873) in place: f12[3..5] (bx,by,bz) from ex..ez, ex0..ez0, bx0..bz0"""
    for a in f12:
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    lib().orc_prefld(C.byref(p), _parr(f12))
    return f12


def update_b(p, f12, smooth):
    """bx,by,bz as emfild leaves them after its solve (F:4238-4302), in place"""
    for a in f12:
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    lib().orc_update_b(C.byref(p), _parr(f12), 1 if smooth else 0)
    return f12


def filt3e(p, ex, ey, ez, dc, sym, ifil=None):
    fx, fy, fz = ifil if ifil is not None else (p.ifilx, p.ifily, p.ifilz)
    lib().orc_filt3e(C.byref(p), _p(ex), _p(ey), _p(ez), dc[0], dc[1], dc[2], fx, fy, fz, sym)


def outmesh3(p, ax, ay, az):
    lib().orc_outmesh3(C.byref(p), _p(ax), _p(ay), _p(az))


def vmesh3(p, ax, ay, az):
    lib().orc_vmesh3(C.byref(p), _p(ax), _p(ay), _p(az))


def vmesh1(p, a):
    lib().orc_vmesh1(C.byref(p), _p(a))


def partbc(p, x, y, z, vy, first=1, stride=1):
    lib().orc_partbc(C.byref(p), _p(x), _p(y), _p(z), _p(vy), len(x), first, stride)


def partbcEST(p, x, y, z, first=1, stride=1):
    lib().orc_partbcEST(C.byref(p), _p(x), _p(y), _p(z), len(x), first, stride)


def fulmov(p, a6, x, y, z, vx, vy, vz, qmult, wmult, ipc, nranks=1, ranfb=None,
           want_raw=False, want_pred=False):
    """One fulmov call by `nranks` simulated ranks.  Particle arrays are
    updated in place for ipc=0.  Returns a dict with wkix, wkih and, for
    ipc>=1, 'mom' (folded qjx,qjy,qjz,q), optionally 'raw' and 'pred'."""
    n = mxyzA(p)
    npr = len(x)
    if ranfb is None:
        ranfb = np.full(nranks, 7331, dtype=np.int32)
    assert ranfb.dtype == np.int32 and len(ranfb) == nranks
    mom = [np.zeros(n) for _ in range(4)]
    raw = [np.zeros(n) for _ in range(4)] if want_raw else [None] * 4
    pred = [np.zeros(npr) for _ in range(6)] if want_pred else [None] * 6
    wk = np.zeros(2)
    lib().orc_fulmov(C.byref(p), _parr(a6), _p(x), _p(y), _p(z), _p(vx), _p(vy), _p(vz),
                     qmult, wmult, npr, ipc, nranks, ranfb.ctypes.data_as(C.POINTER(C.c_int32)),
                     _parr(mom), _parr(raw), _p(wk), _parr(pred))
    out = {"wkix": wk[0], "wkih": wk[1], "ranfb": ranfb}
    if ipc >= 1:
        out["mom"] = mom
        if want_raw:
            out["raw"] = raw
        if want_pred:
            out["pred"] = pred
    return out


def loadpt(p, ppc, vth, vdr, vbeam, ranfa=3021, ranfb=7331):
    """loadpt (F:8735-9080); returns (x,y,z,vx,vy,vz), ranfa_state, ranfb_state."""
    npr = p.mx * p.my * p.mz * ppc
    arrs = [np.empty(npr) for _ in range(6)]
    a, b = C.c_int32(ranfa), C.c_int32(ranfb)
    got = lib().orc_loadpt(C.byref(p), ppc, vth, vdr, vbeam, *[_p(v) for v in arrs],
                           C.byref(a), C.byref(b))
    assert got == npr
    return arrs, a.value, b.value


def loadpt_fv2(vth, vdr):
    fv2 = np.zeros(101)
    v2, dv2 = C.c_double(), C.c_double()
    lib().orc_loadpt_fv2(vth, vdr, _p(fv2), C.byref(v2), C.byref(dv2))
    return fv2, v2.value, dv2.value


def lcg_skip(state, n):
    return lib().orc_lcg_skip(state, n)


def ranfp_stream(state, n):
    s = C.c_int32(state)
    out = np.empty(n)
    for i in range(n):
        out[i] = lib().orc_ranfp(C.byref(s))
    return out, s.value


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP threads of the simulated ranks (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().orc_set_num_threads(int(n))


def time_step(p, a6p, a6c, arrs, qmult, wmult, nranks):
    """Seconds for one species' predictor + corrector pass by `nranks` threads
    (each with private particle arrays, like the reference's MPI ranks)."""
    tp, tc = C.c_double(), C.c_double()
    t = lib().orc_time_step(C.byref(p), _parr(a6p), _parr(a6c), *[_p(a) for a in arrs], qmult, wmult,
                            len(arrs[0]), nranks, C.byref(tp), C.byref(tc))
    return t, tp.value, tc.value
