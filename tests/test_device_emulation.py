"""CPU check of the DEVICE arithmetic: csrc/mrg_device.cuh (cell index via the
round-down add, TSC/linear weights, packed 128-bit gather, closed-form
rotation, scatter factors) is compiled for the host with the CUDA intrinsics
mapped to libm/fenv (tests/host_emul/emul_device.cpp) and compared with the
oracle particle by particle.  No GPU needed; nothing here is a product path."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import util as U

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "macro-particle_simulation_for_magnetic_reconnection_b200", "csrc")


@pytest.fixture(scope="module")
def emul():
    d = tempfile.mkdtemp(prefix="mrg_emul_")
    hdr = open(os.path.join(CSRC, "mrg_device.cuh")).read().replace("#include <cuda_runtime.h>", "")
    open(os.path.join(d, "mrg_device_emul.h"), "w").write(hdr)
    so = os.path.join(d, "libemul.so")
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-frounding-math", "-std=c++17", "-fPIC", "-shared",
                    "-I" + d, "-o", so, os.path.join(HERE, "host_emul", "emul_device.cpp")], check=True)
    L = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.emul_push.argtypes = [dp, ip, dp, dp] + [C.c_double] * 6 + [C.c_int, dp, dp, ip, dp]
    L.emul_gather_plane_diff.argtypes = [dp, ip] + [C.c_double] * 7 + [ip, ip]
    return L


def _run(L, p, F6, arrs, qmult, wmult, ipc, dt, adt, hdt):
    n = len(arrs[0])
    gp = np.array([p.xmax, p.ymax, p.zmax])
    gi = np.array([p.mx, p.my, p.mz], dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    out = np.zeros((n, 6)); wk = np.zeros((n, 2)); keys = np.zeros(n, dtype=np.int32); fac = np.zeros((n, 17))
    part = np.ascontiguousarray(np.stack(arrs, axis=1))
    for l in range(n):
        L.emul_push(gp.ctypes.data_as(dp), gi.ctypes.data_as(ip), F6.ctypes.data_as(dp),
                    part[l].ctypes.data_as(dp), dt, adt, hdt, p.aimpl, qmult, wmult, ipc,
                    out[l].ctypes.data_as(dp), wk[l].ctypes.data_as(dp), keys[l:].ctypes.data_as(ip),
                    fac[l].ctypes.data_as(dp))
    return out, wk, keys, fac


def _edge_particles(p, rng, n):
    """particles on / next to the periodic seams, the walls and cell boundaries"""
    x = rng.uniform(-p.hx / 2, p.xmax - p.hx / 2, n)
    y = rng.uniform(0, p.ymax, n)
    z = rng.uniform(-p.hz / 2, p.zmax - p.hz / 2, n)
    v = [rng.normal(scale=0.2, size=n) for _ in range(3)]
    sp = [np.nextafter(-p.hx / 2, 1), p.xmax - p.hx / 2 - 1e-9, 0.5 * p.hx, np.nextafter(0.5 * p.hx, 0), 2.5 * p.hx]
    for q, val in enumerate(sp):
        x[q] = val
        z[q + 5] = val * p.hz / p.hx
    y[10:16] = [1e-12, p.ymax - 1e-12, p.hy, np.nextafter(p.hy, 0), np.nextafter(p.ymax, 0), 3 * p.hy]
    v[1][10] = -0.3; v[1][11] = 0.3      # cross the walls during the step
    v[0][0] = -0.3; v[0][1] = 0.3        # cross the x seams
    return [x, y, z] + v


@pytest.mark.parametrize("ksp", [1, 2])
def test_device_arithmetic_matches_oracle(emul, ksp):
    p = U.make_parm(8, 6, 8)
    f12 = U.smooth_fields(p, seed=11)
    a6 = O.field_prep(p, f12)
    F6 = np.ascontiguousarray(np.stack(a6, axis=1))          # [node][6]
    rng = np.random.default_rng(ksp)
    arrs = _edge_particles(p, rng, 400)
    q, w = U.QSPEC[ksp], U.WSPEC[ksp]
    # corrector
    ref = [a.copy() for a in arrs]
    r0 = O.fulmov(p, a6, *ref, q, w, 0, nranks=1, ranfb=np.array([7331], dtype=np.int32))
    # the oracle also applied the drive kick; switch it off by comparing with Ez00=0
    pz = U.make_parm(8, 6, 8, Ez00=0.0)
    ref = [a.copy() for a in arrs]
    r0 = O.fulmov(pz, a6, *ref, q, w, 0, nranks=1)
    out, wk, _, _ = _run(emul, p, F6, arrs, q, w, 0, p.dt, p.adt, p.hdt)
    err = U.particle_err([out[:, c] for c in range(6)], ref, p.hx, U.vth(ksp))
    assert err < 1e-12, err
    assert abs(wk[:, 0].sum() - r0["wkix"]) < 1e-12 * abs(r0["wkix"])
    assert abs(wk[:, 1].sum() - r0["wkih"]) < 1e-12 * abs(r0["wkih"]) + 1e-300
    # predictor: predicted state, then scatter factors -> raw moments
    r1 = O.fulmov(p, a6, *[a.copy() for a in arrs], q, w, 1, nranks=1, want_raw=True, want_pred=True)
    out, wk, keys, fac = _run(emul, p, F6, arrs, q, w, 1, p.dt, p.adt, p.hdt)
    err = U.particle_err([out[:, c] for c in range(6)], r1["pred"], p.hx, U.vth(ksp))
    assert err < 1e-12, err
    n = O.mxyzA(p)
    nx, nxy = p.mx + 4, (p.mx + 4) * (p.my + 3)
    M = np.zeros((4, n))
    for l in range(len(keys)):
        for g9 in range(8):
            jy, m = g9 >> 2, g9 & 3
            for r in range(9):
                kz, ix = divmod(r, 3)
                M[m, keys[l] + ix + jy * nx + kz * nxy] += fac[l, g9] * fac[l, 8 + r]
    for m in range(4):
        assert U.rel_l2(M[m], r1["raw"][m]) < 1e-12


def test_gather_plane_is_the_stencil_plane(emul):
    """gather_plane (what the sort records for the restricted field preparation) is the kp the next pass'
    make_stencil computes, seams included; gather_plane_fast (what the corrector records, contracted and
    unwrapped) is within one plane of it, or next to the seam when the exact value is the clamp plane mz."""
    p = U.make_parm(8, 6, 8)
    rng = np.random.default_rng(5)
    arrs = _edge_particles(p, rng, 600)
    arrs[5][5:10] = [-0.4, 0.4, -0.4, 0.4, 0.0]          # cross the z seams
    # wrapped positions right at the seams, where the exact path wraps or clamps and the fast one folds
    zs = [-p.hz / 2, np.nextafter(-p.hz / 2, 1), p.zmax - p.hz / 2 - 1e-10 * p.hz, np.nextafter(p.zmax - p.hz / 2, 0),
          p.zmax - p.hz / 2 - 5e-10 * p.hz, 0.5 * p.hz, np.nextafter(0.5 * p.hz, 0), np.nextafter(1.5 * p.hz, 2 * p.hz)]
    arrs[2][20:20 + len(zs)] = zs
    arrs[5][20:20 + len(zs)] = 0.0
    gp = np.array([p.xmax, p.ymax, p.zmax])
    gi = np.array([p.mx, p.my, p.mz], dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    ex, fa = C.c_int(), C.c_int()
    mz = p.mz
    for hdt in (0.0, 0.6, 5.0):
        for l in range(len(arrs[0])):
            d = emul.emul_gather_plane_diff(gp.ctypes.data_as(dp), gi.ctypes.data_as(ip), *[float(a[l]) for a in arrs], hdt,
                                            C.byref(ex), C.byref(fa))
            assert d == 0, (hdt, l)
            z_ok = -p.hz / 2 <= arrs[2][l] < p.zmax - p.hz / 2          # the corrector only sees wrapped z
            if not z_ok or abs(hdt * arrs[5][l]) >= p.hz:
                continue
            assert -1 <= fa.value <= mz, (hdt, l, fa.value)
            f = fa.value % mz
            if ex.value == mz:
                assert f in (0, 1, mz - 2, mz - 1), (hdt, l, ex.value, fa.value)
            else:
                assert min((ex.value - f) % mz, (f - ex.value) % mz) <= 1, (hdt, l, ex.value, fa.value)
