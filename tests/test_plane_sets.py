"""CPU check of the dependency analysis behind option "planes" (restricted
field preparation, csrc/mrg_api.cu: plane_sets / ensure_prep).

The stages of section 0 of fulmov (F:1127-1148: blend, outmesh3 ghost fill,
filt3e z/x/y sweeps) are re-run here in numpy on ONLY the planes
mrg_plane_sets lists for each stage -- every other plane of every
intermediate array is NaN -- and the planes a particle can gather from must
then equal the oracle's full preparation bit for bit.  If a stage needed a
plane the analysis left out, the NaN would surface.  No GPU, no product
compute path: only the host-side helper of the library is called."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import util as U

GUARD = 1 << 30


def plane_sets(mz, occ):
    import mrg_b200
    lib = mrg_b200.capi.load()
    i32 = C.c_int32
    o = np.ascontiguousarray(occ, dtype=np.uint8)
    lb, lgi, lg = [(i32 * (mz + 4))() for _ in range(3)]
    n = (i32 * 3)()
    rc = lib.mrg_plane_sets(mz, o.ctypes.data_as(C.POINTER(C.c_uint8)), lb, lgi, lg, n)
    assert rc == 0
    return list(lb[:n[0]]), list(lgi[:n[1]]), list(lg[:n[2]])


def restricted_prep(p, f12, listB, listGI, listG):
    """numpy restatement of k_blend / k_filter<axis> / k_finalize on plane lists; untouched planes stay NaN"""
    mx, my, mz = p.mx, p.my, p.mz
    shp = (mz + 4, my + 3, mx + 4)
    f = [a.reshape(shp) for a in f12]
    dc = [0.0, 0.0, 0.0, p.bxc, p.byc, p.bzc]
    I = slice(2, mx + 2)
    J = slice(1, my + 2)
    A = [np.full(shp, np.nan) for _ in range(6)]
    T = [np.full(shp, np.nan) for _ in range(6)]
    om = 1.0 - p.aimpl
    for c in range(6):
        for k in listB:
            a = p.aimpl * f[c][k + 2, J, I] + om * f[c + 6][k + 2, J, I]
            if c >= 3:
                a = a + dc[c]
            A[c][k + 2, J, I] = a
            T[c][k + 2, J, I] = a - dc[c]

    def sweep(a0, a1, a2, a3, a4):
        t = -0.0625 * a0
        t = t + 0.25 * a1
        t = t + 0.625 * a2
        t = t + 0.25 * a3
        return t - 0.0625 * a4

    Z = [np.full(shp, np.nan) for _ in range(6)]
    for c in range(6):
        for k in listGI:
            kr, kl = (k + 1) % mz, (k - 1) % mz
            krr, kll = (kr + 1) % mz, (kl - 1) % mz
            Z[c][k + 2, J, I] = sweep(T[c][krr + 2, J, I], T[c][kr + 2, J, I], T[c][k + 2, J, I], T[c][kl + 2, J, I],
                                      T[c][kll + 2, J, I])
    X = [np.full(shp, np.nan) for _ in range(6)]
    for c in range(6):
        for k in listGI:
            s = Z[c][k + 2, J, I]
            X[c][k + 2, J, I] = sweep(np.roll(s, 2, axis=1), np.roll(s, 1, axis=1), s, np.roll(s, -1, axis=1),
                                      np.roll(s, -2, axis=1))
    Y = [np.full(shp, np.nan) for _ in range(6)]
    for c in range(6):
        sg = (-1.0 if c < 3 else 1.0) * (-1.0 if c % 3 == 1 else 1.0)
        for k in listGI:
            s = X[c][k + 2, J, I]                       # rows j = 0..my
            e = np.empty((my + 5, mx))                  # rows j = -2..my+2, only -1..my+1 used
            e[2:my + 3] = s
            e[1] = sg * s[1]
            e[my + 3] = sg * s[my - 1]
            out = s.copy()
            jj = np.arange(1, my)
            out[jj] = sweep(e[jj + 4], e[jj + 3], e[jj + 2], e[jj + 1], e[jj])
            Y[c][k + 2, J, I] = out
    F = [np.full(shp, np.nan) for _ in range(6)]
    i_src = np.array([(i % mx) + 2 for i in range(-2, mx + 2)])
    for c in range(6):
        for e in listG:
            if e & GUARD:
                continue
            k = e - 2
            ks = k % mz
            if 0 <= k < mz:
                F[c][e, J, I] = Y[c][e, J, I] + dc[c]
                F[c][e, J, :2] = A[c][ks + 2, J, :][:, i_src[:2]]
                F[c][e, J, mx + 2:] = A[c][ks + 2, J, :][:, i_src[mx + 2:]]
            else:
                F[c][e, J, :] = A[c][ks + 2, J, :][:, i_src]
            F[c][e, 0, :] = 0.0
            F[c][e, my + 2, :] = 0.0
    return F


@pytest.mark.parametrize("occ_planes", [[5, 6, 7], [0], [15], [0, 15], [14, 15, 0, 1], [3, 9], [16], list(range(16))])
def test_restricted_preparation_needs_no_other_plane(occ_planes):
    p = U.make_parm(8, 6, 16)
    f12 = U.smooth_fields(p, seed=11)
    a6 = [a.reshape(p.mz + 4, p.my + 3, p.mx + 4) for a in O.field_prep(p, f12)]
    occ = np.zeros(p.mz + 1, dtype=np.uint8)
    occ[occ_planes] = 1
    listB, listGI, listG = plane_sets(p.mz, occ)
    F = restricted_prep(p, f12, listB, listGI, listG)
    need = sorted({kp + 2 + d for kp in occ_planes for d in (-1, 0, 1)})     # extended planes a gather can touch
    assert sorted(e for e in listG if not e & GUARD) == need
    for c in range(6):
        for e in need:
            np.testing.assert_array_equal(F[c][e], a6[c][e], err_msg="component %d plane %d" % (c, e - 2))
    # guard planes sit right outside the prepared set
    for e in listG:
        if e & GUARD:
            e &= ~GUARD
            assert e not in need and ((e - 1) in need or (e + 1) in need)
    # economy: an interior slab blends its own planes + the stencil + the z-sweep halo, nothing else
    if occ_planes == [5, 6, 7]:
        assert listGI == [4, 5, 6, 7, 8] and listB == [2, 3, 4, 5, 6, 7, 8, 9, 10]
