"""The C oracle against a second, independent transcription of the same
Fortran (oracle/np_restatement.py, numpy).  The reference cannot be compiled
here (no Fortran compiler, no MPI) and ships no golden vectors, so parity
stays "unpinned" by the reference itself; two separate restatements that agree
to the last bit on seeded inputs -- field preparation, corrector with partbc
and the serial drive kick, predictor with the 18-node scatter and the
vmesh fold -- is the strongest pin available in this container."""
import numpy as np
import pytest

from oracle import np_restatement as N
from oracle import pyoracle as O
from tests import util as U


@pytest.fixture(scope="module")
def case():
    p = U.make_parm(10, 8, 12, Ez00=0.05)
    sp, ranfb = U.load_species(p, 80)
    f12 = U.smooth_fields(p, seed=17, ghost_nan=False)
    return p, sp, ranfb, f12


@pytest.mark.parametrize("ifil", [(1, 1, 1), (2, 0, 1)])
def test_field_preparation_bit_identical(case, ifil):
    p, _, _, f12 = case
    q = U.make_parm(p.mx, p.my, p.mz)
    q.ifilx, q.ifily, q.ifilz = ifil
    a6 = O.field_prep(q, f12)
    b6 = N.field_prep(q, f12)
    for c in range(6):
        np.testing.assert_array_equal(a6[c], b6[c], err_msg="component %d" % c)


@pytest.mark.parametrize("ksp", [1, 2])
def test_corrector_bit_identical(case, ksp):
    p, sp, ranfb, f12 = case
    a6 = O.field_prep(p, f12)
    qm, wm = U.QSPEC[ksp], U.WSPEC[ksp]
    # push a few particles onto the walls and seams so that every partbc branch runs
    arrs = [a.copy() for a in sp[ksp]]
    arrs[1][:4] = [1e-3, p.ymax - 1e-3, 0.5 * p.ymax, 0.2 * p.ymax]
    arrs[4][:4] = [-0.5, 0.5, 0.0, 0.0]
    arrs[0][4:6] = [-p.hx / 2 + 1e-3, p.xmax - p.hx / 2 - 1e-3]
    arrs[3][4:6] = [-0.5, 0.5]
    arrs[2][6:8] = [-p.hz / 2 + 1e-3, p.zmax - p.hz / 2 - 1e-3]
    arrs[5][6:8] = [-0.5, 0.5]
    a = [v.copy() for v in arrs]
    b = [v.copy() for v in arrs]
    # a ranfp state whose first few hundred draws contain values above 0.999, so that kicks do happen
    ranfb, found = 7331, False
    while not found:
        ranfb += 2
        s_, found = ranfb, False
        for _ in range(300):
            s_, u = N.ranfp_next(s_)
            found = found or u > 0.999
    st = np.array([ranfb], dtype=np.int32)
    r = O.fulmov(p, a6, *a, qm, wm, 0, nranks=1, ranfb=st)
    wkix, wkih, state = N.fulmov(p, a6, *b, qm, wm, 0, ranfb=ranfb)
    for c in range(6):
        np.testing.assert_array_equal(a[c], b[c], err_msg="coordinate %d" % c)
    assert int(st[0]) == state                                  # same number of ranfp draws
    assert r["wkix"] == wkix and r["wkih"] == wkih
    assert N.fulmov.kicks >= 1                                  # the kick itself was exercised


@pytest.mark.parametrize("ksp", [1, 2])
def test_predictor_moments_bit_identical(case, ksp):
    p, sp, ranfb, f12 = case
    a6 = O.field_prep(p, f12)
    qm, wm = U.QSPEC[ksp], U.WSPEC[ksp]
    arrs = [a.copy() for a in sp[ksp]]
    arrs[1][:2] = [1e-3, p.ymax - 1e-3]                         # reflected during the predicted move
    arrs[4][:2] = [-0.5, 0.5]
    r = O.fulmov(p, a6, *[v.copy() for v in arrs], qm, wm, 1, nranks=1, want_raw=True)
    wkix, wkih, raw, folded = N.fulmov(p, a6, *[v.copy() for v in arrs], qm, wm, 1)
    for m in range(4):
        np.testing.assert_array_equal(r["raw"][m], raw[m], err_msg="raw moment %d" % m)
        np.testing.assert_array_equal(r["mom"][m], folded[m], err_msg="folded moment %d" % m)
    assert r["wkix"] == wkix and r["wkih"] == wkih
    # what the fold does (Q1): interior planes next to the seams are overwritten, not accumulated
    f = folded[3].reshape(p.mz + 4, p.my + 3, p.mx + 4)
    assert np.all(f[2:p.mz + 2, 1:p.my + 2, 1 + 2] == 0.0) and np.all(f[2:p.mz + 2, 1:p.my + 2, p.mx - 2 + 2] == 0.0)


@pytest.mark.parametrize("ksp", [1, 2])
def test_loadpt_bit_identical(ksp):
    """The synthetic two-flux-bundle load (positions from ranfp, speeds from the tabulated inverse CDF and ranf,
    beam drift inside the bundles): the C oracle, which bench.py and the device loader are checked against, and the
    numpy transcription of F:8885-9040 give the same particles and leave both LCGs in the same state."""
    p = U.make_parm(6, 5, 8)
    a, sa, sb = O.loadpt(p, 7, U.vth(ksp), 0.0, U.VBEAM[ksp])
    b, ta, tb = N.loadpt(p, 7, U.vth(ksp), 0.0, U.VBEAM[ksp])
    assert (sa, sb) == (ta, tb)
    for c in range(6):
        np.testing.assert_array_equal(a[c], b[c], err_msg="component %d" % c)
    assert np.count_nonzero(a[3] != b[3] - 0.0) == 0 and np.any(np.abs(a[3]) > 0)
