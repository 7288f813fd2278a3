"""Known-answer tests of the CPU oracle (SURVEY.md §4 invariants 1-5 and the
reference quirks Q1-Q3).  The reference ships no tests or golden vectors, so
these analytic properties are what pins the oracle ("parity unpinned")."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests import util as U


@pytest.fixture(scope="module")
def case():
    p = U.make_parm(8, 6, 8)
    sp, ranfb = U.load_species(p, 12)
    f12 = U.smooth_fields(p, seed=3)
    a6 = O.field_prep(p, f12)
    return p, sp, f12, a6


def test_prepared_fields_are_fully_defined(case):
    p, sp, f12, a6 = case
    for a in a6:
        assert np.isfinite(a).all()          # ghosts rebuilt from interior only


def test_partition_of_unity_raw_deposit(case):
    """Invariant 1: sum of the raw (pre-fold) q grid = qmult*N, of qjx = qmult*sum(vxj)."""
    p, sp, f12, a6 = case
    x, y, z, vx, vy, vz = [a.copy() for a in sp[2]]
    r = O.fulmov(p, a6, x, y, z, vx, vy, vz, -1.0, 1.0, 1, nranks=1, want_raw=True, want_pred=True)
    n = len(x)
    assert abs(r["raw"][3].sum() - (-1.0 * n)) < 1e-9 * n
    for c in range(3):
        s = -1.0 * r["pred"][3 + c].sum()
        assert abs(r["raw"][c].sum() - s) < 1e-10 * max(1.0, abs(s)) + 1e-9


def test_rank_count_independence(case):
    p, sp, f12, a6 = case
    arrs = sp[1]
    r1 = O.fulmov(p, a6, *[a.copy() for a in arrs], 1.0, 100.0, 1, nranks=1, want_raw=True)
    for nr in (2, 4, 8):
        rn = O.fulmov(p, a6, *[a.copy() for a in arrs], 1.0, 100.0, 1, nranks=nr, want_raw=True)
        for c in range(4):
            assert U.rel_l2(rn["raw"][c], r1["raw"][c]) < 1e-13
            assert U.rel_l2(rn["mom"][c], r1["mom"][c]) < 1e-13
        assert abs(rn["wkix"] - r1["wkix"]) < 1e-12 * abs(r1["wkix"])


def test_pure_rotation_conserves_speed():
    """Invariant 2: E=0 => |v'| = |v| for any ht*|B| (incl. Dt*wce > 10)."""
    p = U.make_parm(8, 6, 8, wce=0.0)
    n = O.mxyzA(p)
    rng = np.random.default_rng(5)
    N = 2000
    for bmag in (0.1, 1.0, 10.0, 100.0):
        f12 = [np.zeros(n) for _ in range(12)]
        for c, b in zip((3, 4, 5), (0.6 * bmag, -0.48 * bmag, 0.64 * bmag)):
            f12[c][:] = b
            f12[c + 6][:] = b
        a6 = O.field_prep(p, f12)
        x = rng.uniform(0.2 * p.xmax, 0.8 * p.xmax, N)
        y = rng.uniform(0.3 * p.ymax, 0.7 * p.ymax, N)   # away from the zero rows j=-1,my+1
        z = rng.uniform(0.2 * p.zmax, 0.8 * p.zmax, N)
        v = [rng.normal(scale=0.01, size=N) for _ in range(3)]
        v0 = np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
        # qmult/wmult = 1, dt = 1.2 -> ht = 0.6
        O.fulmov(p, a6, x, y, z, v[0], v[1], v[2], 1.0, 1.0, 0, nranks=1)
        v1 = np.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
        assert np.max(np.abs(v1 - v0) / v0) < 1e-13


def test_zero_dt_is_identity(case):
    """Invariant 3 (the it=0 call, F:678-689): dt=adt=hdt=0 leaves x,v alone and
    deposits the current state."""
    p0, sp, f12, a6 = case
    p = U.make_parm(8, 6, 8, dt=0.0)
    p.adt = 0.0
    p.hdt = 0.0
    arrs = [a.copy() for a in sp[2]]
    r = O.fulmov(p, a6, *arrs, -1.0, 1.0, 1, nranks=1, want_pred=True)
    for c in range(6):
        np.testing.assert_array_equal(arrs[c], sp[2][c])
        np.testing.assert_array_equal(r["pred"][c], sp[2][c])
    arrs0 = [a.copy() for a in sp[2]]
    O.fulmov(p, a6, *arrs0, -1.0, 1.0, 0, nranks=1)
    for c in range(6):
        np.testing.assert_array_equal(arrs0[c], sp[2][c])


def test_uniform_field_gather_and_unfiltered_ghosts():
    """Invariant 4 + Q2: a uniform field stays uniform through blend/outmesh/
    filter on rows 1..my-1; ghost rows j=-1,my+1 are zero."""
    p = U.make_parm(8, 6, 8, wce=0.2)
    n = O.mxyzA(p)
    f12 = [np.full(n, 0.03 * (c + 1)) for c in range(12)]
    a6 = O.field_prep(p, f12)
    shp = O.grid_shape(p)
    for c in range(6):
        a = a6[c].reshape(shp)
        expect = p.aimpl * 0.03 * (c + 1) + (1 - p.aimpl) * 0.03 * (c + 7) + (p.bxc if c == 3 else 0.0)
        np.testing.assert_allclose(a[:, 3:p.my - 1, :], expect, rtol=1e-13)   # rows j=2..my-2: pure 5-point average
        assert np.all(a[:, 0, :] == 0.0) and np.all(a[:, p.my + 2, :] == 0.0)  # j=-1, my+1


def test_lcg_skip_ahead_is_exact():
    """Invariant 5: ir_n = lambda^n * ir_0 mod 2^31 (F:9301)."""
    vals, s = O.ranfp_stream(7331, 1000)
    assert s == O.lcg_skip(7331, 1000)
    assert O.lcg_skip(7331, 0) == 7331
    assert O.lcg_skip(3021, 123456789) == O.lcg_skip(O.lcg_skip(3021, 123456000), 789)
    assert 0.0 < vals.min() and vals.max() < 1.0


def test_fold_quirk_q1():
    """Q1: vmesh x/z steps assign (planes i=1, mx-2, k=1, mz-2 end up zero for a
    particle deposit), the y step adds."""
    p = U.make_parm(8, 6, 8)
    sp, _ = U.load_species(p, 10)
    f12 = U.smooth_fields(p, seed=1)
    a6 = O.field_prep(p, f12)
    r = O.fulmov(p, a6, *[a.copy() for a in sp[2]], -1.0, 1.0, 1, nranks=1, want_raw=True)
    shp = O.grid_shape(p)
    q = r["mom"][3].reshape(shp)
    raw = r["raw"][3].reshape(shp)
    assert np.all(q[2:p.mz + 2, 1:p.my + 2, 2 + 1] == 0.0)            # i = 1
    assert np.all(q[2:p.mz + 2, 1:p.my + 2, 2 + p.mx - 2] == 0.0)      # i = mx-2
    assert np.all(q[2 + 1, 1:p.my + 2, 2:p.mx + 2] == 0.0)             # k = 1
    assert np.all(q[2 + p.mz - 2, 1:p.my + 2, 2:p.mx + 2] == 0.0)      # k = mz-2
    # y walls are sums (interior i, k not touched by the z step)
    np.testing.assert_allclose(q[5, 1, 5], raw[5, 1, 5] + raw[5, 0, 5], rtol=1e-15)
    assert abs(q.sum()) < abs(raw.sum())                                 # charge is not conserved by the fold


def test_reversed_y_weights_q3():
    """Q3: node jl=jp carries the fraction, node jr=jp+1 carries 1-fraction."""
    p = U.make_parm(8, 6, 8)
    n = O.mxyzA(p)
    a6 = [np.zeros(n) for _ in range(6)]
    x = np.array([3.0 * p.hx]); z = np.array([3.0 * p.hz])
    y = np.array([(2 + 0.25) * p.hy])
    v = [np.zeros(1) for _ in range(3)]
    pz = U.make_parm(8, 6, 8, dt=0.0); pz.adt = 0.0; pz.hdt = 0.0
    r = O.fulmov(pz, a6, x, y, z, v[0], v[1], v[2], 1.0, 1.0, 1, nranks=1, want_raw=True)
    q = r["raw"][3].reshape(O.grid_shape(p))
    row_jp = q[:, 1 + 2, :].sum()
    row_jp1 = q[:, 1 + 3, :].sum()
    assert abs(row_jp - 0.25) < 1e-9 and abs(row_jp1 - 0.75) < 1e-9


def test_loadpt_statistics_and_rng_state():
    p = U.make_parm(8, 6, 8)
    arrs, a, b = O.loadpt(p, 20, 0.2, 0.0, -0.35e-2)
    npr = 8 * 6 * 8 * 20
    assert len(arrs[0]) == npr
    assert b == O.lcg_skip(7331, 3 * npr) and a == O.lcg_skip(3021, 4 * npr)
    assert arrs[0].min() >= -p.hx / 2 and arrs[0].max() < p.xmax - p.hx / 2
    assert arrs[1].min() >= 0 and arrs[1].max() < p.ymax
