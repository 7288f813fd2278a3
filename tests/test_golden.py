"""Committed fixtures (tests/golden/fulmov_small.npz, made by make_golden.py):
the oracle must keep reproducing them bit for bit (CPU), and the CUDA path
must match them within the north-star tolerances (GPU, through the C ABI).
They pin the oracle against drift; the reference itself has no vectors."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import util as U

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fulmov_small.npz"))


def _parm():
    mx, my, mz = [int(v) for v in G["grid"]]
    dt, aimpl, wce, ez00 = [float(v) for v in G["scalars"]]
    p = O.make_parm(mx, my, mz, *[float(v) for v in G["box"]], dt, aimpl, wce, ez00)
    return p


def test_oracle_reproduces_golden_bit_exact():
    p = _parm()
    f_pred = [np.ascontiguousarray(a) for a in G["f_pred"]]
    f_corr = [np.ascontiguousarray(a) for a in G["f_corr"]]
    a6p = O.field_prep(p, f_pred)
    np.testing.assert_array_equal(np.stack(a6p), G["a6_pred"])
    a6c = O.field_prep(p, f_corr)
    st = np.array([int(G["ranfb_in"][0])], dtype=np.int32)
    for k in (1, 2):
        arrs = [np.ascontiguousarray(a) for a in G["in_%d" % k]]
        r = O.fulmov(p, a6p, *arrs, U.QSPEC[k], U.WSPEC[k], 1, nranks=1, ranfb=st, want_raw=True)
        np.testing.assert_array_equal(np.stack(r["mom"]), G["mom_%d" % k])
        np.testing.assert_array_equal(np.stack(r["raw"]), G["raw_%d" % k])
        assert r["wkix"] == G["wk_pred_%d" % k][0]
    for k in (1, 2):
        arrs = [np.ascontiguousarray(a) for a in G["in_%d" % k]]
        O.fulmov(p, a6c, *arrs, U.QSPEC[k], U.WSPEC[k], 0, nranks=1, ranfb=st)
        np.testing.assert_array_equal(np.stack(arrs), G["out_%d" % k])
    assert int(st[0]) == int(G["ranfb_out"][0])


def test_loader_reproduces_golden_inputs():
    p = _parm()
    sp, ranfb = U.load_species(p, 6)
    for k in (1, 2):
        np.testing.assert_array_equal(np.stack(sp[k]), G["in_%d" % k])
    assert ranfb == int(G["ranfb_in"][0])


@pytest.mark.gpu
def test_cuda_path_matches_golden():
    import mrg_b200 as mrg
    p = _parm()
    ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
    par = mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)
    ctx.set_fields([np.ascontiguousarray(a) for a in G["f_pred"]])
    got = ctx.prepared_fields(par)
    np.testing.assert_array_equal(np.stack(got), G["a6_pred"])            # bit-exact
    st = int(G["ranfb_in"][0])
    for k in (1, 2):
        ctx.upload(k, *[np.ascontiguousarray(a) for a in G["in_%d" % k]])
        ctx.sort(k, p.adt)
        wkix, wkih, _ = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 1, par, st)
        mom = ctx.moments(k)
        raw = ctx.moments(k, folded=False)
        for c in range(4):
            assert U.rel_l2(mom[c], G["mom_%d" % k][c]) < 1e-10
            assert U.rel_l2(raw[c], G["raw_%d" % k][c]) < 1e-10
        assert abs(wkix - G["wk_pred_%d" % k][0]) < 1e-10 * abs(G["wk_pred_%d" % k][0])
    ctx.set_fields([np.ascontiguousarray(a) for a in G["f_corr"]])
    for k in (1, 2):
        _, _, st = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 0, par, st)
        out = ctx.download(k, G["in_%d" % k].shape[1])
        assert U.particle_err(out, list(G["out_%d" % k]), p.hx, U.vth(k)) < 1e-12
    assert st == int(G["ranfb_out"][0])
    ctx.close()
