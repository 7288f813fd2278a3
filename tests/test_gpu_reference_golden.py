"""CUDA path vs REFERENCE OUTPUT: tests/golden/ref_*.npz were written by the reference's own fulmov (translated from
/root/reference/@mrg37-080A.f03, see tests/golden/make_golden_ref.py), run by 1, 2 and 4 simulated MPI ranks over several
steps.  Tolerances are the north star's: moments <= 1e-10 relative L2, particles <= 1e-12 relative per step, the kicked
set / ranfp states exact.  Ranks are emulated by one context per rank on this GPU (round-robin ownership l = rank+1,
rank+1+N, ..., F:1162); the rank sum of the raw moments is done on the host in rank order, the fold by the oracle's vmesh."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from tests import refcases as RC
from tests import util as U

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"loader_4r": lambda: RC.loader_case(6, 4, 6, 32, 3), "loader_1r": lambda: RC.loader_case(6, 4, 6, 32, 2),
         "edge_2r": RC.edge_case}


@pytest.mark.parametrize("tile", [1, 0])
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_matches_reference_output(name, tile):
    import mrg_b200 as mrg
    G = np.load(os.path.join(GOLD, "ref_%s.npz" % name))
    p, sp, ranfb, fsets = CASES[name]()
    nranks, steps, sample = int(G["nranks"][0]), int(G["steps"][0]), int(G["sample"][0])
    gpu = RC.gpu_steps(mrg, p, sp, ranfb, fsets, nranks, tile=tile)
    worst_m = 0.0
    for s in range(steps):
        for k in (1, 2):
            ref = G["mom_%d_%d" % (s, k)]
            worst_m = max(worst_m, max(U.rel_l2(gpu["mom"][s][k][c], ref[c]) for c in range(4)))
            wref = G["wk_%d_%d" % (s, k)]
            got = gpu["wk_pred"][s][k] + gpu["wk_corr"][s][k]
            assert all(abs(a - b) <= 1e-10 * abs(b) for a, b in zip(got, wref))
    assert worst_m < 1e-10, worst_m
    worst_p = 0.0
    for k in (1, 2):
        worst_p = max(worst_p, U.particle_err([a[::sample] for a in gpu["final"][k]], list(G["out_%d" % k]), p.hx, U.vth(k)))
    assert worst_p < 1e-12 * steps, worst_p
    st = gpu["ranfb"]
    assert st == [int(v) for v in G["ranfb_out"]]      # same number of draws on every rank => same kicked set sizes


@pytest.mark.parametrize("tile", [1, 0])
def test_cuda_path_matches_reference_startup(tile):
    """The reference's own initial condition (tests/golden/ref_startup_2r.npz: init, the it = 0 moment pass with
    dt = adt = hdt = 0 (F:673-689), emfld0) and one full step on the fields emfld0 defined, by 2 round-robin ranks."""
    import mrg_b200 as mrg
    G = np.load(os.path.join(GOLD, "ref_startup_2r.npz"))
    grid = tuple(int(v) for v in G["grid"])
    p, p0 = U.make_parm(*grid), U.make_parm(*grid, dt=0.0)
    sp, ranfb = U.load_species(p, int(G["ppc"][0]))
    f12 = [np.ascontiguousarray(f) for f in G["fields"]]
    nranks, sample = int(G["nranks"][0]), int(G["sample"][0])
    # it = 0: every rank's contexts see zero fields and dt = 0; nothing moves, the moments are those emfld0 solves from
    zero = [np.zeros(O.mxyzA(p)) for _ in range(12)]
    par0 = mrg.StepParams(0.0, 0.0, 0.0, p0.aimpl, p0.bxc, p0.byc, p0.bzc, 1, 1, 1, 1, p0.Ez00, p0.zcent, p0.ycent1, p0.ycent2)
    for k in (1, 2):
        raw = [np.zeros(O.mxyzA(p)) for _ in range(4)]
        wk = [0.0, 0.0]
        for r in range(nranks):
            ctx = mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax)
            ctx.set_option("tile", tile)
            ctx.upload(k, *sp[k], first=r + 1, stride=nranks)
            ctx.sort(k, 0.0)
            ctx.set_fields(zero)
            wx, wh, _ = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 1, par0, ranfb)
            part = ctx.moments(k, folded=False)
            for c in range(4):
                raw[c] += part[c]
            wk[0] += wx
            wk[1] += wh
            ctx.close()
        O.vmesh3(p, raw[0], raw[1], raw[2])
        O.vmesh1(p, raw[3])
        ref = G["mom0_%d" % k]
        for c in range(4):
            den = float(np.linalg.norm(ref[c]))
            assert float(np.linalg.norm(raw[c] - ref[c])) <= 1e-10 * den + 1e-300, (k, c)
        assert abs(wk[0] - G["wk0_%d" % k][0]) <= 1e-10 * abs(G["wk0_%d" % k][0]) and abs(wk[1]) <= 1e-30 and G["wk0_%d" % k][1] == 0.0
    gpu = RC.gpu_steps(mrg, p, sp, ranfb, [(f12, f12)], nranks, tile=tile)
    for k in (1, 2):
        ref = G["mom_0_%d" % k]
        assert max(U.rel_l2(gpu["mom"][0][k][c], ref[c]) for c in range(4)) < 1e-10
        got = gpu["wk_pred"][0][k] + gpu["wk_corr"][0][k]
        assert all(abs(a - b) <= 1e-10 * abs(b) for a, b in zip(got, G["wk_0_%d" % k]))
        assert U.particle_err([a[::sample] for a in gpu["final"][k]], list(G["out_%d" % k]), p.hx, U.vth(k)) < 1e-12
    assert gpu["ranfb"] == [int(v) for v in G["ranfb_out"]]
