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
    n = len(sp[1][0])
    par = mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)
    ctxs = [mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax) for _ in range(nranks)]
    st = [ranfb] * nranks
    for r, ctx in enumerate(ctxs):
        ctx.set_option("tile", tile)
        for k in (1, 2):
            ctx.upload(k, *sp[k], first=r + 1, stride=nranks)
            ctx.sort(k, p.hdt)
    worst_m = worst_p = 0.0
    for s in range(steps):
        for ctx in ctxs:
            ctx.set_fields(fsets[s][0])
        for k in (1, 2):
            raw = [np.zeros(O.mxyzA(p)) for _ in range(4)]
            wk = [0.0, 0.0]
            for r, ctx in enumerate(ctxs):
                wx, wh, _ = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 1, par, st[r])
                part = ctx.moments(k, folded=False)
                for c in range(4):
                    raw[c] += part[c]
                wk[0] += wx
                wk[1] += wh
            O.vmesh3(p, raw[0], raw[1], raw[2])
            O.vmesh1(p, raw[3])
            ref = G["mom_%d_%d" % (s, k)]
            for c in range(4):
                worst_m = max(worst_m, U.rel_l2(raw[c], ref[c]))
            wref = G["wk_%d_%d" % (s, k)]
            assert abs(wk[0] - wref[0]) <= 1e-10 * abs(wref[0]) and abs(wk[1] - wref[1]) <= 1e-10 * abs(wref[1])
        for ctx in ctxs:
            ctx.set_fields(fsets[s][1])
        for k in (1, 2):
            for r, ctx in enumerate(ctxs):
                _, _, st[r] = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 0, par, st[r])
                ctx.sort(k, p.hdt)
    assert worst_m < 1e-10, worst_m
    for k in (1, 2):
        got = [np.zeros(n) for _ in range(6)]
        for r, ctx in enumerate(ctxs):
            ctx.download(k, n, r + 1, nranks, out=got)
        ref = list(G["out_%d" % k])
        worst_p = max(worst_p, U.particle_err([a[::sample] for a in got], ref, p.hx, U.vth(k)))
    assert worst_p < 1e-12 * steps, worst_p
    assert st == [int(v) for v in G["ranfb_out"]]      # same number of draws on every rank => same kicked set sizes
    for ctx in ctxs:
        ctx.close()
