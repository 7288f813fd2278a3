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
