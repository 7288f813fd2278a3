"""BASELINE configs[0] exactly: the two-flux-bundle load on a 32 x 32 x 32 grid, 20 particles per cell per species, 10 full
steps (predictor pair, new fields, corrector pair), run by 4 round-robin ranks as `mpiexec -n 4` of the reference would --
the CUDA path (4 rank contexts on this GPU) against the CPU oracle (pinned bit for bit to the reference's own fulmov by
tests/test_ref_pin.py).  Tolerances of the north star: moments <= 1e-10 relative L2 at every step, particles <= 1e-12
relative per step, every rank's ranfp state equal.  The measured errors go into BASELINE.md section 4."""
import json
import os

import numpy as np
import pytest

from tests import refcases as RC
from tests import util as U

pytestmark = pytest.mark.gpu


def test_config1_ten_steps_four_ranks():
    import mrg_b200 as mrg
    steps, nranks = 10, 4
    p, sp, ranfb, fsets = RC.loader_case(32, 32, 32, 20, steps, seed0=1000)
    assert len(sp[1][0]) == 655360                                   # np0 of config 1
    orc = RC.oracle_steps(p, sp, ranfb, fsets, nranks)
    gpu = RC.gpu_steps(mrg, p, sp, ranfb, fsets, nranks)
    worst_m, worst_wk = 0.0, 0.0
    for s in range(steps):
        for k in (1, 2):
            e = max(U.rel_l2(gpu["mom"][s][k][c], orc["mom"][s][k][c]) for c in range(4))
            assert e < 1e-10, (s, k, e)
            worst_m = max(worst_m, e)
            for a, b in zip(gpu["wk_pred"][s][k] + gpu["wk_corr"][s][k], tuple(orc["wk_pred"][s][k]) + tuple(orc["wk_corr"][s][k])):
                worst_wk = max(worst_wk, abs(a - b) / abs(b))
    assert worst_wk < 1e-10
    worst_p = max(U.particle_err(gpu["final"][k], orc["final"][k], p.hx, U.vth(k)) for k in (1, 2))
    assert worst_p < 1e-12 * steps, worst_p
    assert gpu["ranfb"] == orc["ranfb"]                               # same kicked set on every rank, every step
    rec = {"config": "32x32x32, 20 ppc, 10 steps, 4 round-robin ranks", "moments_rel_l2_max": worst_m,
           "wkix_wkih_rel_max": worst_wk, "particles_rel_err_max_after_10_steps": worst_p, "ranfp_states_equal": True}
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump(rec, open(os.path.join(out, "config1_parity.json"), "w"))
    print(rec)
