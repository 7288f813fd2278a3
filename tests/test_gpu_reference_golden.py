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


@pytest.mark.parametrize("tile", [1, 0])
def test_cuda_path_matches_reference_time_cycle(tile):
    """tests/golden/ref_trans_2r.npz: the fields are the reference's own self-consistent solution (prefld + the implicit solve
    of emfild between the fulmov pairs); the CUDA path, given the fields each call saw, must match every step."""
    import mrg_b200 as mrg
    G = np.load(os.path.join(GOLD, "ref_trans_2r.npz"))
    grid = tuple(int(v) for v in G["grid"])
    p = U.make_parm(*grid)
    sp, ranfb = U.load_species(p, int(G["ppc"][0]))
    nranks, steps, sample = int(G["nranks"][0]), int(G["steps"][0]), int(G["sample"][0])
    fsets = [([np.ascontiguousarray(f) for f in G["fpred_%d" % s]], [np.ascontiguousarray(f) for f in G["fcorr_%d" % s]]) for s in range(steps)]
    gpu = RC.gpu_steps(mrg, p, sp, ranfb, fsets, nranks, tile=tile)
    for s in range(steps):
        for k in (1, 2):
            ref = G["mom_%d_%d" % (s, k)]
            assert max(U.rel_l2(gpu["mom"][s][k][c], ref[c]) for c in range(4)) < 1e-10
            got = gpu["wk_pred"][s][k] + gpu["wk_corr"][s][k]
            assert all(abs(a - b) <= 1e-10 * abs(b) for a, b in zip(got, G["wk_%d_%d" % (s, k)]))
    for k in (1, 2):
        assert U.particle_err([a[::sample] for a in gpu["final"][k]], list(G["out_%d" % k]), p.hx, U.vth(k)) < 1e-12 * steps
    assert gpu["ranfb"] == [int(v) for v in G["ranfb_out"]]


@pytest.mark.parametrize("grid,nranks,steps", [((8, 6, 8), 2, 3), ((32, 32, 32), 4, 10)], ids=["small", "config1_scale"])
def test_cuda_path_as_drop_in_inside_the_reference_time_cycle(grid, nranks, steps):
    """The drop-in claim on the GPU: the reference's own time cycle with its own field solver (oracle/_ref: prefld, emfild ->
    emcoef, cfpsol, bcgstb, emfld0; 2 simulated ranks), with the CUDA path in fulmov's place -- fields out of COMMON
    /fields/, summed folded moments into COMMON /srimp7/, particles resident on the GPU -- against the same cycle run
    entirely by the reference.  The closed loop goes through an iterative solver (Bi-CGSTAB, eps = 1e-5, F:4540), so the
    tolerance is not the particle path's: rounding-level noise on the moments (4e-16, injected into the C oracle in the same
    loop) moves E and B by ~1e-13 of their scale and the particles by ~4e-15 after three steps on the small grid: bound 1e-9,
    kicked sets and RNG states equal.  At BASELINE config-1 scale (32^3, 4 ranks, 10 steps) the bound is the solver's."""
    from oracle import pyref as PR
    if not PR.available():
        pytest.skip("oracle/_ref is not built and /root/reference is not here")
    import mrg_b200 as mrg
    p, p0 = U.make_parm(*grid), U.make_parm(*grid, dt=0.0)
    box = (p.xmax, p.ymax, p.zmax)
    with PR.ReferenceLoop(grid, box, nranks) as A:
        A.startup()
        for _ in range(steps):
            A.begin_step(); A.fulmov(1); A.emfild(); A.fulmov(0); A.renew()
        fa, pa, ra = A.fields(), A.particles(), A.ranfb()
    sp, ranfb = U.load_species(p, 32)
    gpu = RC.GpuRanks(mrg, p, sp, ranfb, nranks, lookahead=0.0)
    with PR.ReferenceLoop(grid, box, nranks) as B:
        def it0(L):
            mom, _ = gpu.predict(L.fields(), p0)
            for k in (1, 2):
                L.set_moments(k, mom[k])
        B.startup(it0)
        for k in (1, 2):
            for ctx in gpu.ctxs:
                ctx.sort(k, p.hdt)
        for _ in range(steps):
            B.begin_step()
            mom, _ = gpu.predict(B.fields())
            for k in (1, 2):
                B.set_moments(k, mom[k])
            B.emfild()
            gpu.correct(B.fields())
            B.renew()
        fb = B.fields()
    got = gpu.download()
    st = list(gpu.st)
    gpu.close()
    e_scale = max(float(np.abs(f).max()) for f in fa[:3])
    b_scale = max(float(np.abs(f).max()) for f in fa[3:6])
    assert e_scale > 1e-3
    err_e = max(float(np.abs(a - b).max()) for a, b in zip(fa[:3], fb[:3])) / e_scale
    err_b = max(float(np.abs(a - b).max()) for a, b in zip(fa[3:6], fb[3:6])) / b_scale
    err_p = max(U.particle_err(got[k], pa[k], p.hx, U.vth(k)) for k in (1, 2))
    print("closed loop %s x %d ranks after %d steps: E %.2e  B %.2e  particles %.2e" % (grid, nranks, steps, err_e, err_b, err_p))
    # 8 x 6 x 8: the solve converges hard and the loop stays at rounding level.  32^3: the reference's Bi-CGSTAB stops at
    # eps = 1e-5 (F:4540) and turns 4e-16 of noise on the moments into 2e-7 .. 2e-6 on E within ten steps -- measured with the
    # C oracle in the same loop (tests/test_ref_pin.py::test_closed_loop_amplifies_rounding_noise_at_32_cubed); the CUDA path
    # landed at 2.4e-6 there.  The bound for that case is the solver's tolerance, not the particle path's.
    bound = 1e-9 if grid == (8, 6, 8) else 1e-4       # ten solves at eps = 1e-5
    assert err_e < bound and err_b < bound and err_p < bound, (err_e, err_b, err_p)
    if grid == (8, 6, 8):       # at solver-level deviations a particle next to the drive slab's edge may change sides: one more draw
        assert st == ra
