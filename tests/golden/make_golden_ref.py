"""Generates tests/golden/ref_*.npz with THE REFERENCE'S OWN CODE: oracle/pyref runs /root/reference/@mrg37-080A.f03's
init, loadpt, fulmov, partbc*, srimp1/2, outmesh3, filt3e, vmesh3/1, ranfp, emfld0/poissn (translated to C by oracle/f03c.py, compiled
by oracle/build_ref.py) on the seeded cases of tests/refcases.py, by simulated MPI ranks.  These fixtures are reference
output: tests/test_ref_pin.py holds the C oracle to them bit for bit, tests/test_gpu_reference_golden.py holds the CUDA path
to them within the north-star tolerances.  Needs /root/reference (this container only):
    python tests/golden/make_golden_ref.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as PR            # noqa: E402
from tests import refcases as RC          # noqa: E402
from tests import util as U               # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def digest(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def pack(name, case, nranks, store_inputs, sample):
    p, sp, ranfb, fsets = case
    ref = PR.reference_steps((p.mx, p.my, p.mz), (p.xmax, p.ymax, p.zmax), sp, fsets, nranks=nranks, ranfb_in=ranfb)
    out = {"grid": np.array([p.mx, p.my, p.mz]), "box": np.array([p.xmax, p.ymax, p.zmax]),
           "scalars": np.array([p.dt, p.aimpl, p.bxc, p.Ez00]), "nranks": np.array([nranks]), "ranfb_in": np.array([ranfb]),
           "ranfb_out": np.array(ref["ranfb"]), "steps": np.array([len(fsets)]), "sample": np.array([sample]),
           "consts": np.array([ref["consts"][k] for k in ("hxi", "hyi", "hzi", "xmaxe", "zmaxe", "adt", "hdt", "bxc")])}
    for k in (1, 2):
        out["in_sha_%d" % k] = digest(sp[k])
        if store_inputs:
            out["in_%d" % k] = np.stack(sp[k])
        out["out_sha_%d" % k] = digest(ref["final"][k])
        out["out_%d" % k] = np.stack([a[::sample] for a in ref["final"][k]])
    for s in range(len(fsets)):
        for k in (1, 2):
            out["mom_%d_%d" % (s, k)] = np.stack(ref["mom"][s][k])
            out["wk_%d_%d" % (s, k)] = np.array(list(ref["wk_pred"][s][k]) + list(ref["wk_corr"][s][k]))
    path = os.path.join(HERE, "ref_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(path, "%.0f KB" % (os.path.getsize(path) / 1024))


def pack_startup(name, grid, nranks, sample):
    """the reference's own initial condition: init, the it = 0 moment pass (dt = 0), emfld0 -- then one full step on the fields
    emfld0 defined (prefld / emfild are not on the path: the same fields serve the predictor and the corrector)"""
    p = U.make_parm(*grid)
    S = PR.reference_startup(grid, (p.xmax, p.ymax, p.zmax), nranks=nranks)
    assert S["ranks_agree"] and S["particles_unmoved"]
    f12 = S["fields"]
    ref = PR.reference_steps(grid, (p.xmax, p.ymax, p.zmax), S["particles"], [(f12, f12)], nranks=nranks, ranfb_in=S["ranfb"])
    out = {"grid": np.array(grid), "box": np.array([p.xmax, p.ymax, p.zmax]), "scalars": np.array([p.dt, p.aimpl, p.bxc, p.Ez00]),
           "nranks": np.array([nranks]), "ppc": np.array([32]), "ranfb_in": np.array([S["ranfb"]]), "ranfb_out": np.array(ref["ranfb"]),
           "sample": np.array([sample]), "fields": np.stack(f12)}
    for k in (1, 2):
        out["in_sha_%d" % k] = digest(S["particles"][k])
        out["mom0_%d" % k] = np.stack(S["mom0"][k])
        out["wk0_%d" % k] = np.array(S["wk0"][k])
        out["mom_0_%d" % k] = np.stack(ref["mom"][0][k])
        out["wk_0_%d" % k] = np.array(list(ref["wk_pred"][0][k]) + list(ref["wk_corr"][0][k]))
        out["out_sha_%d" % k] = digest(ref["final"][k])
        out["out_%d" % k] = np.stack([a[::sample] for a in ref["final"][k]])
    path = os.path.join(HERE, "ref_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(path, "%.0f KB" % (os.path.getsize(path) / 1024))


def pack_trans(name, grid, nranks, steps, sample):
    """the reference's own time cycle (oracle/pyref.ReferenceLoop: init, it = 0 pass, emfld0, then prefld -> fulmov x2 ->
    emfild (emcoef, cfpsol, bcgstb) -> fulmov x2 -> renewal): the fields every fulmov call saw are the reference's own
    self-consistent solution, not synthetic modes"""
    p = U.make_parm(*grid)
    out = {"grid": np.array(grid), "box": np.array([p.xmax, p.ymax, p.zmax]), "scalars": np.array([p.dt, p.aimpl, p.bxc, p.Ez00]),
           "nranks": np.array([nranks]), "ppc": np.array([32]), "steps": np.array([steps]), "sample": np.array([sample])}
    with PR.ReferenceLoop(grid, (p.xmax, p.ymax, p.zmax), nranks) as A:
        A.startup()
        out["ranfb_in"] = np.array([A.ranfb()[0]])
        for k in (1, 2):
            out["in_sha_%d" % k] = digest(A.particles()[k])
        for s in range(steps):
            A.begin_step()
            out["fpred_%d" % s] = np.stack(A.fields())
            wkp = A.fulmov(1)
            for k in (1, 2):
                out["mom_%d_%d" % (s, k)] = np.stack(A.moments(k))
            A.emfild()
            out["fcorr_%d" % s] = np.stack(A.fields())
            wkc = A.fulmov(0)
            for k in (1, 2):
                out["wk_%d_%d" % (s, k)] = np.array(list(wkp[k]) + list(wkc[k]))
            A.renew()
            assert A.ranks_agree()
        final = A.particles()
        out["ranfb_out"] = np.array(A.ranfb())
        out["e_max"] = np.array([max(float(np.abs(f).max()) for f in A.fields()[:3])])
    for k in (1, 2):
        out["out_sha_%d" % k] = digest(final[k])
        out["out_%d" % k] = np.stack([a[::sample] for a in final[k]])
    path = os.path.join(HERE, "ref_%s.npz" % name)
    np.savez_compressed(path, **out)
    print(path, "%.0f KB" % (os.path.getsize(path) / 1024), "max|E| %.3e" % out["e_max"][0])


def main():
    pack("loader_4r", RC.loader_case(6, 4, 6, 32, 3), 4, store_inputs=False, sample=4)
    pack("loader_1r", RC.loader_case(6, 4, 6, 32, 2), 1, store_inputs=False, sample=16)
    pack("edge_2r", RC.edge_case(), 2, store_inputs=True, sample=1)
    pack_startup("startup_2r", (8, 6, 8), 2, sample=8)
    pack_trans("trans_2r", (8, 6, 8), 2, steps=2, sample=8)


if __name__ == "__main__":
    main()
