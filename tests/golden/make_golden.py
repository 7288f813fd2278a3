"""Generates tests/golden/fulmov_small.npz from the CPU oracle.

A REGRESSION pin of the oracle restatement (inputs and outputs of one predictor
+ one corrector call per species on an 8x6x8 grid).  The independent pin --
fixtures written by the reference's own code -- is tests/golden/ref_*.npz, made
by make_golden_ref.py through oracle/_ref.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O          # noqa: E402
from tests import util as U              # noqa: E402


def main():
    p = U.make_parm(8, 6, 8)
    sp, ranfb = U.load_species(p, 6)
    f_pred = U.smooth_fields(p, seed=31, ghost_nan=False)
    f_corr = U.smooth_fields(p, seed=32, ghost_nan=False)
    out = {"grid": np.array([p.mx, p.my, p.mz]), "box": np.array([p.xmax, p.ymax, p.zmax]),
           "scalars": np.array([p.dt, p.aimpl, p.bxc, p.Ez00]), "ranfb_in": np.array([ranfb]),
           "f_pred": np.stack(f_pred), "f_corr": np.stack(f_corr)}
    st = np.array([ranfb], dtype=np.int32)
    a6p = O.field_prep(p, f_pred)
    a6c = O.field_prep(p, f_corr)
    out["a6_pred"] = np.stack(a6p)
    for k in (1, 2):
        out["in_%d" % k] = np.stack(sp[k])
        arrs = [a.copy() for a in sp[k]]
        r = O.fulmov(p, a6p, *arrs, U.QSPEC[k], U.WSPEC[k], 1, nranks=1, ranfb=st, want_raw=True)
        out["mom_%d" % k] = np.stack(r["mom"])
        out["raw_%d" % k] = np.stack(r["raw"])
        out["wk_pred_%d" % k] = np.array([r["wkix"], r["wkih"]])
    for k in (1, 2):
        arrs = [a.copy() for a in sp[k]]
        r = O.fulmov(p, a6c, *arrs, U.QSPEC[k], U.WSPEC[k], 0, nranks=1, ranfb=st)
        out["out_%d" % k] = np.stack(arrs)
        out["wk_corr_%d" % k] = np.array([r["wkix"], r["wkih"]])
    out["ranfb_out"] = np.array([int(st[0])])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fulmov_small.npz"), **out)
    print("written", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
