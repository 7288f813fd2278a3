"""Seeded inputs shared by the reference-pin tests, the golden-vector generator and the GPU parity tests.
Nothing here touches the reference; the cases are run through oracle/pyref (the translated reference), the C oracle and
the CUDA path by their respective tests."""
import numpy as np

from oracle import pyoracle as O
from tests import util as U


def loader_case(mx, my, mz, ppc, steps, seed0=100):
    """two-flux-bundle load of the oracle's loader (pinned bit for bit to the reference's loadpt) + `steps` pairs of
    smooth field sets (before / after the field solve)"""
    p = U.make_parm(mx, my, mz)
    sp, ranfb = U.load_species(p, ppc)
    fsets = [(U.smooth_fields(p, seed=seed0 + 2 * s), U.smooth_fields(p, seed=seed0 + 2 * s + 1)) for s in range(steps)]
    return p, sp, ranfb, fsets


def edge_case(mx=8, my=6, mz=8, n=2048, seed=42):
    """particles on / next to the periodic seams, the walls and cell boundaries, fast enough to
    wrap and reflect; same construction as tests/test_gpu_parity.py::test_edge_particles"""
    p = U.make_parm(mx, my, mz)
    rng = np.random.default_rng(seed)
    x = rng.uniform(-p.hx / 2, p.xmax - p.hx / 2, n)
    y = rng.uniform(0, p.ymax, n)
    z = rng.uniform(-p.hz / 2, p.zmax - p.hz / 2, n)
    v = [rng.normal(scale=0.3, size=n) for _ in range(3)]
    sx = [np.nextafter(-p.hx / 2, 1), np.nextafter(p.xmax - p.hx / 2, 0), 0.5 * p.hx, np.nextafter(0.5 * p.hx, 0),
          np.nextafter(0.5 * p.hx, 1), 0.0, 1.5 * p.hx, -p.hx / 2, p.xmax - p.hx / 2]
    for q, val in enumerate(sx):
        x[q] = val
        z[q + 16] = val * p.hz / p.hx
    y[32:42] = [np.nextafter(0, 1), np.nextafter(p.ymax, 0), p.hy, np.nextafter(p.hy, 0), np.nextafter(p.hy, 1),
                p.ymax - 1e-9, 1e-9, (p.my - 1) * p.hy, 0.0, p.ymax]
    v[1][32] = -0.5; v[1][33] = 0.5; v[1][37] = 0.5; v[1][38] = -0.5
    v[0][0] = -0.5; v[0][1] = 0.5; v[2][16] = -0.5; v[2][17] = 0.5
    v[1][40] = 0.0; v[1][41] = 0.0          # exactly on the walls with no motion: the jp >= my branch of F:1191 / F:2285
    # a few particles inside the drive slab so that the kick draws (F:1343-1353)
    k0 = 64
    m = 400
    z[k0:k0 + m] = p.zcent + rng.uniform(-0.14, 0.14, m) * p.zmax
    y[k0:k0 + m] = np.where(rng.uniform(size=m) < 0.5, p.ycent1, p.ycent2) + rng.uniform(-0.02, 0.02, m) * p.ymax
    for c in range(3):
        v[c][k0:k0 + m] *= 0.05
    arrs = [np.ascontiguousarray(a) for a in [x, y, z] + v]
    sp = {1: arrs, 2: [a.copy() for a in arrs]}
    fsets = [(U.smooth_fields(p, seed=7), U.smooth_fields(p, seed=8))]
    return p, sp, 7331, fsets


def oracle_steps(p, sp, ranfb, fsets, nranks):
    """the same call sequence as pyref.reference_steps through the C oracle"""
    arrs = {k: [a.copy() for a in sp[k]] for k in (1, 2)}
    st = np.full(nranks, ranfb, dtype=np.int32)
    out = {"mom": [], "wk_pred": [], "wk_corr": []}
    for f_pred, f_corr in fsets:
        a6 = O.field_prep(p, f_pred)
        mom, wk = {}, {}
        for k in (1, 2):
            r = O.fulmov(p, a6, *arrs[k], U.QSPEC[k], U.WSPEC[k], 1, nranks=nranks, ranfb=st)
            mom[k], wk[k] = r["mom"], (r["wkix"], r["wkih"])
        out["mom"].append(mom)
        out["wk_pred"].append(wk)
        a6 = O.field_prep(p, f_corr)
        wk = {}
        for k in (1, 2):
            r = O.fulmov(p, a6, *arrs[k], U.QSPEC[k], U.WSPEC[k], 0, nranks=nranks, ranfb=st)
            wk[k] = (r["wkix"], r["wkih"])
        out["wk_corr"].append(wk)
    out["final"] = arrs
    out["ranfb"] = [int(v) for v in st]
    return out


class GpuRanks:
    """`nranks` round-robin ranks of the CUDA path emulated by one context per rank on this GPU (l = rank+1, rank+1+N, ...,
    F:1162): the rank sum of the raw moments is done on the host in rank order (mpi_allreduce, F:2379-2384, 2533), the fold by
    the oracle's vmesh3/vmesh1; every rank keeps its own ranfp state."""

    def __init__(self, mrg, p, sp, ranfb, nranks, tile=1, lookahead=None):
        self.mrg, self.p, self.nranks, self.n = mrg, p, nranks, len(sp[1][0])
        self.ctxs = [mrg.MrgContext(p.mx, p.my, p.mz, p.xmax, p.ymax, p.zmax) for _ in range(nranks)]
        self.st = [ranfb] * nranks
        for r, ctx in enumerate(self.ctxs):
            ctx.set_option("tile", tile)
            for k in (1, 2):
                ctx.upload(k, *sp[k], first=r + 1, stride=nranks)
                ctx.sort(k, p.hdt if lookahead is None else lookahead)

    def params(self, p):
        return self.mrg.StepParams(p.dt, p.adt, p.hdt, p.aimpl, p.bxc, p.byc, p.bzc, 1, 1, 1, 1, p.Ez00, p.zcent, p.ycent1, p.ycent2)

    def predict(self, f12, p=None):
        """ipc = 1 for ions then electrons on the fields f12: ({ksp: folded summed moments}, {ksp: (wkix, wkih)})"""
        p = p or self.p
        par = self.params(p)
        for ctx in self.ctxs:
            ctx.set_fields(f12)
        mom, wk = {}, {}
        for k in (1, 2):
            raw = [np.zeros(O.mxyzA(p)) for _ in range(4)]
            w = [0.0, 0.0]
            for r, ctx in enumerate(self.ctxs):
                wx, wh, _ = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 1, par, self.st[r])
                part = ctx.moments(k, folded=False)
                for c in range(4):
                    raw[c] += part[c]
                w[0] += wx
                w[1] += wh
            O.vmesh3(p, raw[0], raw[1], raw[2])
            O.vmesh1(p, raw[3])
            mom[k], wk[k] = raw, tuple(w)
        return mom, wk

    def correct(self, f12, p=None):
        p = p or self.p
        par = self.params(p)
        for ctx in self.ctxs:
            ctx.set_fields(f12)
        wk = {}
        for k in (1, 2):
            w = [0.0, 0.0]
            for r, ctx in enumerate(self.ctxs):
                wx, wh, self.st[r] = ctx.fulmov(k, U.QSPEC[k], U.WSPEC[k], 0, par, self.st[r])
                ctx.sort(k, p.hdt)
                w[0] += wx
                w[1] += wh
            wk[k] = tuple(w)
        return wk

    def download(self):
        final = {}
        for k in (1, 2):
            final[k] = [np.zeros(self.n) for _ in range(6)]
            for r, ctx in enumerate(self.ctxs):
                ctx.download(k, self.n, r + 1, self.nranks, out=final[k])
        return final

    def close(self):
        for ctx in self.ctxs:
            ctx.close()


def gpu_steps(mrg, p, sp, ranfb, fsets, nranks, tile=1):
    """the same call sequence as oracle_steps through the CUDA path (GpuRanks).  Returns the same dict."""
    G = GpuRanks(mrg, p, sp, ranfb, nranks, tile=tile)
    out = {"mom": [], "wk_pred": [], "wk_corr": []}
    for f_pred, f_corr in fsets:
        mom, wk = G.predict(f_pred)
        out["mom"].append(mom)
        out["wk_pred"].append(wk)
        out["wk_corr"].append(G.correct(f_corr))
    out["final"] = G.download()
    out["ranfb"] = list(G.st)
    G.close()
    return out
