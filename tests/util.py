"""Shared builders for the parity tests (synthetic two-flux-bundle loads and
smooth analytic fields; seeds 3021/7331 as in the reference, F:9255-9256)."""
import numpy as np

from oracle import pyoracle as O

# rec_3d80A:4-9 (datum1)
HX, HY, HZ = 300.0 / 40, 600.0 / 72, 300.0 / 40
QSPEC = {1: 1.0, 2: -1.0}
WSPEC = {1: 100.0, 2: 1.0}
VBEAM = {1: 0.35e-2, 2: -0.35e-2}
VETH, TE_BY_TI = 0.2, 1.0


def vth(ksp):
    """F:8589-8594"""
    return VETH / np.sqrt(TE_BY_TI * WSPEC[1]) if ksp == 1 else VETH


def make_parm(mx, my, mz, dt=1.2, aimpl=0.6, wce=0.2, Ez00=0.25e-2):
    return O.make_parm(mx, my, mz, HX * mx, HY * my, HZ * mz, dt, aimpl, wce, Ez00)


def load_species(p, ppc):
    """Both species exactly as init does (F:8649-8658): rantbl before each
    loadpt, electrons take the ion positions (ipleql)."""
    out = {}
    for ksp in (1, 2):
        arrs, a, b = O.loadpt(p, ppc, vth(ksp), 0.0, VBEAM[ksp])
        out[ksp] = arrs
        ranfb = b
    for c in range(3):
        out[2][c][:] = out[1][c]
    return out, ranfb


def idx(p, i, j, k):
    return (i + 2) + (p.mx + 4) * ((j + 1) + (p.my + 3) * (k + 2))


def smooth_fields(p, seed=0, amp_e=1e-2, amp_b=0.03, ghost_nan=True):
    """12 field arrays (ex..bz, ex0..bz0): smooth flux-bundle-like modes on the
    interior points the reference defines (i<mx, j<=my, k<mz).  Ghost elements
    are NaN: fulmov must never read them (F:1127-1139 touches the interior
    only and outmesh3 rebuilds the ghosts)."""
    rng = np.random.default_rng(seed)
    n = O.mxyzA(p)
    shp = O.grid_shape(p)
    k, j, i = np.meshgrid(np.arange(-2, p.mz + 2), np.arange(-1, p.my + 2), np.arange(-2, p.mx + 2), indexing="ij")
    X = 2 * np.pi * i / p.mx
    Y = np.pi * j / p.my
    Z = 2 * np.pi * k / p.mz
    inside = (i >= 0) & (i < p.mx) & (j >= 0) & (j <= p.my) & (k >= 0) & (k < p.mz)
    out = []
    for c in range(12):
        amp = amp_e if (c % 6) < 3 else amp_b
        a = rng.normal(size=6)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        f = amp * (a[0] * np.sin(X + ph[0]) * np.cos(Y) + a[1] * np.cos(2 * Z + ph[1]) * np.sin(Y)
                   + a[2] * np.sin(X + Z + ph[2]) + 0.3 * a[3] * np.cos(3 * X) * np.cos(2 * Y) * np.sin(Z)
                   + 0.05 * a[4] * rng.normal(size=shp))
        if c >= 6:  # the "0" copies differ a little from the new fields
            f = f * (1.0 + 0.05 * a[5])
        f = np.where(inside, f, np.nan if ghost_nan else 0.0)
        assert f.size == n
        out.append(np.ascontiguousarray(f.reshape(-1)))
    return out


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def particle_err(got, ref, pos_floor, vel_floor):
    """max over particles and components of |got-ref| / max(|ref|, floor):
    positions are floored at one cell size, velocities at the species'
    thermal speed (a relative error of a component that happens to be ~0
    says nothing)."""
    e = 0.0
    for c in range(6):
        fl = pos_floor if c < 3 else vel_floor
        e = max(e, float(np.max(np.abs(got[c] - ref[c]) / np.maximum(np.abs(ref[c]), fl))))
    return e
