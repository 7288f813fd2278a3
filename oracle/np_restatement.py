"""Second, independent restatement of the /fulmov/ path in numpy.

TEST INFRASTRUCTURE ONLY (tests/test_oracle_crosscheck.py).  The reference
cannot be compiled here (Fortran 2003 + MPI, no compiler), so the C oracle
(fulmov_oracle.c) cannot be pinned against reference output.  What can be done
is to transcribe the same Fortran a second time, separately and in another
language, and demand that the two transcriptions agree to the last bit: a slip
in either one (a swapped index, a wrong weight, a different association) shows
up as a mismatch.  Everything below was written from @mrg37-080A.f03 directly
(F:n = its line n), not from the C file; expressions keep the association of
the source, numpy evaluates them in IEEE double without contraction, and the
scatter uses np.add.at, which applies the updates in particle order like the
serial loop of the source.

Arrays are the reference's (-2:mx+1,-1:my+1,-2:mz+1), i fastest, flattened;
here viewed as a[k+2, j+1, i+2].
"""
import numpy as np


def _view(p, a):
    return a.reshape(p.mz + 4, p.my + 3, p.mx + 4)


def outmesh3(p, a):
    """F:3088-3148, one array"""
    mx, my, mz = p.mx, p.my, p.mz
    K, J = slice(2, mz + 2), slice(1, my + 2)                 # k = 0..mz-1, j = 0..my
    for i in (-2, -1):                                        # a(i,j,k) = a(i+mx,j,k)
        a[K, J, i + 2] = a[K, J, i + mx + 2]
    for i in (mx, mx + 1):                                    # a(i,j,k) = a(i-mx,j,k)
        a[K, J, i + 2] = a[K, J, i - mx + 2]
    a[K, 0, :] = 0.0                                          # j = -1, all i
    a[K, my + 2, :] = 0.0                                     # j = my+1
    for k in (-2, -1):                                        # a(i,j,k) = a(i,j,k+mz), all i, j
        a[k + 2, :, :] = a[k + mz + 2, :, :]
    for k in (mz, mz + 1):
        a[k + 2, :, :] = a[k - mz + 2, :, :]


def filt3e(p, e3, dc, sym):
    """F:7351-7506 on three arrays (views); weights (-1,4,10,4,-1)/16"""
    mx, my, mz = p.mx, p.my, p.mz
    K, J, I = slice(2, mz + 2), slice(1, my + 2), slice(2, mx + 2)
    for c in range(3):
        e3[c][K, J, I] = e3[c][K, J, I] - dc[c]
    for _ in range(p.ifilz):
        for c in range(3):
            a = e3[c][K, J, I].copy()                         # a[k, j, i], k = 0..mz-1
            kr = np.roll(np.arange(mz), -1); kl = np.roll(np.arange(mz), 1)      # pzr, pzl
            krr, kll = kr[kr], kl[kl]
            e3[c][K, J, I] = (-0.0625 * a[krr] + 0.25 * a[kr] + 0.625 * a + 0.25 * a[kl]) - 0.0625 * a[kll]
    for _ in range(p.ifilx):
        for c in range(3):
            a = e3[c][K, J, I].copy()
            ir = np.roll(np.arange(mx), -1); il = np.roll(np.arange(mx), 1)
            irr, ill = ir[ir], il[il]
            e3[c][K, J, I] = (-0.0625 * a[:, :, ill] + 0.25 * a[:, :, il] + 0.625 * a + 0.25 * a[:, :, ir]) - 0.0625 * a[:, :, irr]
    sgn = (sym, -sym, sym)
    for _ in range(p.ifily):
        for c in range(3):
            a = np.zeros((mz, my + 5, mx))                    # rows j = -2..my+2 (only -1..my+1 are read)
            a[:, 2:my + 3] = e3[c][K, J, I]
            a[:, 1] = sgn[c] * e3[c][K, 1 + 1, I]             # a(-1) = +-sym * e(1)
            a[:, my + 3] = sgn[c] * e3[c][K, my - 1 + 1, I]   # a(my+1) = +-sym * e(my-1)
            j = np.arange(1, my)                              # j = 1..my-1
            e3[c][K, 2:my + 1, I] = (-0.0625 * a[:, j + 4] + 0.25 * a[:, j + 3] + 0.625 * a[:, j + 2]
                                     + 0.25 * a[:, j + 1]) - 0.0625 * a[:, j]
    for c in range(3):
        e3[c][K, J, I] = e3[c][K, J, I] + dc[c]


def field_prep(p, f12):
    """F:1127-1148: the six prepared arrays exa..bza (flattened, extended layout)"""
    mx, my, mz = p.mx, p.my, p.mz
    K, J, I = slice(2, mz + 2), slice(1, my + 2), slice(2, mx + 2)
    f = [_view(p, a) for a in f12]
    out = [np.zeros((mz + 4, my + 3, mx + 4)) for _ in range(6)]
    dc = (0.0, 0.0, 0.0, p.bxc, p.byc, p.bzc)
    for c in range(6):
        v = p.aimpl * f[c][K, J, I] + (1.0 - p.aimpl) * f[c + 6][K, J, I]
        out[c][K, J, I] = v + dc[c] if c >= 3 else v
    for c in range(6):
        outmesh3(p, out[c])
    filt3e(p, out[0:3], (0.0, 0.0, 0.0), -1.0)
    filt3e(p, out[3:6], (p.bxc, p.byc, p.bzc), 1.0)
    return [a.reshape(-1) for a in out]


def partbc(p, x, y, z, vy=None):
    """F:1856-1879 (vy given) / F:1928-1949 (partbcEST): in place, applied once"""
    dx, dz = p.hx / 2, p.hz / 2
    hi = x >= p.xmax - dx
    lo = ~hi & (x <= -dx)
    x[hi] = x[hi] - p.xmaxe
    x[lo] = x[lo] + p.xmaxe
    hi = y >= p.ymax
    lo = ~hi & (y <= 0.0)
    y[hi] = 2.0 * p.ymax - y[hi]
    y[lo] = -y[lo]
    if vy is not None:
        vy[hi | lo] = -vy[hi | lo]
    hi = z >= p.zmax - dz
    lo = ~hi & (z <= -dz)
    z[hi] = z[hi] - p.zmaxe
    z[lo] = z[lo] + p.zmaxe


def _cells(p, rx, ry, rz):
    ip = (p.hxi * rx + 0.500000001).astype(np.int64)          # int() truncates; the arguments are >= 0
    jp = (p.hyi * ry + 0.000000001).astype(np.int64)
    kp = (p.hzi * rz + 0.500000001).astype(np.int64)
    return ip, jp, kp


def _weights(p, rx, ry, rz, ip, jp, kp):
    xx = p.hxi * rx - ip
    fx = (0.5 * (0.5 - xx) * (0.5 - xx), 0.75 - xx * xx, 0.5 * (0.5 + xx) * (0.5 + xx))     # l, c, r
    zz = p.hzi * rz - kp
    fz = (0.5 * (0.5 - zz) * (0.5 - zz), 0.75 - zz * zz, 0.5 * (0.5 + zz) * (0.5 + zz))
    fyl = p.hyi * ry - jp
    fyr = 1.0 - fyl
    return fx, fyl, fyr, fz


def ranfp_next(ir):
    """F:9286-9305: ir <- iand(lambda*ir, 2^31-1) in int32 arithmetic; returns (ir, ir * 2^-31)"""
    ir = (48828125 * ir) & 0xFFFFFFFF
    ir &= 0x7FFFFFFF
    return ir, ir * 0.5 ** 31


def fulmov(p, a6, x, y, z, vx, vy, vz, qmult, wmult, ipc, ranfb=7331):
    """One rank (ipar = 1, size = 1) of F:1150-1390.  ipc = 0 updates x..vz in place and returns
    (wkix, wkih, ranfb_state); ipc >= 1 returns (wkix, wkih, raw[4], folded[4])."""
    mx, my, mz = p.mx, p.my, p.mz
    A = [_view(p, a) for a in a6]
    hh = p.dt * qmult / wmult
    ht = 0.5 * hh
    ht2 = ht ** 2
    rx = x + p.hdt * vx
    ry = y + p.hdt * vy
    rz = z + p.hdt * vz
    partbc(p, rx, ry, rz)
    ip, jp, kp = _cells(p, rx, ry, rz)
    (fxl, fxc, fxr), fyl, fyr, (fzl, fzc, fzr) = _weights(p, rx, ry, rz, ip, jp, kp)
    jl, jr = jp.copy(), jp + 1
    top, bot = jp >= my, jp < 0
    jr[top], jl[top], fyr[top], fyl[top] = my + 1, my, 0.0, 1.0
    jr[bot], jl[bot], fyr[bot], fyl[bot] = 0, -1, 1.0, 0.0
    il, ic, ir = ip - 1 + 2, ip + 2, ip + 1 + 2
    kl, kc, kr = kp - 1 + 2, kp + 2, kp + 1 + 2

    def g(F):
        def row(j, k):
            return F[k, j + 1, ir] * fxr + F[k, j + 1, ic] * fxc + F[k, j + 1, il] * fxl
        return (fyr * (row(jr, kr) * fzr + row(jr, kc) * fzc + row(jr, kl) * fzl)
                + fyl * (row(jl, kr) * fzr + row(jl, kc) * fzc + row(jl, kl) * fzl))

    exi, eyi, ezi, bxi, byi, bzi = (g(F) for F in A)
    bsqi = bxi ** 2 + byi ** 2 + bzi ** 2
    acx = exi + vy * bzi - vz * byi
    acy = eyi + vz * bxi - vx * bzi
    acz = ezi + vx * byi - vy * bxi
    ach = exi * bxi + eyi * byi + ezi * bzi
    den = 1.0 + ht2 * bsqi
    dvx = (acx + ht2 * ach * bxi + ht * (acy * bzi - acz * byi)) / den
    dvy = (acy + ht2 * ach * byi + ht * (acz * bxi - acx * bzi)) / den
    dvz = (acz + ht2 * ach * bzi + ht * (acx * byi - acy * bxi)) / den
    wkix = wkih = 0.0
    for t in 0.5 * (acx ** 2 + acy ** 2 + acz ** 2):          # the source accumulates particle by particle
        wkix = wkix + t
    for t in 0.5 * ach ** 2:
        wkih = wkih + t
    if ipc == 0:
        x[:] = x + p.dt * (vx + 0.5 * hh * dvx)
        y[:] = y + p.dt * (vy + 0.5 * hh * dvy)
        z[:] = z + p.dt * (vz + 0.5 * hh * dvz)
        vx[:] = vx + hh * dvx
        vy[:] = vy + hh * dvy
        vz[:] = vz + hh * dvz
        partbc(p, x, y, z, vy)
        bxa = A[3]
        state = ranfb
        slab = (np.abs(z - p.zcent) < 0.15 * p.zmax) & ((np.abs(y - p.ycent2) < 0.025 * p.ymax) | (np.abs(y - p.ycent1) < 0.025 * p.ymax))
        fulmov.kicks = 0
        for l in np.nonzero(slab)[0]:                         # F:1342-1364, l order, one draw per slab particle
            state, u = ranfp_next(state)
            if u > 0.999:
                fulmov.kicks += 1
                i_, j_, k_ = int(p.hxi * x[l] + 0.500000001), int(p.hyi * y[l] + 0.000000001), int(p.hzi * z[l] + 0.500000001)
                vy0 = p.Ez00 / bxa[k_ + 2, j_ + 1, i_ + 2]
                if abs(y[l] - p.ycent2) < 0.05 * p.ymax:
                    vy[l] = vy[l] - vy0
                elif abs(y[l] - p.ycent1) < 0.05 * p.ymax:
                    vy[l] = vy[l] + vy0
        return wkix, wkih, state
    vxj = vx + p.aimpl * hh * dvx
    vyj = vy + p.aimpl * hh * dvy
    vzj = vz + p.aimpl * hh * dvz
    rx = x + p.adt * (vx + 0.5 * hh * dvx)
    ry = y + p.adt * (vy + 0.5 * hh * dvy)
    rz = z + p.adt * (vz + 0.5 * hh * dvz)
    partbc(p, rx, ry, rz, vyj)
    # srimp1 (F:2273-2374) and srimp2 (F:2471-2529): same cells and weights, no fy override in the wall branches
    ip, jp, kp = _cells(p, rx, ry, rz)
    fx, fyl, fyr, fz = _weights(p, rx, ry, rz, ip, jp, kp)
    jl, jr = jp.copy(), jp + 1
    top, bot = jp >= my, jp < 0
    jr[top], jl[top] = my + 1, my
    jr[bot], jl[bot] = 0, -1
    ii = (ip - 1 + 2, ip + 2, ip + 1 + 2)
    kk = (kp - 1 + 2, kp + 2, kp + 1 + 2)
    nx, ny = mx + 4, my + 3
    raw = [np.zeros((mz + 4) * ny * nx) for _ in range(4)]
    # the 18 statements of one particle in source order (jl block then jr block; il, i, ir; kl, k, kr), particle
    # after particle: np.add.at applies the flattened updates in exactly that sequence
    node = np.stack([(kk[c] * ny + (jj + 1)) * nx + ii[a] for jj in (jl, jr) for a in range(3) for c in range(3)], axis=1)
    for m, qg in enumerate((qmult * vxj, qmult * vyj, qmult * vzj, np.full_like(vxj, qmult))):
        val = np.stack([qg * fx[a] * fy * fz[c] for fy in (fyl, fyr) for a in range(3) for c in range(3)], axis=1)
        np.add.at(raw[m], node.reshape(-1), val.reshape(-1))
    raw = [a.reshape(mz + 4, ny, nx) for a in raw]
    folded = [vmesh(p, a.copy()) for a in raw]
    return wkix, wkih, [a.reshape(-1) for a in raw], [a.reshape(-1) for a in folded]


def vmesh(p, a):
    """F:3243-3305 (vmesh3) = F:3327-3377 (vmesh1), one array: x and z by assignment, y by addition"""
    mx, my, mz = p.mx, p.my, p.mz
    for i in (-2, -1):
        a[:, :, mx + i + 2] = a[:, :, i + 2]
    for i in (mx, mx + 1):
        a[:, :, i - mx + 2] = a[:, :, i + 2]
    I = slice(2, mx + 2)
    a[:, 0 + 1, I] = a[:, 0 + 1, I] + a[:, -1 + 1, I]
    a[:, my + 1, I] = a[:, my + 1, I] + a[:, my + 1 + 1, I]
    J = slice(1, my + 2)
    for k in (-2, -1):
        a[mz + k + 2, J, I] = a[k + 2, J, I]
    for k in (mz, mz + 1):
        a[k - mz + 2, J, I] = a[k + 2, J, I]
    return a


def loadpt(p, ppc, vth, vdr, vbeam, ranfa=3021, ranfb=7331):
    """F:8885-8909 (table fv2) + F:8937-9040, one species; `ppc` particles per cell stand for the source's
    hard-coded 32 (F:8941).  Returns ([x, y, z, vx, vy, vz], ranfa_state, ranfb_state)."""
    import math
    mx, my, mz = p.mx, p.my, p.mz
    # inverse-CDF table of exp(-v^2) (v + vrg1), Simpson's rule, 100 intervals x 1000 sub-steps
    vrg1 = vdr / vth
    fun2 = lambda v: math.exp(-v ** 2) * (v + vrg1)
    vv = max(-3.0, -vrg1)
    dv = (3.0 - vv) / 100.0
    v2 = vv * vth
    dv2 = dv * vth
    fv2 = [0.0] * 102                                         # fv2(1..101), fv2(1) = 0
    for j in range(1, 101):
        s = 0.0
        sdv = dv / 1000.0
        for _ in range(500):
            vv = vv + 2.0 * sdv
            s = s + 4.0 * fun2(vv - sdv) + 2.0 * fun2(vv)
        s = (s + 4.0 * fun2(vv + sdv) + fun2(vv + 2.0 * sdv)) * sdv / 3.0
        fv2[j + 1] = fv2[j] + s
    top = fv2[101]
    for j in range(1, 102):
        fv2[j] = fv2[j] / top
    npr = mx * my * mz * ppc
    x, y, z = np.empty(npr), np.empty(npr), np.empty(npr)
    vx, vy, vz = np.empty(npr), np.empty(npr), np.empty(npr)
    sb = ranfb
    for l in range(npr):                                      # positions: three ranfp draws per particle
        sb, u = ranfp_next(sb); x[l] = p.xmax * u - p.hx / 2
        sb, u = ranfp_next(sb); y[l] = p.ymax * u
        sb, u = ranfp_next(sb); z[l] = p.zmax * u - p.hz / 2
    sa = ranfa
    zcent, dzcent, dzsmt = 0.50 * p.zmax, 0.125 * p.zmax, 0.15 * p.zmax
    ycent1, ycent2, dycent = 0.30 * p.ymax, 0.70 * p.ymax, 0.05 * p.ymax
    rrz, rry = 0.25 * p.zmax, 0.075 * p.ymax
    for l in range(npr):                                      # velocities: four ranf draws per particle (same LCG)
        sa, eps = ranfp_next(sa)
        k2 = 100
        for k in range(1, 101):
            k2 = k
            if fv2[k] > eps:
                break
        y1, y2 = fv2[k2 - 1], fv2[k2]
        x2 = (eps - y2) / (y2 - y1) + k2
        vmag = v2 + dv2 * (x2 - 1.0) + vdr
        sa, u = ranfp_next(sa); vxo = vmag * (u - 0.5)
        sa, u = ranfp_next(sa); vyo = vmag * (u - 0.5)
        sa, u = ranfp_next(sa); vzo = vmag * (u - 0.5)
        az = abs(z[l] - zcent)
        if az < dzcent or az < dzsmt:
            ycnt1, ycnt2 = ycent1 + dycent, ycent2 - dycent
        else:
            ycnt1, ycnt2 = ycent1, ycent2
        vdrift = 0.0
        if az <= rrz and (abs(y[l] - ycnt1) <= rry or abs(y[l] - ycnt2) <= rry):
            vdrift = vbeam
        vx[l], vy[l], vz[l] = vxo + vdrift, vyo, vzo
    return [x, y, z, vx, vy, vz], sa, sb
