"""CPU check of the slab-wise rank sum of the moments (option "compact",
csrc/mrg_api.cu: compact_layout / compact_planes_ok / compact_sum).

N simulated ranks deposit random values on exactly the planes the eligibility
rule allows (gather planes within 4 planes of the rank's block, wrapped across
the periodic seam; a deposit cell is within one plane of the gather cell and
its nodes within one plane of that, ghost planes not wrapped -- F:2273-2374),
the exchange of compact_sum is replayed in numpy with the strip positions the
library reports (two neighbour strips added into the own block, in-place
all-gather of the blocks, ghost planes broadcast by ranks 0 and N-1), and every
rank must end up with the plain sum over ranks, which is what the whole-grid
allreduce delivers.  No GPU, no product compute path."""
import ctypes as C

import numpy as np
import pytest


def layout(mz, n, r, occ=None):
    import mrg_b200
    lib = mrg_b200.capi.load()
    out = (C.c_int32 * 7)()
    o = None
    if occ is not None:
        o = np.ascontiguousarray(occ, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))
    assert lib.mrg_compact_layout(mz, n, r, o, out) == 0
    return list(out)


def deposit_planes(mz, gather_planes):
    """extended planes (k+2) a rank can deposit to, given the z planes of its gather cells"""
    out = set()
    for kp in gather_planes:
        cells = [mz] if kp == mz else [(kp + d) % mz for d in (-1, 0, 1)]     # predicted position, wrapped by partbc
        for c in cells:
            for d in (-1, 0, 1):
                out.add(c + d + 2)                                             # nodes are not wrapped: ghosts
    return sorted(out)


@pytest.mark.parametrize("mz,n", [(32, 2), (64, 4), (128, 8), (48, 3)])
def test_exchange_equals_the_sum_over_ranks(mz, n):
    rng = np.random.default_rng(mz + n)
    nz, L = mz + 4, mz // n
    assert layout(mz, n, 0)[0] == 1
    M = []
    for r in range(n):
        lo, hi = r * L, (r + 1) * L - 1
        gp = {(k % mz) for k in range(lo - 3, hi + 4)}                        # block +- 3, then the +-1 widening of add_occupancy
        gp |= {(k + d) % mz for k in list(gp) for d in (-1, 1)}
        occ = np.zeros(mz + 1, dtype=np.uint8)
        occ[sorted(gp)] = 1
        if r in (0, n - 1):
            occ[mz] = 1                                                        # the clamp plane next to the seam
        assert layout(mz, n, r, occ)[6] == 1, (r, sorted(gp))
        a = np.zeros((nz, 5))
        planes = deposit_planes(mz, [k for k in range(mz + 1) if occ[k]])
        a[planes] = rng.normal(size=(len(planes), 5))
        M.append(a)
    total = sum(M)
    H = layout(mz, n, 0)[5]
    lay = [layout(mz, n, r) for r in range(n)]
    # 1. strips to the ring neighbours, added into the own block
    rx = []
    for r in range(n):
        dn, up = (r - 1) % n, (r + 1) % n
        from_dn = M[dn][lay[dn][1]:lay[dn][1] + H].copy()                      # the lower neighbour's upward strip
        from_up = M[up][lay[up][2]:lay[up][2] + H].copy()                      # the upper neighbour's downward strip
        rx.append((from_dn, from_up))
    for r in range(n):
        M[r][lay[r][3]:lay[r][3] + H] += rx[r][0]
        M[r][lay[r][4]:lay[r][4] + H] += rx[r][1]
    # 2. in-place all-gather of the blocks, 3. ghost planes from their owners
    out = np.zeros((nz, 5))
    for r in range(n):
        out[2 + r * L:2 + (r + 1) * L] = M[r][2 + r * L:2 + (r + 1) * L]
    out[0:2] = M[0][0:2]
    out[mz + 2:mz + 4] = M[n - 1][mz + 2:mz + 4]
    np.testing.assert_allclose(out, total, rtol=0, atol=1e-12)


def test_eligibility_rejects_far_planes_and_odd_grids():
    mz, n = 64, 4
    occ = np.zeros(mz + 1, dtype=np.uint8)
    occ[16:32] = 1
    assert layout(mz, n, 1, occ)[6] == 1
    occ[36] = 1                                       # 5 planes above the block of rank 1: too far
    assert layout(mz, n, 1, occ)[6] == 0
    occ[36] = 0
    occ[mz] = 1                                       # an interior rank cannot own the clamp plane
    assert layout(mz, n, 1, occ)[6] == 0
    occ[:] = 0
    occ[[0, 1, 2, 61, 62, 63]] = 1                    # rank 0 with particles that crossed the seam
    assert layout(mz, n, 0, occ)[6] == 1
    assert layout(50, 4, 0)[0] == 0                   # mz not divisible by the ranks
    assert layout(32, 4, 0)[0] == 0                   # blocks thinner than two strips
