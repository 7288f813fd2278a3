// mrg_tile.cuh -- the fast particle passes.
//
//   k_predict_run   cell-run deposition with warp-level pre-reduction; fields
//                   gathered through L1 (works for ANY particle order).
//   k_predict_tile  same deposition, but a CTA owns a pencil of TILE_CELLS
//                   cells along x: the 6 stencil rows of the six prepared
//                   fields are staged in shared memory with 1-D bulk TMA
//                   (cp.async.bulk + mbarrier), the particles stream through a
//                   ring of 2-D tensor-TMA stages, cell-run totals leave with
//                   red.global.add.f64; also emits the next order's sort keys.
//   k_correct_tile  corrector with the same staging; scatters the updated
//                   particles straight into the next cell order (fused sort),
//                   records the z planes of the next gather, applies the drive
//                   kick inline under slab ownership.
//
// The tiled kernels need the cell index built by mrg_sort (cell_end[]); a
// particle whose stencil is not inside its CTA's tile takes the L1 / global
// atomic path, so results never depend on how well the order fits.
// F:n = /root/reference/@mrg37-080A.f03 line n.
#pragma once
#include <cuda.h>   // CUtensorMap (type only; the descriptors are encoded on the host in mrg_api.cu)
#include "mrg_kernels.cuh"

namespace mrg {

constexpr int TILE_CELLS = 32;
constexpr int TILE_NODES = TILE_CELLS + 2;
constexpr int TILE_ROW_D = TILE_NODES * 6;   // doubles per staged field row
constexpr int TILE_ACC_D = TILE_NODES * 4;   // doubles per accumulator row
constexpr unsigned FULL = 0xffffffffu;
// resident CTAs per SM the tiled kernels are compiled for (register budget)
#ifndef MRG_PRED_MINB
#define MRG_PRED_MINB 4
#endif
#ifndef MRG_PRED_SMEM_TILE
#define MRG_PRED_SMEM_TILE 0
#endif
#ifndef MRG_CORR_MINB
#define MRG_CORR_MINB 5
#endif

// ---------------------------------------------------------------------------
// Deposit target: global moment array M4[node][4], optionally fronted by a
// shared-memory accumulator tile sM[row = kz*2+jy][node][4] that covers the
// stencils of cells (i0 .. i0+ncell-1, j, k); key = stencil base node n0.
//
// Lane roles in the pre-reduction (phase B): a QUAD of lanes serves one
// particle; lane q = lane&3 owns the value rows g9 = 2q, 2q+1 (g9 = jy*4 + m).
// After the cross-quad reduction of flush_quad a lane holds three totals of
// row g9 = 2q + hi (hi = lane bit 4): r = r0, r0+1 (r0 = 4*bit3 + 2*bit2) and
// r = 8; the addresses of those three, relative to the cell, are lane
// constants kept in the target.
// ---------------------------------------------------------------------------
template <bool TILED>
struct Target {
  const GP& g;
  double* __restrict__ M4;
  double* sM;
  int n0_first, ncell;
  int so[3];      // shared-memory offsets (doubles) of the lane's three flush values, relative to sM + 4*d
  int go[3];      // global offsets (doubles) relative to M4 + 4*n0
  __device__ __forceinline__ Target(const GP& g_, double* M4_, double* sM_, int n0f, int nc, int lane)
      : g(g_), M4(M4_), sM(sM_), n0_first(n0f), ncell(nc) {
    const int g9 = 2 * (lane & 3) + ((lane >> 4) & 1);
    const int jy = g9 >> 2, m = g9 & 3;
    const int r0 = ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int r = (k == 2) ? 8 : r0 + k;
      const int kz = r / 3, ix = r - 3 * kz;
      so[k] = ((kz * 2 + jy) * TILE_NODES + ix) * 4 + m;
      go[k] = 4 * (ix + jy * g.nx + kz * g.nxy) + m;
    }
  }
  // general (g9, r) element of the cell whose stencil base node is `key`
  __device__ __forceinline__ void add(int key, int g9, int r, double v) const {
    if (TILED) {
      const unsigned d = (unsigned)(key - n0_first);
      if (d < (unsigned)ncell) {
        const int jy = g9 >> 2, m = g9 & 3, kz = r / 3, ix = r - 3 * kz;
        atomicAdd(sM + ((kz * 2 + jy) * TILE_NODES + (int)d + ix) * 4 + m, v);
        return;
      }
    }
    atomicAdd(mom_addr(M4, g, key, g9, r), v);
  }
  // the lane's k-th flush value of cell `key`
  __device__ __forceinline__ void add_flush(int key, int k, double v) const {
    if (TILED) {
      const unsigned d = (unsigned)(key - n0_first);
      if (d < (unsigned)ncell) {
        atomicAdd(sM + 4 * (int)d + so[k], v);
        return;
      }
    }
    atomicAdd(M4 + 4 * (size_t)key + go[k], v);
  }
};

#ifndef MRG_DEPOSIT_UNROLL
#define MRG_DEPOSIT_UNROLL 2
#endif
constexpr int DEPOSIT_UNROLL = MRG_DEPOSIT_UNROLL;   // unroll of the four sub-iterations of deposit_parked
constexpr int PR_WARPS = 4;        // warps per block
constexpr int PR_W_STRIDE = 10;    // doubles per particle in the W slab: wxz[9] + key
// Q slab: four rows (one per lane role q) of 32 double2.  Rows are 34 double2 apart: in the phase-B read the 8 lanes
// of a quarter-warp are 4 roles x 2 particles, and a 32-double2 (512-byte) row stride would put the four roles on
// the same banks (a 4-way conflict, 16 wavefronts per LDS.128 -- measured); 34 spreads them over all 32 banks.
constexpr int PR_Q_ROW = 34;       // double2 per row
constexpr int PR_Q_D = 4 * PR_Q_ROW * 2;   // doubles per warp

// Sum the quad-distributed accumulators over the warp with a transposing
// butterfly (18 -> 9 -> 4+1 -> 2+1 values per lane) and add the 72 totals of
// cell `n0` to the target.
template <bool TILED>
__device__ __forceinline__ void flush_quad(double* acc, int n0, const Target<TILED>& tg) {
  const int lane = threadIdx.x & 31;
  tr_round<18>(acc, (lane & 16) != 0, 16);     // acc[0..8]: row g9 = 2q + hi, r = 0..8
  double a8 = acc[8];
  a8 += __shfl_xor_sync(FULL, a8, 8);
  a8 += __shfl_xor_sync(FULL, a8, 4);
  tr_round<8>(acc, (lane & 8) != 0, 8);        // acc[0..3]: r = 4*bit3 + 0..3
  tr_round<4>(acc, (lane & 4) != 0, 4);        // acc[0..1]: r = 4*bit3 + 2*bit2 + 0..1
  tg.add_flush(n0, 0, acc[0]);
  tg.add_flush(n0, 1, acc[1]);
  if ((lane & 12) == 0) tg.add_flush(n0, 2, a8);
}

// phase A tail: park the 17 scatter factors + key of this lane's particle
__device__ __forceinline__ void park_factors(double* W, double* Q, int lane, const double qvy[8], const double wxz[9], int key) {
  double2* Wp = reinterpret_cast<double2*>(W + lane * PR_W_STRIDE);
  Wp[0] = make_double2(wxz[0], wxz[1]);
  Wp[1] = make_double2(wxz[2], wxz[3]);
  Wp[2] = make_double2(wxz[4], wxz[5]);
  Wp[3] = make_double2(wxz[6], wxz[7]);
  Wp[4] = make_double2(wxz[8], __longlong_as_double((long long)key));
  double2* Qp = reinterpret_cast<double2*>(Q);
#pragma unroll
  for (int qq = 0; qq < 4; qq++) Qp[qq * PR_Q_ROW + lane] = make_double2(qvy[2 * qq], qvy[2 * qq + 1]);
}

// phase B: four sub-iterations of 8 particles; a QUAD of lanes serves one
// particle, lane q owning value rows g9 = 2q, 2q+1.  Fast path: all eight
// particles continue the warp's current cell (key < 0 marks an empty slot
// whose factors are zero) -> 18 FMAs per lane.  Otherwise lanes are grouped by
// key; a group that continues the current cell, is large, or reaches the last
// lane is summed in registers across (sub-)iterations, other groups (strays)
// go straight to the target.
//
// own_key is the key of the lane's OWN particle (still in a register from
// phase A): one ballot against the current cell tells which sub-iterations
// may take the fast path, so that path reads no key and votes nothing.  Wq/Qq
// are the lane's read pointers into the parked factors (particle pl = lane/4
// of sub-iteration 0); they advance by a constant per sub-iteration.
template <bool TILED>
__device__ __forceinline__ void deposit_parked(const double2* Wq, const double2* Qq, int own_key, int lane, double* acc, int& cur,
                                               int group_min, const Target<TILED>& tg) {
  const int q = lane & 3;
  unsigned bad = __ballot_sync(FULL, own_key >= 0 && own_key != cur);   // bit p: particle p does not continue the current cell
#pragma unroll DEPOSIT_UNROLL
  for (int sub = 0; sub < 4; sub++, Wq += 8 * PR_W_STRIDE / 2, Qq += 8, bad >>= 8) {
    const double2 w01 = Wq[0], w23 = Wq[1], w45 = Wq[2], w67 = Wq[3], w8k = Wq[4];
    const double2 qv = Qq[0];
    const double wxz[9] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y, w8k.x};
    if ((bad & 0xffu) == 0u) {
#pragma unroll
      for (int r = 0; r < 9; r++) {
        acc[r] = fma(qv.x, wxz[r], acc[r]);
        acc[9 + r] = fma(qv.y, wxz[r], acc[9 + r]);
      }
      continue;
    }
    const int key = (int)__double_as_longlong(w8k.y);
    const bool valid = key >= 0;
    unsigned remaining = __ballot_sync(FULL, valid);
    while (remaining) {
      const int leader = __ffs(remaining) - 1;
      const int kk = __shfl_sync(FULL, key, leader);
      const unsigned grp = __ballot_sync(FULL, valid && key == kk) & remaining;
      const bool member = (grp >> lane) & 1u;
      const bool accumulate = (kk == cur) || (__popc(grp) >= group_min) || (grp >> 31);
      if (accumulate) {
        if (kk != cur) {
          if (cur >= 0) flush_quad<TILED>(acc, cur, tg);
#pragma unroll
          for (int n = 0; n < 18; n++) acc[n] = 0.0;
          cur = kk;
        }
        if (member) {
#pragma unroll
          for (int r = 0; r < 9; r++) {
            acc[r] = fma(qv.x, wxz[r], acc[r]);
            acc[9 + r] = fma(qv.y, wxz[r], acc[9 + r]);
          }
        }
      } else if (member) {
#pragma unroll
        for (int r = 0; r < 9; r++) {
          tg.add(key, 2 * q, r, qv.x * wxz[r]);
          tg.add(key, 2 * q + 1, r, qv.y * wxz[r]);
        }
      }
      remaining &= ~grp;
    }
    // the current cell may have changed: the remaining sub-iterations are judged against the new one
    bad = __ballot_sync(FULL, own_key >= 0 && own_key != cur) >> (8 * sub);
  }
}

// ---------------------------------------------------------------------------
// Predictor, any particle order (fields through L1): a warp owns 32*ITERS
// consecutive particles.  F:1162-1283, 1300-1306, 1375, srimp1+srimp2 scatter.
// ---------------------------------------------------------------------------
template <int ITERS>
__global__ void __launch_bounds__(PR_WARPS * 32)
k_predict_run(GP g, PushParams pp, ParticleSoA P, const double* __restrict__ F6, double* __restrict__ M4,
              double* __restrict__ wk_partial, int group_min) {
  __shared__ __align__(16) double smW[PR_WARPS][32 * PR_W_STRIDE];
  __shared__ __align__(16) double smQ[PR_WARPS][PR_Q_D];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* W = smW[w];
  double* Q = smQ[w];
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long base = warp * (32LL * ITERS);
  const Target<false> tg(g, M4, nullptr, 0, 0, lane);
  const double2* Wq = reinterpret_cast<const double2*>(W + (lane >> 2) * PR_W_STRIDE);
  const double2* Qq = reinterpret_cast<const double2*>(Q) + (lane & 3) * PR_Q_ROW + (lane >> 2);
  double wx = 0.0, wh = 0.0;
  double acc[18];
#pragma unroll
  for (int n = 0; n < 18; n++) acc[n] = 0.0;
  int cur = -1;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    const long long t = base + 32LL * it + lane;
    if (base + 32LL * it >= P.n) break;                     // warp-uniform
    int key = -1;
    {
      double qvy[8], wxz[9];
      if (t < P.n) {
        const Predicted o = predict_one(g, pp, P, t, F6, wx, wh);
        key = scatter_factors(g, pp.qmult, o, qvy, wxz);
      } else {
#pragma unroll
        for (int n = 0; n < 8; n++) qvy[n] = 0.0;
#pragma unroll
        for (int n = 0; n < 9; n++) wxz[n] = 0.0;
      }
      park_factors(W, Q, lane, qvy, wxz, key);
    }
    __syncwarp();
    deposit_parked<false>(Wq, Qq, key, lane, acc, cur, group_min, tg);
    __syncwarp();
  }
  if (cur >= 0) flush_quad<false>(acc, cur, tg);
  block_wk_store(wx, wh, wk_partial);
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D bulk TMA (global -> shared), sm_90+ PTX.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// same, with an L2 evict-first policy: the particle streams (GBs per pass) must not
// push the field / moment grids out of L2 (CUTLASS CacheHintSm90::EVICT_FIRST encoding)
__device__ __forceinline__ void bulk_g2s_stream(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned ok = 0;
#pragma unroll 1
  for (int spin = 0; spin < (1 << 26); spin++) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    if (ok) return;
  }
  __trap();   // a lost TMA completion must not hang the GPU (each try_wait suspends the warp for a hardware time slice, so 2^26 of them is minutes, not a profiler replay or a preemption)
}

// ---------------------------------------------------------------------------
// Tile bookkeeping.  Tile t covers cells (i0 .. i0+ncell-1, j, k) of the sort
// key space (i fastest, F:1175-1177 indices); its particles are the slots
// [p0,p1) of the cell index built by the last mrg_sort.
// ---------------------------------------------------------------------------
struct Tile {
  int i0, ncell, j, k;
  int p0, p1;
  int n0_first;     // stencil base node of cell (i0,j,k) = node (i0-1, j, k-1)
};
__device__ __forceinline__ Tile tile_of(const GP& g, const int* __restrict__ cell_end, int tile) {
  const int ntx = (g.mx + TILE_CELLS - 1) / TILE_CELLS;
  Tile t;
  const int tx = tile % ntx, r = tile / ntx;
  t.j = r % g.my;
  t.k = g.kz0 + r / g.my;                                     // the launch covers planes kz0 .. kz0+nkz-1 (mod mz)
  if (t.k >= g.mz) t.k -= g.mz;
  t.i0 = tx * TILE_CELLS;
  t.ncell = min(TILE_CELLS, g.mx - t.i0);
  const int c0 = t.i0 + g.mx * (t.j + g.my * t.k);
  t.p0 = (c0 == 0) ? 0 : cell_end[c0 - 1];
  t.p1 = cell_end[c0 + t.ncell - 1];
  t.n0_first = node_of(g, t.i0 - 1, t.j, t.k - 1);
  return t;
}

// stage the 6 stencil rows (jy = 0,1; kz = 0,1,2) of the packed fields: each
// row is contiguous in F6, (ncell+2)*48 bytes, 16-byte aligned
__device__ __forceinline__ void stage_fields(const GP& g, const Tile& t, const double* __restrict__ F6, double* sF,
                                             unsigned long long* bar) {
  if (threadIdx.x == 0) {
    const unsigned row_bytes = (unsigned)(t.ncell + 2) * 48u;
    mbar_expect_tx(bar, 6u * row_bytes);
#pragma unroll
    for (int kz = 0; kz < 3; kz++)
#pragma unroll
      for (int jy = 0; jy < 2; jy++) {
        const size_t node = (size_t)t.n0_first + (size_t)jy * g.nx + (size_t)kz * g.nxy;
        bulk_g2s(sF + (kz * 2 + jy) * TILE_ROW_D, F6 + node * 6, row_bytes, bar);
      }
  }
}

// gather of the six fields from the staged tile: same arithmetic as gather6,
// the 54 loads are warp-broadcast LDS.128 when the lanes share a cell
__device__ __forceinline__ void gather6_tile(const double* sF, int delta, const Stencil& s, double f[6]) {
#pragma unroll
  for (int c = 0; c < 6; c++) f[c] = 0.0;
  const double2* base = reinterpret_cast<const double2*>(sF + delta * 6);
#pragma unroll
  for (int kz = 0; kz < 3; kz++) {
#pragma unroll
    for (int jy = 0; jy < 2; jy++) {
      const double2* r = base + (kz * 2 + jy) * (TILE_ROW_D / 2);
      const double wyz = s.fy[jy] * s.fz[kz];
      double2 v[9];
#pragma unroll
      for (int q = 0; q < 9; q++) v[q] = r[q];
#pragma unroll
      for (int ix = 0; ix < 3; ix++) {
        const double w = s.fx[ix] * wyz;
        f[0] = fma(w, v[3 * ix + 0].x, f[0]);
        f[1] = fma(w, v[3 * ix + 0].y, f[1]);
        f[2] = fma(w, v[3 * ix + 1].x, f[2]);
        f[3] = fma(w, v[3 * ix + 1].y, f[3]);
        f[4] = fma(w, v[3 * ix + 2].x, f[4]);
        f[5] = fma(w, v[3 * ix + 2].y, f[5]);
      }
    }
  }
}

// half-step position + gather (tile or L1) + rotation for one particle
__device__ __forceinline__ Kick gather_rotate(const GP& g, const PushParams& pp, const Tile& t, const double* sF,
                                              const double* __restrict__ F6, double x, double y, double z, double vx,
                                              double vy, double vz) {
  double rx = __dadd_rn(x, __dmul_rn(pp.hdt, vx));          // F:1163-1165
  double ry = __dadd_rn(y, __dmul_rn(pp.hdt, vy));
  double rz = __dadd_rn(z, __dmul_rn(pp.hdt, vz));
  if (__any_sync(FULL, maybe_wrap(g, rx, ry, rz))) wrap_pos(g, rx, ry, rz);   // partbcEST, F:1168 (whole warp calls this)
  Stencil s;
  make_stencil<true>(g, rx, ry, rz, s);
  double f[6];
  const unsigned d = (unsigned)(s.n0 - t.n0_first);
  if (d < (unsigned)t.ncell) gather6_tile(sF, (int)d, s, f);
  else gather6(F6, g, s, f);
  return rotate(f, vx, vy, vz, pp.ht, pp.ht2);
}

// ---------------------------------------------------------------------------
// Particle streams.  The NW warps of a CTA split the tile's iterations (32
// particles each) evenly; every warp pulls its own slice through an NS-deep
// ring of shared-memory stages.  A stage is filled by ONE 2-D tensor TMA copy:
// the six SoA arrays of a particle set are rows of one allocation, so the box
// {32 slots} x {6 rows} of the [6][cap] fp64 tensor brings the 32 particles of
// an iteration (1536 B).  A box must start on a 16-byte boundary of the inner
// dimension (an unaligned coordinate is an illegal-instruction fault), so the
// tile's slot range is extended downwards to a multiple of 4 slots and the up
// to 3 extra leading lanes of its first iteration are masked like the trailing
// ones of its last.  The corrector adds two 1-D boxes of 32 int32 to the
// same stage and mbarrier: the slots' original indices and their next-order
// sort keys.  Issuing a stage costs one elected lane a dozen instructions;
// the earlier six 1-D bulk copies with their 64-bit address arithmetic were
// a quarter of the corrector's instruction stream.
// ---------------------------------------------------------------------------
#ifndef MRG_PRED_NSTAGE
#define MRG_PRED_NSTAGE 3
#endif
#ifndef MRG_CORR_NSTAGE
#define MRG_CORR_NSTAGE 4
#endif
constexpr int TSTAGE_P = 6 * 32 * 8;              // particle box of a stage
constexpr int TSTAGE_PIK = TSTAGE_P + 128 + 128;  // + ids + keys

// The stream keeps 32-bit shared-window addresses of its ring and barriers (computed once per warp): converting a
// generic pointer for every PTX operand costs an S2R + LEA pair each time.
__device__ __forceinline__ void mbar_expect_tx_s(unsigned bar_s, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(unsigned bar_s, unsigned phase) {
  unsigned ok = 0;
#pragma unroll 1
  for (int spin = 0; spin < (1 << 26); spin++) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar_s), "r"(phase)
        : "memory");
    if (ok) return;
  }
  __trap();   // a lost TMA completion must not hang the GPU
}
#ifndef MRG_STREAM_HINT
#define MRG_STREAM_HINT 0x12F0000000000000ull   // L2 evict-first (0x14F0... = evict-last, 0x10F0... = normal)
#endif
__device__ __forceinline__ void tma_box_2d_s(unsigned dst_s, const CUtensorMap* tm, int c0, int c1, unsigned bar_s) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst_s),
      "l"(tm), "r"(c0), "r"(c1), "r"(bar_s), "l"(MRG_STREAM_HINT)
      : "memory");
}
__device__ __forceinline__ void tma_box_1d_s(unsigned dst_s, const CUtensorMap* tm, int c0, unsigned bar_s) {
  asm volatile(
      "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2}], [%3], %4;" ::"r"(dst_s),
      "l"(tm), "r"(c0), "r"(bar_s), "l"(MRG_STREAM_HINT)
      : "memory");
}
__device__ __forceinline__ void tma_box_2d(void* dst, const CUtensorMap* tm, int c0, int c1, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull)   // evict-first, as bulk_g2s_stream
      : "memory");
}
__device__ __forceinline__ void tma_box_1d(void* dst, const CUtensorMap* tm, int c0, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2}], [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(smem_u32(bar)), "l"(0x12F0000000000000ull)
      : "memory");
}

struct Stream {
  int a, b;         // this warp's particle slots [a, b); a is a multiple of 4 slots ...
  int lo;           // ... and slots below lo (only in the tile's first iteration) belong to the previous tile
  int nit;          // iterations
  int issued;       // iterations whose copies have been issued
  unsigned char* ring;       // [NS][STAGE]
  unsigned ring_s, bar_s;    // the ring and its NS mbarriers as shared-window addresses
  // ring cursors of a ring whose depth is not a power of two (loop-carried; a power-of-two ring derives slot and
  // phase from the iteration number with a mask and a shift, which is cheaper than carrying them)
  int is_slot, rd_slot;
  unsigned rd_phase;
};
// the descriptors of a launch: particles always; ids / keys may be absent (nullptr)
struct StreamMaps { const CUtensorMap* p; const CUtensorMap* id; const CUtensorMap* key; };

template <int NS> struct RingPow2 { static constexpr bool value = (NS & (NS - 1)) == 0; };

template <int NS, int STAGE>
__device__ __forceinline__ void stream_issue(const StreamMaps& M, Stream& st, int lane) {
  if (st.issued < st.nit) {                                    // warp-uniform
    const int slot = RingPow2<NS>::value ? (st.issued & (NS - 1)) : st.is_slot;
    if (lane == 0) {
      const int e = st.a + 32 * st.issued;
      const unsigned dst = st.ring_s + slot * STAGE;
      const unsigned bar = st.bar_s + slot * 8;
      mbar_expect_tx_s(bar, (unsigned)TSTAGE_P + (M.id ? 128u : 0u) + (M.key ? 128u : 0u));
      tma_box_2d_s(dst, M.p, e, 0, bar);
      if (M.id) tma_box_1d_s(dst + TSTAGE_P, M.id, e, bar);
      if (M.key) tma_box_1d_s(dst + TSTAGE_P + 128, M.key, e, bar);
    }
    st.issued++;
    if (!RingPow2<NS>::value) st.is_slot = (slot + 1 == NS) ? 0 : slot + 1;
  }
}

// iterations [w*N/NW, (w+1)*N/NW) of the tile's N = ceil(len/32) go to warp w; starts the ring
template <int NS, int STAGE>
__device__ __forceinline__ void stream_open(const StreamMaps& M, const Tile& t, int w, int nwarps, int lane, unsigned char* ring,
                                            unsigned long long* bar, Stream& st) {
  const int base = t.p0 & ~3;                                  // box alignment: 4 slots = 16 bytes of int32, 32 of fp64
  const int N = (t.p1 - base + 31) >> 5;
  const int i0 = (w * N) / nwarps, i1 = ((w + 1) * N) / nwarps;
  st.a = base + 32 * i0;
  st.b = min(base + 32 * i1, t.p1);
  st.lo = t.p0;
  st.nit = i1 - i0;
  st.issued = 0;
  st.ring = ring;
  st.ring_s = smem_u32(ring);
  st.bar_s = smem_u32(bar);
  st.is_slot = 0; st.rd_slot = 0; st.rd_phase = 0u;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NS; s++) mbar_init(bar + s, 1);
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < NS - 1; s++) stream_issue<NS, STAGE>(M, st, lane);
}

// one particle's six phase-space coordinates
struct P6 { double x, y, z, vx, vy, vz; };

// ring slot / mbarrier phase of iteration `it` (it = the iteration under the read cursor, or the one after it)
template <int NS>
__device__ __forceinline__ int ring_slot(const Stream& st, int it, bool next) {
  if (RingPow2<NS>::value) return (it + (next ? 1 : 0)) & (NS - 1);
  return next ? ((st.rd_slot + 1 == NS) ? 0 : st.rd_slot + 1) : st.rd_slot;
}
template <int NS>
__device__ __forceinline__ unsigned ring_phase(const Stream& st, int it, bool next) {
  if (RingPow2<NS>::value) return (unsigned)(((it + (next ? 1 : 0)) / NS) & 1);
  return (next && st.rd_slot + 1 == NS) ? (st.rd_phase ^ 1u) : st.rd_phase;
}
template <int NS>
__device__ __forceinline__ void stream_wait(const Stream& st, int it) {
  mbar_wait_s(st.bar_s + ring_slot<NS>(st, it, false) * 8, ring_phase<NS>(st, it, false));
}
template <int NS>
__device__ __forceinline__ void stream_wait_next(const Stream& st, int it) {
  mbar_wait_s(st.bar_s + ring_slot<NS>(st, it, true) * 8, ring_phase<NS>(st, it, true));
}
// move the read cursor to the next iteration (a no-op for power-of-two rings)
template <int NS>
__device__ __forceinline__ void stream_advance(Stream& st) {
  if (!RingPow2<NS>::value) {
    if (st.rd_slot + 1 == NS) { st.rd_slot = 0; st.rd_phase ^= 1u; }
    else st.rd_slot++;
  }
}
// lane's particle of iteration `it` (whose stage has been waited for).  A lane outside [lo, b) -- the first / last
// iteration of a slice may be partial -- reads the nearest particle of the slice instead (a shared-memory broadcast):
// it then runs the same arithmetic on finite data and takes the same gather path as its neighbour; its results are masked.
template <int NS, int STAGE>
__device__ __forceinline__ void stream_read(const Stream& st, int it, int lane, P6& o) {
  const int first = st.a + 32 * it;
  const int le = min(max(first + lane, st.lo), st.b - 1) - first;
  const double* src = reinterpret_cast<const double*>(st.ring + ring_slot<NS>(st, it, false) * STAGE) + le;
  o.x = src[0]; o.y = src[32]; o.z = src[64];
  o.vx = src[96]; o.vy = src[128]; o.vz = src[160];
}
template <int NS, int STAGE>
__device__ __forceinline__ int stream_read_id(const Stream& st, int it, int lane) {
  return reinterpret_cast<const int*>(st.ring + ring_slot<NS>(st, it, false) * STAGE + TSTAGE_P)[lane];
}
template <int NS, int STAGE>
__device__ __forceinline__ int stream_read_key(const Stream& st, int it, int lane, bool next) {
  return reinterpret_cast<const int*>(st.ring + ring_slot<NS>(st, it, next) * STAGE + TSTAGE_P + 128)[lane];
}

// per-warp partial sums of wkix/wkih (F:1282-1283) -> wk_partial[2*(block*nwarps + w)]
__device__ __forceinline__ void warp_wk_store(double wx, double wh, double* __restrict__ partial, int nwarps) {
  wx = warp_sum(wx);
  wh = warp_sum(wh);
  if ((threadIdx.x & 31) == 0) {
    const size_t e = 2 * ((size_t)blockIdx.x * nwarps + (threadIdx.x >> 5));
    partial[e + 0] = wx;
    partial[e + 1] = wh;
  }
}

// Rank of the lane inside its run of EQUAL, CONTIGUOUS keys among the valid lanes (cell-sorted lanes carry
// mostly equal keys; equal keys in separate runs are simply claimed separately).  Returns the rank, the
// run length in `count` and whether the lane heads its run.  The valid lanes must be contiguous.
__device__ __forceinline__ int run_rank(int key, bool valid, int lane, int& count, bool& head) {
  const int prev = __shfl_up_sync(FULL, key, 1);
  const unsigned vmask = __ballot_sync(FULL, valid);
  head = valid && (lane == 0 || key != prev || !((vmask >> (lane - 1)) & 1u));
  const unsigned heads = __ballot_sync(FULL, head);
  const unsigned below = heads & ((2u << lane) - 1u);          // heads at or below this lane
  const int start = 31 - __clz(below | 1u);
  const unsigned above = heads & ~((2u << lane) - 1u);         // heads above this lane
  const int end = above ? (__ffs(above) - 1) : (32 - __clz(vmask));   // first lane of the next run / one past the last valid lane
  count = end - start;
  return lane - start;
}

// wkix/wkih of the warp -> two global accumulators (zeroed by the host before the launch)
__device__ __forceinline__ void warp_wk_atomic(double wx, double wh, double* __restrict__ acc2) {
  wx = warp_sum(wx);
  wh = warp_sum(wh);
  if ((threadIdx.x & 31) == 0 && (wx != 0.0 || wh != 0.0)) { atomicAdd(acc2, wx); atomicAdd(acc2 + 1, wh); }
}

// ---------------------------------------------------------------------------
// Predictor on TMA-staged tiles.
// ---------------------------------------------------------------------------
constexpr int PNS = MRG_PRED_NSTAGE, CNS = MRG_CORR_NSTAGE;
constexpr int ZOCC_VIOLATION = 0x7ffffff0;
constexpr int PRED_ACC_D = MRG_PRED_SMEM_TILE ? 6 * TILE_ACC_D : 0;      // the accumulator tile exists only when it is used
constexpr int PRED_RING_BYTES = PR_WARPS * PNS * TSTAGE_P;               // first in the carve-up: tensor TMA wants 128-byte aligned boxes
constexpr int PRED_SMEM_BYTES = PRED_RING_BYTES + (6 * TILE_ROW_D + PRED_ACC_D + PR_WARPS * (32 * PR_W_STRIDE + PR_Q_D)) * 8 +
                                (PR_WARPS * PNS + 1) * 8;
__global__ void __launch_bounds__(PR_WARPS * 32, MRG_PRED_MINB)
k_predict_tile(GP g, PushParams pp, const __grid_constant__ CUtensorMap tmP, const double* __restrict__ F6, double* __restrict__ M4,
               const int* __restrict__ cell_end, double* __restrict__ wk_partial, int group_min,
               int* __restrict__ prekey, int* __restrict__ prehist) {
  // dynamic shared memory (PRED_SMEM_BYTES > 48 KB static limit), carved up by hand
  extern __shared__ __align__(1024) unsigned char smem_pred[];
  unsigned char* sRing = smem_pred;                            // [warps][PNS][TSTAGE_P]  particle stages
  double* sF = reinterpret_cast<double*>(smem_pred + PRED_RING_BYTES);   // [6][TILE_ROW_D]  staged fields
  double* sM = sF + 6 * TILE_ROW_D;                            // [6][TILE_ACC_D]      moment accumulators (optional)
  double* smW = sM + PRED_ACC_D;                               // [warps][32*PR_W_STRIDE]
  double* smQ = smW + PR_WARPS * 32 * PR_W_STRIDE;             // [warps][PR_Q_D]
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(smQ + PR_WARPS * PR_Q_D);   // [warps][PNS] + 1
  unsigned long long& bar = sBar[PR_WARPS * PNS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Tile t = tile_of(g, cell_end, blockIdx.x);
  const bool busy = t.p1 > t.p0;                              // block-uniform
  double wx = 0.0, wh = 0.0;
  if (busy) {
    Stream st;
    const StreamMaps maps{&tmP, nullptr, nullptr};
    stream_open<PNS, TSTAGE_P>(maps, t, w, PR_WARPS, lane, sRing + w * (PNS * TSTAGE_P), sBar + w * PNS, st);   // particles in flight during the field staging
    if (threadIdx.x == 0) mbar_init(&bar, 1);
#if MRG_PRED_SMEM_TILE
    for (int e = threadIdx.x; e < 6 * TILE_ACC_D; e += blockDim.x) sM[e] = 0.0;
#endif
    __syncthreads();
    stage_fields(g, t, F6, sF, &bar);
    mbar_wait(&bar, 0);
    double* W = smW + w * (32 * PR_W_STRIDE);
    double* Q = smQ + w * PR_Q_D;
    const double2* Wq = reinterpret_cast<const double2*>(W + (lane >> 2) * PR_W_STRIDE);
    const double2* Qq = reinterpret_cast<const double2*>(Q) + (lane & 3) * PR_Q_ROW + (lane >> 2);
    // MRG_PRED_SMEM_TILE = 0: cell-run totals go straight to global memory with red.global.add.f64 (fire and
    // forget; shared-memory fp64 atomics are compare-and-swap loops on sm_100a)
    const Target<(MRG_PRED_SMEM_TILE != 0)> tg(g, M4, sM, t.n0_first, t.ncell, lane);
    double acc[18];
#pragma unroll
    for (int n = 0; n < 18; n++) acc[n] = 0.0;
    int cur = -1;
    const double ah = pp.aimpl * pp.hh, hh2 = 0.5 * pp.hh;
#pragma unroll 1
    for (int it = 0; it < st.nit; it++) {
      P6 c;
      stream_issue<PNS, TSTAGE_P>(maps, st, lane);           // refill the slot consumed in iteration it-1
      stream_wait<PNS>(st, it);
      stream_read<PNS, TSTAGE_P>(st, it, lane, c);
      const int p = st.a + 32 * it + lane;
      const bool valid = p >= st.lo && p < st.b;
      stream_advance<PNS>(st);
      int key = -1;
      {
        double qvy[8], wxz[9];
        const Kick k = gather_rotate(g, pp, t, sF, F6, c.x, c.y, c.z, c.vx, c.vy, c.vz);
        Predicted o;
        o.vxj = fma(ah, k.dvx, c.vx);                         // F:1300-1302
        o.vyj = fma(ah, k.dvy, c.vy);
        o.vzj = fma(ah, k.dvz, c.vz);
        o.rx = fma(pp.adt, fma(hh2, k.dvx, c.vx), c.x);       // F:1304-1306
        o.ry = fma(pp.adt, fma(hh2, k.dvy, c.vy), c.y);
        o.rz = fma(pp.adt, fma(hh2, k.dvz, c.vz), c.z);
        if (__any_sync(FULL, maybe_wrap(g, o.rx, o.ry, o.rz))) {
          if (wrap_pos(g, o.rx, o.ry, o.rz)) o.vyj = -o.vyj;  // partbc, F:1375
        }
        key = scatter_factors(g, valid ? pp.qmult : 0.0, o, qvy, wxz);
        if (valid) { wx += k.wx; wh += k.wh; } else key = -1;
        park_factors(W, Q, lane, qvy, wxz, key);
        if (prekey) {
          // Order of the NEXT step, decided one pass early: cell of x' + hdt*v' with x', v' from this
          // pass' dv (the corrector's dv differs only through the field update, so nearly every key is
          // exact; a key is a sorting hint and never changes a result).  The corrector scatters by it.
          const double xn = fma(pp.dt, fma(hh2, k.dvx, c.vx), c.x), un = fma(pp.hh, k.dvx, c.vx);
          const double yn = fma(pp.dt, fma(hh2, k.dvy, c.vy), c.y), vn = fma(pp.hh, k.dvy, c.vy);
          const double zn = fma(pp.dt, fma(hh2, k.dvz, c.vz), c.z), wn = fma(pp.hh, k.dvz, c.vz);
          const int kcell = sort_cell_folded(g, fma(pp.hdt, un, xn), fma(pp.hdt, vn, yn), fma(pp.hdt, wn, zn));
          int cnt;
          bool head;
          run_rank(kcell, valid, lane, cnt, head);
          if (valid) prekey[p] = kcell;
          if (head) atomicAdd(prehist + kcell, cnt);
        }
      }
      __syncwarp();
      deposit_parked<(MRG_PRED_SMEM_TILE != 0)>(Wq, Qq, key, lane, acc, cur, group_min, tg);
      __syncwarp();
    }
    if (cur >= 0) flush_quad<(MRG_PRED_SMEM_TILE != 0)>(acc, cur, tg);
#if MRG_PRED_SMEM_TILE
    __syncthreads();
    // flush the accumulator tile: 4 moments of a node = one 32-byte sector
    const int nodes = t.ncell + 2;
    for (int e = threadIdx.x; e < 6 * nodes * 4; e += blockDim.x) {
      const int row = e / (nodes * 4), rem = e - row * (nodes * 4);
      const double v = sM[row * TILE_ACC_D + rem];
      if (v != 0.0) {
        const int kz = row >> 1, jy = row & 1;
        atomicAdd(M4 + 4 * ((size_t)t.n0_first + (size_t)jy * g.nx + (size_t)kz * g.nxy) + rem, v);
      }
    }
#endif
  }
  warp_wk_atomic(wx, wh, wk_partial);
}

// ---------------------------------------------------------------------------
// Corrector on TMA-staged tiles: F:1162-1295, partbc F:1337, slab test of the
// drive kick F:1343-1345.  With key_out != nullptr it also writes the cell of
// wrap(x' + lookahead*v') (next step's sort key) and the cell histogram, which
// lets mrg_sort skip its key pass.  The key is a sorting hint only (any value
// gives the same results), so it is formed with contracted arithmetic.
//
// Fused cell sort (prekey): the keys of the next order arrive in the ring with
// the particles.  The stage of iteration it+1 is waited for during iteration
// it (it was issued CNS-1 = 3 iterations earlier), its keys are ranked and the
// slots of its cell runs claimed with one atomic per run, so neither the key
// nor the claim (an atomic with a return value) is waited for when iteration
// it+1 needs them.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PR_WARPS * 32, MRG_CORR_MINB)
k_correct_tile(GP g, PushParams pp, const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmId,
               const __grid_constant__ CUtensorMap tmKey, int have_id, const double* __restrict__ F6,
               const int* __restrict__ cell_end, double* __restrict__ wk_partial, unsigned* __restrict__ slab_bits,
               int* __restrict__ slab_list, int* __restrict__ slab_count, int* __restrict__ key_out, int* __restrict__ hist,
               double lookahead, int scatter, int* __restrict__ cursor, SortArrays D, unsigned* __restrict__ zocc,
               const unsigned* __restrict__ lcg_tab) {
  __shared__ __align__(128) unsigned char sRing[PR_WARPS][CNS * TSTAGE_PIK];
  __shared__ __align__(128) double sF[6 * TILE_ROW_D];
  __shared__ __align__(8) unsigned long long sBar[PR_WARPS][CNS];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Tile t = tile_of(g, cell_end, blockIdx.x);
  const bool busy = t.p1 > t.p0;
  double wx = 0.0, wh = 0.0;
  if (busy) {
    Stream st;
    const StreamMaps maps{&tmP, have_id ? &tmId : nullptr, scatter ? &tmKey : nullptr};
    stream_open<CNS, TSTAGE_PIK>(maps, t, w, PR_WARPS, lane, sRing[w], sBar[w], st);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    stage_fields(g, t, F6, sF, &bar);
    mbar_wait(&bar, 0);
    const double hh2 = 0.5 * pp.hh;
    int cb = 0, crk = 0;         // claim of the current iteration: base (in the run's head lane), rank in the run
    int rlo = 0x7fffffff, rhi = -0x7fffffff;   // range of next-pass gather planes seen by this lane, relative to t.k
    if (st.nit > 0) stream_wait<CNS>(st, 0);                    // warp-uniform; a warp of a thin tile may have no iteration
    if (scatter && st.nit > 0) {
      const int pk0 = stream_read_key<CNS, TSTAGE_PIK>(st, 0, lane, false);
      int cnt;
      bool head;
      crk = run_rank(pk0, st.a + lane >= st.lo && st.a + lane < st.b, lane, cnt, head);
      if (head) cb = atomicAdd(cursor + pk0, cnt);
    }
#pragma unroll 1
    for (int it = 0; it < st.nit; it++) {
      P6 c;
      stream_issue<CNS, TSTAGE_PIK>(maps, st, lane);         // refill the slot consumed in iteration it-1
      stream_read<CNS, TSTAGE_PIK>(st, it, lane, c);         // stage `it` was waited for one iteration ago
      const int p = st.a + 32 * it + lane;
      const bool valid = p >= st.lo && p < st.b;
      int kcell = -1;
      const int idv = have_id ? stream_read_id<CNS, TSTAGE_PIK>(st, it, lane) : p;
      int nb = 0, nrk = 0;
      if (it + 1 < st.nit) {                                  // warp-uniform
        stream_wait_next<CNS>(st, it);
        if (scatter) {                                        // claim the slots of iteration it + 1
          const int pk1 = stream_read_key<CNS, TSTAGE_PIK>(st, it, lane, true);
          int cnt;
          bool head;
          nrk = run_rank(pk1, p + 32 < st.b, lane, cnt, head);
          if (head) nb = atomicAdd(cursor + pk1, cnt);
        }
      }
      const Kick k = gather_rotate(g, pp, t, sF, F6, c.x, c.y, c.z, c.vx, c.vy, c.vz);
      double x = fma(pp.dt, fma(hh2, k.dvx, c.vx), c.x);      // F:1289-1291
      double y = fma(pp.dt, fma(hh2, k.dvy, c.vy), c.y);
      double z = fma(pp.dt, fma(hh2, k.dvz, c.vz), c.z);
      const double vx = fma(pp.hh, k.dvx, c.vx);              // F:1293-1295
      double vy = fma(pp.hh, k.dvy, c.vy);
      const double vz = fma(pp.hh, k.dvz, c.vz);
      if (__any_sync(FULL, maybe_wrap(g, x, y, z))) {
        if (wrap_pos(g, x, y, z)) vy = -vy;                   // partbc, F:1337
      }
      if (zocc) {
        // z plane of the next pass' gather cell, relative to this tile's plane, good to +-1 (contracted arithmetic,
        // folded in index space); ensure_prep widens the recorded planes by one
        int kq = gather_plane_fast(g, z, vz, lookahead) - t.k;
        const int hmz = g.mz >> 1;
        kq = kq > hmz ? kq - g.mz : (kq < -hmz ? kq + g.mz : kq);
        // precondition of the +-1 plane bound and of the slab-wise exchange: |vz| dt < hz.  A violation rides in rhi as a
        // sentinel and ends up in the flag word behind the bitmap; the host then ignores the record (full preparation,
        // whole-grid allreduce)
        if (valid) { rlo = min(rlo, kq); rhi = max(rhi, (fabs(vz) * pp.dt >= g.hz) ? ZOCC_VIOLATION : kq); }
      }
      if (pp.drive_on && pp.kick_inline) {                    // E x B drive kick with a per-particle draw, F:1343-1364
        if (valid && (fabs(z - pp.zcent) < pp.zw) && ((fabs(y - pp.ycent2) < pp.yw) || (fabs(y - pp.ycent1) < pp.yw))) {
          const unsigned ir = (lcg_pow_tab(lcg_tab, (unsigned long long)(unsigned)idv + 1ull) * pp.kick_state) & 0x7fffffffu;
          if ((double)ir * (1.0 / 2147483648.0) > 0.999) {    // F:1353
            int ip, jp, kp;
            cell_of(g, x, y, z, ip, jp, kp);                  // F:1347-1349
            const double vy0 = __ddiv_rn(pp.Ez00, F6[(size_t)node_of(g, ip, jp, kp) * 6 + 3]);   // F:1354
            if (fabs(y - pp.ycent2) < pp.yw2) vy = __dsub_rn(vy, vy0);
            else if (fabs(y - pp.ycent1) < pp.yw2) vy = __dadd_rn(vy, vy0);
          }
        }
      }
      int d = p;                                              // slot the updated particle is written to
      if (scatter) {
        d = __shfl_sync(FULL, cb, lane - crk) + crk;          // claimed one iteration ago
        cb = nb; crk = nrk;
      }
      if (valid) {
        wx += k.wx; wh += k.wh;
        const int id = idv;
        if (scatter) {
          __stcs(D.dst[0] + d, x); __stcs(D.dst[1] + d, y); __stcs(D.dst[2] + d, z);
          __stcs(D.dst[3] + d, vx); __stcs(D.dst[4] + d, vy); __stcs(D.dst[5] + d, vz);
          D.id_dst[d] = id;
        } else {
          __stcs(D.src_rw[0] + p, x); __stcs(D.src_rw[1] + p, y); __stcs(D.src_rw[2] + p, z);
          __stcs(D.src_rw[3] + p, vx); __stcs(D.src_rw[4] + p, vy); __stcs(D.src_rw[5] + p, vz);
        }
        if (key_out) {
          kcell = sort_cell_folded(g, fma(lookahead, vx, x), fma(lookahead, vy, y), fma(lookahead, vz, z));
          key_out[p] = kcell;
        }
      }
      if (pp.drive_on && !pp.kick_inline) {                   // slab of the E x B drive kick, F:1343-1345 (kicked by k_kick)
        const bool in_slab = valid && (fabs(z - pp.zcent) < pp.zw) &&
                             ((fabs(y - pp.ycent2) < pp.yw) || (fabs(y - pp.ycent1) < pp.yw));
        const unsigned m = __ballot_sync(FULL, in_slab);
        if (m) {                                              // one counter bump per warp: inside the slab every lane qualifies
          int base = 0;
          if (lane == __ffs(m) - 1) base = atomicAdd(slab_count, __popc(m));
          base = __shfl_sync(FULL, base, __ffs(m) - 1);
          if (in_slab) {
            atomicOr(slab_bits + (idv >> 5), 1u << (idv & 31));
            slab_list[base + __popc(m & ((1u << lane) - 1u))] = d;
          }
        }
      }
      if (key_out) {
        const unsigned act = __ballot_sync(FULL, valid);
        if (valid) {
          const unsigned m = __match_any_sync(act, kcell);
          if (lane == __ffs(m) - 1) atomicAdd(hist + kcell, __popc(m));
        }
      }
      stream_advance<CNS>(st);
      __syncwarp();
    }
    if (zocc) {                                               // one range per warp; bits already set cost a load only
      rlo = __reduce_min_sync(FULL, rlo);
      rhi = __reduce_max_sync(FULL, rhi);
      if (rhi == ZOCC_VIOLATION) {
        if (lane == 0) atomicOr(zocc + ((g.mz + 1 + 31) >> 5), 1u);
        rhi = g.mz;
      }
      if (lane == 0 && rlo <= rhi) {
        for (int r = max(rlo, -g.mz); r <= min(rhi, g.mz); r++) {
          int k = t.k + r;
          k = k < 0 ? k + g.mz : (k >= g.mz ? k - g.mz : k);
          k = min(max(k, 0), g.mz - 1);
          if (!((__ldcg(zocc + (k >> 5)) >> (k & 31)) & 1u)) atomicOr(zocc + (k >> 5), 1u << (k & 31));
        }
      }
    }
  }
  warp_wk_atomic(wx, wh, wk_partial);
}

}  // namespace mrg
