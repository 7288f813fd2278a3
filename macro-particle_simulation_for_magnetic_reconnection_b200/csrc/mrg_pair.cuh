// mrg_pair.cuh -- particle passes that push TWO particles per thread.
//
//   k_predict_pair   ipc >= 1: gather, implicit rotation, predicted position /
//                    velocity, partbc, fused srimp1+srimp2 deposition
//   k_correct_pair   ipc == 0: gather, implicit rotation, in-place update,
//                    partbc, drive-slab test, next sort key
//
// Why pairs: both passes are bound by three SM resources at once -- fp64 issue
// (64 lanes/clk/SM), the shared-memory data pipe (the 54 LDS.128 of the 18-node
// x 6-field gather) and plain instruction issue.  Cell-sorted neighbours
// (slots 2m, 2m+1) are in the same cell 63 times out of 64, so one thread
// gathers once for both: the field rows are loaded for particle A, used for A,
// re-loaded (predicated, only in lanes whose B sits in another cell) and used
// for B.  That halves the LDS traffic and the per-particle bookkeeping, and the
// two independent dependency chains hide the fp64 latency.
//
// Data movement: a CTA owns a pencil of TILE_CELLS cells along x (as in
// mrg_tile.cuh); the six stencil rows of the packed fields arrive by 1-D bulk
// TMA; every warp streams its contiguous slice of the tile's particles through
// a ring of 64-particle shared-memory stages filled by bulk TMA (one copy per
// SoA array, issued by lane 0 with an L2 evict-first hint).  Slices are walked
// from the even slot below their start, so a lane's pair (2m, 2m+1) is one
// aligned 128-bit shared load per array and one 128-bit coalesced global store
// per array in the corrector.
// F:n = /root/reference/@mrg37-080A.f03 line n.
#pragma once
#include "mrg_tile.cuh"

// resident CTAs per SM the pair kernels are compiled for (register budget)
#ifndef MRG_PPAIR_MINB
#define MRG_PPAIR_MINB 2
#endif
#ifndef MRG_CPAIR_MINB
#define MRG_CPAIR_MINB 3
#endif

namespace mrg {

// ---------------------------------------------------------------------------
// Particle stream: 64-particle stages.
// ---------------------------------------------------------------------------
constexpr int PSTAGE = 64;                       // particles (doubles per array) per stage
constexpr int PSTAGE_BYTES = PSTAGE * 8;
constexpr int PRING_D = 6 * PSTAGE;              // doubles per stage

struct PairStream {
  int a, b;       // this warp's particle slots [a, b)
  int a_al;       // a rounded down to even
  int nit;        // iterations of 64 slots starting at a_al
  int issued;
  double* ring;                // [NST][6][PSTAGE]
  unsigned long long* bar;     // [NST]
};

template <int NST>
__device__ __forceinline__ void pstream_issue(const ParticleSoA& P, PairStream& st, int lane) {
  if (st.issued < st.nit) {                      // warp-uniform
    if (lane == 0) {
      const int slot = st.issued % NST;
      const size_t e = (size_t)(st.a_al + PSTAGE * st.issued);
      double* dst = st.ring + slot * PRING_D;
      unsigned long long* bar = st.bar + slot;
      mbar_expect_tx(bar, 6u * PSTAGE_BYTES);
      bulk_g2s_stream(dst + 0 * PSTAGE, P.x + e, PSTAGE_BYTES, bar);
      bulk_g2s_stream(dst + 1 * PSTAGE, P.y + e, PSTAGE_BYTES, bar);
      bulk_g2s_stream(dst + 2 * PSTAGE, P.z + e, PSTAGE_BYTES, bar);
      bulk_g2s_stream(dst + 3 * PSTAGE, P.vx + e, PSTAGE_BYTES, bar);
      bulk_g2s_stream(dst + 4 * PSTAGE, P.vy + e, PSTAGE_BYTES, bar);
      bulk_g2s_stream(dst + 5 * PSTAGE, P.vz + e, PSTAGE_BYTES, bar);
    }
    st.issued++;
  }
}

// The tile's ceil(len/64) iterations are split evenly over the warps.  All
// stream scalars are broadcast from lane 0 so that the compiler can keep them
// (and the TMA operands derived from them) in uniform registers.
template <int NST>
__device__ __forceinline__ void pstream_open(const ParticleSoA& P, const Tile& t, int w, int nwarps, int lane, double* ring,
                                             unsigned long long* bar, PairStream& st) {
  const int p0 = __shfl_sync(FULL, t.p0, 0), p1 = __shfl_sync(FULL, t.p1, 0);
  const int base = p0 & ~1;
  const int N = (p1 - base + PSTAGE - 1) / PSTAGE;
  const int i0 = (w * N) / nwarps, i1 = ((w + 1) * N) / nwarps;
  st.a = max(base + PSTAGE * i0, p0);
  st.b = min(base + PSTAGE * i1, p1);
  st.a_al = base + PSTAGE * i0;
  st.nit = i1 - i0;
  st.issued = 0;
  st.ring = ring;
  st.bar = bar;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NST; s++) mbar_init(bar + s, 1);
  }
  __syncwarp();
#pragma unroll
  for (int s = 0; s < NST - 1; s++) pstream_issue<NST>(P, st, lane);
}

// Wait for iteration `it`, hand the lane its pair A = slot g0, B = slot g0+1
// (g0 = a_al + 64*it + 2*lane) and keep the ring full.  Slots outside [a,b)
// are replaced by a copy of a valid particle of the same warp so that every
// lane follows the common path; their results are masked by okA / okB.
// Call with the whole warp, after the __syncwarp() that ended the previous
// iteration's reads of the ring.
template <int NST>
__device__ __forceinline__ void pstream_next(const ParticleSoA& P, PairStream& st, int it, int lane, P6& A, P6& B, int& g0,
                                             bool& okA, bool& okB) {
  pstream_issue<NST>(P, st, lane);               // refills the stage consumed in iteration it-1
  const int slot = it % NST;
  mbar_wait(st.bar + slot, (unsigned)((it / NST) & 1));
  const double2* src = reinterpret_cast<const double2*>(st.ring + slot * PRING_D) + lane;
  const double2 X = src[0 * (PSTAGE / 2)], Y = src[1 * (PSTAGE / 2)], Z = src[2 * (PSTAGE / 2)];
  const double2 U = src[3 * (PSTAGE / 2)], V = src[4 * (PSTAGE / 2)], Wv = src[5 * (PSTAGE / 2)];
  A.x = X.x; A.y = Y.x; A.z = Z.x; A.vx = U.x; A.vy = V.x; A.vz = Wv.x;
  B.x = X.y; B.y = Y.y; B.z = Z.y; B.vx = U.y; B.vy = V.y; B.vz = Wv.y;
  g0 = st.a_al + PSTAGE * it + 2 * lane;
  okA = (g0 >= st.a) && (g0 < st.b);
  okB = (g0 + 1 < st.b);                          // g0 + 1 >= a always holds
  if (__any_sync(FULL, !(okA && okB))) {          // first / last iteration of a slice only
    // lane 0 always holds at least one valid particle of this iteration
    const bool s0 = __shfl_sync(FULL, (int)okA, 0) != 0;
    P6 S;
    S.x = __shfl_sync(FULL, s0 ? A.x : B.x, 0); S.y = __shfl_sync(FULL, s0 ? A.y : B.y, 0);
    S.z = __shfl_sync(FULL, s0 ? A.z : B.z, 0); S.vx = __shfl_sync(FULL, s0 ? A.vx : B.vx, 0);
    S.vy = __shfl_sync(FULL, s0 ? A.vy : B.vy, 0); S.vz = __shfl_sync(FULL, s0 ? A.vz : B.vz, 0);
    const P6 A0 = A;
    if (!okA) A = okB ? B : S;
    if (!okB) B = okA ? A0 : S;
  }
}

// gather weights of one particle (subset of Stencil, F:1175-1215)
struct GW {
  int d;              // stencil base node relative to the tile's first base node (valid when in-tile)
  int n0;             // absolute stencil base node
  double fx[3], fy[2], fz[3];
};
__device__ __forceinline__ void gather_weights(const GP& g, double rx, double ry, double rz, GW& o) {
  Stencil s;
  make_stencil<true>(g, rx, ry, rz, s);
  o.n0 = s.n0;
#pragma unroll
  for (int k = 0; k < 3; k++) { o.fx[k] = s.fx[k]; o.fz[k] = s.fz[k]; }
  o.fy[0] = s.fy[0]; o.fy[1] = s.fy[1];
}

// Gather + rotation for the pair.  F:1162-1283.
// Returns true in lanes that had to gather outside the tile.
__device__ __forceinline__ bool push_pair(const GP& g, const PushParams& pp, const Tile& t, const double* sF,
                                          const double* __restrict__ F6, const P6& A, const P6& B, Kick& kA, Kick& kB) {
  double ax = __dadd_rn(A.x, __dmul_rn(pp.hdt, A.vx));          // F:1163-1165
  double ay = __dadd_rn(A.y, __dmul_rn(pp.hdt, A.vy));
  double az = __dadd_rn(A.z, __dmul_rn(pp.hdt, A.vz));
  double bx = __dadd_rn(B.x, __dmul_rn(pp.hdt, B.vx));
  double by = __dadd_rn(B.y, __dmul_rn(pp.hdt, B.vy));
  double bz = __dadd_rn(B.z, __dmul_rn(pp.hdt, B.vz));
  if (__any_sync(FULL, maybe_wrap(g, ax, ay, az) | maybe_wrap(g, bx, by, bz))) {   // partbcEST, F:1168
    wrap_pos(g, ax, ay, az);
    wrap_pos(g, bx, by, bz);
  }
  GW wa, wb;
  gather_weights(g, ax, ay, az, wa);
  gather_weights(g, bx, by, bz, wb);
  const unsigned dA = (unsigned)(wa.n0 - t.n0_first), dB = (unsigned)(wb.n0 - t.n0_first);
  const bool outA = dA >= (unsigned)t.ncell, outB = dB >= (unsigned)t.ncell;
  const int da = outA ? 0 : (int)dA, db = outB ? 0 : (int)dB;
  const bool reload = (db != da);
  const bool any_reload = __any_sync(FULL, reload);
  double fA[6], fB[6];
#pragma unroll
  for (int c = 0; c < 6; c++) { fA[c] = 0.0; fB[c] = 0.0; }
  const double2* bA = reinterpret_cast<const double2*>(sF + da * 6);
  const double2* bB = reinterpret_cast<const double2*>(sF + db * 6);
#pragma unroll
  for (int kz = 0; kz < 3; kz++) {
#pragma unroll
    for (int jy = 0; jy < 2; jy++) {
      const int ro = (kz * 2 + jy) * (TILE_ROW_D / 2);
      double2 v[9];
#pragma unroll
      for (int q = 0; q < 9; q++) v[q] = bA[ro + q];
      const double wyzA = wa.fy[jy] * wa.fz[kz];
#pragma unroll
      for (int ix = 0; ix < 3; ix++) {
        const double w = wa.fx[ix] * wyzA;
        fA[0] = fma(w, v[3 * ix + 0].x, fA[0]);
        fA[1] = fma(w, v[3 * ix + 0].y, fA[1]);
        fA[2] = fma(w, v[3 * ix + 1].x, fA[2]);
        fA[3] = fma(w, v[3 * ix + 1].y, fA[3]);
        fA[4] = fma(w, v[3 * ix + 2].x, fA[4]);
        fA[5] = fma(w, v[3 * ix + 2].y, fA[5]);
      }
      if (any_reload) {                          // warp-uniform; ~40 % of the iterations, one or two lanes
        if (reload) {
#pragma unroll
          for (int q = 0; q < 9; q++) v[q] = bB[ro + q];
        }
      }
      const double wyzB = wb.fy[jy] * wb.fz[kz];
#pragma unroll
      for (int ix = 0; ix < 3; ix++) {
        const double w = wb.fx[ix] * wyzB;
        fB[0] = fma(w, v[3 * ix + 0].x, fB[0]);
        fB[1] = fma(w, v[3 * ix + 0].y, fB[1]);
        fB[2] = fma(w, v[3 * ix + 1].x, fB[2]);
        fB[3] = fma(w, v[3 * ix + 1].y, fB[3]);
        fB[4] = fma(w, v[3 * ix + 2].x, fB[4]);
        fB[5] = fma(w, v[3 * ix + 2].y, fB[5]);
      }
    }
  }
  // a particle whose stencil is not inside this CTA's tile gathers through L1 (any order stays correct)
  if (outA | outB) {
    Stencil s;
    if (outA) { make_stencil<true>(g, ax, ay, az, s); gather6(F6, g, s, fA); }
    if (outB) { make_stencil<true>(g, bx, by, bz, s); gather6(F6, g, s, fB); }
  }
  kA = rotate(fA, A.vx, A.vy, A.vz, pp.ht, pp.ht2);
  kB = rotate(fB, B.vx, B.vy, B.vz, pp.ht, pp.ht2);
  return outA | outB;
}

// ---------------------------------------------------------------------------
// Corrector, two particles per thread.  F:1162-1295, partbc F:1337, slab test
// of the drive kick F:1343-1345; writes the next sort key (cell of
// x' + lookahead*v', periodic / reflecting images folded in index space -- a
// sorting hint only).
// ---------------------------------------------------------------------------
constexpr int CORR_NST = 3;
template <int NW>
struct CorrSmem {
  static constexpr int ring_d = NW * CORR_NST * PRING_D;
  static constexpr int bytes = (6 * TILE_ROW_D + ring_d) * 8 + (NW * CORR_NST + 1) * 8;
};

struct Slab {                 // drive-kick slab of F:1343-1345
  unsigned* bits; int* list; int* count;
};
__device__ __forceinline__ void slab_test(const PushParams& pp, const ParticleSoA& P, const Slab& sl, int p, double y, double z) {
  if ((fabs(z - pp.zcent) < pp.zw) && ((fabs(y - pp.ycent2) < pp.yw) || (fabs(y - pp.ycent1) < pp.yw))) {
    const int id = P.id ? P.id[p] : p;
    atomicOr(sl.bits + (id >> 5), 1u << (id & 31));
    sl.list[atomicAdd(sl.count, 1)] = p;
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, MRG_CPAIR_MINB)
k_correct_pair(GP g, PushParams pp, ParticleSoA P, const double* __restrict__ F6, const int* __restrict__ cell_end,
               double* __restrict__ wk_partial, Slab sl, int* __restrict__ key_out, double lookahead) {
  extern __shared__ __align__(128) double smem_dyn[];
  double* sF = smem_dyn;
  double* sRing = sF + 6 * TILE_ROW_D;
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sRing + CorrSmem<NW>::ring_d);
  unsigned long long& bar = sBar[NW * CORR_NST];
  const int lane = threadIdx.x & 31;
  const int w = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
  const Tile t = tile_of(g, cell_end, blockIdx.x);
  const bool busy = t.p1 > t.p0;
  double wx = 0.0, wh = 0.0;
  if (busy) {
    PairStream st;
    pstream_open<CORR_NST>(P, t, w, NW, lane, sRing + w * (CORR_NST * PRING_D), sBar + w * CORR_NST, st);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    stage_fields(g, t, F6, sF, &bar);
    mbar_wait(&bar, 0);
    // only tiles within one cell of the drive slab can hold slab particles after the move (|dt*v| < 1 cell);
    // particles that were gathered outside the tile are tested unconditionally
    bool near_slab = false;
    if (pp.drive_on) {
      const double zl = (t.k - 1.5) * g.hz, zh = (t.k + 1.5) * g.hz, yl = (t.j - 1.0) * g.hy, yh = (t.j + 2.0) * g.hy;
      const bool zin = (zh > pp.zcent - pp.zw) && (zl < pp.zcent + pp.zw);
      const bool y1 = (yh > pp.ycent1 - pp.yw) && (yl < pp.ycent1 + pp.yw);
      const bool y2 = (yh > pp.ycent2 - pp.yw) && (yl < pp.ycent2 + pp.yw);
      near_slab = zin && (y1 || y2);
    }
    const double hh2 = 0.5 * pp.hh;
#pragma unroll 1
    for (int it = 0; it < st.nit; it++) {
      P6 A, B;
      int g0;
      bool okA, okB;
      pstream_next<CORR_NST>(P, st, it, lane, A, B, g0, okA, okB);
      Kick kA, kB;
      const bool stray = push_pair(g, pp, t, sF, F6, A, B, kA, kB);
      if (okA) { wx += kA.wx; wh += kA.wh; }
      if (okB) { wx += kB.wx; wh += kB.wh; }
      double2 X, Y, Z, U, V, Wv;
      X.x = fma(pp.dt, fma(hh2, kA.dvx, A.vx), A.x);          // F:1289-1291
      Y.x = fma(pp.dt, fma(hh2, kA.dvy, A.vy), A.y);
      Z.x = fma(pp.dt, fma(hh2, kA.dvz, A.vz), A.z);
      U.x = fma(pp.hh, kA.dvx, A.vx);                          // F:1293-1295
      V.x = fma(pp.hh, kA.dvy, A.vy);
      Wv.x = fma(pp.hh, kA.dvz, A.vz);
      X.y = fma(pp.dt, fma(hh2, kB.dvx, B.vx), B.x);
      Y.y = fma(pp.dt, fma(hh2, kB.dvy, B.vy), B.y);
      Z.y = fma(pp.dt, fma(hh2, kB.dvz, B.vz), B.z);
      U.y = fma(pp.hh, kB.dvx, B.vx);
      V.y = fma(pp.hh, kB.dvy, B.vy);
      Wv.y = fma(pp.hh, kB.dvz, B.vz);
      if (__any_sync(FULL, maybe_wrap(g, X.x, Y.x, Z.x) | maybe_wrap(g, X.y, Y.y, Z.y))) {   // partbc, F:1337
        if (wrap_pos(g, X.x, Y.x, Z.x)) V.x = -V.x;
        if (wrap_pos(g, X.y, Y.y, Z.y)) V.y = -V.y;
      }
      int2 key;
      if (key_out) {
        key.x = sort_cell_folded(g, fma(lookahead, U.x, X.x), fma(lookahead, V.x, Y.x), fma(lookahead, Wv.x, Z.x));
        key.y = sort_cell_folded(g, fma(lookahead, U.y, X.y), fma(lookahead, V.y, Y.y), fma(lookahead, Wv.y, Z.y));
      }
      if (__all_sync(FULL, okA && okB)) {
        __stcs(reinterpret_cast<double2*>(P.x + g0), X); __stcs(reinterpret_cast<double2*>(P.y + g0), Y);
        __stcs(reinterpret_cast<double2*>(P.z + g0), Z); __stcs(reinterpret_cast<double2*>(P.vx + g0), U);
        __stcs(reinterpret_cast<double2*>(P.vy + g0), V); __stcs(reinterpret_cast<double2*>(P.vz + g0), Wv);
        if (key_out) *reinterpret_cast<int2*>(key_out + g0) = key;
      } else {
        if (okA) {
          P.x[g0] = X.x; P.y[g0] = Y.x; P.z[g0] = Z.x; P.vx[g0] = U.x; P.vy[g0] = V.x; P.vz[g0] = Wv.x;
          if (key_out) key_out[g0] = key.x;
        }
        if (okB) {
          P.x[g0 + 1] = X.y; P.y[g0 + 1] = Y.y; P.z[g0 + 1] = Z.y; P.vx[g0 + 1] = U.y; P.vy[g0 + 1] = V.y; P.vz[g0 + 1] = Wv.y;
          if (key_out) key_out[g0 + 1] = key.y;
        }
      }
      if (pp.drive_on) {
        // in-tile particles can only be in the slab when the tile is near it; out-of-tile ones are always tested
        if (near_slab || __any_sync(FULL, stray)) {
          if (okA) slab_test(pp, P, sl, g0, Y.x, Z.x);
          if (okB) slab_test(pp, P, sl, g0 + 1, Y.y, Z.y);
        }
      }
      __syncwarp();
    }
  }
  warp_wk_store(wx, wh, wk_partial, NW);
}

// ---------------------------------------------------------------------------
// Predictor, two particles per thread.  F:1162-1283, 1300-1306, partbc F:1375,
// srimp1 + srimp2 scatter (F:2273-2374, 2471-2529) through the cell-run
// pre-reduction of mrg_tile.cuh: factors of the 64 particles of an iteration
// are parked in shared memory in slot order, quads of lanes accumulate runs of
// equal cells in registers.
// ---------------------------------------------------------------------------
constexpr int PRED_NST = 2;
constexpr int PARK_D = 64 * 18;                  // doubles parked per warp: 64 x (wxz[9] + key + qvy[8])
template <int NW>
struct PredSmem {
  static constexpr int ring_d = NW * PRED_NST * PRING_D;
  static constexpr int bytes = (6 * TILE_ROW_D + 6 * TILE_ACC_D + NW * PARK_D + ring_d) * 8 + (NW * PRED_NST + 1) * 8;
};

// park layout per warp: W[64][10] then Q[4][64] double2 (value rows 2q, 2q+1 of slot s at Q[q][s])
__device__ __forceinline__ void park_slot(double* W, double* Q, int slot, const double qvy[8], const double wxz[9], int key) {
  double2* Wp = reinterpret_cast<double2*>(W + slot * PR_W_STRIDE);
  Wp[0] = make_double2(wxz[0], wxz[1]);
  Wp[1] = make_double2(wxz[2], wxz[3]);
  Wp[2] = make_double2(wxz[4], wxz[5]);
  Wp[3] = make_double2(wxz[6], wxz[7]);
  Wp[4] = make_double2(wxz[8], __longlong_as_double((long long)key));
  double2* Qp = reinterpret_cast<double2*>(Q);
#pragma unroll
  for (int qq = 0; qq < 4; qq++) Qp[qq * 64 + slot] = make_double2(qvy[2 * qq], qvy[2 * qq + 1]);
}

// predicted position / velocity, partbc and the scatter factors of one particle (F:1300-1306, 1375, 2274-2313)
__device__ __forceinline__ void predict_factors(const GP& g, const PushParams& pp, const P6& c, const Kick& k, bool ok,
                                                bool do_wrap, double ah, double hh2, double qvy[8], double wxz[9], int& key) {
  Predicted o;
  o.vxj = fma(ah, k.dvx, c.vx);
  o.vyj = fma(ah, k.dvy, c.vy);
  o.vzj = fma(ah, k.dvz, c.vz);
  o.rx = fma(pp.adt, fma(hh2, k.dvx, c.vx), c.x);
  o.ry = fma(pp.adt, fma(hh2, k.dvy, c.vy), c.y);
  o.rz = fma(pp.adt, fma(hh2, k.dvz, c.vz), c.z);
  if (do_wrap) {
    if (wrap_pos(g, o.rx, o.ry, o.rz)) o.vyj = -o.vyj;
  }
  key = scatter_factors(g, ok ? pp.qmult : 0.0, o, qvy, wxz);   // masked slots deposit exact zeros
  if (!ok) key = -1;
}

// phase B over `nsub` sub-iterations of 8 parked slots; same contract as deposit_parked
template <int NSUB, bool TILED>
__device__ __forceinline__ void deposit_parked64(const double* W, const double* Q, int lane, double* acc, int& cur,
                                                 const Target<TILED>& tg) {
  const int q = lane & 3, pl = lane >> 2;
#pragma unroll 1
  for (int sub = 0; sub < NSUB; sub++) {
    const int p = sub * 8 + pl;
    const double2* Wp = reinterpret_cast<const double2*>(W + p * PR_W_STRIDE);
    const double2 w01 = Wp[0], w23 = Wp[1], w45 = Wp[2], w67 = Wp[3], w8k = Wp[4];
    double2 qv = reinterpret_cast<const double2*>(Q)[q * 64 + p];
    const int key = (int)__double_as_longlong(w8k.y);
    const double wxz[9] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y, w8k.x};
    bool pending = (key >= 0) && (key != cur);     // slots of another cell than the current run
    if (__any_sync(FULL, pending)) {
      // run boundary (or a stray): finish the members of the current run, then switch cell by cell
      bool done = (key < 0);
      for (;;) {
        const bool member = !done && (key == cur);
        const double ax = member ? qv.x : 0.0, ay = member ? qv.y : 0.0;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          acc[r] = fma(ax, wxz[r], acc[r]);
          acc[9 + r] = fma(ay, wxz[r], acc[9 + r]);
        }
        done = done || member;
        const unsigned rest = __ballot_sync(FULL, !done);
        if (rest == 0u) break;
        const int kk = __shfl_sync(FULL, key, __ffs(rest) - 1);
        if (cur >= 0) flush_quad<TILED>(acc, cur, tg);
#pragma unroll
        for (int n = 0; n < 18; n++) acc[n] = 0.0;
        cur = kk;
      }
    } else {
#pragma unroll
      for (int r = 0; r < 9; r++) {
        acc[r] = fma(qv.x, wxz[r], acc[r]);
        acc[9 + r] = fma(qv.y, wxz[r], acc[9 + r]);
      }
    }
  }
}

template <int NW>
__global__ void __launch_bounds__(NW * 32, MRG_PPAIR_MINB)
k_predict_pair(GP g, PushParams pp, ParticleSoA P, const double* __restrict__ F6, double* __restrict__ M4,
               const int* __restrict__ cell_end, double* __restrict__ wk_partial) {
  extern __shared__ __align__(128) double smem_dyn[];
  double* sF = smem_dyn;                                       // [6][TILE_ROW_D]   staged fields
  double* sM = sF + 6 * TILE_ROW_D;                            // [6][TILE_ACC_D]   moment accumulators
  double* sPark = sM + 6 * TILE_ACC_D;                         // [NW][PARK_D]
  double* sRing = sPark + NW * PARK_D;                         // [NW][PRED_NST][PRING_D]
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sRing + PredSmem<NW>::ring_d);
  unsigned long long& bar = sBar[NW * PRED_NST];
  const int lane = threadIdx.x & 31;
  const int w = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
  const Tile t = tile_of(g, cell_end, blockIdx.x);
  const bool busy = t.p1 > t.p0;                              // block-uniform
  double wx = 0.0, wh = 0.0;
  if (busy) {
    PairStream st;
    pstream_open<PRED_NST>(P, t, w, NW, lane, sRing + w * (PRED_NST * PRING_D), sBar + w * PRED_NST, st);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
#if MRG_PRED_SMEM_TILE
    for (int e = threadIdx.x; e < 6 * TILE_ACC_D; e += blockDim.x) sM[e] = 0.0;
#endif
    __syncthreads();
    stage_fields(g, t, F6, sF, &bar);
    mbar_wait(&bar, 0);
    double* W = sPark + w * PARK_D;
    double* Q = W + 64 * PR_W_STRIDE;
    const Target<(MRG_PRED_SMEM_TILE != 0)> tg(g, M4, sM, t.n0_first, t.ncell, lane);
    double acc[18];
#pragma unroll
    for (int n = 0; n < 18; n++) acc[n] = 0.0;
    int cur = -1;
    const double ah = pp.aimpl * pp.hh, hh2 = 0.5 * pp.hh;
#pragma unroll 1
    for (int it = 0; it < st.nit; it++) {
      P6 A, B;
      int g0;
      bool okA, okB;
      pstream_next<PRED_NST>(P, st, it, lane, A, B, g0, okA, okB);
      Kick kA, kB;
      push_pair(g, pp, t, sF, F6, A, B, kA, kB);
      if (okA) { wx += kA.wx; wh += kA.wh; }
      if (okB) { wx += kB.wx; wh += kB.wh; }
      // partbc of the predicted positions is needed by a few lanes per step: decide per warp
      const bool mw = maybe_wrap(g, fma(pp.adt, fma(hh2, kA.dvx, A.vx), A.x), fma(pp.adt, fma(hh2, kA.dvy, A.vy), A.y),
                                 fma(pp.adt, fma(hh2, kA.dvz, A.vz), A.z)) |
                      maybe_wrap(g, fma(pp.adt, fma(hh2, kB.dvx, B.vx), B.x), fma(pp.adt, fma(hh2, kB.dvy, B.vy), B.y),
                                 fma(pp.adt, fma(hh2, kB.dvz, B.vz), B.z));
      const bool do_wrap = __any_sync(FULL, mw);
      {
        double qvy[8], wxz[9];
        int key;
        predict_factors(g, pp, A, kA, okA, do_wrap, ah, hh2, qvy, wxz, key);
        park_slot(W, Q, 2 * lane, qvy, wxz, key);
      }
      {
        double qvy[8], wxz[9];
        int key;
        predict_factors(g, pp, B, kB, okB, do_wrap, ah, hh2, qvy, wxz, key);
        park_slot(W, Q, 2 * lane + 1, qvy, wxz, key);
      }
      __syncwarp();
      deposit_parked64<8, (MRG_PRED_SMEM_TILE != 0)>(W, Q, lane, acc, cur, tg);
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
  warp_wk_store(wx, wh, wk_partial, NW);
}

// cell histogram of the sort keys emitted by the corrector (warp-aggregated)
__global__ void k_key_hist(long long n, const int* __restrict__ key, int* __restrict__ hist) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool valid = t < n;
  const unsigned act = __ballot_sync(FULL, valid);
  if (!valid) return;
  const int kcell = key[t];
  const unsigned m = __match_any_sync(act, kcell);
  if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(hist + kcell, __popc(m));
}

}  // namespace mrg
