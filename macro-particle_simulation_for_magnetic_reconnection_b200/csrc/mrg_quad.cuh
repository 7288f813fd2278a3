// mrg_quad.cuh -- particle passes with a QUAD-COOPERATIVE polynomial gather.
//
//   k_correct_quad   ipc == 0: gather, implicit rotation, in-place update,
//                    partbc, drive-slab test, next sort key
//   k_predict_quad   ipc >= 1: gather, implicit rotation, predicted position /
//                    velocity, partbc, fused srimp1+srimp2 deposition (the
//                    cell-run pre-reduction of mrg_tile.cuh)
//
// Why: ncu shows the gather of mrg_tile.cuh (54 warp-broadcast LDS.128 per
// particle, 18 nodes x 6 fields) keeping the shared-memory data pipe at 74 % of
// its peak while the fp64 pipe is 35 % busy.  Two changes remove most of that
// traffic and a fifth of the arithmetic:
//
//   * The interpolant of one field on one cell (TSC in x,z, linear in y,
//     F:1208-1270) is a polynomial of degree (2,1,2) in the in-cell offsets
//     (xx,yy,zz).  The CTA converts the staged stencil rows of its 16-cell
//     pencil once into a coefficient table; a field value is then 17 FMAs
//     (Horner) instead of 18 FMAs + 6 shared weight products.
//   * The four lanes of a quad serve four consecutive (cell-sorted) particles
//     TOGETHER: lane role r = (F,b) holds the nine coefficients of yy^b of the
//     three fields of E (F=0) or B (F=1) of the quad's cell -- 14 LDS.128 per
//     lane, and since all quads of a warp usually sit in one or two cells the
//     loads are broadcasts of four addresses -- evaluates its three half
//     polynomials at all four particles (8 FMAs each) and hands the results to
//     their owners with xor-shuffles.  Per particle that is 3.5 LDS.128 + 7.5
//     64-bit shuffles instead of 54 LDS.128.
//   * A quad whose particles straddle two cells takes a second pass (the loop
//     below runs once per distinct cell of the quad); particles whose stencil is
//     not in the CTA's pencil (stale sort, wall row jp >= my) use the generic
//     gather through L1, so results never depend on the particle order.
//
// Particle streams, field staging, deposition and the tile bookkeeping are the
// ones of mrg_tile.cuh (pencils of QT_CELLS = 16 cells here: the coefficient
// table costs 912 bytes per cell).  F:n = /root/reference/@mrg37-080A.f03 line n.
#pragma once
#include "mrg_lane.cuh"

namespace mrg {

constexpr int QT_CELLS = 16;                  // cells per CTA pencil
constexpr int QT_NODES = QT_CELLS + 2;
constexpr int QT_ROW_D = QT_NODES * 6;        // doubles per staged field row
constexpr int QT_ROLE_D = 28;                 // 27 coefficients of a role, padded to 14 x 16 bytes
constexpr int QT_CELL_D = 4 * QT_ROLE_D + 2;  // +2: neighbouring cells start 4 banks apart (conflict-free broadcasts)
constexpr int QT_TAB_D = QT_CELLS * QT_CELL_D;
#ifndef MRG_QCORR_MINB
#define MRG_QCORR_MINB 5
#endif
#ifndef MRG_QPRED_MINB
#define MRG_QPRED_MINB 4
#endif
constexpr int QC_NST = 3, QP_NST = 2;         // particle ring stages of the corrector / predictor

template <int CELLS>
__device__ __forceinline__ Tile tile_of_n(const GP& g, const int* __restrict__ cell_end, int tile) {
  const int ntx = (g.mx + CELLS - 1) / CELLS;
  Tile t;
  const int tx = tile % ntx, r = tile / ntx;
  t.j = r % g.my;
  t.k = r / g.my;
  t.i0 = tx * CELLS;
  t.ncell = min(CELLS, g.mx - t.i0);
  const int c0 = t.i0 + g.mx * (t.j + g.my * t.k);
  t.p0 = (c0 == 0) ? 0 : cell_end[c0 - 1];
  t.p1 = cell_end[c0 + t.ncell - 1];
  t.n0_first = node_of(g, t.i0 - 1, t.j, t.k - 1);
  return t;
}

// stage the 6 stencil rows (jy = 0,1; kz = 0,1,2) of the packed fields, QT_ROW_D doubles apart
__device__ __forceinline__ void qstage_fields(const GP& g, const Tile& t, const double* __restrict__ F6, double* sF,
                                              unsigned long long* bar) {
  if (threadIdx.x == 0) {
    const unsigned row_bytes = (unsigned)(t.ncell + 2) * 48u;
    mbar_expect_tx(bar, 6u * row_bytes);
#pragma unroll
    for (int kz = 0; kz < 3; kz++)
#pragma unroll
      for (int jy = 0; jy < 2; jy++) {
        const size_t node = (size_t)t.n0_first + (size_t)jy * g.nx + (size_t)kz * g.nxy;
        bulk_g2s(sF + (kz * 2 + jy) * QT_ROW_D, F6 + node * 6, row_bytes, bar);
      }
  }
}

// Coefficient table entry of (cell, role r = 2F + b): T[f*9 + a*3 + cz] multiplies xx^a zz^cz in the
// yy^b part of field 3F + f.  x and z: TSC (tsc_poly); y (quirk Q3): node jl carries yy, node jr 1 - yy.
__device__ __forceinline__ void quad_table_entry(const double* sF, int cell, int role, double* T) {
  const int F = role >> 1, b = role & 1;
#pragma unroll
  for (int f = 0; f < 3; f++) {
    double Y[3][3];   // [a][kz]
#pragma unroll
    for (int kz = 0; kz < 3; kz++) {
      const double* r0 = sF + (kz * 2 + 0) * QT_ROW_D + cell * 6 + 3 * F + f;   // row jl
      const double* r1 = r0 + QT_ROW_D;                                          // row jr
      double a0, a1, a2, c0, c1, c2;
      tsc_poly(r1[0], r1[6], r1[12], c0, c1, c2);
      if (b) {
        tsc_poly(r0[0], r0[6], r0[12], a0, a1, a2);
        Y[0][kz] = a0 - c0; Y[1][kz] = a1 - c1; Y[2][kz] = a2 - c2;
      } else {
        Y[0][kz] = c0; Y[1][kz] = c1; Y[2][kz] = c2;
      }
    }
#pragma unroll
    for (int a = 0; a < 3; a++) tsc_poly(Y[a][0], Y[a][1], Y[a][2], T[f * 9 + a * 3 + 0], T[f * 9 + a * 3 + 1], T[f * 9 + a * 3 + 2]);
  }
}
__device__ __forceinline__ void quad_build_table(const double* sF, double* sT, int ncell) {
  for (int e = threadIdx.x; e < ncell * 4; e += blockDim.x) {
    double T[28];
    quad_table_entry(sF, e >> 2, e & 3, T);
    T[27] = 0.0;
    double2* dst = reinterpret_cast<double2*>(sT + (e >> 2) * QT_CELL_D + (e & 3) * QT_ROLE_D);
#pragma unroll
    for (int q = 0; q < 14; q++) dst[q] = make_double2(T[2 * q], T[2 * q + 1]);
  }
}

// one half polynomial: sum_{a,cz} T[a*3+cz] xx^a zz^cz  (8 FMA)
__device__ __forceinline__ double half_poly(const double* T, double xx, double zz) {
  const double t0 = fma(fma(T[6], xx, T[3]), xx, T[0]);
  const double t1 = fma(fma(T[7], xx, T[4]), xx, T[1]);
  const double t2 = fma(fma(T[8], xx, T[5]), xx, T[2]);
  return fma(fma(t2, zz, t1), zz, t0);
}

// Gather of the six prepared fields for the lane's particle (in-cell offsets gc, cell delta d inside the
// pencil or anything else when the generic path must be used), cooperatively per quad.  `want` = the lane
// holds a particle that needs field values.  Returns false for lanes that must gather generically.
__device__ __forceinline__ bool quad_gather(const double* sT, int ncell, int lane, bool want, int d, const GCoord& gc, double f[6]) {
  const int r = lane & 3, qb = lane & ~3;
  const bool in_tile = want && ((unsigned)d < (unsigned)ncell);
  bool done = !in_tile;
  // per-lane constants of the assembly: own role (F,b); partner k has role r^k
  const bool b = r & 1, F = (r >> 1) & 1;
  const double wa = b ? gc.yy : 1.0, wb = b ? 1.0 : gc.yy;   // weights of the (own-b, other-b) halves
#pragma unroll 1
  for (;;) {
    const unsigned todo = __ballot_sync(FULL, !done);
    if (todo == 0u) break;
    const unsigned qm = (todo >> qb) & 0xFu;
    const int leader = qb + (qm ? (__ffs(qm) - 1) : 0);
    const int target = __shfl_sync(FULL, d, leader);          // quad-uniform; garbage-free only when qm != 0
    const int cell = qm ? target : 0;
    double T[28];
    {
      const double2* src = reinterpret_cast<const double2*>(sT + cell * QT_CELL_D + r * QT_ROLE_D);
#pragma unroll
      for (int q = 0; q < 14; q++) { const double2 v = src[q]; T[2 * q] = v.x; T[2 * q + 1] = v.y; }
    }
    double R[4][3];   // R[k][f]: half polynomial of role r^k of field f at MY particle
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double xk = k ? __shfl_xor_sync(FULL, gc.xx, k) : gc.xx;
      const double zk = k ? __shfl_xor_sync(FULL, gc.zz, k) : gc.zz;
#pragma unroll
      for (int fi = 0; fi < 3; fi++) {
        const double v = half_poly(T + 9 * fi, xk, zk);
        R[k][fi] = k ? __shfl_xor_sync(FULL, v, k) : v;
      }
    }
    if (!done && d == cell) {
#pragma unroll
      for (int fi = 0; fi < 3; fi++) {
        const double same = fma(R[0][fi], wa, R[1][fi] * wb);   // field fi of my own set F
        const double othr = fma(R[2][fi], wa, R[3][fi] * wb);   // field fi of the other set
        f[fi] = F ? othr : same;                                 // exa,eya,eza
        f[3 + fi] = F ? same : othr;                             // bxa,bya,bza
      }
      done = true;
    }
  }
  return in_tile;
}

// ---------------------------------------------------------------------------
// Particle stream with a compile-time ring depth (same scheme as Stream in mrg_tile.cuh).
// ---------------------------------------------------------------------------
template <int NST>
struct QStream {
  int a, b, nit, issued;
  double* ring;                // [NST][6][STAGE_D]
  unsigned long long* bar;     // [NST]
  __device__ __forceinline__ void issue(const ParticleSoA& P, int lane) {
    if (issued < nit) {
      if (lane == 0) {
        const int slot = issued % NST;
        const int e = (a + 32 * issued) & ~1;
        double* dst = ring + slot * 6 * STAGE_D;
        unsigned long long* bb = bar + slot;
        mbar_expect_tx(bb, 6u * STAGE_BYTES);
        bulk_g2s_stream(dst + 0 * STAGE_D, P.x + e, STAGE_BYTES, bb);
        bulk_g2s_stream(dst + 1 * STAGE_D, P.y + e, STAGE_BYTES, bb);
        bulk_g2s_stream(dst + 2 * STAGE_D, P.z + e, STAGE_BYTES, bb);
        bulk_g2s_stream(dst + 3 * STAGE_D, P.vx + e, STAGE_BYTES, bb);
        bulk_g2s_stream(dst + 4 * STAGE_D, P.vy + e, STAGE_BYTES, bb);
        bulk_g2s_stream(dst + 5 * STAGE_D, P.vz + e, STAGE_BYTES, bb);
      }
      issued++;
    }
  }
  __device__ __forceinline__ void open(const ParticleSoA& P, const Tile& t, int w, int nwarps, int lane, double* ring_, unsigned long long* bar_) {
    const int N = (t.p1 - t.p0 + 31) >> 5;
    const int i0 = (w * N) / nwarps, i1 = ((w + 1) * N) / nwarps;
    a = t.p0 + 32 * i0;
    b = min(t.p0 + 32 * i1, t.p1);
    nit = i1 - i0;
    issued = 0;
    ring = ring_;
    bar = bar_;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < NST; s++) mbar_init(bar + s, 1);
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NST - 1; s++) issue(P, lane);
  }
  __device__ __forceinline__ void next(const ParticleSoA& P, int it, int lane, P6& o) {
    issue(P, lane);                                            // refill the slot consumed in iteration it-1
    const int slot = it % NST;
    mbar_wait(bar + slot, (unsigned)((it / NST) & 1));
    const int start = a + 32 * it;
    const double* src = ring + slot * 6 * STAGE_D + (start & 1) + lane;
    o.x = src[0 * STAGE_D]; o.y = src[1 * STAGE_D]; o.z = src[2 * STAGE_D];
    o.vx = src[3 * STAGE_D]; o.vy = src[4 * STAGE_D]; o.vz = src[5 * STAGE_D];
  }
};

// half-step position + cooperative gather (+ generic fallback) + rotation of the lane's particle
__device__ __forceinline__ Kick quad_gather_rotate(const GP& g, const PushParams& pp, const Tile& t, const double* sT,
                                                   const double* __restrict__ F6, int lane, bool valid, const P6& c, bool& generic) {
  double rx = __dadd_rn(c.x, __dmul_rn(pp.hdt, c.vx));        // F:1163-1165
  double ry = __dadd_rn(c.y, __dmul_rn(pp.hdt, c.vy));
  double rz = __dadd_rn(c.z, __dmul_rn(pp.hdt, c.vz));
  wrap_pos(g, rx, ry, rz);                                    // partbcEST, F:1168
  GCoord gc;
  gather_coords(g, rx, ry, rz, gc);
  double f[6];
#pragma unroll
  for (int e = 0; e < 6; e++) f[e] = 0.0;
  const bool fast = quad_gather(sT, t.ncell, lane, valid, gc.n0 - t.n0_first, gc, f);
  generic = valid && !fast;
  if (generic) {                                       // stencil outside the pencil / wall row: through L1
    Stencil s;
    make_stencil<true>(g, rx, ry, rz, s);
    gather6(F6, g, s, f);
  }
  return rotate(f, c.vx, c.vy, c.vz, pp.ht, pp.ht2);
}

// ---------------------------------------------------------------------------
// Corrector.  F:1162-1295, partbc F:1337, slab test of the drive kick F:1343-1345, next sort key.
// ---------------------------------------------------------------------------
constexpr int QCORR_SMEM_BYTES = (6 * QT_ROW_D + QT_TAB_D + PR_WARPS * QC_NST * 6 * STAGE_D) * 8 + (PR_WARPS * QC_NST + 1) * 8;
__global__ void __launch_bounds__(PR_WARPS * 32, MRG_QCORR_MINB)
k_correct_quad(GP g, PushParams pp, ParticleSoA P, const double* __restrict__ F6, const int* __restrict__ cell_end,
               double* __restrict__ wk_out, Slab sl, int* __restrict__ key_out, double lookahead) {
  extern __shared__ __align__(128) double smem_dyn[];
  double* sF = smem_dyn;                                       // [6][QT_ROW_D]
  double* sT = sF + 6 * QT_ROW_D;                              // coefficient table
  double* sRing = sT + QT_TAB_D;                               // [warps][QC_NST*6*STAGE_D]
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sRing + PR_WARPS * QC_NST * 6 * STAGE_D);
  unsigned long long& bar = sBar[PR_WARPS * QC_NST];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Tile t = tile_of_n<QT_CELLS>(g, cell_end, blockIdx.x);
  if (t.p1 <= t.p0) return;                                    // block-uniform
  QStream<QC_NST> st;
  st.open(P, t, w, PR_WARPS, lane, sRing + w * (QC_NST * 6 * STAGE_D), sBar + w * QC_NST);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  qstage_fields(g, t, F6, sF, &bar);
  mbar_wait(&bar, 0);
  quad_build_table(sF, sT, t.ncell);
  __syncthreads();
  bool near_slab = false;
  if (pp.drive_on) {
    const double zl = (t.k - 1.5) * g.hz, zh = (t.k + 1.5) * g.hz, yl = (t.j - 1.0) * g.hy, yh = (t.j + 2.0) * g.hy;
    const bool zin = (zh > pp.zcent - pp.zw) && (zl < pp.zcent + pp.zw);
    const bool y1 = (yh > pp.ycent1 - pp.yw) && (yl < pp.ycent1 + pp.yw);
    const bool y2 = (yh > pp.ycent2 - pp.yw) && (yl < pp.ycent2 + pp.yw);
    near_slab = zin && (y1 || y2);
  }
  double wx = 0.0, wh = 0.0;
  const double hh2 = 0.5 * pp.hh;
#pragma unroll 1
  for (int it = 0; it < st.nit; it++) {
    P6 c;
    st.next(P, it, lane, c);
    const int p = st.a + 32 * it + lane;
    const bool valid = p < st.b;
    bool generic;
    const Kick k = quad_gather_rotate(g, pp, t, sT, F6, lane, valid, c, generic);
    if (valid) {
      wx += k.wx; wh += k.wh;
      double x = fma(pp.dt, fma(hh2, k.dvx, c.vx), c.x);      // F:1289-1291
      double y = fma(pp.dt, fma(hh2, k.dvy, c.vy), c.y);
      double z = fma(pp.dt, fma(hh2, k.dvz, c.vz), c.z);
      const double vx = fma(pp.hh, k.dvx, c.vx);              // F:1293-1295
      double vy = fma(pp.hh, k.dvy, c.vy);
      const double vz = fma(pp.hh, k.dvz, c.vz);
      if (wrap_pos(g, x, y, z)) vy = -vy;                     // partbc, F:1337
      __stcs(P.x + p, x); __stcs(P.y + p, y); __stcs(P.z + p, z);
      __stcs(P.vx + p, vx); __stcs(P.vy + p, vy); __stcs(P.vz + p, vz);
      if (key_out) key_out[p] = sort_cell_folded(g, fma(lookahead, vx, x), fma(lookahead, vy, y), fma(lookahead, vz, z));
      // in-pencil particles can only reach the slab from a pencil next to it (|dt*v| < 1 cell); the
      // few that were gathered generically are always tested
      if (pp.drive_on && (near_slab || generic)) slab_test(pp, P, sl, p, y, z);
    }
    __syncwarp();
  }
  wx = warp_sum(wx);
  wh = warp_sum(wh);
  if (lane == 0) { atomicAdd(wk_out, wx); atomicAdd(wk_out + 1, wh); }
}

// ---------------------------------------------------------------------------
// Predictor.  F:1162-1283, 1300-1306, partbc F:1375, srimp1 + srimp2 scatter through the cell-run
// pre-reduction of mrg_tile.cuh (factors parked in shared memory, quads accumulate runs of equal cells,
// transposing butterfly, shared-memory moment tile, one red.global.add.f64 per touched value).
// ---------------------------------------------------------------------------
constexpr int QPRED_SMEM_BYTES =
    (QT_TAB_D + 6 * TILE_ACC_D + PR_WARPS * (32 * PR_W_STRIDE + PR_Q_D + QP_NST * 6 * STAGE_D)) * 8 + (PR_WARPS * QP_NST + 1) * 8;
static_assert(PR_WARPS * (32 * PR_W_STRIDE + PR_Q_D) >= 6 * QT_ROW_D, "staged field rows alias the park area");
__global__ void __launch_bounds__(PR_WARPS * 32, MRG_QPRED_MINB)
k_predict_quad(GP g, PushParams pp, ParticleSoA P, const double* __restrict__ F6, double* __restrict__ M4,
               const int* __restrict__ cell_end, double* __restrict__ wk_out, int group_min) {
  extern __shared__ __align__(128) double smem_dyn[];
  double* sT = smem_dyn;                                       // coefficient table
  double* sM = sT + QT_TAB_D;                                  // [6][TILE_ACC_D] moment accumulators
  double* smW = sM + 6 * TILE_ACC_D;                           // [warps][32*PR_W_STRIDE]; during set-up: staged field rows
  double* smQ = smW + PR_WARPS * 32 * PR_W_STRIDE;             // [warps][256]
  double* sRing = smQ + PR_WARPS * PR_Q_D;                        // [warps][QP_NST*6*STAGE_D]
  unsigned long long* sBar = reinterpret_cast<unsigned long long*>(sRing + PR_WARPS * QP_NST * 6 * STAGE_D);
  unsigned long long& bar = sBar[PR_WARPS * QP_NST];
  double* sF = smW;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const Tile t = tile_of_n<QT_CELLS>(g, cell_end, blockIdx.x);
  if (t.p1 <= t.p0) return;                                    // block-uniform
  QStream<QP_NST> st;
  st.open(P, t, w, PR_WARPS, lane, sRing + w * (QP_NST * 6 * STAGE_D), sBar + w * QP_NST);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  for (int e = threadIdx.x; e < 6 * TILE_ACC_D; e += blockDim.x) sM[e] = 0.0;
  __syncthreads();
  qstage_fields(g, t, F6, sF, &bar);
  mbar_wait(&bar, 0);
  quad_build_table(sF, sT, t.ncell);
  __syncthreads();                                             // sF is dead from here on: the park area takes over
  double* W = smW + w * (32 * PR_W_STRIDE);
  double* Q = smQ + w * PR_Q_D;
  const Target<true> tg(g, M4, sM, t.n0_first, t.ncell, lane);
  double acc[18];
#pragma unroll
  for (int n = 0; n < 18; n++) acc[n] = 0.0;
  int cur = -1;
  double wx = 0.0, wh = 0.0;
  const double ah = pp.aimpl * pp.hh, hh2 = 0.5 * pp.hh;
#pragma unroll 1
  for (int it = 0; it < st.nit; it++) {
    P6 c;
    st.next(P, it, lane, c);
    const bool valid = st.a + 32 * it + lane < st.b;
    int key = -1;
    {
      bool generic;
      const Kick k = quad_gather_rotate(g, pp, t, sT, F6, lane, valid, c, generic);
      double qvy[8], wxz[9];
      if (valid) {
        wx += k.wx; wh += k.wh;
        Predicted o;
        o.vxj = fma(ah, k.dvx, c.vx);                         // F:1300-1302
        o.vyj = fma(ah, k.dvy, c.vy);
        o.vzj = fma(ah, k.dvz, c.vz);
        o.rx = fma(pp.adt, fma(hh2, k.dvx, c.vx), c.x);       // F:1304-1306
        o.ry = fma(pp.adt, fma(hh2, k.dvy, c.vy), c.y);
        o.rz = fma(pp.adt, fma(hh2, k.dvz, c.vz), c.z);
        if (wrap_pos(g, o.rx, o.ry, o.rz)) o.vyj = -o.vyj;    // partbc, F:1375
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
    deposit_parked<true>(reinterpret_cast<const double2*>(W + (lane >> 2) * PR_W_STRIDE),
                         reinterpret_cast<const double2*>(Q) + (lane & 3) * PR_Q_ROW + (lane >> 2), key, lane, acc, cur, group_min, tg);
    __syncwarp();
  }
  if (cur >= 0) flush_quad<true>(acc, cur, tg);
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
  wx = warp_sum(wx);
  wh = warp_sum(wh);
  if (lane == 0) { atomicAdd(wk_out, wx); atomicAdd(wk_out + 1, wh); }
}

}  // namespace mrg
