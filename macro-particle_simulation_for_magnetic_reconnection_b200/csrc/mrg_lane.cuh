// mrg_lane.cuh -- particle passes with REGISTER-STATIONARY operands.
//
//   k_lane<true>    ipc >= 1: gather, implicit rotation, predicted position /
//                   velocity, partbc, fused srimp1+srimp2 deposition
//   k_lane<false>   ipc == 0: gather, implicit rotation, in-place update,
//                   partbc, drive-slab test, next sort key
//
// Why: ncu showed the broadcast-gather kernels (mrg_tile.cuh / mrg_pair.cuh)
// bound by the shared-memory data pipe (74 % of peak wavefronts; fp64 pipe only
// 35 % busy): every particle pulls its 108 field values (54 LDS.128) and parks /
// re-reads 18 deposit factors.  Here the operands that are shared by all the
// particles of a cell stay in REGISTERS and the particles stream past them:
//
//   * One warp owns a tile of LT_CELLS = 16 cells along x.  Its particles are
//     stored (by mrg_sort, layout 1) as 16 interleaved RUNS: slot(run s, row r)
//     = p0 + 16 r + s; run s holds the cell-sorted particles [s n/16, (s+1) n/16)
//     of the tile, so a run stays in one cell for ~ppc consecutive rows.
//   * Lane pair (2s, 2s+1) walks run s, two rows per round: lane h = 0 owns
//     particle A (row c), lane h = 1 owns particle B (row c+1).  A warp load of
//     one coordinate is 256 contiguous bytes.
//   * Gather (F:1217-1270): the trilinear-in-y / TSC-in-x,z interpolant of a
//     cell is a polynomial of degree (2,1,2) in the in-cell offsets (xx,yy,zz):
//     18 coefficients per field, built once per tile into shared memory.  Lane
//     h = 0 keeps the 54 coefficients of exa,eya,eza of its run's current cell
//     in registers, lane h = 1 those of bxa,bya,bza; both evaluate their three
//     fields for A and for B by Horner (17 FMA per field instead of 18 FMA + 6
//     weight products), then swap three values with one xor-shuffle each.
//   * Deposit (F:2273-2374, 2471-2529): lane h keeps the 36 accumulators
//     (9 nodes x 4 moments) of stencil row jy = h of the run's current cell in
//     registers; every particle adds its own row and, after a 13-value
//     shuffle, its partner's.  Accumulators go to the warp's shared-memory
//     moment tile when the run changes cell (plain read-modify-write: the tile
//     is private to the warp, conflicting pairs take turns) and the tile goes to
//     global memory once, with red.global.add.f64.
//   * Anything that does not fit (stencil outside the tile, the jp >= my wall
//     row, a scatter cell different from the run's cell) takes the generic
//     L1 / global-atomic path, so results never depend on how stale the sort is.
//
// Per 32 particles this needs ~305 fp64 instructions (predictor) and < 100
// shared-memory wavefronts instead of 380 and 320.
// F:n = /root/reference/@mrg37-080A.f03 line n.
#pragma once
#include "mrg_pair.cuh"

namespace mrg {

#ifndef MRG_LT_CELLS
#define MRG_LT_CELLS 8
#endif
constexpr int LT_CELLS = MRG_LT_CELLS;       // cells per warp tile (<= 16: one table cell per lane pair)
constexpr int LT_NODES = LT_CELLS + 2;
constexpr int LT_ROW_D = LT_NODES * 6;       // doubles per staged field row
constexpr int LT_ACC_D = LT_NODES * 4;       // doubles per moment-tile row
constexpr int LT_RUNS = 16;                  // interleaved runs per tile (= lane pairs)
constexpr int LT_COEF_D = LT_CELLS * 108;    // coefficient table: [cell][half][3 fields][18]

// slot <-> (run, row) of the interleaved tile layout (also used by k_sort_scatter_lane)
__host__ __device__ __forceinline__ int lt_run_rows(int n, int s) { return (n > s) ? ((n - s + 15) >> 4) : 0; }

struct LTile {
  int i0, ncell, j, k;
  int p0, n;
  int n0_first;     // stencil base node of cell (i0,j,k)
};
__device__ __forceinline__ LTile ltile_of(const GP& g, const int* __restrict__ cell_end, int tile) {
  const int ntx = (g.mx + LT_CELLS - 1) / LT_CELLS;
  LTile t;
  const int tx = tile % ntx, r = tile / ntx;
  t.j = r % g.my;
  t.k = r / g.my;
  t.i0 = tx * LT_CELLS;
  t.ncell = min(LT_CELLS, g.mx - t.i0);
  const int c0 = t.i0 + g.mx * (t.j + g.my * t.k);
  t.p0 = (c0 == 0) ? 0 : cell_end[c0 - 1];
  t.n = cell_end[c0 + t.ncell - 1] - t.p0;
  t.n0_first = node_of(g, t.i0 - 1, t.j, t.k - 1);
  return t;
}

// Polynomial coefficients of one field on one cell from its 18 stencil nodes
// N[ix][jy][kz].  With xx,zz in [-1/2,1/2) and yy in [0,1):
//   x (TSC, F:1208-1210): fxl = 1/8 - xx/2 + xx^2/2, fxc = 3/4 - xx^2, fxr = 1/8 + xx/2 + xx^2/2
//   y (F:1212-1213, quirk Q3): node jl=jp carries fyl = yy, node jr=jp+1 carries 1-yy
//   z like x.   C[a*6 + b*3 + c] multiplies xx^a yy^b zz^c.
__device__ __forceinline__ void tsc_poly(double l, double c, double r, double& p0, double& p1, double& p2) {
  const double s = l + r;
  p0 = fma(0.125, s, 0.75 * c);
  p1 = 0.5 * (r - l);
  p2 = fma(0.5, s, -c);
}
__device__ __forceinline__ void cell_coefficients(const double* sF, int cell, int comp, double* C) {
  double X[3][2][3];   // [a][jy][kz]
#pragma unroll
  for (int kz = 0; kz < 3; kz++)
#pragma unroll
    for (int jy = 0; jy < 2; jy++) {
      const double* row = sF + (kz * 2 + jy) * LT_ROW_D + cell * 6 + comp;
      tsc_poly(row[0], row[6], row[12], X[0][jy][kz], X[1][jy][kz], X[2][jy][kz]);
    }
#pragma unroll
  for (int a = 0; a < 3; a++) {
    double Y[2][3];
#pragma unroll
    for (int kz = 0; kz < 3; kz++) {
      Y[0][kz] = X[a][1][kz];
      Y[1][kz] = X[a][0][kz] - X[a][1][kz];
    }
#pragma unroll
    for (int b = 0; b < 2; b++) tsc_poly(Y[b][0], Y[b][1], Y[b][2], C[a * 6 + b * 3 + 0], C[a * 6 + b * 3 + 1], C[a * 6 + b * 3 + 2]);
  }
}

// three fields at (xx,yy,zz): 17 FMA each
__device__ __forceinline__ void horner3(const double* C, double xx, double yy, double zz, double out[3]) {
#pragma unroll
  for (int f = 0; f < 3; f++) {
    const double* c = C + f * 18;
    double u[3];
#pragma unroll
    for (int cz = 0; cz < 3; cz++) {
      const double t0 = fma(fma(c[12 + cz], xx, c[6 + cz]), xx, c[cz]);
      const double t1 = fma(fma(c[15 + cz], xx, c[9 + cz]), xx, c[3 + cz]);
      u[cz] = fma(t1, yy, t0);
    }
    out[f] = fma(fma(u[2], zz, u[1]), zz, u[0]);
  }
}

// cell index + in-cell offsets of the gather position (F:1175-1177, 1203-1213)
struct GCoord {
  int n0;           // stencil base node; INT_MAX-like marker when the generic path must be used
  double xx, yy, zz;
};
__device__ __forceinline__ void gather_coords(const GP& g, double rx, double ry, double rz, GCoord& s) {
  const double tx = __dmul_rn(g.hxi, rx);
  const double ty = __dmul_rn(g.hyi, ry);
  const double tz = __dmul_rn(g.hzi, rz);
  double ipd, jpd, kpd;
  int ip = floor_pos(__dadd_rn(tx, 0.500000001), ipd);
  int jp = floor_pos(__dadd_rn(ty, 0.000000001), jpd);
  int kp = floor_pos(__dadd_rn(tz, 0.500000001), kpd);
  const bool odd = (ip < 0) | (ip > g.mx) | (jp < 0) | (jp >= g.my) | (kp < 0) | (kp > g.mz);   // wall row F:1191-1195 / clamped
  s.n0 = odd ? 0x3fffffff : (ip + 1) + g.nx * ((jp + 1) + g.ny * (kp + 1));
  s.xx = __dsub_rn(tx, ipd);
  s.yy = __dsub_rn(ty, jpd);
  s.zz = __dsub_rn(tz, kpd);
}

__device__ __forceinline__ double shx(double v) { return __shfl_xor_sync(FULL, v, 1); }

// accumulators of stencil row jy = h of cell `cur` -> the warp's moment tile (plain RMW; caller serialises conflicts)
__device__ __forceinline__ void acc_to_tile(const double* acc, double* sM, int cur, int h) {
#pragma unroll
  for (int r = 0; r < 9; r++) {
    const int kz = r / 3, ix = r - 3 * kz;
    double2* p = reinterpret_cast<double2*>(sM + ((kz * 2 + h) * LT_NODES + cur + ix) * 4);
    double2 a = p[0], b = p[1];
    a.x += acc[r * 4 + 0]; a.y += acc[r * 4 + 1];
    b.x += acc[r * 4 + 2]; b.y += acc[r * 4 + 3];
    p[0] = a; p[1] = b;
  }
}

// Flush the accumulators of every pair with need = true (pair-uniform; cur >= 0).  Pairs whose cells are
// at least 3 apart touch disjoint tile nodes, so one class cur % 3 goes at a time; pairs of a class that
// share the same cell take turns.
__device__ __forceinline__ void flush_pairs(bool need, int cur, int h, int lane, const double* acc, double* sM) {
  unsigned pending = __ballot_sync(FULL, need);
  while (pending) {
    const int l0 = __ffs(pending) - 1;
    const int cls = __shfl_sync(FULL, cur, l0) % 3;
    const bool cand = ((pending >> lane) & 1u) && (cur % 3 == cls);
    const unsigned cm = __ballot_sync(FULL, cand);
    bool go = false;
    if (cand) {
      const unsigned grp = __match_any_sync(cm, cur);
      go = (lane >> 1) == ((__ffs(grp) - 1) >> 1);
    }
    if (go) acc_to_tile(acc, sM, cur, h);
    pending &= ~__ballot_sync(FULL, go);
    __syncwarp();
  }
}

// ---- asynchronous copies ------------------------------------------------------------------------
constexpr int LT_RING = 4;                   // particle ring stages (rounds in flight), power of two
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// six SoA base pointers (particles, or the predicted-state scratch of the split predictor)
struct Six { double* p[6]; };

// Lane-private ring of the lane's next rows: stage k holds row (2*idx + h) of run s for idx % LT_RING == k,
// as ring[k][array][lane].  `off` = slot of (run s, row 0) in the arrays.
__device__ __forceinline__ void ring_issue(double* sRing, const Six& A, int off, int idx, int h, int len, int lane) {
  const int r = 2 * idx + h;
  if (r < len) {
    double* d = sRing + (idx & (LT_RING - 1)) * 192 + lane;
    const int o = off + r * LT_RUNS;
#pragma unroll
    for (int a = 0; a < 6; a++) cp_async8(d + 32 * a, A.p[a] + o);
  }
}
__device__ __forceinline__ void ring_read(const double* sRing, int idx, int lane, double v[6]) {
  const double* rp = sRing + (idx & (LT_RING - 1)) * 192 + lane;
#pragma unroll
  for (int a = 0; a < 6; a++) v[a] = rp[32 * a];
}
// lanes of runs without particles never copy anything: give them a harmless particle
__device__ __forceinline__ void ring_clear(double* sRing, int lane) {
#pragma unroll
  for (int k = 0; k < LT_RING * 6; k++) sRing[k * 32 + lane] = 0.0;
}

// -------------------------------------------------------------------------------------------------
// k_lane<MODE>: gather + implicit rotation with the cell's field polynomials in registers.
//   MODE 0  corrector (ipc = 0): in-place update, partbc, next sort key, drive-slab test
//   MODE 1  first half of the predictor (ipc >= 1): predicted position (after partbc) and velocity of
//           every particle -> scratch arrays Q (same slot); k_lane_deposit scatters them.
// Persistent: warp (= CTA) b handles tiles b, b + gridDim.x, ...  Every lane streams its own rows through
// a ring of 8-byte cp.async copies (lane-private slots, so the pair-local cursors need no cross-lane
// bookkeeping); with >= 10 resident warps per SM the ring covers the HBM latency.
//
// Round of pair s (cursor c even): lane h owns row c + h.  If A = row c and B = row c + 1 have the same
// home cell both are pushed in this round; otherwise A goes first and B alone in the next round
// (`second`), so rows always advance in aligned pairs.
// -------------------------------------------------------------------------------------------------
#ifndef MRG_LANE_MINB
#define MRG_LANE_MINB 12
#endif
template <int MODE>
__global__ void __launch_bounds__(32, MRG_LANE_MINB)
k_lane(GP g, PushParams pp, ParticleSoA P, Six Q, const double* __restrict__ F6, const int* __restrict__ cell_end,
       double* __restrict__ wk_out, Slab sl, int* __restrict__ key_out, double lookahead, int ntiles) {
  __shared__ __align__(128) double sF[6 * LT_ROW_D];           // staged field rows
  __shared__ __align__(16) double sC[LT_COEF_D];
  __shared__ __align__(16) double sRing[LT_RING * 6 * 32];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x;
  const int s = lane >> 1, h = lane & 1;
  if (lane == 0) mbar_init(&bar, 1);
  __syncwarp();
  unsigned phase = 0;
  double wx = 0.0, wh = 0.0;
  const double hh2 = 0.5 * pp.hh, ah = pp.aimpl * pp.hh;
  Six A;
  A.p[0] = P.x; A.p[1] = P.y; A.p[2] = P.z; A.p[3] = P.vx; A.p[4] = P.vy; A.p[5] = P.vz;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const LTile t = ltile_of(g, cell_end, tile);
    if (t.n <= 0) continue;                                    // warp-uniform

    // ---- stage the six stencil rows of the packed fields (1-D bulk TMA) ---------------------------
    __syncwarp();
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const unsigned row_bytes = (unsigned)(t.ncell + 2) * 48u;
      mbar_expect_tx(&bar, 6u * row_bytes);
#pragma unroll
      for (int kz = 0; kz < 3; kz++)
#pragma unroll
        for (int jy = 0; jy < 2; jy++) {
          const size_t node = (size_t)t.n0_first + (size_t)jy * g.nx + (size_t)kz * g.nxy;
          bulk_g2s(sF + (kz * 2 + jy) * LT_ROW_D, F6 + node * 6, row_bytes, &bar);
        }
    }
    // ---- particle ring ------------------------------------------------------------------------
    const int len = lt_run_rows(t.n, s);                       // rows of run s
    const int off = t.p0 + s;
    if (len == 0) ring_clear(sRing, lane);
#pragma unroll
    for (int k = 0; k < LT_RING; k++) {
      ring_issue(sRing, A, off, k, h, len, lane);
      cp_async_commit();
    }
    mbar_wait(&bar, phase);
    phase ^= 1u;

    // ---- coefficient table: lane (cell s, half h) converts its three fields -----------------------
    double C[54];
    if (s < t.ncell) {
#pragma unroll
      for (int f = 0; f < 3; f++) cell_coefficients(sF, s, 3 * h + f, C + 18 * f);
      double2* dst = reinterpret_cast<double2*>(sC + (s * 2 + h) * 54);
#pragma unroll
      for (int e = 0; e < 27; e++) dst[e] = make_double2(C[2 * e], C[2 * e + 1]);
    }
    __syncwarp();

    int cur = -1;                                              // cell (0..ncell-1) of C, pair-uniform
    int c = 0;                                                 // cursor: row of particle A (even), pair-uniform
    bool second = false;                                       // this round handles B alone, pair-uniform
    bool near_slab = false;
    if (MODE == 0 && pp.drive_on) {
      const double zl = (t.k - 1.5) * g.hz, zh = (t.k + 1.5) * g.hz, yl = (t.j - 1.0) * g.hy, yh = (t.j + 2.0) * g.hy;
      const bool zin = (zh > pp.zcent - pp.zw) && (zl < pp.zcent + pp.zw);
      const bool y1 = (yh > pp.ycent1 - pp.yw) && (yl < pp.ycent1 + pp.yw);
      const bool y2 = (yh > pp.ycent2 - pp.yw) && (yl < pp.ycent2 + pp.yw);
      near_slab = zin && (y1 || y2);
    }

#pragma unroll 1
    for (;;) {
      const bool activeA = c < len;                            // pair-uniform
      if (!__any_sync(FULL, activeA)) break;
      cp_async_wait<LT_RING - 1>();
      double q[6];                                             // own particle: row c + h of run s
      ring_read(sRing, c >> 1, lane, q);
      // half-step position, partbcEST, cell + offsets                                F:1163-1177
      double rx = __dadd_rn(q[0], __dmul_rn(pp.hdt, q[3]));
      double ry = __dadd_rn(q[1], __dmul_rn(pp.hdt, q[4]));
      double rz = __dadd_rn(q[2], __dmul_rn(pp.hdt, q[5]));
      wrap_pos(g, rx, ry, rz);
      GCoord gc;
      gather_coords(g, rx, ry, rz, gc);
      const int d_own = gc.n0 - t.n0_first;                    // 0..ncell-1 inside the tile
      const int dA = __shfl_sync(FULL, d_own, lane & ~1), dB = __shfl_sync(FULL, d_own, lane | 1);
      const bool validB = c + 1 < len;
      const bool strayA = (unsigned)dA >= (unsigned)t.ncell, strayB = (unsigned)dB >= (unsigned)t.ncell;
      bool doA, doB, adv;
      if (!second) {
        doA = activeA;
        doB = activeA && validB && (strayA ? strayB : (dB == dA));
        adv = doB || !validB;
      } else {
        doA = false; doB = true; adv = true;
      }
      const int home = second ? dB : dA;
      const bool change = activeA && ((unsigned)home < (unsigned)t.ncell) && (home != cur);
      second = activeA && !adv;
      if (__any_sync(FULL, change)) {
        if (change) {
          const double2* src = reinterpret_cast<const double2*>(sC + (home * 2 + h) * 54);
#pragma unroll
          for (int e = 0; e < 27; e++) { const double2 v = src[e]; C[2 * e] = v.x; C[2 * e + 1] = v.y; }
          cur = home;
        }
      }
      const bool proc = h ? doB : doA;                         // this lane's own particle is handled in this round
      const int slot = off + (c + h) * LT_RUNS;
      // refill the ring stage of this round as soon as its values are in registers
      if (activeA && adv) {
        c += 2;
        ring_issue(sRing, A, off, (c >> 1) + LT_RING - 1, h, len, lane);
      }
      cp_async_commit();
      // ---- gather: own three fields at the own and at the partner's particle, then swap  F:1217-1270
      double fo[3], fp[3], f[6];
      horner3(C, gc.xx, gc.yy, gc.zz, fo);
      horner3(C, shx(gc.xx), shx(gc.yy), shx(gc.zz), fp);
#pragma unroll
      for (int e = 0; e < 3; e++) {
        const double got = shx(fp[e]);                         // the partner's three fields at my particle
        f[e] = h ? got : fo[e];                                // exa,eya,eza
        f[3 + e] = h ? fo[e] : got;                            // bxa,bya,bza
      }
      const bool stray = (unsigned)d_own >= (unsigned)t.ncell;
      if (proc && stray) {                                     // generic gather through L1 (any order stays correct)
        Stencil st;
        make_stencil<true>(g, rx, ry, rz, st);
        gather6(F6, g, st, f);
      }
      const Kick k = rotate(f, q[3], q[4], q[5], pp.ht, pp.ht2);   // F:1272-1283
      if (proc) { wx += k.wx; wh += k.wh; }

      if (MODE == 1) {
        // predicted velocity / position, partbc                                      F:1300-1306, 1375
        double vxj = fma(ah, k.dvx, q[3]);
        double vyj = fma(ah, k.dvy, q[4]);
        double vzj = fma(ah, k.dvz, q[5]);
        double x = fma(pp.adt, fma(hh2, k.dvx, q[3]), q[0]);
        double y = fma(pp.adt, fma(hh2, k.dvy, q[4]), q[1]);
        double z = fma(pp.adt, fma(hh2, k.dvz, q[5]), q[2]);
        if (wrap_pos(g, x, y, z)) vyj = -vyj;
        if (proc) {
          __stcs(Q.p[0] + slot, x); __stcs(Q.p[1] + slot, y); __stcs(Q.p[2] + slot, z);
          __stcs(Q.p[3] + slot, vxj); __stcs(Q.p[4] + slot, vyj); __stcs(Q.p[5] + slot, vzj);
        }
      } else {
        // in-place update, partbc                                                    F:1289-1295, 1337
        double x = fma(pp.dt, fma(hh2, k.dvx, q[3]), q[0]);
        double y = fma(pp.dt, fma(hh2, k.dvy, q[4]), q[1]);
        double z = fma(pp.dt, fma(hh2, k.dvz, q[5]), q[2]);
        const double vx = fma(pp.hh, k.dvx, q[3]);
        double vy = fma(pp.hh, k.dvy, q[4]);
        const double vz = fma(pp.hh, k.dvz, q[5]);
        if (wrap_pos(g, x, y, z)) vy = -vy;
        if (proc) {
          __stcs(P.x + slot, x); __stcs(P.y + slot, y); __stcs(P.z + slot, z);
          __stcs(P.vx + slot, vx); __stcs(P.vy + slot, vy); __stcs(P.vz + slot, vz);
          if (key_out) key_out[slot] = sort_cell_folded(g, fma(lookahead, vx, x), fma(lookahead, vy, y), fma(lookahead, vz, z));
          if (pp.drive_on && (near_slab || stray)) slab_test(pp, P, sl, slot, y, z);
        }
      }
    }
    cp_async_wait<0>();
  }
  wx = warp_sum(wx);
  wh = warp_sum(wh);
  if (lane == 0) { atomicAdd(wk_out, wx); atomicAdd(wk_out + 1, wh); }
}

// -------------------------------------------------------------------------------------------------
// k_lane_deposit: second half of the predictor -- srimp1 + srimp2 (F:2273-2374, 2471-2529) of the
// predicted states in Q.  Lane h of pair s keeps the 36 accumulators (9 nodes x 4 moments) of stencil
// row jy = h of the run's current cell in registers; a particle adds its own row and, after a 13-value
// xor-shuffle, its partner's.  When the run moves to another cell the accumulators are added to the
// warp's shared-memory moment tile (flush_pairs), the tile goes to global memory once per tile.  A
// particle whose cell differs from its run's cell (B of a pair that straddles two cells, a particle
// outside the tile) is deposited with 72 global atomics.  Rows advance two per round for every pair.
// -------------------------------------------------------------------------------------------------
#ifndef MRG_DEP_MINB
#define MRG_DEP_MINB 16
#endif
__global__ void __launch_bounds__(32, MRG_DEP_MINB)
k_lane_deposit(GP g, double qmult, Six Q, double* __restrict__ M4, const int* __restrict__ cell_end, int ntiles) {
  __shared__ __align__(16) double sM[6 * LT_ACC_D];
  __shared__ __align__(16) double sRing[LT_RING * 6 * 32];
  const int lane = threadIdx.x;
  const int s = lane >> 1, h = lane & 1;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const LTile t = ltile_of(g, cell_end, tile);
    if (t.n <= 0) continue;                                    // warp-uniform
    __syncwarp();
    const int len = lt_run_rows(t.n, s);
    const int off = t.p0 + s;
    if (len == 0) ring_clear(sRing, lane);
#pragma unroll
    for (int k = 0; k < LT_RING; k++) {
      ring_issue(sRing, Q, off, k, h, len, lane);
      cp_async_commit();
    }
    for (int e = lane; e < 6 * LT_ACC_D; e += 32) sM[e] = 0.0;
    __syncwarp();
    double acc[36];
#pragma unroll
    for (int e = 0; e < 36; e++) acc[e] = 0.0;
    int cur = -1;
    const int rounds = (lt_run_rows(t.n, 0) + 1) >> 1;         // run 0 is the longest

#pragma unroll 1
    for (int it = 0; it < rounds; it++) {
      cp_async_wait<LT_RING - 1>();
      double q[6];
      ring_read(sRing, it, lane, q);
      const bool valid = 2 * it + h < len;
      ring_issue(sRing, Q, off, it + LT_RING, h, len, lane);
      cp_async_commit();
      Stencil st;
      make_stencil<false>(g, q[0], q[1], q[2], st);            // F:2274-2308
      const int d_own = st.n0 - t.n0_first;
      const int dA = __shfl_sync(FULL, d_own, lane & ~1);
      const bool activeA = 2 * it < len;
      const bool change = activeA && ((unsigned)dA < (unsigned)t.ncell) && (dA != cur);
      if (__any_sync(FULL, change)) {
        flush_pairs(change && cur >= 0, cur, h, lane, acc, sM);
        if (change) {
#pragma unroll
          for (int e = 0; e < 36; e++) acc[e] = 0.0;
          cur = dA;
        }
      }
      double wxz[9];
#pragma unroll
      for (int kz = 0; kz < 3; kz++)
#pragma unroll
        for (int ix = 0; ix < 3; ix++) wxz[kz * 3 + ix] = st.fx[ix] * st.fz[kz];
      const bool mine = valid && (cur >= 0) && (d_own == cur);
      if (valid && !mine) {                                    // other cell: 72 global atomics
        double qvy[8];
#pragma unroll
        for (int jy = 0; jy < 2; jy++) {
          const double qf = qmult * st.fy[jy];
          qvy[jy * 4 + 0] = qf * q[3]; qvy[jy * 4 + 1] = qf * q[4]; qvy[jy * 4 + 2] = qf * q[5]; qvy[jy * 4 + 3] = qf;
        }
        deposit_direct72(qvy, wxz, st.n0, g, M4);
      }
      const double qm = mine ? qmult : 0.0;
      const double qo = qm * (h ? st.fy[1] : st.fy[0]);        // own stencil row jy = h
      const double qp = qm * (h ? st.fy[0] : st.fy[1]);        // partner's row
      const double own[4] = {qo * q[3], qo * q[4], qo * q[5], qo};
      double got[4];
      got[0] = shx(qp * q[3]); got[1] = shx(qp * q[4]); got[2] = shx(qp * q[5]); got[3] = shx(qp);
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const double w = wxz[r], pwr = shx(w);
#pragma unroll
        for (int m = 0; m < 4; m++) acc[r * 4 + m] = fma(got[m], pwr, fma(own[m], w, acc[r * 4 + m]));
      }
    }
    cp_async_wait<0>();
    flush_pairs(cur >= 0, cur, h, lane, acc, sM);
    // the moment tile -> global: 4 moments of a node = one 32-byte sector
    const int nodes = t.ncell + 2;
    for (int e = lane; e < 6 * nodes * 4; e += 32) {
      const int rw = e / (nodes * 4), rem = e - rw * (nodes * 4);
      const double v = sM[rw * LT_ACC_D + rem];
      if (v != 0.0) {
        const int kz = rw >> 1, jy = rw & 1;
        atomicAdd(M4 + 4 * ((size_t)t.n0_first + (size_t)jy * g.nx + (size_t)kz * g.nxy) + rem, v);
      }
    }
  }
}

// Scatter pass of the cell sort for the interleaved tile layout: the q-th particle (cell order) of a
// 16-cell tile with n particles goes to run s, row j (runs are consecutive pieces of the cell order;
// the first n - 16 (R-1) runs have R = ceil(n/16) rows, the others R-1), i.e. slot p0 + 16 j + s.
__global__ void k_sort_scatter_lane(long long n, int mx, const int* __restrict__ key, const int* __restrict__ start,
                                    int* __restrict__ cursor, SortArrays A) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool valid = t < n;
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const int kcell = key[t];
  const unsigned m = __match_any_sync(act, kcell);
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  int b = 0;
  if (lane == leader) b = atomicAdd(cursor + kcell, __popc(m));
  b = __shfl_sync(m, b, leader);
  const int qg = b + __popc(m & ((1u << lane) - 1u));          // position in cell order
  const int ci = kcell % mx;
  const int i0 = (ci / LT_CELLS) * LT_CELLS;
  const int c0 = kcell - (ci - i0);
  const int ncl = min(LT_CELLS, mx - i0);
  const int p0 = start[c0], nt = start[c0 + ncl] - p0;
  const int q = qg - p0;
  const int R = (nt + 15) >> 4, rf = nt - 16 * (R - 1);
  int s, j;
  if (q < rf * R) { s = q / R; j = q - s * R; }
  else { const int q2 = q - rf * R; const int s2 = q2 / (R - 1); s = rf + s2; j = q2 - s2 * (R - 1); }
  const int d = p0 + j * LT_RUNS + s;
#pragma unroll
  for (int c = 0; c < 6; c++) A.dst[c][d] = A.src[c][t];
  A.id_dst[d] = A.id_src ? A.id_src[t] : (int)t;
}

}  // namespace mrg
