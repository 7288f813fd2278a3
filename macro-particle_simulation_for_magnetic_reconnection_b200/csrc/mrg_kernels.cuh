// mrg_kernels.cuh -- sm_100a kernels of the /fulmov/ path.
//
//   field preparation   k_blend, k_dcsub, k_filter<axis>, k_finalize   F:1127-1148
//   particle passes     k_correct (ipc=0), k_predict<..> (ipc>=1)       F:1162-1309,1375
//   moment fold         k_fold_unpack                                   F:3243-3377
//   drive kick          k_popc, k_kick                                  F:1342-1364
//   maintenance         k_sort_keys, k_sort_scatter, scan kernels
//   synthetic load      k_loadpt                                        F:8937-9040
//
// Device layouts (all fp64):
//   particles   SoA x,y,z,vx,vy,vz (+ int32 id = original local index)
//   fields      F6[node][6]  = exa,eya,eza,bxa,bya,bza  (48 B / node)
//   moments     M4[node][4]  = qjx,qjy,qjz,q            (32 B / node = 1 sector)
//   node        = (i+2) + (mx+4)*((j+1) + (my+3)*(k+2))  (reference order)
#pragma once
#include "mrg_device.cuh"

namespace mrg {

struct Ptr6 { double* p[6]; };
struct CPtr6 { const double* p[6]; };
struct CPtr12 { const double* p[12]; };
struct Ptr4 { double* p[4]; };

// Warp-aggregated marking of occupied planes: bit kp of occ[] (all lanes call; lanes with valid = false mark
// nothing).  `cache` remembers the plane the warp marked last, so a cell-sorted warp issues one atomic per run.
__device__ __forceinline__ void mark_plane(unsigned* __restrict__ occ, int kp, bool valid, int& cache) {
  const int lo = __reduce_min_sync(0xffffffffu, valid ? kp : 0x7fffffff);
  const int hi = __reduce_max_sync(0xffffffffu, valid ? kp : -1);
  if (hi < 0 || (lo == hi && lo == cache)) return;              // warp-uniform
  const int lane = threadIdx.x & 31;
  if (lo == hi) {
    if (lane == 0) atomicOr(occ + (lo >> 5), 1u << (lo & 31));
    cache = lo;
  } else {
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned m = __match_any_sync(act, kp);
      if (lane == __ffs(m) - 1) atomicOr(occ + (kp >> 5), 1u << (kp & 31));
    }
    cache = -1;
  }
}

// ---------------------------------------------------------------------------
// Field preparation.  Interior = i in [0,mx), j in [0,my], k in [0,mz).
// All arithmetic follows the source association with _rn intrinsics so the
// prepared fields are bit-identical to the CPU restatement.
// ---------------------------------------------------------------------------
// `planes` != nullptr restricts the work to the listed z planes (k values): the
// particle passes of a rank that owns a z slab only read the prepared fields
// near its particles, so the other planes need not be prepared (PlaneSet in
// mrg_api.cu works out which planes each stage needs).
__device__ __forceinline__ bool interior_ijk(const GP& g, long long t, const int* __restrict__ planes, int nplanes,
                                             int& i, int& j, int& k) {
  const long long n = (long long)g.mx * (g.my + 1) * nplanes;
  if (t >= n) return false;
  i = (int)(t % g.mx);
  long long r = t / g.mx;
  j = (int)(r % (g.my + 1));
  k = (int)(r / (g.my + 1));
  if (planes) k = planes[k];
  return true;
}

// F:1127-1139: A = aimpl*f + (1-aimpl)*f0 (+dc for B); then F:7351-7359: T = A - dc.  A itself is only needed on the
// few nodes whose periodic images are ghosts (k_finalize), which recompute it, so no kernel stores it.
__device__ __forceinline__ double blend_A(const CPtr12& f, int c, int m, double aimpl, double om, double dcc) {
  double a = __dadd_rn(__dmul_rn(aimpl, f.p[c][m]), __dmul_rn(om, f.p[c + 6][m]));
  if (c >= 3) a = __dadd_rn(a, dcc);
  return a;
}
__device__ __forceinline__ double filter5(double a0, double a1, double a2, double a3, double a4) {   // source order of the sum
  double t = __dmul_rn(-0.0625, a0);
  t = __dadd_rn(t, __dmul_rn(0.25, a1));
  t = __dadd_rn(t, __dmul_rn(0.625, a2));
  t = __dadd_rn(t, __dmul_rn(0.25, a3));
  return __dsub_rn(t, __dmul_rn(0.0625, a4));
}
__global__ void k_blend(GP g, CPtr12 f, Ptr6 T, double aimpl, double om, double bxc, double byc, double bzc,
                        const int* __restrict__ planes, int nplanes) {
  int i, j, k;
  if (!interior_ijk(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, planes, nplanes, i, j, k)) return;
  const int m = node_of(g, i, j, k);
  const double dc[6] = {0.0, 0.0, 0.0, bxc, byc, bzc};
#pragma unroll
  for (int c = 0; c < 6; c++) T.p[c][m] = __dsub_rn(blend_A(f, c, m, aimpl, om, dc[c]), dc[c]);
}
// the blend fused with the first z sweep (F:7365-7395): the five planes of T are blended on the fly (they sit in L2
// between neighbouring planes of the launch), so neither A nor the unfiltered T is written
__global__ void k_blend_filter_z(GP g, CPtr12 f, Ptr6 D, double aimpl, double om, double bxc, double byc, double bzc,
                                 const int* __restrict__ planes, int nplanes) {
  int i, j, k;
  if (!interior_ijk(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, planes, nplanes, i, j, k)) return;
  const int kr = (k == g.mz - 1) ? 0 : k + 1, kl = (k == 0) ? g.mz - 1 : k - 1;
  const int krr = (kr == g.mz - 1) ? 0 : kr + 1, kll = (kl == 0) ? g.mz - 1 : kl - 1;
  const int m = node_of(g, i, j, k);
  const int m0 = node_of(g, i, j, krr), m1 = node_of(g, i, j, kr), m3 = node_of(g, i, j, kl), m4 = node_of(g, i, j, kll);
  const double dc[6] = {0.0, 0.0, 0.0, bxc, byc, bzc};
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const double a0 = __dsub_rn(blend_A(f, c, m0, aimpl, om, dc[c]), dc[c]);
    const double a1 = __dsub_rn(blend_A(f, c, m1, aimpl, om, dc[c]), dc[c]);
    const double a2 = __dsub_rn(blend_A(f, c, m, aimpl, om, dc[c]), dc[c]);
    const double a3 = __dsub_rn(blend_A(f, c, m3, aimpl, om, dc[c]), dc[c]);
    const double a4 = __dsub_rn(blend_A(f, c, m4, aimpl, om, dc[c]), dc[c]);
    D.p[c][m] = filter5(a0, a1, a2, a3, a4);
  }
}

// entry prefld of emfild (F:3820-3873): b = b0 + dt*(-curl ea) on the interior nodes, ea = aimpl*e + (1-aimpl)*e0 blended
// on the fly at the four neighbours each component reads.  One-sided differences with the mirror rows folded in (the
// factor 2) on the walls j = 0 and j = my, where by = 0.  Every operation is the source's, in its order, with _rn
// intrinsics (true division): the result is bit-identical to the host's prefld, so the host need not upload bx, by, bz.
__device__ __forceinline__ double blend_E(const CPtr12& f, int c, int m, double aimpl, double om) {
  return __dadd_rn(__dmul_rn(aimpl, f.p[c][m]), __dmul_rn(om, f.p[c + 6][m]));
}
__global__ void k_prefld(GP g, CPtr12 f, double* __restrict__ bx, double* __restrict__ by, double* __restrict__ bz,
                         double aimpl, double om, double dt, double hx2, double hy2, double hz2,
                         const int* __restrict__ planes, int nplanes) {
  int i, j, k;
  if (!interior_ijk(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, planes, nplanes, i, j, k)) return;
  const int kr = (k == g.mz - 1) ? 0 : k + 1, kl = (k == 0) ? g.mz - 1 : k - 1;      // pzr, pzl (F:8399-8422)
  const int ir = (i == g.mx - 1) ? 0 : i + 1, il = (i == 0) ? g.mx - 1 : i - 1;      // pxr, pxl (F:8341-8364)
  const int m = node_of(g, i, j, k);
  const double dey_z = __dsub_rn(blend_E(f, 1, node_of(g, i, j, kr), aimpl, om), blend_E(f, 1, node_of(g, i, j, kl), aimpl, om));
  const double dey_x = __dsub_rn(blend_E(f, 1, node_of(g, ir, j, k), aimpl, om), blend_E(f, 1, node_of(g, il, j, k), aimpl, om));
  double tx, tz;
  if (j >= 1 && j <= g.my - 1) {                                                     // F:3834-3853
    const double dez_y = __dsub_rn(blend_E(f, 2, node_of(g, i, j + 1, k), aimpl, om), blend_E(f, 2, node_of(g, i, j - 1, k), aimpl, om));
    const double dez_x = __dsub_rn(blend_E(f, 2, node_of(g, ir, j, k), aimpl, om), blend_E(f, 2, node_of(g, il, j, k), aimpl, om));
    const double dex_z = __dsub_rn(blend_E(f, 0, node_of(g, i, j, kr), aimpl, om), blend_E(f, 0, node_of(g, i, j, kl), aimpl, om));
    const double dex_y = __dsub_rn(blend_E(f, 0, node_of(g, i, j + 1, k), aimpl, om), blend_E(f, 0, node_of(g, i, j - 1, k), aimpl, om));
    tx = __dsub_rn(__ddiv_rn(dey_z, hz2), __ddiv_rn(dez_y, hy2));
    const double ty = __dsub_rn(__ddiv_rn(dez_x, hx2), __ddiv_rn(dex_z, hz2));
    tz = __dsub_rn(__ddiv_rn(dex_y, hy2), __ddiv_rn(dey_x, hx2));
    by[m] = __dadd_rn(f.p[10][m], __dmul_rn(dt, ty));
  } else {                                                                           // F:3856-3880
    const int j1 = (j == 0) ? 1 : g.my;               // the row the one-sided term reads: (i,1,k) at j = 0, (i,my,k) at j = my
    const double ez2 = __ddiv_rn(__dmul_rn(2.0, blend_E(f, 2, node_of(g, i, j1, k), aimpl, om)), hy2);
    const double ex2 = __ddiv_rn(__dmul_rn(2.0, blend_E(f, 0, node_of(g, i, j1, k), aimpl, om)), hy2);
    const double qz = __ddiv_rn(dey_z, hz2), qx = -__ddiv_rn(dey_x, hx2);
    tx = (j == 0) ? __dsub_rn(qz, ez2) : __dadd_rn(qz, ez2);
    tz = (j == 0) ? __dadd_rn(qx, ex2) : __dsub_rn(qx, ex2);
    by[m] = 0.0;
  }
  bx[m] = __dadd_rn(f.p[9][m], __dmul_rn(dt, tx));
  bz[m] = __dadd_rn(f.p[11][m], __dmul_rn(dt, tz));
}

// One (-1,4,10,4,-1)/16 sweep, F:7365-7395 (AXIS=2, z), F:7401-7434 (AXIS=0, x),
// F:7438-7492 (AXIS=1, y with wall mirror rows; rows j=0 and j=my are copied).
// y sweep of one node (F:7438-7492): rows j = 0 and j = my are copied; mirror rows a(-1) = sg*e(1), a(my+1) = sg*e(my-1)
// (F:7455-7471): the E call has sym=-1 (F:1144), the B call sym=+1 (F:1147); the y component takes -sym.
__device__ __forceinline__ double filter_y(const GP& g, const double* __restrict__ s, int c, int i, int j, int k, int m) {
  if (j < 1 || j > g.my - 1) return s[m];
  const double sg = ((c < 3) ? -1.0 : 1.0) * ((c % 3 == 1) ? -1.0 : 1.0);
  const int jp2 = j + 2, jm2 = j - 2;
  const double a0 = (jp2 == g.my + 1) ? sg * s[node_of(g, i, g.my - 1, k)] : s[node_of(g, i, jp2, k)];
  const double a1 = s[node_of(g, i, j + 1, k)];
  const double a2 = s[m];
  const double a3 = s[node_of(g, i, j - 1, k)];
  const double a4 = (jm2 == -1) ? sg * s[node_of(g, i, 1, k)] : s[node_of(g, i, jm2, k)];
  return filter5(a0, a1, a2, a3, a4);
}
template <int AXIS>
__global__ void k_filter(GP g, CPtr6 S, Ptr6 D, const int* __restrict__ planes, int nplanes) {
  int i, j, k;
  if (!interior_ijk(g, blockIdx.x * (long long)blockDim.x + threadIdx.x, planes, nplanes, i, j, k)) return;
  const int m = node_of(g, i, j, k);
#pragma unroll
  for (int c = 0; c < 6; c++) {
    const double* s = S.p[c];
    double a0, a1, a2, a3, a4;   // in source order of the sum
    if (AXIS == 2) {
      const int kr = (k == g.mz - 1) ? 0 : k + 1, kl = (k == 0) ? g.mz - 1 : k - 1;
      const int krr = (kr == g.mz - 1) ? 0 : kr + 1, kll = (kl == 0) ? g.mz - 1 : kl - 1;
      a0 = s[node_of(g, i, j, krr)]; a1 = s[node_of(g, i, j, kr)]; a2 = s[m];
      a3 = s[node_of(g, i, j, kl)]; a4 = s[node_of(g, i, j, kll)];
    } else if (AXIS == 0) {
      const int ir = (i == g.mx - 1) ? 0 : i + 1, il = (i == 0) ? g.mx - 1 : i - 1;
      const int irr = (ir == g.mx - 1) ? 0 : ir + 1, ill = (il == 0) ? g.mx - 1 : il - 1;
      a0 = s[node_of(g, ill, j, k)]; a1 = s[node_of(g, il, j, k)]; a2 = s[m];
      a3 = s[node_of(g, ir, j, k)]; a4 = s[node_of(g, irr, j, k)];
    } else {
      D.p[c][m] = filter_y(g, s, c, i, j, k, m);
      continue;
    }
    D.p[c][m] = filter5(a0, a1, a2, a3, a4);
  }
}

// Compose the packed gather array over the whole extended grid:
// interior nodes take the filtered value + dc (F:7498-7506); every ghost node
// takes what outmesh3 (F:3088-3148) left there BEFORE the filter ran, i.e. the
// unfiltered blend A of the periodic image, or zero on rows j=-1, my+1.
// With a plane list the entries are EXTENDED plane indices k+2; an entry with
// bit 30 set is a guard plane (just outside the prepared set) and is filled
// with NaN, so a gather that strays beyond the prepared planes cannot pass
// unnoticed.
constexpr int PLANE_GUARD = 1 << 30;
// FY: the last y sweep (F:7438-7492) is done here instead of in a pass of its own: T then holds the sweep's input
template <bool FY>
__global__ void k_finalize(GP g, CPtr12 f, CPtr6 T, double* __restrict__ F6, double aimpl, double om, double bxc, double byc,
                           double bzc, const int* __restrict__ planes, int nplanes) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nxy * nplanes) return;
  bool guard = false;
  if (planes) {
    const int e = planes[t / g.nxy];
    guard = (e & PLANE_GUARD) != 0;
    t = (t % g.nxy) + (long long)(e & ~PLANE_GUARD) * g.nxy;
  }
  const int i = (int)(t % g.nx) - 2;
  const int j = (int)((t / g.nx) % g.ny) - 1;
  const int k = (int)(t / g.nxy) - 2;
  const bool in = (i >= 0 && i < g.mx && j >= 0 && j <= g.my && k >= 0 && k < g.mz);
  const double dc[6] = {0.0, 0.0, 0.0, bxc, byc, bzc};
  double out[6];
  if (guard) {
#pragma unroll
    for (int c = 0; c < 6; c++) out[c] = __longlong_as_double(0x7ff8000000000000ll);
  } else if (in) {
#pragma unroll
    for (int c = 0; c < 6; c++) out[c] = __dadd_rn(FY ? filter_y(g, T.p[c], c, i, j, k, (int)t) : T.p[c][t], dc[c]);
  } else if (j < 0 || j > g.my) {
#pragma unroll
    for (int c = 0; c < 6; c++) out[c] = 0.0;
  } else {
    const int is = i < 0 ? i + g.mx : (i >= g.mx ? i - g.mx : i);
    const int ks = k < 0 ? k + g.mz : (k >= g.mz ? k - g.mz : k);
    const int m = node_of(g, is, j, ks);
#pragma unroll
    for (int c = 0; c < 6; c++) out[c] = blend_A(f, c, m, aimpl, om, dc[c]);
  }
  double2* o = reinterpret_cast<double2*>(F6) + t * 3;
  o[0] = make_double2(out[0], out[1]);
  o[1] = make_double2(out[2], out[3]);
  o[2] = make_double2(out[4], out[5]);
}

// slab-wise moment exchange: add the two neighbour strips into the rank's own block
__global__ void k_add_strips(double* __restrict__ lo, const double* __restrict__ rx_lo, double* __restrict__ hi,
                             const double* __restrict__ rx_hi, long long n) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n) return;
  lo[t] += rx_lo[t];
  hi[t] += rx_hi[t];
}

// Slab-wise exchange over NVLink peer memory: ONE kernel adds the two strips received from the ring neighbours into
// the rank's own block of the raw moments and pushes the finished block straight into every peer's copy of the array
// (plain 128-bit stores to the peers' mapped buffers, cudaIpc) -- what ncclAllGather + two ncclBroadcasts did after a
// separate add kernel.  [g0, g0 + cnt) is the block in doubles (ghost planes included on the two end ranks); the strips
// [lo0, lo0 + strip) and [hi0, hi0 + strip) lie inside it.  The 2-double allreduce that follows on the same stream is the
// barrier that tells every rank all blocks have landed.
struct PeerPtrs { double* p[8]; int n; };
// [hole0, hole0 + hole_cnt) inside the block is left out: an earlier launch of this kernel (strip = 0) already pushed it
// while the last part of the particle kernel was still running (split launch, mrg_api.cu)
__global__ void __launch_bounds__(256) k_add_push(double* __restrict__ M4, size_t g0, size_t cnt, size_t lo0, const double* __restrict__ rx_lo,
                                                  size_t hi0, const double* __restrict__ rx_hi, size_t strip, PeerPtrs peers,
                                                  size_t hole0, size_t hole_cnt) {
  const size_t n2 = (cnt - hole_cnt) >> 1;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n2; t += (size_t)gridDim.x * blockDim.x) {
    size_t g = g0 + 2 * t;
    if (g >= hole0) g += hole_cnt;
    double2 v = *reinterpret_cast<const double2*>(M4 + g);
    bool changed = false;
    if (g - lo0 < strip) { const double2 a = *reinterpret_cast<const double2*>(rx_lo + (g - lo0)); v.x += a.x; v.y += a.y; changed = true; }
    if (g - hi0 < strip) { const double2 a = *reinterpret_cast<const double2*>(rx_hi + (g - hi0)); v.x += a.x; v.y += a.y; changed = true; }
    if (changed) *reinterpret_cast<double2*>(M4 + g) = v;
#pragma unroll
    for (int q = 0; q < 8; q++)
      if (q < peers.n) *reinterpret_cast<double2*>(peers.p[q] + g) = v;
  }
  __threadfence_system();
}

// packed F6 -> six reference-layout arrays (mrg_get_prepared_fields)
__global__ void k_unpack6(GP g, const double* __restrict__ F6, Ptr6 out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= g.ntot) return;
#pragma unroll
  for (int c = 0; c < 6; c++) out.p[c][t] = F6[t * 6 + c];
}

// ---------------------------------------------------------------------------
// Per-block partial sums of wkix/wkih -> fixed-order final sum (deterministic
// for a fixed launch shape).  F:1282-1283, 1312-1317.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void block_wk_store(double wx, double wh, double* __restrict__ partial) {
  __shared__ double sh[2][32];
  wx = warp_sum(wx);
  wh = warp_sum(wh);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sh[0][w] = wx; sh[1][w] = wh; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int q = 0; q < nw; q++) { a += sh[0][q]; b += sh[1][q]; }
    partial[2 * blockIdx.x + 0] = a;
    partial[2 * blockIdx.x + 1] = b;
  }
}

__global__ void k_wk_final(const double* __restrict__ partial, int nblocks, double* __restrict__ out2) {
  __shared__ double sh[2][256];
  double a = 0.0, b = 0.0;
  for (int q = threadIdx.x; q < nblocks; q += 256) { a += partial[2 * q]; b += partial[2 * q + 1]; }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sh[0][threadIdx.x] += sh[0][threadIdx.x + s]; sh[1][threadIdx.x] += sh[1][threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = sh[0][0]; out2[1] = sh[1][0]; }
}

// ---------------------------------------------------------------------------
// Corrector pass, ipc = 0: F:1162-1295 + partbc F:1337 + slab test of the
// drive kick F:1343-1345 (the kick itself needs the serial RNG order and runs
// in k_kick).  One thread per particle; particle streams use evict-first
// loads/stores so the field working set stays in L1/L2.
// ---------------------------------------------------------------------------
struct PushParams {
  double dt, adt, hdt, aimpl;
  double hh, ht, ht2;       // F:1150-1152
  double qmult;
  // drive-kick slab (F:1343-1345): |z-zcent| < zw and (|y-ycent2| < yw or |y-ycent1| < yw)
  double zcent, ycent1, ycent2, zw, yw;
  int drive_on;
  // drive kick inside the tiled corrector (kick_inline): the particle whose original local index is id draws
  // ranfp number id + 1 of the stream that starts at kick_state, instead of the draw its position among the rank's
  // slab particles in l order selects (F:1342-1364 with a serial stream).  Only for ownerships the reference does
  // not have (z slabs): there is no reference stream to reproduce, and the serial order costs a bitmap over the
  // particle indices, a scan and a second kernel -- mostly on the ranks that own the slab.
  int kick_inline;
  unsigned kick_state;
  double Ez00, yw2;
};

struct ParticleSoA {
  double* x; double* y; double* z; double* vx; double* vy; double* vz;
  const int* id;            // original local index of the slot; nullptr = identity
  long long n;
};

__global__ void __launch_bounds__(256)
k_correct(GP g, PushParams pp, ParticleSoA P, const double* __restrict__ F6, double* __restrict__ wk_partial,
          unsigned* __restrict__ slab_bits, int* __restrict__ slab_list, int* __restrict__ slab_count) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double wx = 0.0, wh = 0.0;
  if (t < P.n) {
    double x = __ldcs(P.x + t), y = __ldcs(P.y + t), z = __ldcs(P.z + t);
    double vx = __ldcs(P.vx + t), vy = __ldcs(P.vy + t), vz = __ldcs(P.vz + t);
    double rx = __dadd_rn(x, __dmul_rn(pp.hdt, vx));      // F:1163-1165
    double ry = __dadd_rn(y, __dmul_rn(pp.hdt, vy));
    double rz = __dadd_rn(z, __dmul_rn(pp.hdt, vz));
    wrap_pos(g, rx, ry, rz);                                // partbcEST, F:1168
    Stencil s;
    make_stencil<true>(g, rx, ry, rz, s);
    double f[6];
    gather6(F6, g, s, f);
    const Kick k = rotate(f, vx, vy, vz, pp.ht, pp.ht2);
    wx = k.wx; wh = k.wh;
    const double hh2 = 0.5 * pp.hh;
    x = fma(pp.dt, fma(hh2, k.dvx, vx), x);                 // F:1289-1291
    y = fma(pp.dt, fma(hh2, k.dvy, vy), y);
    z = fma(pp.dt, fma(hh2, k.dvz, vz), z);
    vx = fma(pp.hh, k.dvx, vx);                             // F:1293-1295
    vy = fma(pp.hh, k.dvy, vy);
    vz = fma(pp.hh, k.dvz, vz);
    if (wrap_pos(g, x, y, z)) vy = -vy;                     // partbc, F:1337
    __stcs(P.x + t, x); __stcs(P.y + t, y); __stcs(P.z + t, z);
    __stcs(P.vx + t, vx); __stcs(P.vy + t, vy); __stcs(P.vz + t, vz);
    if (pp.drive_on) {
      if ((fabs(z - pp.zcent) < pp.zw) && ((fabs(y - pp.ycent2) < pp.yw) || (fabs(y - pp.ycent1) < pp.yw))) {
        const int id = P.id ? P.id[t] : (int)t;
        atomicOr(slab_bits + (id >> 5), 1u << (id & 31));
        slab_list[atomicAdd(slab_count, 1)] = (int)t;
      }
    }
  }
  block_wk_store(wx, wh, wk_partial);
}

// ---------------------------------------------------------------------------
// Deposit machinery (srimp1 F:2273-2374 + srimp2 F:2471-2529 fused: the
// 18-node stencil is shared; the four quantities qmult*vxj, qmult*vyj,
// qmult*vzj, qmult ride together).
//
// Value index n = g9*9 + r with g9 = jy*4 + m (m = moment 0..3), r = kz*3 + ix:
//   contribution c[n] = (qv[m]*fy[jy]) * (fx[ix]*fz[kz]) = qvy[g9] * wxz[r]
// Moment address of (g9, r) for stencil base node n0:
//   M4 + 4*(n0 + ix + jy*nx + kz*nxy) + m
// ---------------------------------------------------------------------------
__device__ __forceinline__ double sel(bool b, double t, double f) { return b ? t : f; }

__device__ __forceinline__ double* mom_addr(double* __restrict__ M4, const GP& g, int n0, int g9, int r) {
  const int jy = g9 >> 2, m = g9 & 3, kz = r / 3, ix = r - 3 * kz;
  return M4 + 4 * ((size_t)n0 + ix + (size_t)jy * g.nx + (size_t)kz * g.nxy) + m;
}

// MODE 0 baseline: one particle per lane straight to memory, 72 red.global.add.f64
__device__ __forceinline__ void deposit_direct72(const double qvy[8], const double wxz[9], int n0, const GP& g, double* __restrict__ M4) {
#pragma unroll
  for (int r = 0; r < 9; r++)
#pragma unroll
    for (int g9 = 0; g9 < 8; g9++) atomicAdd(mom_addr(M4, g, n0, g9, r), qvy[g9] * wxz[r]);
}

// transposing butterfly round: N values on every lane -> N/2 (the lane keeps
// the half selected by `bit` and receives the partner's partial sums for it)
template <int N>
__device__ __forceinline__ void tr_round(double* v, bool bit, int xormask) {
#pragma unroll
  for (int n = 0; n < N / 2; n++) {
    const double a = v[n], b = v[n + N / 2];
    const double send = sel(bit, a, b);
    const double keep = sel(bit, b, a);
    v[n] = keep + __shfl_xor_sync(0xffffffffu, send, xormask);
  }
}

// ---------------------------------------------------------------------------
// Predictor pass, ipc >= 1: F:1162-1283, 1300-1306, partbc F:1375, then the
// fused srimp1+srimp2 scatter.
//
// k_predict_direct: every particle issues its 72 atomics (baseline, MODE 0).
// ---------------------------------------------------------------------------
struct Predicted { double rx, ry, rz, vxj, vyj, vzj; };

__device__ __forceinline__ Predicted predict_one(const GP& g, const PushParams& pp, const ParticleSoA& P, long long t,
                                                 const double* __restrict__ F6, double& wx, double& wh) {
  const double x = __ldcs(P.x + t), y = __ldcs(P.y + t), z = __ldcs(P.z + t);
  const double vx = __ldcs(P.vx + t), vy = __ldcs(P.vy + t), vz = __ldcs(P.vz + t);
  double rx = __dadd_rn(x, __dmul_rn(pp.hdt, vx));        // F:1163-1165
  double ry = __dadd_rn(y, __dmul_rn(pp.hdt, vy));
  double rz = __dadd_rn(z, __dmul_rn(pp.hdt, vz));
  wrap_pos(g, rx, ry, rz);                                  // partbcEST, F:1168
  Stencil s;
  make_stencil<true>(g, rx, ry, rz, s);
  double f[6];
  gather6(F6, g, s, f);
  const Kick k = rotate(f, vx, vy, vz, pp.ht, pp.ht2);
  wx += k.wx; wh += k.wh;
  const double ah = pp.aimpl * pp.hh, hh2 = 0.5 * pp.hh;
  Predicted o;
  o.vxj = fma(ah, k.dvx, vx);                               // F:1300-1302
  o.vyj = fma(ah, k.dvy, vy);
  o.vzj = fma(ah, k.dvz, vz);
  o.rx = fma(pp.adt, fma(hh2, k.dvx, vx), x);               // F:1304-1306
  o.ry = fma(pp.adt, fma(hh2, k.dvy, vy), y);
  o.rz = fma(pp.adt, fma(hh2, k.dvz, vz), z);
  if (wrap_pos(g, o.rx, o.ry, o.rz)) o.vyj = -o.vyj;        // partbc, F:1375
  return o;
}

// scatter stencil factors of one predicted particle: key = base node,
// qvy[g9] = qmult*(vxj|vyj|vzj|1)*fy[jy], wxz[r] = fx[ix]*fz[kz]
__device__ __forceinline__ int scatter_factors(const GP& g, double qmult, const Predicted& o, double qvy[8], double wxz[9]) {
  Stencil s;
  make_stencil<false>(g, o.rx, o.ry, o.rz, s);              // F:2274-2308
#pragma unroll
  for (int jy = 0; jy < 2; jy++) {
    const double qf = qmult * s.fy[jy];
    qvy[jy * 4 + 0] = qf * o.vxj;                           // F:2311-2313
    qvy[jy * 4 + 1] = qf * o.vyj;
    qvy[jy * 4 + 2] = qf * o.vzj;
    qvy[jy * 4 + 3] = qf;                                   // F:2509
  }
#pragma unroll
  for (int kz = 0; kz < 3; kz++)
#pragma unroll
    for (int ix = 0; ix < 3; ix++) wxz[kz * 3 + ix] = s.fx[ix] * s.fz[kz];
  return s.n0;
}

__global__ void __launch_bounds__(128)
k_predict_direct(GP g, PushParams pp, ParticleSoA P, const double* __restrict__ F6, double* __restrict__ M4,
                 double* __restrict__ wk_partial) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double wx = 0.0, wh = 0.0;
  if (t < P.n) {
    const Predicted o = predict_one(g, pp, P, t, F6, wx, wh);
    double qvy[8], wxz[9];
    const int n0 = scatter_factors(g, pp.qmult, o, qvy, wxz);
    deposit_direct72(qvy, wxz, n0, g, M4);
  }
  block_wk_store(wx, wh, wk_partial);
}

// ---------------------------------------------------------------------------
// vmesh3 / vmesh1 (F:3243-3305, 3327-3377) as one gather per output element,
// fused with the AoS -> reference-layout unpack.  x and z steps ASSIGN, the y
// step ADDS (kept bug-compatible; needs mx,mz >= 4).  fold=0 only unpacks.
// ---------------------------------------------------------------------------
__global__ void k_fold_unpack(GP g, const double* __restrict__ M4, Ptr4 out, int fold) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= g.ntot) return;
  const int i = (int)(t % g.nx) - 2;
  const int j = (int)((t / g.nx) % g.ny) - 1;
  const int k = (int)(t / g.nxy) - 2;
  double v[4];
  if (!fold) {
    const double2* s = reinterpret_cast<const double2*>(M4) + t * 2;
    const double2 a = s[0], b = s[1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    const bool ii = (i >= 0 && i <= g.mx - 1);
    const bool jj = (j >= 0 && j <= g.my);
    int ks = k;                                   // z step, F:3287-3305
    if (ii && jj) {
      if (k == g.mz - 2 || k == g.mz - 1) ks = k - g.mz;
      else if (k == 0 || k == 1) ks = k + g.mz;
    }
    int is = i;                                   // x step, F:3243-3261 (all j,k)
    if (i == g.mx - 2 || i == g.mx - 1) is = i - g.mx;
    else if (i == 0 || i == 1) is = i + g.mx;
    const double2* s = reinterpret_cast<const double2*>(M4) + (size_t)node_of(g, is, j, ks) * 2;
    double2 a = s[0], b = s[1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    if (ii && (j == 0 || j == g.my)) {            // y step, F:3266-3282
      const int jg = (j == 0) ? -1 : g.my + 1;
      s = reinterpret_cast<const double2*>(M4) + (size_t)node_of(g, is, jg, ks) * 2;
      a = s[0]; b = s[1];
      v[0] = __dadd_rn(v[0], a.x); v[1] = __dadd_rn(v[1], a.y);
      v[2] = __dadd_rn(v[2], b.x); v[3] = __dadd_rn(v[3], b.y);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; c++) out.p[c][t] = v[c];
}

// ---------------------------------------------------------------------------
// Exclusive scan of int arrays (cell histogram, slab popcounts): three
// kernels, 2048 elements per block.
// ---------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

__device__ __forceinline__ int block_excl_scan(int v, int* total) {
  __shared__ int wsum[SCAN_BLOCK / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    int s = (lane < SCAN_BLOCK / 32) ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_BLOCK / 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += u;
    }
    if (lane < SCAN_BLOCK / 32) wsum[lane] = s;
  }
  __syncthreads();
  const int woff = (w == 0) ? 0 : wsum[w - 1];
  if (total) *total = wsum[SCAN_BLOCK / 32 - 1];
  __syncthreads();
  return woff + incl - v;
}

// pass 1: per-tile sums
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_reduce(const int* __restrict__ in, long long n, int* __restrict__ tile_sum) {
  const long long b0 = (long long)blockIdx.x * SCAN_TILE;
  int s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) {
    const long long t = b0 + q * SCAN_BLOCK + threadIdx.x;
    if (t < n) s += in[t];
  }
  int total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
// pass 2: scan of the tile sums by ONE block (sequential over chunks)
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_tiles(int* __restrict__ tile_sum, int ntiles, int* __restrict__ grand_total) {
  int carry = 0;
  for (int c0 = 0; c0 < ntiles; c0 += SCAN_BLOCK) {
    const int t = c0 + threadIdx.x;
    const int v = (t < ntiles) ? tile_sum[t] : 0;
    int total;
    const int e = block_excl_scan(v, &total);
    if (t < ntiles) tile_sum[t] = carry + e;
    carry += total;
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}
// pass 3: per-tile exclusive scan + tile offset
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(const int* __restrict__ in, long long n, const int* __restrict__ tile_off, int* __restrict__ out) {
  const long long b0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) { v[q] = (b0 + q < n) ? in[b0 + q] : 0; s += v[q]; }
  int e = block_excl_scan(s, nullptr) + tile_off[blockIdx.x];
#pragma unroll
  for (int q = 0; q < SCAN_ITEMS; q++) { if (b0 + q < n) out[b0 + q] = e; e += v[q]; }
}

// ---------------------------------------------------------------------------
// Cell sort (maintenance).  key = cell of wrap(x + lookahead*v), i fastest.
// ---------------------------------------------------------------------------
__global__ void k_sort_keys(GP g, ParticleSoA P, double lookahead, int* __restrict__ key, int* __restrict__ hist,
                            unsigned* __restrict__ zocc) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool valid = t < P.n;
  int kcell = -1;
  if (zocc) {   // planes the next pass gathers from (exact, unlike the key)
    int cache = -1;
    mark_plane(zocc, valid ? gather_plane(g, P.z[t], P.vz[t], lookahead) : 0, valid, cache);
  }
  if (valid) {
    double x = fma(lookahead, P.vx[t], P.x[t]);
    double y = fma(lookahead, P.vy[t], P.y[t]);
    double z = fma(lookahead, P.vz[t], P.z[t]);
    wrap_pos(g, x, y, z);
    kcell = sort_cell(g, x, y, z);
    key[t] = kcell;
  }
  // warp-aggregated histogram increment
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const unsigned m = __match_any_sync(act, kcell);
    if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(hist + kcell, __popc(m));
  }
}

struct SortArrays {
  const double* src[6]; double* dst[6];
  const int* id_src; int* id_dst;
  double* src_rw[6];     // the set being read, writable (corrector that updates in place)
};
__global__ void k_sort_scatter(long long n, const int* __restrict__ key, int* __restrict__ cursor, SortArrays A) {
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
  const int d = b + __popc(m & ((1u << lane) - 1u));
#pragma unroll
  for (int c = 0; c < 6; c++) A.dst[c][d] = A.src[c][t];
  A.id_dst[d] = A.id_src ? A.id_src[t] : (int)t;
}

// ---------------------------------------------------------------------------
// Self-checks (mrg_self_check): sums of the raw moments over the extended grid
// (srimp1/srimp2 weights are a partition of unity, F:2296-2308, so the raw q sums
// to qmult * N exactly up to rounding), and sum / sum of squares (mod 2^64) of the
// slots' original indices (a permutation of 0..n-1 after any number of sorts).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_moment_sums(const double* __restrict__ M4, long long ntot, double* __restrict__ out4) {
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  const double2* m = reinterpret_cast<const double2*>(M4);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < ntot; t += (long long)gridDim.x * blockDim.x) {
    const double2 u = m[2 * t], v = m[2 * t + 1];
    a[0] += u.x; a[1] += u.y; a[2] += v.x; a[3] += v.y;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const double w = warp_sum(a[c]);
    if ((threadIdx.x & 31) == 0) atomicAdd(out4 + c, w);
  }
}
__global__ void __launch_bounds__(256) k_id_sums(const int* __restrict__ id, long long n, unsigned long long* __restrict__ out2) {
  unsigned long long s1 = 0ull, s2 = 0ull;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const unsigned long long v = (unsigned long long)(unsigned)id[t];
    s1 += v; s2 += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out2, s1); atomicAdd(out2 + 1, s2); }
}

// Dependent-free DFMA stream for the second roofline (mrg_dfma_peak): 8 independent chains per thread.
__global__ void __launch_bounds__(256) k_dfma_peak(double* __restrict__ out, int iters, double a, double b) {
  double v[8];
#pragma unroll
  for (int q = 0; q < 8; q++) v[q] = (double)(threadIdx.x + q);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = fma(v[q], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; q++) s += v[q];
  if (s == 123.456) out[0] = s;      // never true: keeps the chains alive
}

// histogram of existing sort keys (mrg_sort, when the keys came without one)
__global__ void k_key_hist(long long n, const int* __restrict__ key, int* __restrict__ hist) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool valid = t < n;
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (!valid) return;
  const int kcell = key[t];
  const unsigned m = __match_any_sync(act, kcell);
  if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(hist + kcell, __popc(m));
}

// out[id[slot]] = in[slot]  (download in original order)
__global__ void k_unpermute(long long n, const int* __restrict__ id, const double* __restrict__ in, double* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) out[id[t]] = in[t];
}

// ---------------------------------------------------------------------------
// Drive kick, F:1342-1364.  The reference draws ranfp once for every owned
// particle inside the slab, in l order.  slab_bits holds one bit per original
// local index; word_off = exclusive popcount scan; the n-th slab particle (in
// l order) uses state*lambda^(n+1).
// ---------------------------------------------------------------------------
// z planes of a cell-sorted order that hold particles: bit k of occ[] is set when the cells of plane k own slots.
// `start` is the exclusive scan of the cell histogram (ncell + 1 entries), read BEFORE the scatter advances it.
__global__ void k_plane_occupancy(const int* __restrict__ start, int cells_per_plane, int mz, unsigned* __restrict__ occ) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= mz) return;
  if (start[(size_t)(k + 1) * cells_per_plane] > start[(size_t)k * cells_per_plane]) atomicOr(occ + (k >> 5), 1u << (k & 31));
}

// planes the next pass gathers from, for an order whose keys were not made by k_sort_keys
__global__ void k_mark_planes(GP g, ParticleSoA P, double lookahead, unsigned* __restrict__ zocc) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool valid = t < P.n;
  int cache = -1;
  mark_plane(zocc, valid ? gather_plane(g, P.z[t], P.vz[t], lookahead) : 0, valid, cache);
}

__global__ void k_popc(const unsigned* __restrict__ bits, long long nwords, int* __restrict__ out) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < nwords) out[t] = __popc(bits[t]);
}

// lambda^n mod 2^32 for n < 2^33 from three 2048-entry tables of lambda^(j * 2048^t) (built by the host): the kick
// needs one skip-ahead per slab particle, and the square-and-multiply loop of lcg_skip was most of k_kick.
__device__ __forceinline__ unsigned lcg_pow_tab(const unsigned* __restrict__ tab, unsigned long long n) {
  return __ldg(tab + (unsigned)(n & 2047ull)) * __ldg(tab + 2048 + (unsigned)((n >> 11) & 2047ull)) *
         __ldg(tab + 4096 + (unsigned)((n >> 22) & 2047ull));
}

__global__ void k_kick(GP g, ParticleSoA P, const double* __restrict__ F6, const unsigned* __restrict__ bits,
                       const int* __restrict__ word_off, const int* __restrict__ slab_list,
                       const int* __restrict__ slab_count, unsigned state0, double Ez00, double ycent1,
                       double ycent2, double yw2, const unsigned* __restrict__ lcg_tab) {
  const int count = *slab_count;   // read on the device: the host sizes the grid without knowing it
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const int slot = slab_list[e];
    const int id = P.id ? P.id[slot] : slot;
    const unsigned w = bits[id >> 5];
    const int rank = word_off[id >> 5] + __popc(w & ((1u << (id & 31)) - 1u));
    const unsigned ir = (lcg_pow_tab(lcg_tab, (unsigned long long)rank + 1ull) * state0) & 0x7fffffffu;   // = lcg_skip(state0, rank + 1)
    const double u = (double)ir * (1.0 / 2147483648.0);     // F:9302
    if (u > 0.999) {                                          // F:1353
      const double x = P.x[slot], y = P.y[slot], z = P.z[slot];
      int ip, jp, kp;
      cell_of(g, x, y, z, ip, jp, kp);                        // F:1347-1349
      const double bxa = F6[(size_t)node_of(g, ip, jp, kp) * 6 + 3];
      const double vy0 = __ddiv_rn(Ez00, bxa);                // F:1354
      if (fabs(y - ycent2) < yw2) P.vy[slot] = __dsub_rn(P.vy[slot], vy0);
      else if (fabs(y - ycent1) < yw2) P.vy[slot] = __dadd_rn(P.vy[slot], vy0);
    }
  }
}

// ---------------------------------------------------------------------------
// Synthetic two-flux-bundle load, F:8937-9040, for local index m <-> global
// l = first + m*stride (1-based).  Positions use ranfp draws 3(l-1)+1..3 from
// seed sb; velocities use ranf draws 4(l-1)+1..4 from seed sa.  Bit-exact
// with the serial loader (all products/sums use _rn intrinsics).
// ---------------------------------------------------------------------------
struct LoadParams {
  double fv2[101];
  double v2, dv2, vdr, vbeam;
  double half_hx, half_hz;
  double zcent, dzcent, dzsmt, ycent1, ycent2, dycent, rrz, rry;
  unsigned sa, sb;
  long long first, stride;
};
// one particle of the serial loader: l0 = l - 1 (global index), written to local slot m
__device__ __forceinline__ void loadpt_one(const GP& g, const LoadParams& L, const ParticleSoA& P, unsigned long long l0, long long m);

__global__ void k_loadpt(GP g, LoadParams L, ParticleSoA P) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= P.n) return;
  loadpt_one(g, L, P, (unsigned long long)(L.first - 1 + m * L.stride), m);
}

// z-slab ownership (option "shard" = 1): rank r owns the particles whose INITIAL z lies in slab r of nslab equal
// slabs of [-hz/2, zmax-hz/2).  z of particle l is the third ranfp draw, F:8949.
__device__ __forceinline__ bool loadpt_in_slab(const GP& g, const LoadParams& L, unsigned long long l0, int nslab, int slab) {
  unsigned s = lcg_skip(L.sb, 3ull * l0 + 2ull);
  s = lcg_next(s);
  const double z = __dsub_rn(__dmul_rn(g.zmax, (double)s * (1.0 / 2147483648.0)), L.half_hz);
  int q = (int)((z + L.half_hz) / g.zmax * nslab);
  q = min(max(q, 0), nslab - 1);
  return q == slab;
}
// pass 1: per-block (256 consecutive l) count of owned particles
__global__ void __launch_bounds__(256) k_loadpt_slab_count(GP g, LoadParams L, long long npr, int nslab, int slab, int* __restrict__ block_count) {
  const long long l0 = blockIdx.x * 256LL + threadIdx.x;
  const bool mine = l0 < npr && loadpt_in_slab(g, L, (unsigned long long)l0, nslab, slab);
  const int c = __syncthreads_count(mine);
  if (threadIdx.x == 0) block_count[blockIdx.x] = c;
}
// pass 2: owned particles of a block go to consecutive local slots in l order (block_off = exclusive scan of pass 1)
__global__ void __launch_bounds__(256) k_loadpt_slab_fill(GP g, LoadParams L, ParticleSoA P, long long npr, int nslab, int slab,
                                                          const int* __restrict__ block_off) {
  __shared__ int wcnt[8];
  const long long l0 = blockIdx.x * 256LL + threadIdx.x;
  const bool mine = l0 < npr && loadpt_in_slab(g, L, (unsigned long long)l0, nslab, slab);
  const unsigned b = __ballot_sync(0xffffffffu, mine);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) wcnt[w] = __popc(b);
  __syncthreads();
  int off = block_off[blockIdx.x];
  for (int q = 0; q < w; q++) off += wcnt[q];
  if (mine) loadpt_one(g, L, P, (unsigned long long)l0, off + __popc(b & ((1u << lane) - 1u)));
}

__device__ __forceinline__ void loadpt_one(const GP& g, const LoadParams& L, const ParticleSoA& P, unsigned long long l0, long long m) {
  const double inv = 1.0 / 2147483648.0;
  unsigned s = lcg_skip(L.sb, 3ull * l0);
  s = lcg_next(s); const double x = __dsub_rn(__dmul_rn(g.xmax, (double)s * inv), L.half_hx);  // F:8947
  s = lcg_next(s); const double y = __dmul_rn(g.ymax, (double)s * inv);                        // F:8948
  s = lcg_next(s); const double z = __dsub_rn(__dmul_rn(g.zmax, (double)s * inv), L.half_hz);  // F:8949
  unsigned a = lcg_skip(L.sa, 4ull * l0);
  a = lcg_next(a); const double eps = (double)a * inv;      // F:8977
  int k2 = 100;
  for (int k = 1; k <= 100; k++) { k2 = k; if (L.fv2[k - 1] > eps) break; }   // F:8979-8982
  const double y1 = L.fv2[k2 - 2], y2 = L.fv2[k2 - 1];
  const double x2 = __dadd_rn(__ddiv_rn(__dsub_rn(eps, y2), __dsub_rn(y2, y1)), (double)k2);   // F:8986
  const double vmag = __dadd_rn(__dadd_rn(L.v2, __dmul_rn(L.dv2, __dsub_rn(x2, 1.0))), L.vdr); // F:8988
  a = lcg_next(a); const double vxo = __dmul_rn(vmag, __dsub_rn((double)a * inv, 0.5));        // F:8993-8995
  a = lcg_next(a); const double vyo = __dmul_rn(vmag, __dsub_rn((double)a * inv, 0.5));
  a = lcg_next(a); const double vzo = __dmul_rn(vmag, __dsub_rn((double)a * inv, 0.5));
  double ycnt1, ycnt2;
  const double az = fabs(__dsub_rn(z, L.zcent));
  if (az < L.dzcent || az < L.dzsmt) {                       // F:9012-9018
    ycnt1 = __dadd_rn(L.ycent1, L.dycent); ycnt2 = __dsub_rn(L.ycent2, L.dycent);
  } else { ycnt1 = L.ycent1; ycnt2 = L.ycent2; }            // F:9020-9021
  double vdrift = 0.0;
  if (az <= L.rrz && (fabs(__dsub_rn(y, ycnt1)) <= L.rry || fabs(__dsub_rn(y, ycnt2)) <= L.rry)) vdrift = L.vbeam;  // F:9031-9035
  P.x[m] = x; P.y[m] = y; P.z[m] = z;
  P.vx[m] = __dadd_rn(vxo, vdrift); P.vy[m] = vyo; P.vz[m] = vzo;     // F:9037-9039
}

}  // namespace mrg
