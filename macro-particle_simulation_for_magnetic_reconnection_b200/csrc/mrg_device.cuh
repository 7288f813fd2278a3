// mrg_device.cuh -- device-side building blocks of the /fulmov/ path.
//
// F:n = /root/reference/@mrg37-080A.f03 line n (what each piece reproduces).
// Lines that decide an integer cell index or a branch use the _rn/_rd
// intrinsics so that no FMA contraction can change the decision; everything
// else is free to contract.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mrg {

// Grid constants (COMMON /parm2/,/ptable/ subset), passed by value to kernels.
struct GP {
  int mx, my, mz;
  int nx, ny, nz;          // extended sizes mx+4, my+3, mz+4        (F:1061)
  int nxy;                 // nx*ny
  long long ntot;          // mxyzA                                  (P:33)
  double xmax, ymax, zmax;
  double hx, hy, hz;       // F:8454,8467,8484
  double hxi, hyi, hzi;    // F:8567-8569
  double xmaxe, zmaxe;     // F:8575,8577
  double xlo, xhi;         // -hx/2, xmax-hx/2   (F:1856-1862)
  double zlo, zhi;
  double ymax2;            // 2*ymax             (F:1867)
  int xhi_h, xlo_h, ymax_h, zhi_h, zlo_h;   // high words of the partbc limits (maybe_wrap pre-test)
  int kz0, nkz;            // z planes of the sort order that hold particles: kz0 .. kz0+nkz-1 (mod mz); tiled launches cover only these
};

// node index of (i,j,k) in the (-2:mx+1,-1:my+1,-2:mz+1) layout
__host__ __device__ __forceinline__ int node_of(const GP& g, int i, int j, int k) {
  return (i + 2) + g.nx * ((j + 1) + g.ny * (k + 2));
}

#define MRG_TWO52 4503599627370496.0

// floor(s) for 0 <= s < 2^31 without F2I/I2F: adding 2^52 in round-down mode
// leaves floor(s) in the low mantissa bits.  Returns the integer and its
// exact double value.  (The reference truncates with int(), F:1175-1177; its
// arguments are >= 0 after partbc/partbcEST.)
__device__ __forceinline__ int floor_pos(double s, double& as_double) {
  double u = __dadd_rd(s, MRG_TWO52);
  as_double = __dsub_rn(u, MRG_TWO52);
  return __double2loint(u);
}

// partbc (F:1856-1879) / partbcEST (F:1928-1949): applied once, no loop.
// Returns true when the y wall reflected the particle (partbc then flips vy).
// The six comparisons are done on the IEEE bit patterns with integer
// instructions (the fp64 pipe is the bottleneck of the particle passes): for
// non-NaN x, `x >= a` with a > 0 is a signed compare of the patterns, `x <= b`
// with b < 0 an unsigned one, `y <= 0` is "pattern <= 0" (covers -0.0).  The
// decisions are identical to the floating-point compares of the source.
__device__ __forceinline__ bool ge_pos(double x, double a) { return __double_as_longlong(x) >= __double_as_longlong(a); }
__device__ __forceinline__ bool le_neg(double x, double b) {
  return (unsigned long long)__double_as_longlong(x) >= (unsigned long long)__double_as_longlong(b);
}
__device__ __forceinline__ bool wrap_pos(const GP& g, double& x, double& y, double& z) {
  const bool xge = ge_pos(x, g.xhi), xle = le_neg(x, g.xlo);
  x = __dadd_rn(x, xge ? -g.xmaxe : (xle ? g.xmaxe : 0.0));      // x - xmaxe | x + xmaxe | x
  const bool yge = ge_pos(y, g.ymax), yle = __double_as_longlong(y) <= 0ll;
  const bool flip = yge || yle;
  const double yr = __dsub_rn(yge ? g.ymax2 : 0.0, y);            // 2*ymax - y | -y
  y = flip ? yr : y;
  const bool zge = ge_pos(z, g.zhi), zle = le_neg(z, g.zlo);
  z = __dadd_rn(z, zge ? -g.zmaxe : (zle ? g.zmaxe : 0.0));
  return flip;
}

// ---------------------------------------------------------------------------
// Conservative "might need partbc" test on the high words (integer pipe).  True
// whenever any of the six comparisons of wrap_pos could be true; false
// positives only for coordinates within 2^-20 (relative) of a limit.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool maybe_wrap(const GP& g, double x, double y, double z) {
  const int xh = __double2hiint(x), yh = __double2hiint(y), zh = __double2hiint(z);
  return (xh >= g.xhi_h) | ((unsigned)xh >= (unsigned)g.xlo_h) | (yh >= g.ymax_h) | (yh <= 0) | (zh >= g.zhi_h) |
         ((unsigned)zh >= (unsigned)g.zlo_h);
}

// Cell index + weights.  GATHER=true: F:1175-1215, GATHER=false: F:2274-2308
// (the scatter does not override fyl/fyr in the jp>=my branch).
//   n0 = node of (il,jl,kl); the 18 nodes are n0 + ix + jy*nx + kz*nxy.
//   fx = (fxl,fxc,fxr), fz = (fzl,fzc,fzr), fy[0]=fyl (row jl), fy[1]=fyr.
struct Stencil {
  int n0;
  int ip, jp, kp;
  double fx[3], fy[2], fz[3];
};

template <bool GATHER>
__device__ __forceinline__ void make_stencil(const GP& g, double rx, double ry, double rz, Stencil& s) {
  const double tx = __dmul_rn(g.hxi, rx);
  const double ty = __dmul_rn(g.hyi, ry);
  const double tz = __dmul_rn(g.hzi, rz);
  double ipd, jpd, kpd;
  int ip = floor_pos(__dadd_rn(tx, 0.500000001), ipd);
  int jp = floor_pos(__dadd_rn(ty, 0.000000001), jpd);
  int kp = floor_pos(__dadd_rn(tz, 0.500000001), kpd);
  // Memory safety only: valid (wrapped) positions already satisfy these.
  ip = min(max(ip, 0), g.mx);
  jp = min(max(jp, 0), g.my);
  kp = min(max(kp, 0), g.mz);
  s.ip = ip; s.jp = jp; s.kp = kp;
  s.n0 = (ip + 1) + g.nx * ((jp + 1) + g.ny * (kp + 1));
  double fyl = __dsub_rn(ty, jpd);
  double fyr = __dsub_rn(1.0, fyl);
  if (GATHER && jp >= g.my) { fyr = 0.0; fyl = 1.0; }   // F:1191-1195
  s.fy[0] = fyl; s.fy[1] = fyr;
  const double xx = __dsub_rn(tx, ipd);
  const double xm = 0.5 - xx, xp = 0.5 + xx;
  s.fx[0] = (0.5 * xm) * xm;
  s.fx[1] = fma(-xx, xx, 0.75);
  s.fx[2] = (0.5 * xp) * xp;
  const double zz = __dsub_rn(tz, kpd);
  const double zm = 0.5 - zz, zp = 0.5 + zz;
  s.fz[0] = (0.5 * zm) * zm;
  s.fz[1] = fma(-zz, zz, 0.75);
  s.fz[2] = (0.5 * zp) * zp;
}

// Nearest-node indices only (drive kick, F:1347-1349; sort key).
__device__ __forceinline__ void cell_of(const GP& g, double x, double y, double z, int& ip, int& jp, int& kp) {
  double d;
  ip = floor_pos(__dadd_rn(__dmul_rn(g.hxi, x), 0.500000001), d);
  jp = floor_pos(__dadd_rn(__dmul_rn(g.hyi, y), 0.000000001), d);
  kp = floor_pos(__dadd_rn(__dmul_rn(g.hzi, z), 0.500000001), d);
  ip = min(max(ip, 0), g.mx);
  jp = min(max(jp, 0), g.my);
  kp = min(max(kp, 0), g.mz);
}

// Sort key: linear cell index (i fastest) of an already wrapped position.  A
// sorting hint only -- every result is independent of the particle order -- so
// it may contract to FMA (it can differ from the exact F:1175-1177 index only
// for positions within an ulp of a cell boundary).
__device__ __forceinline__ int sort_cell(const GP& g, double x, double y, double z) {
  int ip = __double2loint(__dadd_rd(fma(g.hxi, x, 0.500000001), MRG_TWO52));
  int jp = __double2loint(__dadd_rd(fma(g.hyi, y, 0.000000001), MRG_TWO52));
  int kp = __double2loint(__dadd_rd(fma(g.hzi, z, 0.500000001), MRG_TWO52));
  ip = min(max(ip, 0), g.mx - 1);
  jp = min(max(jp, 0), g.my - 1);
  kp = min(max(kp, 0), g.mz - 1);
  return ip + g.mx * (jp + g.my * kp);
}

// Sort key of an UNWRAPPED position: periodic / reflecting images are folded in index space (a sorting
// hint only, like sort_cell).
__device__ __forceinline__ int sort_cell_folded(const GP& g, double x, double y, double z) {
  int ip = __double2loint(__dadd_rd(fma(g.hxi, x, 0.500000001 + 65536.0), MRG_TWO52)) - 65536;
  int jp = __double2loint(__dadd_rd(fma(g.hyi, y, 0.000000001 + 65536.0), MRG_TWO52)) - 65536;
  int kp = __double2loint(__dadd_rd(fma(g.hzi, z, 0.500000001 + 65536.0), MRG_TWO52)) - 65536;
  // one periodic image either side in x and z, one mirror image at either wall in y -- with masks instead of branches
  // (the compiler turned the conditional form into three divergent regions per particle)
  ip += (ip >> 31) & g.mx;
  ip -= ((g.mx - 1 - ip) >> 31) & g.mx;
  kp += (kp >> 31) & g.mz;
  kp -= ((g.mz - 1 - kp) >> 31) & g.mz;
  jp ^= jp >> 31;                                        // jp < 0  -> -1 - jp
  jp += ((g.my - 1 - jp) >> 31) & (2 * g.my - 1 - 2 * jp);   // jp >= my -> 2 my - 1 - jp
  ip = min(max(ip, 0), g.mx - 1);
  jp = min(max(jp, 0), g.my - 1);
  kp = min(max(kp, 0), g.mz - 1);
  return ip + g.mx * (jp + g.my * kp);
}

// z plane kp of the gather cell of the NEXT pass over this particle: the exact z part of F:1165 (half-step
// estimate), partbcEST (F:1941-1947) and F:1177, with make_stencil's clamp.  The passes read the prepared
// fields on planes kp-1..kp+1 only, which lets a rank prepare just the planes near its particles.
__device__ __forceinline__ int gather_plane(const GP& g, double z, double vz, double hdt) {
  double rz = __dadd_rn(z, __dmul_rn(hdt, vz));
  const bool zge = ge_pos(rz, g.zhi), zle = le_neg(rz, g.zlo);
  rz = __dadd_rn(rz, zge ? -g.zmaxe : (zle ? g.zmaxe : 0.0));
  double d;
  const int kp = floor_pos(__dadd_rn(__dmul_rn(g.hzi, rz), 0.500000001), d);
  return min(max(kp, 0), g.mz);
}

// The same plane with contracted arithmetic and no wrap (three fp64 instructions instead of seven, no branch): for
// a wrapped z the result lies in [-1, mz] and, folded into [0, mz), is within one plane of gather_plane -- or
// gather_plane is the clamp value mz and the folded result is a plane next to the seam.  The corrector records
// this one; ensure_prep widens accordingly (add_occupancy in mrg_api.cu).
__device__ __forceinline__ int gather_plane_fast(const GP& g, double z, double vz, double hdt) {
  return __double2loint(__dadd_rd(fma(g.hzi, fma(hdt, vz, z), 0.500000001 + 65536.0), MRG_TWO52)) - 65536;
}

// Gather of the six prepared fields (F:1217-1270) from the packed array
// F6[node][6] = (exa,eya,eza,bxa,bya,bza): 9 x 128-bit loads per stencil row.
// The 18 weights are formed as fx*(fy*fz); the sum order differs from the
// Fortran nesting by rounding only.
__device__ __forceinline__ void gather6(const double* __restrict__ F6, const GP& g, const Stencil& s, double f[6]) {
#pragma unroll
  for (int c = 0; c < 6; c++) f[c] = 0.0;
  const double2* base = reinterpret_cast<const double2*>(F6) + (size_t)s.n0 * 3;
  const int sy = g.nx * 3;
  const size_t sz = (size_t)g.nxy * 3;
#pragma unroll
  for (int kz = 0; kz < 3; kz++) {
#pragma unroll
    for (int jy = 0; jy < 2; jy++) {
      const double2* r = base + jy * sy + kz * sz;
      const double wyz = s.fy[jy] * s.fz[kz];
      double2 v[9];
#pragma unroll
      for (int q = 0; q < 9; q++) v[q] = __ldg(r + q);
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

// Closed-form implicit rotation, F:1272-1283.  One reciprocal instead of the
// three divisions of the source.
struct Kick {
  double dvx, dvy, dvz, wx, wh;
};
__device__ __forceinline__ Kick rotate(const double f[6], double vx, double vy, double vz, double ht, double ht2) {
  const double exi = f[0], eyi = f[1], ezi = f[2], bxi = f[3], byi = f[4], bzi = f[5];
  const double bsqi = fma(bxi, bxi, fma(byi, byi, bzi * bzi));
  const double acx = exi + (vy * bzi - vz * byi);
  const double acy = eyi + (vz * bxi - vx * bzi);
  const double acz = ezi + (vx * byi - vy * bxi);
  const double ach = fma(exi, bxi, fma(eyi, byi, ezi * bzi));
  const double rden = 1.0 / fma(ht2, bsqi, 1.0);
  const double t = ht2 * ach;
  Kick k;
  k.dvx = (acx + fma(t, bxi, ht * (acy * bzi - acz * byi))) * rden;
  k.dvy = (acy + fma(t, byi, ht * (acz * bxi - acx * bzi))) * rden;
  k.dvz = (acz + fma(t, bzi, ht * (acx * byi - acy * bxi))) * rden;
  k.wx = 0.5 * fma(acx, acx, fma(acy, acy, acz * acz));
  k.wh = 0.5 * (ach * ach);
  return k;
}

// 31-bit multiplicative LCG of ranf/ranfp (F:9263-9305): state*lambda^n.
__host__ __device__ __forceinline__ uint32_t lcg_skip(uint32_t state, unsigned long long n) {
  uint32_t base = 48828125u, acc = 1u;
  while (n) {
    if (n & 1ull) acc *= base;
    base *= base;
    n >>= 1;
  }
  return (acc * state) & 0x7fffffffu;
}
__host__ __device__ __forceinline__ uint32_t lcg_next(uint32_t s) { return (48828125u * s) & 0x7fffffffu; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mrg
