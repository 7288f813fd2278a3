// Host-side emulation of the device helpers in csrc/mrg_device.cuh: the same
// header is compiled with g++ (CUDA intrinsics mapped to libm / fenv) so the
// index/weight/gather/rotation arithmetic can be checked against the CPU
// oracle without a GPU.  Test infrastructure only.
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#define __host__
#define __device__
#define __forceinline__ inline
#define __restrict__
struct double2 { double x, y; };
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dadd_rd(double a, double b) {
  std::fesetround(FE_DOWNWARD);
  volatile double va = a, vb = b;
  volatile double r = va + vb;
  std::fesetround(FE_TONEAREST);
  return r;
}
static inline int __double2loint(double d) { int64_t u; std::memcpy(&u, &d, 8); return (int)(uint32_t)(u & 0xffffffffu); }
static inline int __double2hiint(double d) { int64_t u; std::memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline long long __double_as_longlong(double d) { long long u; std::memcpy(&u, &d, 8); return u; }
static inline double2 __ldg(const double2* p) { return *p; }
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
using std::min; using std::max; using std::fma;
#define MRG_EMUL_NO_CUDA_RUNTIME
#include "mrg_device_emul.h"

using namespace mrg;

extern "C" {
// one corrector / predictor push of a single particle using the device helpers
void emul_push(const double* gp /*packed GP doubles*/, const int* gi, const double* F6, const double* part /*x,y,z,vx,vy,vz*/,
               double dt, double adt, double hdt, double aimpl, double qmult, double wmult, int ipc, double* out /*6*/,
               double* wk /*2*/, int* key, double* qvy_wxz /*17*/) {
  GP g;
  g.mx = gi[0]; g.my = gi[1]; g.mz = gi[2];
  g.nx = g.mx + 4; g.ny = g.my + 3; g.nz = g.mz + 4; g.nxy = g.nx * g.ny; g.ntot = (long long)g.nxy * g.nz;
  g.xmax = gp[0]; g.ymax = gp[1]; g.zmax = gp[2];
  g.hx = g.xmax / g.mx; g.hy = g.ymax / g.my; g.hz = g.zmax / g.mz;
  g.hxi = 0.9999999999999 / g.hx; g.hyi = 0.9999999999999 / g.hy; g.hzi = 0.9999999999999 / g.hz;
  g.xmaxe = 0.9999999999999 * g.xmax; g.zmaxe = 0.9999999999999 * g.zmax;
  g.xlo = -(g.hx / 2); g.xhi = g.xmax - g.hx / 2; g.zlo = -(g.hz / 2); g.zhi = g.zmax - g.hz / 2; g.ymax2 = 2.0 * g.ymax;
  double x = part[0], y = part[1], z = part[2], vx = part[3], vy = part[4], vz = part[5];
  const double hh = dt * qmult / wmult, ht = 0.5 * hh, ht2 = ht * ht;
  double rx = __dadd_rn(x, __dmul_rn(hdt, vx)), ry = __dadd_rn(y, __dmul_rn(hdt, vy)), rz = __dadd_rn(z, __dmul_rn(hdt, vz));
  wrap_pos(g, rx, ry, rz);
  Stencil s;
  make_stencil<true>(g, rx, ry, rz, s);
  double f[6];
  gather6(F6, g, s, f);
  Kick k = rotate(f, vx, vy, vz, ht, ht2);
  wk[0] = k.wx; wk[1] = k.wh;
  if (ipc == 0) {
    const double hh2 = 0.5 * hh;
    x = fma(dt, fma(hh2, k.dvx, vx), x); y = fma(dt, fma(hh2, k.dvy, vy), y); z = fma(dt, fma(hh2, k.dvz, vz), z);
    vx = fma(hh, k.dvx, vx); vy = fma(hh, k.dvy, vy); vz = fma(hh, k.dvz, vz);
    if (wrap_pos(g, x, y, z)) vy = -vy;
    out[0] = x; out[1] = y; out[2] = z; out[3] = vx; out[4] = vy; out[5] = vz;
  } else {
    const double ah = aimpl * hh, hh2 = 0.5 * hh;
    double vxj = fma(ah, k.dvx, vx), vyj = fma(ah, k.dvy, vy), vzj = fma(ah, k.dvz, vz);
    rx = fma(adt, fma(hh2, k.dvx, vx), x); ry = fma(adt, fma(hh2, k.dvy, vy), y); rz = fma(adt, fma(hh2, k.dvz, vz), z);
    if (wrap_pos(g, rx, ry, rz)) vyj = -vyj;
    out[0] = rx; out[1] = ry; out[2] = rz; out[3] = vxj; out[4] = vyj; out[5] = vzj;
    make_stencil<false>(g, rx, ry, rz, s);
    *key = s.n0;
    for (int jy = 0; jy < 2; jy++) {
      const double qf = qmult * s.fy[jy];
      qvy_wxz[jy * 4 + 0] = qf * vxj; qvy_wxz[jy * 4 + 1] = qf * vyj; qvy_wxz[jy * 4 + 2] = qf * vzj; qvy_wxz[jy * 4 + 3] = qf;
    }
    for (int kz = 0; kz < 3; kz++) for (int ix = 0; ix < 3; ix++) qvy_wxz[8 + kz * 3 + ix] = s.fx[ix] * s.fz[kz];
  }
}

// gather_plane (plane bookkeeping of the restricted preparation) against the kp make_stencil<true> finds for the
// same particle: returns gather_plane - stencil kp (must be 0)
int emul_gather_plane_diff(const double* gp, const int* gi, double x, double y, double z, double vx, double vy, double vz, double hdt,
                           int* exact, int* fast) {
  GP g;
  g.mx = gi[0]; g.my = gi[1]; g.mz = gi[2];
  g.nx = g.mx + 4; g.ny = g.my + 3; g.nz = g.mz + 4; g.nxy = g.nx * g.ny; g.ntot = (long long)g.nxy * g.nz;
  g.xmax = gp[0]; g.ymax = gp[1]; g.zmax = gp[2];
  g.hx = g.xmax / g.mx; g.hy = g.ymax / g.my; g.hz = g.zmax / g.mz;
  g.hxi = 0.9999999999999 / g.hx; g.hyi = 0.9999999999999 / g.hy; g.hzi = 0.9999999999999 / g.hz;
  g.xmaxe = 0.9999999999999 * g.xmax; g.zmaxe = 0.9999999999999 * g.zmax;
  g.xlo = -(g.hx / 2); g.xhi = g.xmax - g.hx / 2; g.zlo = -(g.hz / 2); g.zhi = g.zmax - g.hz / 2; g.ymax2 = 2.0 * g.ymax;
  double rx = __dadd_rn(x, __dmul_rn(hdt, vx)), ry = __dadd_rn(y, __dmul_rn(hdt, vy)), rz = __dadd_rn(z, __dmul_rn(hdt, vz));
  wrap_pos(g, rx, ry, rz);
  Stencil s;
  make_stencil<true>(g, rx, ry, rz, s);
  if (fast) *fast = gather_plane_fast(g, z, vz, hdt);
  if (exact) *exact = gather_plane(g, z, vz, hdt);
  return gather_plane(g, z, vz, hdt) - s.kp;
}
}
