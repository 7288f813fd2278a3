/*
 * fulmov_oracle.c -- CPU restatement (plain C99) of the reference's /fulmov/
 * particle hot path.  TEST INFRASTRUCTURE ONLY; see fulmov_oracle.h.
 * PARITY UNPINNED by the reference's own tests (it has none) -- see header.
 *
 * Every routine cites the reference lines it follows (F:n =
 * /root/reference/@mrg37-080A.f03 line n).  Expression association follows
 * the Fortran source (left to right, explicit parentheses kept); build with
 * -O2 -ffp-contract=off so no FMA contraction changes the rounding.
 */
#include "fulmov_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* (-2:mx+1,-1:my+1,-2:mz+1), i fastest: F:1061 */
#define NX(p) ((int64_t)(p)->mx + 4)
#define NY(p) ((int64_t)(p)->my + 3)
#define NZ(p) ((int64_t)(p)->mz + 4)
#define IDX(p, i, j, k) \
  (((int64_t)(i) + 2) + NX(p) * (((int64_t)(j) + 1) + NY(p) * ((int64_t)(k) + 2)))

int64_t orc_mxyzA(const orc_parm* p) { return NX(p) * NY(p) * NZ(p); }

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline legs of bench.py set the count explicitly */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* F:8454-8484 (hx,hy,hz), F:8567-8580 (hxi.., xmaxe.., adt, hdt),
 * F:8601-8603 (bxc), F:9001-9006 (zcent, ycent1, ycent2), F:368-370 (ifil). */
void orc_parm_init(orc_parm* p, int mx, int my, int mz, double xmax,
                   double ymax, double zmax, double dt, double aimpl,
                   double wce_by_wpe, double Ez00) {
  memset(p, 0, sizeof(*p));
  p->mx = mx; p->my = my; p->mz = mz;
  p->ifilx = 1; p->ifily = 1; p->ifilz = 1;
  p->xmax = xmax; p->ymax = ymax; p->zmax = zmax;
  p->hx = xmax / mx; p->hy = ymax / my; p->hz = zmax / mz;
  p->hxi = 0.9999999999999 / p->hx;
  p->hyi = 0.9999999999999 / p->hy;
  p->hzi = 0.9999999999999 / p->hz;
  p->xmaxe = 0.9999999999999 * xmax;
  p->ymaxe = 0.9999999999999 * ymax;
  p->zmaxe = 0.9999999999999 * zmax;
  p->dt = dt; p->aimpl = aimpl;
  p->adt = aimpl * dt;
  p->hdt = 0.5 * dt;
  p->bxc = wce_by_wpe; p->byc = 0.0; p->bzc = 0.0;
  p->Ez00 = Ez00;
  p->zcent = 0.50 * zmax;
  p->ycent1 = 0.30 * ymax;
  p->ycent2 = 0.70 * ymax;
}

/* ------------------------------------------------------------------------ */
/* F:9263-9305: ir = iand(lambda*ir, 2^31-1) with int32 wrap; value ir*2^-31 */
static int32_t lcg_next(int32_t ir) {
  uint32_t prod = (uint32_t)48828125u * (uint32_t)ir; /* wraps mod 2^32 */
  return (int32_t)(prod & 0x7fffffffu);
}
double orc_ranf(int32_t* state) {
  *state = lcg_next(*state);
  return (double)(*state) * (1.0 / 2147483648.0);
}
double orc_ranfp(int32_t* state) {
  *state = lcg_next(*state);
  return (double)(*state) * (1.0 / 2147483648.0);
}
int32_t orc_lcg_skip(int32_t state, uint64_t n) {
  uint32_t base = 48828125u, acc = 1u;
  while (n) {
    if (n & 1u) acc *= base;
    base *= base;
    n >>= 1;
  }
  return (int32_t)((acc * (uint32_t)state) & 0x7fffffffu);
}

/* ------------------------------------------------------------------------ */
/* outmesh3, F:3088-3148 */
static void outmesh1_one(const orc_parm* p, double* a) {
  const int mx = p->mx, my = p->my, mz = p->mz;
  for (int i = -2; i <= -1; i++)                       /* F:3088-3096 */
    for (int j = 0; j <= my; j++)
      for (int k = 0; k <= mz - 1; k++) a[IDX(p, i, j, k)] = a[IDX(p, i + mx, j, k)];
  for (int i = mx; i <= mx + 1; i++)                   /* F:3098-3106 */
    for (int j = 0; j <= my; j++)
      for (int k = 0; k <= mz - 1; k++) a[IDX(p, i, j, k)] = a[IDX(p, i - mx, j, k)];
  for (int k = 0; k <= mz - 1; k++)                    /* F:3110-3126 */
    for (int i = -2; i <= mx + 1; i++) {
      a[IDX(p, i, -1, k)] = 0.0;
      a[IDX(p, i, my + 1, k)] = 0.0;
    }
  for (int k = -2; k <= -1; k++)                       /* F:3130-3138 */
    for (int j = -1; j <= my + 1; j++)
      for (int i = -2; i <= mx + 1; i++) a[IDX(p, i, j, k)] = a[IDX(p, i, j, k + mz)];
  for (int k = mz; k <= mz + 1; k++)                   /* F:3140-3148 */
    for (int j = -1; j <= my + 1; j++)
      for (int i = -2; i <= mx + 1; i++) a[IDX(p, i, j, k)] = a[IDX(p, i, j, k - mz)];
}
void orc_outmesh3(const orc_parm* p, double* ax, double* ay, double* az) {
  outmesh1_one(p, ax); outmesh1_one(p, ay); outmesh1_one(p, az);
}

/* vmesh1, F:3327-3377 (vmesh3, F:3243-3305, is the same on three arrays).
 * x and z steps ASSIGN, the y step ADDS. */
void orc_vmesh1(const orc_parm* p, double* a) {
  const int mx = p->mx, my = p->my, mz = p->mz;
  for (int i = -2; i <= -1; i++)                       /* F:3327-3333 */
    for (int j = -1; j <= my + 1; j++)
      for (int k = -2; k <= mz + 1; k++) a[IDX(p, mx + i, j, k)] = a[IDX(p, i, j, k)];
  for (int i = mx; i <= mx + 1; i++)                   /* F:3335-3341 */
    for (int j = -1; j <= my + 1; j++)
      for (int k = -2; k <= mz + 1; k++) a[IDX(p, i - mx, j, k)] = a[IDX(p, i, j, k)];
  for (int k = -2; k <= mz + 1; k++)                   /* F:3346-3351 */
    for (int i = 0; i <= mx - 1; i++)
      a[IDX(p, i, 0, k)] = a[IDX(p, i, 0, k)] + a[IDX(p, i, -1, k)];
  for (int k = -2; k <= mz + 1; k++)                   /* F:3353-3358 */
    for (int i = 0; i <= mx - 1; i++)
      a[IDX(p, i, my, k)] = a[IDX(p, i, my, k)] + a[IDX(p, i, my + 1, k)];
  for (int k = -2; k <= -1; k++)                       /* F:3363-3369 */
    for (int j = 0; j <= my; j++)
      for (int i = 0; i <= mx - 1; i++) a[IDX(p, i, j, mz + k)] = a[IDX(p, i, j, k)];
  for (int k = mz; k <= mz + 1; k++)                   /* F:3371-3377 */
    for (int j = 0; j <= my; j++)
      for (int i = 0; i <= mx - 1; i++) a[IDX(p, i, j, k - mz)] = a[IDX(p, i, j, k)];
}
void orc_vmesh3(const orc_parm* p, double* ax, double* ay, double* az) {
  orc_vmesh1(p, ax); orc_vmesh1(p, ay); orc_vmesh1(p, az);
}

/* Periodic neighbour tables on interior indices, F:8341-8348, 8399-8406. */
static int per_l(int i, int m) { return i == 0 ? m - 1 : i - 1; }
static int per_r(int i, int m) { return i == m - 1 ? 0 : i + 1; }

/* filt3e, F:7351-7506.  s = scratch of mxyzA doubles per component. */
void orc_filt3e(const orc_parm* p, double* ex, double* ey, double* ez,
                double exc, double eyc, double ezc, int ifilx, int ifily,
                int ifilz, int sym) {
  const int mx = p->mx, my = p->my, mz = p->mz;
  const int64_t n = orc_mxyzA(p);
  double* e[3] = {ex, ey, ez};
  const double dc[3] = {exc, eyc, ezc};
  /* y-mirror sign: x,z components +sym, y component -sym (F:7458-7460) */
  const double ysgn[3] = {(double)sym, -(double)sym, (double)sym};
  double* a[3];
  for (int c = 0; c < 3; c++) a[c] = (double*)malloc((size_t)n * sizeof(double));

  for (int c = 0; c < 3; c++)                          /* F:7351-7359 */
    for (int k = 0; k <= mz - 1; k++)
      for (int j = 0; j <= my; j++)
        for (int i = 0; i <= mx - 1; i++) e[c][IDX(p, i, j, k)] = e[c][IDX(p, i, j, k)] - dc[c];

  for (int ntz = 1; ntz <= ifilz; ntz++) {             /* F:7365-7395 */
    for (int c = 0; c < 3; c++) {
      for (int k = 0; k <= mz - 1; k++)
        for (int j = 0; j <= my; j++)
          for (int i = 0; i <= mx - 1; i++) a[c][IDX(p, i, j, k)] = e[c][IDX(p, i, j, k)];
      for (int k = 0; k <= mz - 1; k++)
        for (int j = 0; j <= my; j++)
          for (int i = 0; i <= mx - 1; i++) {
            int kr = per_r(k, mz), kl = per_l(k, mz);
            int krr = per_r(kr, mz), kll = per_l(kl, mz);
            e[c][IDX(p, i, j, k)] =
                -0.0625 * a[c][IDX(p, i, j, krr)] + 0.25 * a[c][IDX(p, i, j, kr)] +
                0.625 * a[c][IDX(p, i, j, k)] + 0.25 * a[c][IDX(p, i, j, kl)] -
                0.0625 * a[c][IDX(p, i, j, kll)];
          }
    }
  }
  for (int ntx = 1; ntx <= ifilx; ntx++) {             /* F:7401-7434 */
    for (int c = 0; c < 3; c++) {
      for (int k = 0; k <= mz - 1; k++)
        for (int j = 0; j <= my; j++)
          for (int i = 0; i <= mx - 1; i++) a[c][IDX(p, i, j, k)] = e[c][IDX(p, i, j, k)];
      for (int k = 0; k <= mz - 1; k++)
        for (int j = 0; j <= my; j++)
          for (int i = 0; i <= mx - 1; i++) {
            int ir = per_r(i, mx), il = per_l(i, mx);
            int irr = per_r(ir, mx), ill = per_l(il, mx);
            e[c][IDX(p, i, j, k)] =
                -0.0625 * a[c][IDX(p, ill, j, k)] + 0.25 * a[c][IDX(p, il, j, k)] +
                0.625 * a[c][IDX(p, i, j, k)] + 0.25 * a[c][IDX(p, ir, j, k)] -
                0.0625 * a[c][IDX(p, irr, j, k)];
          }
    }
  }
  for (int nty = 1; nty <= ifily; nty++) {             /* F:7438-7492 */
    for (int c = 0; c < 3; c++) {
      for (int k = 0; k <= mz - 1; k++)
        for (int j = 0; j <= my; j++)
          for (int i = 0; i <= mx - 1; i++) a[c][IDX(p, i, j, k)] = e[c][IDX(p, i, j, k)];
      for (int k = 0; k <= mz - 1; k++)                /* mirror rows, js=1 */
        for (int i = 0; i <= mx - 1; i++) {
          a[c][IDX(p, i, -1, k)] = ysgn[c] * e[c][IDX(p, i, 1, k)];
          a[c][IDX(p, i, my + 1, k)] = ysgn[c] * e[c][IDX(p, i, my - 1, k)];
        }
      for (int k = 0; k <= mz - 1; k++)
        for (int j = 1; j <= my - 1; j++)
          for (int i = 0; i <= mx - 1; i++)
            e[c][IDX(p, i, j, k)] =
                -0.0625 * a[c][IDX(p, i, j + 2, k)] + 0.25 * a[c][IDX(p, i, j + 1, k)] +
                0.625 * a[c][IDX(p, i, j, k)] + 0.25 * a[c][IDX(p, i, j - 1, k)] -
                0.0625 * a[c][IDX(p, i, j - 2, k)];
    }
  }
  for (int c = 0; c < 3; c++)                          /* F:7498-7506 */
    for (int k = 0; k <= mz - 1; k++)
      for (int j = 0; j <= my; j++)
        for (int i = 0; i <= mx - 1; i++) e[c][IDX(p, i, j, k)] = e[c][IDX(p, i, j, k)] + dc[c];
  for (int c = 0; c < 3; c++) free(a[c]);
}

/* F:1127-1148 */
void orc_field_prep(const orc_parm* p, const double* const f12[12],
                    double* const a6[6]) {
  const int mx = p->mx, my = p->my, mz = p->mz;
  const double aimpl = p->aimpl;
  const double dc[6] = {0.0, 0.0, 0.0, p->bxc, p->byc, p->bzc};
  const int64_t n = orc_mxyzA(p);
  /* The Fortran temporaries are uninitialised automatic arrays; outmesh3
   * defines every extended element before use.  NaN-fill to prove that. */
  for (int c = 0; c < 6; c++)
    for (int64_t m = 0; m < n; m++) a6[c][m] = NAN;
  for (int k = 0; k <= mz - 1; k++)
    for (int j = 0; j <= my; j++)
      for (int i = 0; i <= mx - 1; i++) {
        int64_t m = IDX(p, i, j, k);
        for (int c = 0; c < 3; c++)                    /* F:1130-1132 */
          a6[c][m] = aimpl * f12[c][m] + (1.0 - aimpl) * f12[c + 6][m];
        for (int c = 3; c < 6; c++)                    /* F:1134-1136 */
          a6[c][m] = aimpl * f12[c][m] + (1.0 - aimpl) * f12[c + 6][m] + dc[c];
      }
  orc_outmesh3(p, a6[0], a6[1], a6[2]);                /* F:1141 */
  orc_outmesh3(p, a6[3], a6[4], a6[5]);                /* F:1142 */
  orc_filt3e(p, a6[0], a6[1], a6[2], 0.0, 0.0, 0.0, p->ifilx, p->ifily, p->ifilz, -1);
  orc_filt3e(p, a6[3], a6[4], a6[5], p->bxc, p->byc, p->bzc, p->ifilx, p->ifily, p->ifilz, +1);
}

/* entry prefld of emfild, F:3820-3873: the magnetic field predicted from the time-decentred electric field,
 * b = b0 + dt * (-curl ea), written into f12[3..5] on the interior nodes.  The y walls (j = 0, my) take one-sided
 * differences with the mirror rows folded in (the factor 2) and by = 0.  hx2 = 2 hx etc. (F:8563-8565), the periodic
 * neighbours come from the tables pxl/pxr, pzl/pzr (F:8341-8364, 8399-8422).  SURVEY 8(f1): the first piece of the
 * field-side assembly that reads and writes only what the particle path already holds on the device. */
void orc_prefld(const orc_parm* p, double* const f12[12]) {
  const int mx = p->mx, my = p->my, mz = p->mz;
  const double aimpl = p->aimpl, dt = p->dt;
  const double hx2 = 2.0 * p->hx, hy2 = 2.0 * p->hy, hz2 = 2.0 * p->hz;
  const int64_t n = orc_mxyzA(p);
  double* ea[3];
  for (int c = 0; c < 3; c++) {
    ea[c] = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t m = 0; m < n; m++) ea[c][m] = NAN;           /* only interior nodes are defined and read */
  }
  for (int k = 0; k <= mz - 1; k++)                           /* F:3823-3831 */
    for (int j = 0; j <= my; j++)
      for (int i = 0; i <= mx - 1; i++) {
        const int64_t m = IDX(p, i, j, k);
        for (int c = 0; c < 3; c++) ea[c][m] = aimpl * f12[c][m] + (1.0 - aimpl) * f12[c + 6][m];
      }
  const double *exa = ea[0], *eya = ea[1], *eza = ea[2];
  double *bx = f12[3], *by = f12[4], *bz = f12[5];
  const double *bx0 = f12[9], *by0 = f12[10], *bz0 = f12[11];
  for (int k = 0; k <= mz - 1; k++) {
    const int kr = (k == mz - 1) ? 0 : k + 1, kl = (k == 0) ? mz - 1 : k - 1;
    for (int i = 0; i <= mx - 1; i++) {
      const int ir = (i == mx - 1) ? 0 : i + 1, il = (i == 0) ? mx - 1 : i - 1;
      for (int j = 1; j <= my - 1; j++) {                     /* F:3834-3853 */
        const int64_t m = IDX(p, i, j, k);
        bx[m] = bx0[m] + dt * ((eya[IDX(p, i, j, kr)] - eya[IDX(p, i, j, kl)]) / hz2 - (eza[IDX(p, i, j + 1, k)] - eza[IDX(p, i, j - 1, k)]) / hy2);
        by[m] = by0[m] + dt * ((eza[IDX(p, ir, j, k)] - eza[IDX(p, il, j, k)]) / hx2 - (exa[IDX(p, i, j, kr)] - exa[IDX(p, i, j, kl)]) / hz2);
        bz[m] = bz0[m] + dt * ((exa[IDX(p, i, j + 1, k)] - exa[IDX(p, i, j - 1, k)]) / hy2 - (eya[IDX(p, ir, j, k)] - eya[IDX(p, il, j, k)]) / hx2);
      }
      {                                                       /* F:3856-3880 */
        const int64_t m0 = IDX(p, i, 0, k), m1 = IDX(p, i, my, k);
        bx[m0] = bx0[m0] + dt * ((eya[IDX(p, i, 0, kr)] - eya[IDX(p, i, 0, kl)]) / hz2 - 2 * eza[IDX(p, i, 1, k)] / hy2);
        by[m0] = 0;
        bz[m0] = bz0[m0] + dt * (-(eya[IDX(p, ir, 0, k)] - eya[IDX(p, il, 0, k)]) / hx2 + 2 * exa[IDX(p, i, 1, k)] / hy2);
        bx[m1] = bx0[m1] + dt * ((eya[IDX(p, i, my, kr)] - eya[IDX(p, i, my, kl)]) / hz2 + 2 * eza[IDX(p, i, my, k)] / hy2);
        by[m1] = 0;
        bz[m1] = bz0[m1] + dt * (-(eya[IDX(p, ir, my, k)] - eya[IDX(p, il, my, k)]) / hx2 - 2 * exa[IDX(p, i, my, k)] / hy2);
      }
    }
  }
  for (int c = 0; c < 3; c++) free(ea[c]);
}

/* The magnetic field emfild leaves behind its solve, F:4238-4302: the same update as prefld from the NEW electric field,
 * and on the steps with mod(it,5) = 1 the smoothing outmesh3 + filt3e(sym = +1, no dc) of bx,by,bz (F:4298-4302).  The
 * electric field is whatever cfpsol (and, on those steps, its own smoothing F:4225-4229) left in f12[0..2]. */
void orc_update_b(const orc_parm* p, double* const f12[12], int smooth) {
  orc_prefld(p, f12);
  if (smooth) {
    orc_outmesh3(p, f12[3], f12[4], f12[5]);
    orc_filt3e(p, f12[3], f12[4], f12[5], 0.0, 0.0, 0.0, p->ifilx, p->ifily, p->ifilz, +1);
  }
}

/* ------------------------------------------------------------------------ */
/* partbc F:1856-1879 (vy != NULL) / partbcEST F:1928-1949 (vy == NULL) */
static void wrap_one(const orc_parm* p, double* x, double* y, double* z, double* vy) {
  const double dx = p->hx / 2, dz = p->hz / 2;
  if (*x >= p->xmax - dx) *x = *x - p->xmaxe;
  else if (*x <= -dx) *x = *x + p->xmaxe;
  if (*y >= p->ymax) {
    *y = 2.0 * p->ymax - *y;
    if (vy) *vy = -*vy;
  } else if (*y <= 0.0) {
    *y = -*y;
    if (vy) *vy = -*vy;
  }
  if (*z >= p->zmax - dz) *z = *z - p->zmaxe;
  else if (*z <= -dz) *z = *z + p->zmaxe;
}
void orc_partbc(const orc_parm* p, double* x, double* y, double* z, double* vy,
                int64_t npr, int64_t first, int64_t stride) {
  for (int64_t l = first; l <= npr; l += stride)
    wrap_one(p, &x[l - 1], &y[l - 1], &z[l - 1], &vy[l - 1]);
}
void orc_partbcEST(const orc_parm* p, double* x, double* y, double* z,
                   int64_t npr, int64_t first, int64_t stride) {
  for (int64_t l = first; l <= npr; l += stride)
    wrap_one(p, &x[l - 1], &y[l - 1], &z[l - 1], NULL);
}

/* Cell index and weights.  gather_mode=1: F:1175-1215 (fyl/fyr overridden in
 * the edge branches); gather_mode=0: F:2274-2308 (they are not). */
typedef struct {
  int il, i, ir, jl, jr, kl, k, kr;
  double fxl, fxc, fxr, fyl, fyr, fzl, fzc, fzr;
} stencil;

static void index_weights(const orc_parm* p, double rx, double ry, double rz,
                          int gather_mode, stencil* s) {
  int ip = (int)(p->hxi * rx + 0.500000001);
  int jp = (int)(p->hyi * ry + 0.000000001);
  int kp = (int)(p->hzi * rz + 0.500000001);
  s->il = ip - 1; s->i = ip; s->ir = ip + 1;
  s->jl = jp; s->jr = jp + 1;
  s->fyl = p->hyi * ry - jp;
  s->fyr = 1.0 - s->fyl;
  if (jp >= p->my) {
    s->jr = p->my + 1; s->jl = p->my;
    if (gather_mode) { s->fyr = 0.0; s->fyl = 1.0; }
  } else if (jp < 0) {
    s->jr = 0; s->jl = -1;
    if (gather_mode) { s->fyr = 1.0; s->fyl = 0.0; }
  }
  s->kl = kp - 1; s->k = kp; s->kr = kp + 1;
  double xx = p->hxi * rx - ip;
  s->fxl = 0.5 * (0.5 - xx) * (0.5 - xx);
  s->fxc = 0.75 - xx * xx;
  s->fxr = 0.5 * (0.5 + xx) * (0.5 + xx);
  double zz = p->hzi * rz - kp;
  s->fzl = 0.5 * (0.5 - zz) * (0.5 - zz);
  s->fzc = 0.75 - zz * zz;
  s->fzr = 0.5 * (0.5 + zz) * (0.5 + zz);
}

/* F:1217-1224 (same form for all six fields up to F:1270) */
static double gather_one(const orc_parm* p, const double* a, const stencil* s) {
#define A(i, j, k) a[IDX(p, i, j, k)]
  return s->fyr *
             ((A(s->ir, s->jr, s->kr) * s->fxr + A(s->i, s->jr, s->kr) * s->fxc + A(s->il, s->jr, s->kr) * s->fxl) * s->fzr +
              (A(s->ir, s->jr, s->k) * s->fxr + A(s->i, s->jr, s->k) * s->fxc + A(s->il, s->jr, s->k) * s->fxl) * s->fzc +
              (A(s->ir, s->jr, s->kl) * s->fxr + A(s->i, s->jr, s->kl) * s->fxc + A(s->il, s->jr, s->kl) * s->fxl) * s->fzl) +
         s->fyl *
             ((A(s->ir, s->jl, s->kr) * s->fxr + A(s->i, s->jl, s->kr) * s->fxc + A(s->il, s->jl, s->kr) * s->fxl) * s->fzr +
              (A(s->ir, s->jl, s->k) * s->fxr + A(s->i, s->jl, s->k) * s->fxc + A(s->il, s->jl, s->k) * s->fxl) * s->fzc +
              (A(s->ir, s->jl, s->kl) * s->fxr + A(s->i, s->jl, s->kl) * s->fxc + A(s->il, s->jl, s->kl) * s->fxl) * s->fzl);
#undef A
}

/* 18-node scatter of one quantity, F:2315-2333 (order of the products kept:
 * qgam*fx*fy*fz evaluated left to right). */
static void scatter_one(const orc_parm* p, double* q, const stencil* s, double qgam) {
  const int ii[3] = {s->il, s->i, s->ir};
  const double fx[3] = {s->fxl, s->fxc, s->fxr};
  const int kk[3] = {s->kl, s->k, s->kr};
  const double fz[3] = {s->fzl, s->fzc, s->fzr};
  const int jj[2] = {s->jl, s->jr};
  const double fy[2] = {s->fyl, s->fyr};
  for (int b = 0; b < 2; b++)
    for (int a = 0; a < 3; a++)
      for (int c = 0; c < 3; c++) {
        int64_t m = IDX(p, ii[a], jj[b], kk[c]);
        q[m] = q[m] + qgam * fx[a] * fy[b] * fz[c];
      }
}

void orc_srimp1_scatter(const orc_parm* p, const double* rx, const double* ry,
                        const double* rz, const double* vxj, const double* vyj,
                        const double* vzj, double qmult, double* qjx,
                        double* qjy, double* qjz, int64_t npr, int64_t first,
                        int64_t stride) {
  for (int64_t l = first; l <= npr; l += stride) {     /* F:2273 */
    stencil s;
    index_weights(p, rx[l - 1], ry[l - 1], rz[l - 1], 0, &s);
    double qq = qmult;                                 /* F:2310-2313 */
    scatter_one(p, qjx, &s, qq * vxj[l - 1]);
    scatter_one(p, qjy, &s, qq * vyj[l - 1]);
    scatter_one(p, qjz, &s, qq * vzj[l - 1]);
  }
}
void orc_srimp2_scatter(const orc_parm* p, const double* rx, const double* ry,
                        const double* rz, double qmult, double* q, int64_t npr,
                        int64_t first, int64_t stride) {
  for (int64_t l = first; l <= npr; l += stride) {     /* F:2471 */
    stencil s;
    index_weights(p, rx[l - 1], ry[l - 1], rz[l - 1], 0, &s);
    scatter_one(p, q, &s, qmult);                      /* F:2509-2528 */
  }
}

/* ------------------------------------------------------------------------ */
/* One rank's share of fulmov, F:1150-1365 (+ the scatter loops of srimp1/2
 * into this rank's private raw arrays). */
static void fulmov_rank(const orc_parm* p, const double* const a6[6], double* x,
                        double* y, double* z, double* vx, double* vy,
                        double* vz, double qmult, double wmult, int64_t npr,
                        int ipc, int64_t first, int64_t stride, int32_t* ranfb,
                        double* const raw4[4], double wk[2], double* rxl,
                        double* ryl, double* rzl, double* vxj, double* vyj,
                        double* vzj) {
  const double dt = p->dt, aimpl = p->aimpl, adt = p->adt, hdt = p->hdt;
  const double hh = dt * qmult / wmult;                /* F:1150-1152 */
  const double ht = 0.5 * hh;
  const double ht2 = ht * ht;

  for (int64_t l = first; l <= npr; l += stride) {     /* F:1162-1166 */
    rxl[l - 1] = x[l - 1] + hdt * vx[l - 1];
    ryl[l - 1] = y[l - 1] + hdt * vy[l - 1];
    rzl[l - 1] = z[l - 1] + hdt * vz[l - 1];
  }
  orc_partbcEST(p, rxl, ryl, rzl, npr, first, stride); /* F:1168 */

  double wkix = 0.0, wkih = 0.0;
  for (int64_t l = first; l <= npr; l += stride) {     /* F:1174-1309 */
    const int64_t m = l - 1;
    stencil s;
    index_weights(p, rxl[m], ryl[m], rzl[m], 1, &s);
    double exi = gather_one(p, a6[0], &s), eyi = gather_one(p, a6[1], &s),
           ezi = gather_one(p, a6[2], &s), bxi = gather_one(p, a6[3], &s),
           byi = gather_one(p, a6[4], &s), bzi = gather_one(p, a6[5], &s);
    double bsqi = bxi * bxi + byi * byi + bzi * bzi;   /* F:1272-1280 */
    double acx = exi + vy[m] * bzi - vz[m] * byi;
    double acy = eyi + vz[m] * bxi - vx[m] * bzi;
    double acz = ezi + vx[m] * byi - vy[m] * bxi;
    double ach = exi * bxi + eyi * byi + ezi * bzi;
    double dvx = (acx + ht2 * ach * bxi + ht * (acy * bzi - acz * byi)) / (1.0 + ht2 * bsqi);
    double dvy = (acy + ht2 * ach * byi + ht * (acz * bxi - acx * bzi)) / (1.0 + ht2 * bsqi);
    double dvz = (acz + ht2 * ach * bzi + ht * (acx * byi - acy * bxi)) / (1.0 + ht2 * bsqi);
    wkix = wkix + 0.5 * (acx * acx + acy * acy + acz * acz); /* F:1282-1283 */
    wkih = wkih + 0.5 * (ach * ach);
    if (ipc == 0) {                                    /* F:1289-1295 */
      x[m] = x[m] + dt * (vx[m] + 0.5 * hh * dvx);
      y[m] = y[m] + dt * (vy[m] + 0.5 * hh * dvy);
      z[m] = z[m] + dt * (vz[m] + 0.5 * hh * dvz);
      vx[m] = vx[m] + hh * dvx;
      vy[m] = vy[m] + hh * dvy;
      vz[m] = vz[m] + hh * dvz;
    } else {                                           /* F:1300-1306 */
      vxj[m] = vx[m] + aimpl * hh * dvx;
      vyj[m] = vy[m] + aimpl * hh * dvy;
      vzj[m] = vz[m] + aimpl * hh * dvz;
      rxl[m] = x[m] + adt * (vx[m] + 0.5 * hh * dvx);
      ryl[m] = y[m] + adt * (vy[m] + 0.5 * hh * dvy);
      rzl[m] = z[m] + adt * (vz[m] + 0.5 * hh * dvz);
    }
  }
  wk[0] = wkix; wk[1] = wkih;

  if (ipc == 0) {
    orc_partbc(p, x, y, z, vy, npr, first, stride);    /* F:1337 */
    for (int64_t l = first; l <= npr; l += stride) {   /* F:1342-1364 */
      const int64_t m = l - 1;
      if ((fabs(z[m] - p->zcent) < 0.15 * p->zmax) &&
          ((fabs(y[m] - p->ycent2) < 0.025 * p->ymax) ||
           (fabs(y[m] - p->ycent1) < 0.025 * p->ymax))) {
        int ip = (int)(p->hxi * x[m] + 0.500000001);
        int jp = (int)(p->hyi * y[m] + 0.000000001);
        int kp = (int)(p->hzi * z[m] + 0.500000001);
        if (orc_ranfp(ranfb) > 0.999) {                /* F:1353 */
          double vy0 = p->Ez00 / a6[3][IDX(p, ip, jp, kp)];
          if (fabs(y[m] - p->ycent2) < 0.05 * p->ymax) vy[m] = vy[m] - vy0;
          else if (fabs(y[m] - p->ycent1) < 0.05 * p->ymax) vy[m] = vy[m] + vy0;
        }
      }
    }
  } else {
    orc_partbc(p, rxl, ryl, rzl, vyj, npr, first, stride); /* F:1375 */
    orc_srimp1_scatter(p, rxl, ryl, rzl, vxj, vyj, vzj, qmult, raw4[0], raw4[1], raw4[2], npr, first, stride);
    orc_srimp2_scatter(p, rxl, ryl, rzl, qmult, raw4[3], npr, first, stride);
  }
}

void orc_fulmov(const orc_parm* p, const double* const a6[6], double* x,
                double* y, double* z, double* vx, double* vy, double* vz,
                double qmult, double wmult, int64_t npr, int ipc, int nranks,
                int32_t* ranfb, double* const mom4[4], double* const raw4[4],
                double wk[2], double* const pred6[6]) {
  const int64_t n = orc_mxyzA(p);
  double* tmp[6];
  for (int c = 0; c < 6; c++) {
    if (pred6 && pred6[c]) tmp[c] = pred6[c];
    else tmp[c] = (double*)malloc((size_t)(npr > 0 ? npr : 1) * sizeof(double));
  }
  double* part = NULL;  /* per-rank raw partials, [rank][4][n] */
  if (ipc >= 1) part = (double*)calloc((size_t)nranks * 4 * (size_t)n, sizeof(double));
  double* wkr = (double*)calloc((size_t)nranks * 2, sizeof(double));

#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1)
#endif
  for (int r = 0; r < nranks; r++) {
    double* rr[4] = {NULL, NULL, NULL, NULL};
    if (ipc >= 1)
      for (int c = 0; c < 4; c++) rr[c] = part + ((size_t)r * 4 + c) * (size_t)n;
    fulmov_rank(p, a6, x, y, z, vx, vy, vz, qmult, wmult, npr, ipc, r + 1, nranks,
                &ranfb[r], rr, &wkr[2 * r], tmp[0], tmp[1], tmp[2], tmp[3], tmp[4], tmp[5]);
  }
  /* mpi_allreduce stand-ins: rank-ordered sums (F:1312-1317, 2379-2384, 2533) */
  wk[0] = 0.0; wk[1] = 0.0;
  for (int r = 0; r < nranks; r++) { wk[0] = wk[0] + wkr[2 * r]; wk[1] = wk[1] + wkr[2 * r + 1]; }
  if (ipc >= 1) {
    for (int c = 0; c < 4; c++) {
      double* out = mom4[c];
      for (int64_t m = 0; m < n; m++) {
        double s = 0.0;
        for (int r = 0; r < nranks; r++) s = s + part[((size_t)r * 4 + c) * (size_t)n + m];
        out[m] = s;
      }
      if (raw4 && raw4[c]) memcpy(raw4[c], out, (size_t)n * sizeof(double));
    }
    orc_vmesh3(p, mom4[0], mom4[1], mom4[2]);          /* F:2398 */
    orc_vmesh1(p, mom4[3]);                            /* F:2544 */
  }
  free(wkr);
  free(part);
  for (int c = 0; c < 6; c++)
    if (!(pred6 && pred6[c])) free(tmp[c]);
}

/* ------------------------------------------------------------------------ */
/* loadpt, F:8850-9040 (the unused fdr/fv1 tables of F:8806-8878 are omitted) */
static double fun2(double v, double vrg1) { return exp(-(v * v)) * (v + vrg1); } /* F:9195 */

void orc_loadpt_fv2(double vth, double vdr, double fv2[101], double* v2, double* dv2) {
  double vrg1 = vdr / vth;                             /* F:8885-8890 */
  double vv = (-3.0 > -vrg1) ? -3.0 : -vrg1;
  double dv = (3.0 - vv) / 100.0;
  *v2 = vv * vth;
  *dv2 = dv * vth;
  fv2[0] = 0.0;                                        /* F:8851 */
  for (int j = 1; j <= 100; j++) {                     /* F:8892-8905 */
    double s = 0.0;
    int ns = 1000, k2 = ns / 2;
    double sdv = dv / (double)(float)ns;
    for (int k = 1; k <= k2; k++) {
      vv = vv + 2.0 * sdv;
      s = s + 4.0 * fun2(vv - sdv, vrg1) + 2.0 * fun2(vv, vrg1);
    }
    s = (s + 4.0 * fun2(vv + sdv, vrg1) + fun2(vv + 2.0 * sdv, vrg1)) * sdv / 3.0;
    fv2[j] = fv2[j - 1] + s;
  }
  double norm = fv2[100];
  for (int j = 0; j <= 100; j++) fv2[j] = fv2[j] / norm; /* F:8907-8909 */
}

int64_t orc_loadpt(const orc_parm* p, int ppc, double vth, double vdr,
                   double vbeam, double* x, double* y, double* z, double* vx,
                   double* vy, double* vz, int32_t* ranfa, int32_t* ranfb) {
  double fv2[101], v2, dv2;
  orc_loadpt_fv2(vth, vdr, fv2, &v2, &dv2);
  const double xmax = p->xmax, ymax = p->ymax, zmax = p->zmax;
  int64_t l = 0;
  for (int k = 1; k <= p->mz; k++)                     /* F:8937-8954 */
    for (int j = 1; j <= p->my; j++)
      for (int i = 1; i <= p->mx; i++)
        for (int m = 1; m <= ppc; m++) {
          x[l] = xmax * orc_ranfp(ranfb) - p->hx / 2;
          y[l] = ymax * orc_ranfp(ranfb);
          z[l] = zmax * orc_ranfp(ranfb) - p->hz / 2;
          l++;
        }
  const int64_t npr = l;
  const double zcent = 0.50 * zmax, dzcent = 0.125 * zmax, dzsmt = 0.15 * zmax;
  const double ycent1 = 0.30 * ymax, ycent2 = 0.70 * ymax, dycent = 0.05 * ymax;
  for (l = 0; l < npr; l++) {                          /* F:8976-9040 */
    double eps = orc_ranf(ranfa);
    int k2 = 100;
    for (int k = 1; k <= 100; k++) {                   /* F:8979-8982, 1-based fv2(k) */
      k2 = k;
      if (fv2[k - 1] > eps) break;
    }
    double y1 = fv2[k2 - 2], y2 = fv2[k2 - 1];
    double x2 = (eps - y2) / (y2 - y1) + k2;
    double vmag = v2 + dv2 * (x2 - 1.0) + vdr;
    double vxout = vmag * (orc_ranf(ranfa) - 0.5);
    double vyout = vmag * (orc_ranf(ranfa) - 0.5);
    double vzout = vmag * (orc_ranf(ranfa) - 0.5);
    double ycnt1, ycnt2;
    if (fabs(z[l] - zcent) < dzcent) {                 /* F:9012-9023 */
      ycnt1 = ycent1 + dycent; ycnt2 = ycent2 - dycent;
    } else if (fabs(z[l] - zcent) < dzsmt) {
      ycnt1 = ycent1 + dycent; ycnt2 = ycent2 - dycent;
    } else {
      ycnt1 = ycent1; ycnt2 = ycent2;
    }
    const double rrz = 0.25 * zmax, rry = 0.075 * ymax; /* F:9026-9035 */
    double vdrift = 0.0;
    if ((fabs(z[l] - zcent) <= rrz) &&
        (fabs(y[l] - ycnt1) <= rry || fabs(y[l] - ycnt2) <= rry))
      vdrift = vbeam;
    vx[l] = vxout + vdrift;
    vy[l] = vyout;
    vz[l] = vzout;
  }
  return npr;
}

/* ------------------------------------------------------------------------ */
/* CPU-baseline timer: one full particle step (ipc=1 then ipc=0, F:761-787)
 * of one species executed the way the reference's MPI job does it: every
 * "rank" (an OpenMP thread here) holds its OWN full-size copy of the particle
 * arrays and temporaries (F:121-122, 1057) and walks them with stride nranks.
 * Copies are made before the clock starts.  Returns the wall-clock seconds of
 * the two passes including the rank-ordered moment sums and the folds; the
 * field preparation (a6p for the predictor, a6c for the corrector) is timed
 * separately by the caller.  x..vz are left untouched. */
double orc_time_step(const orc_parm* p, const double* const a6p[6], const double* const a6c[6],
                     const double* x, const double* y, const double* z, const double* vx,
                     const double* vy, const double* vz, double qmult, double wmult, int64_t npr,
                     int nranks, double* t_pred, double* t_corr) {
  const int64_t n = orc_mxyzA(p);
  const size_t np = (size_t)(npr > 0 ? npr : 1);
  double** priv = (double**)malloc((size_t)nranks * 12 * sizeof(double*));
  const double* src[6] = {x, y, z, vx, vy, vz};
  for (int r = 0; r < nranks; r++)
    for (int c = 0; c < 12; c++) {
      priv[r * 12 + c] = (double*)malloc(np * sizeof(double));
      if (c < 6) memcpy(priv[r * 12 + c], src[c], np * sizeof(double));
      else memset(priv[r * 12 + c], 0, np * sizeof(double));
    }
  double* part = (double*)calloc((size_t)nranks * 4 * (size_t)n, sizeof(double));
  double* mom = (double*)calloc((size_t)4 * (size_t)n, sizeof(double));
  double* wkr = (double*)calloc((size_t)nranks * 2, sizeof(double));
  int32_t* st = (int32_t*)malloc((size_t)nranks * sizeof(int32_t));
  for (int r = 0; r < nranks; r++) st[r] = 7331;
  double t0 = 0, t1 = 0, t2 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#pragma omp parallel for schedule(static, 1)
#endif
  for (int r = 0; r < nranks; r++) {
    double** q = priv + r * 12;
    double* rr[4];
    for (int c = 0; c < 4; c++) rr[c] = part + ((size_t)r * 4 + c) * (size_t)n;
    fulmov_rank(p, a6p, q[0], q[1], q[2], q[3], q[4], q[5], qmult, wmult, npr, 1, r + 1, nranks,
                &st[r], rr, &wkr[2 * r], q[6], q[7], q[8], q[9], q[10], q[11]);
  }
  for (int c = 0; c < 4; c++)
    for (int64_t m = 0; m < n; m++) {
      double s = 0.0;
      for (int r = 0; r < nranks; r++) s = s + part[((size_t)r * 4 + c) * (size_t)n + m];
      mom[(size_t)c * (size_t)n + m] = s;
    }
  orc_vmesh3(p, mom, mom + n, mom + 2 * n);
  orc_vmesh1(p, mom + 3 * n);
#ifdef _OPENMP
  t1 = omp_get_wtime();
#pragma omp parallel for schedule(static, 1)
#endif
  for (int r = 0; r < nranks; r++) {
    double** q = priv + r * 12;
    double* rr[4] = {NULL, NULL, NULL, NULL};
    fulmov_rank(p, a6c, q[0], q[1], q[2], q[3], q[4], q[5], qmult, wmult, npr, 0, r + 1, nranks,
                &st[r], rr, &wkr[2 * r], q[6], q[7], q[8], q[9], q[10], q[11]);
  }
#ifdef _OPENMP
  t2 = omp_get_wtime();
#endif
  if (t_pred) *t_pred = t1 - t0;
  if (t_corr) *t_corr = t2 - t1;
  for (int r = 0; r < nranks * 12; r++) free(priv[r]);
  free(priv); free(part); free(mom); free(wkr); free(st);
  return t2 - t0;
}
