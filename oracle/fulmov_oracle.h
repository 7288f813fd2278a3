/*
 * fulmov_oracle.h -- CPU oracle for the /fulmov/ particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or the reported CPU baseline.
 *
 * PINNED TO THE REFERENCE'S OWN CODE: the reference ships no tests, golden
 * vectors or fixtures and the image has no Fortran compiler, so
 * oracle/f03c.py (a Fortran-2003-subset -> C translator) compiles the
 * reference's fulmov, init, loadpt, partbc*, srimp1/2, outmesh3, filt3e,
 * vmesh3/1, ranf(p) from /root/reference/@mrg37-080A.f03 where it lies into
 * oracle/_ref/ (oracle/build_ref.py; git-ignored).  This restatement agrees
 * with that library BIT FOR BIT on loads, folded moments, wkix/wkih,
 * corrector output with the drive kick and every rank's ranfp state, for 1-8
 * simulated ranks, edge placements and the dt*wce > 10 regime
 * (tests/test_ref_pin.py), and reproduces the committed reference-output
 * fixtures tests/golden/ref_*.npz bit for bit.  A second, independent numpy
 * transcription (oracle/np_restatement.py, tests/test_oracle_crosscheck.py)
 * and the analytic invariants (tests/test_oracle_invariants.py) remain as
 * further layers.
 *
 * Citation shorthand: F:n = /root/reference/@mrg37-080A.f03 line n.
 * Arrays use the reference layout real(C_DOUBLE)(-2:mx+1,-1:my+1,-2:mz+1),
 * i fastest (F:1061-1071); particles are 1-based in the source, 0-based here.
 */
#ifndef FULMOV_ORACLE_H
#define FULMOV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Values the reference keeps in COMMON /parm1/,/parm2/,/ptable/,/profl/. */
typedef struct orc_parm {
  int32_t mx, my, mz;
  int32_t ifilx, ifily, ifilz;           /* F:368-370 forces 1,1,1            */
  double xmax, ymax, zmax;
  double hx, hy, hz;                     /* F:8454,8467,8484                  */
  double hxi, hyi, hzi;                  /* F:8567-8569                       */
  double xmaxe, ymaxe, zmaxe;            /* F:8575-8577                       */
  double dt, aimpl, adt, hdt;            /* F:8579-8580                       */
  double bxc, byc, bzc;                  /* F:8601-8603                       */
  double Ez00, zcent, ycent1, ycent2;    /* /profl/, F:9001-9006              */
} orc_parm;

int64_t orc_mxyzA(const orc_parm* p);

void orc_parm_init(orc_parm* p, int mx, int my, int mz, double xmax,
                   double ymax, double zmax, double dt, double aimpl,
                   double wce_by_wpe, double Ez00);

/* LCGs, F:9263-9305.  state is the COMMON integer (ranfa or ranfb).         */
double orc_ranf(int32_t* state);
double orc_ranfp(int32_t* state);
/* state * lambda^n mod 2^31 (exact skip-ahead of either generator).         */
int32_t orc_lcg_skip(int32_t state, uint64_t n);

/* Ghost fill / fold / filter, F:3076-3151, 3227-3380, 7305-7509.            */
void orc_outmesh3(const orc_parm* p, double* ax, double* ay, double* az);
void orc_vmesh3(const orc_parm* p, double* ax, double* ay, double* az);
void orc_vmesh1(const orc_parm* p, double* ax);
void orc_filt3e(const orc_parm* p, double* ex, double* ey, double* ez,
                double exc, double eyc, double ezc, int ifilx, int ifily,
                int ifilz, int sym);

/* F:1127-1148: blend, outmesh3 x2, filt3e x2.  f12 = ex,ey,ez,bx,by,bz,
 * ex0,ey0,ez0,bx0,by0,bz0; a6 = exa,eya,eza,bxa,bya,bza (all mxyzA doubles;
 * a6 is fully overwritten).                                                  */
void orc_field_prep(const orc_parm* p, const double* const f12[12],
                    double* const a6[6]);
/* entry prefld of emfild (F:3820-3873): bx,by,bz <- b0 - dt curl(ea) on the interior nodes; writes f12[3..5] */
void orc_prefld(const orc_parm* p, double* const f12[12]);
/* bx,by,bz as emfild leaves them after its solve (F:4238-4302): prefld's update from the new E, smoothed when mod(it,5) = 1 */
void orc_update_b(const orc_parm* p, double* const f12[12], int smooth);

/* F:1811-1882 and F:1886-1952 on the strided subset l = first, first+stride..*/
void orc_partbc(const orc_parm* p, double* x, double* y, double* z,
                double* vy, int64_t npr, int64_t first, int64_t stride);
void orc_partbcEST(const orc_parm* p, double* x, double* y, double* z,
                   int64_t npr, int64_t first, int64_t stride);

/* Scatter loops of srimp1 (F:2273-2374) and srimp2 (F:2471-2529) for one
 * rank's strided subset, ACCUMULATING into the given raw extended arrays
 * (no zeroing, no allreduce, no fold).                                       */
void orc_srimp1_scatter(const orc_parm* p, const double* rx, const double* ry,
                        const double* rz, const double* vxj, const double* vyj,
                        const double* vzj, double qmult, double* qjx,
                        double* qjy, double* qjz, int64_t npr, int64_t first,
                        int64_t stride);
void orc_srimp2_scatter(const orc_parm* p, const double* rx, const double* ry,
                        const double* rz, double qmult, double* q, int64_t npr,
                        int64_t first, int64_t stride);

/*
 * One call of fulmov (F:1044-1390) executed by `nranks` simulated MPI ranks
 * (rank r owns l = r+1, r+1+nranks, ... as in F:1162); rank partials of the
 * moments and of wkix/wkih are summed in rank order in place of
 * mpi_allreduce.  Ranks run as OpenMP threads when compiled with -fopenmp.
 *
 *  a6        prepared fields (from orc_field_prep), read only
 *  x..vz     particle arrays, npr entries, updated in place when ipc==0
 *  ipc       1: predict + deposit, 0: update + boundary + drive kick
 *  ranfb     per-rank ranfp states (nranks entries), advanced by the kick
 *  mom4      ipc>=1: qjx,qjy,qjz,q after srimp1/srimp2 incl. vmesh fold
 *  raw4      ipc>=1, optional (may be NULL): the same before the fold
 *  wk        wk[0]=wkix, wk[1]=wkih after the rank sum
 *  pred6     ipc>=1, optional: rxl,ryl,rzl,vxj,vyj,vzj after partbc
 */
void orc_fulmov(const orc_parm* p, const double* const a6[6], double* x,
                double* y, double* z, double* vx, double* vy, double* vz,
                double qmult, double wmult, int64_t npr, int ipc, int nranks,
                int32_t* ranfb, double* const mom4[4], double* const raw4[4],
                double wk[2], double* const pred6[6]);

/* loadpt (F:8735-9080) with the per-cell count as a parameter (the shipped
 * source hard-codes 32, F:8941).  ranfa/ranfb are the two LCG states, used
 * and advanced exactly as the source does.  Returns npr.                     */
int64_t orc_loadpt(const orc_parm* p, int ppc, double vth, double vdr,
                   double vbeam, double* x, double* y, double* z, double* vx,
                   double* vy, double* vz, int32_t* ranfa, int32_t* ranfb);
/* The tabulated cumulative distribution fv2(1:101) and v2, dv2 of loadpt
 * (F:8885-8909).                                                             */
void orc_loadpt_fv2(double vth, double vdr, double fv2[101], double* v2,
                    double* dv2);

/* CPU-baseline timer (one species, ipc=1 then ipc=0) with per-rank private
 * particle arrays as in the reference's MPI job; returns seconds.            */
double orc_time_step(const orc_parm* p, const double* const a6p[6], const double* const a6c[6],
                     const double* x, const double* y, const double* z, const double* vx,
                     const double* vy, const double* vz, double qmult, double wmult, int64_t npr,
                     int nranks, double* t_pred, double* t_corr);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
