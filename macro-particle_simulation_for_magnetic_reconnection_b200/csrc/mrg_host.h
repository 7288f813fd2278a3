/*
 * mrg_host.h -- host-side mirror of the reference's fulmov interface, above
 * the C ABI of include/mrg_fulmov.h.
 *
 * The reference is Fortran 2003; this image has no Fortran compiler, so the
 * host logic that the ISO_C_BINDING shim (fortran/mrg_gpu.f03) needs lives
 * here in C++ where it can be compiled and tested: same subroutine name, same
 * argument list passed by reference (F:1044), same COMMON-block side effects
 * (F:1066-1110), same "no status argument" error behaviour (a failure stops
 * the program, as a Fortran `stop` would).  F:n = @mrg37-080A.f03 line n.
 */
#ifndef MRG_HOST_H
#define MRG_HOST_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Pointers into the caller's COMMON blocks (c_loc of each member in the
 * Fortran shim; plain numpy buffers in the tests). */
typedef struct mrg_common_view {
  int32_t mx, my, mz;                       /* param_080A.h:14                */
  /* common/fields/  F:1066 */
  double *ex, *ey, *ez, *bx, *by, *bz, *ex0, *ey0, *ez0, *bx0, *by0, *bz0;
  /* common/srimp7/  F:1067 (the members fulmov writes) */
  double *qix, *qiy, *qiz, *qex, *qey, *qez, *qi, *qe;
  /* common/parm1/   F:1088 */
  int32_t *it, *ldec, *ifilx, *ifily, *ifilz, *nha;
  /* common/parm2/   F:1099-1105 */
  double *xmax, *ymax, *zmax, *dt, *aimpl, *adt, *hdt, *bxc, *byc, *bzc;
  double *edec;                             /* edec(3000,12), column-major    */
  /* common/wkinel/  F:1107 */
  double *wkix, *wkih;
  /* common/profl/   F:1110 */
  double *zcent, *ycent1, *ycent2, *Ez00;
  /* common/ranfb/   F:9294 */
  int32_t *ranfb;
  /* common/iope66/  F:1120 */
  int32_t *io_pe;
} mrg_common_view;

/* Bind the COMMON storage and pick the CUDA device; the device context is
 * created at the first fulmov call (when ipar/size are known). */
int mrg_host_bind(const mrg_common_view* view, int32_t device);
void mrg_host_unbind(void);

/* Optional hints.  Without them every ksp==1 call re-uploads /fields/ (trans
 * always moves ions first after a field change, F:761-766, 782-787). */
void mrg_host_fields_changed(void);
void mrg_host_set_auto_fields(int32_t on);
/* More than the reference's two species (qspec(4), wspec(4) exist, F:1100, but trans / fulmov / emfild handle ksp = 1|2
 * only, F:1321-1327): allow ksp <= n (call before the first fulmov) and say where the moments of species 3, 4 go (the
 * reference has no COMMON member for them; NULL = not delivered).  edec rows are written for ksp = 1, 2 only.          */
int mrg_host_set_nspecies(int32_t n);
int mrg_host_bind_extra_moments(int32_t ksp, double* qjx, double* qjy, double* qjz, double* q);
/* Finer hints for a host that marks its three field updates in trans (turn
 * auto_fields off first): bit i of mask = member i of COMMON /fields/ changed
 * on the host (prefld F:759 -> 0x038 bx,by,bz; emfild F:771 -> 0x03F ex..bz);
 * renewed = the host ran the loop ex0 <- ex (F:796-807), which is then
 * repeated on the device copies instead of uploading ex0..bz0. */
void mrg_host_fields_changed_mask(uint32_t mask);
/* After `call prefld` (F:759): with auto fields off the entry is repeated on the device (mrg_prefld, bit-identical to the
 * host's) instead of uploading bx,by,bz; with auto fields on this is mrg_host_fields_changed_mask(0x038).                */
void mrg_host_prefld_done(void);
/* After `call emfild` (F:771): with auto fields off only ex,ey,ez are uploaded and bx,by,bz are recomputed on the device as
 * emfild does behind its solve (mrg_update_b; smoothed when mod(it,5) = 1); else mrg_host_fields_changed_mask(0x03F).   */
void mrg_host_emfild_done(void);
void mrg_host_fields_renewed(void);
/* Cell-sort every n-th corrector call of a species (0 = never). */
void mrg_host_set_sort_interval(int32_t n);
/* 0: return to the caller after an error (mrg_host_status() != 0) instead of
 * exiting -- used by the tests. */
void mrg_host_set_exit_on_error(int32_t on);
/* Multi-rank hosts: called with the error code before the process exits on a failure, so that the host can take the
 * whole job down (a wrapper around MPI_Abort) instead of leaving the other ranks blocked in a collective.            */
void mrg_host_set_abort(void (*fn)(int));
int mrg_host_status(void);

/* The drop-in: same name and argument list as F:1044. */
void mrg_host_fulmov(double* x, double* y, double* z, double* vx, double* vy, double* vz,
            double* qmult, double* wmult, int32_t* npr, int32_t* ipc, int32_t* ksp,
            int32_t* ipar, int32_t* size);

/* Particles live on the GPU between calls.  Call before any host code reads
 * x..vz (restrt F:9622, diag1 F:7879) ... */
int mrg_host_pull_particles(int32_t ksp, double* x, double* y, double* z, double* vx,
                            double* vy, double* vz, int32_t npr, int32_t ipar, int32_t size);
/* ... and after host code has modified them (restrt read, F:381-387). */
void mrg_host_particles_changed(int32_t ksp);

/* NCCL bootstrap for size > 1: rank 0 fills id, the host broadcasts it
 * (MPI_Bcast of 128 bytes) and every rank passes it here before the first
 * fulmov call. */
int mrg_host_unique_id(unsigned char id[128]);
int mrg_host_set_unique_id(const unsigned char id[128]);

/* Underlying context (for diagnostics / tests); NULL before the first call. */
void* mrg_host_context(void);

#ifdef __cplusplus
}
#endif
#endif
