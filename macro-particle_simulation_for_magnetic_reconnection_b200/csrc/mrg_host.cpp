// mrg_host.cpp -- see mrg_host.h.  Host mirror of subroutine fulmov (F:1044)
// that keeps particles resident on the GPU and feeds COMMON /srimp7/,
// /wkinel/ and edec exactly where the reference writes them.
#include "mrg_host.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/mrg_fulmov.h"

namespace {

struct HostState {
  mrg_common_view v{};
  bool bound = false;
  int device = 0;
  mrg_ctx* ctx = nullptr;
  bool resident[MRG_MAX_SPECIES] = {false, false, false, false};
  int corrector_calls[MRG_MAX_SPECIES] = {0, 0, 0, 0};
  bool auto_fields = true;
  unsigned dirty = 0xFFFu;     // members of COMMON /fields/ the device copy is stale for
  bool renew = false;          // ex0 <- ex still to be repeated on the device
  int b_pending = -1;          // -1 none; 0 = the host ran prefld, 1 = emfild on a smoothing step: bx,by,bz are recomputed on the device
  bool it0 = false;            // the it = 0 pair of calls has been seen and emfld0 may have rewritten every field since
  int sort_interval = 1;
  bool exit_on_error = true;
  int status = 0;
  bool have_id = false;
  int nspecies = 2;            // the reference handles two (F:1321-1327); more only on request (qspec(4), wspec(4), F:1100)
  double* extra[MRG_MAX_SPECIES][4] = {};   // moment arrays of species 3, 4 (the reference has none)
  unsigned char id[MRG_UNIQUE_ID_BYTES];
} H;

void (*g_abort)(int) = nullptr;   // e.g. a wrapper around MPI_Abort: a rank that dies alone leaves its peers in NCCL forever

void die(const char* where, int rc) {
  H.status = rc;
  std::fprintf(stderr, "fulmov(gpu): %s failed (%d): %s\n", where, rc, mrg_last_error());
  if (H.exit_on_error) {               // the reference has no status argument: stop
    if (g_abort) g_abort(rc);
    std::exit(1);
  }
}

}  // namespace

extern "C" {

int mrg_host_bind(const mrg_common_view* view, int32_t device) {
  if (!view) return MRG_ERR_ARG;
  mrg_host_unbind();
  H.v = *view;
  H.device = device;
  H.bound = true;
  H.status = 0;
  return MRG_OK;
}

void mrg_host_unbind(void) {
  if (H.ctx) mrg_destroy(H.ctx);
  H = HostState();
}

void mrg_host_fields_changed(void) { H.dirty = 0xFFFu; }
void mrg_host_fields_changed_mask(uint32_t mask) { H.dirty |= (mask & 0xFFFu); }
void mrg_host_fields_renewed(void) {
  if (H.auto_fields) H.dirty |= 0xFC0u;
  else H.renew = true;
}
void mrg_host_prefld_done(void) {                 // after `call prefld` (F:759)
  if (H.auto_fields) H.dirty |= 0x038u;
  else H.b_pending = 0;
}
void mrg_host_emfild_done(void) {                 // after `call emfild` (F:771): only ex,ey,ez cross PCIe
  if (H.auto_fields || !H.bound) { H.dirty |= 0x03Fu; return; }
  H.dirty |= 0x007u;
  H.b_pending = (*H.v.it % 5 == 1) ? 1 : 0;       // F:4298: the smoothing steps
}
void mrg_host_set_auto_fields(int32_t on) { H.auto_fields = on != 0; }
int mrg_host_set_nspecies(int32_t n) {
  if (n < 2 || n > MRG_MAX_SPECIES || H.ctx) return MRG_ERR_ARG;   // before the first fulmov call
  H.nspecies = n;
  return MRG_OK;
}
int mrg_host_bind_extra_moments(int32_t ksp, double* qjx, double* qjy, double* qjz, double* q) {
  if (ksp < 3 || ksp > MRG_MAX_SPECIES) return MRG_ERR_ARG;
  H.extra[ksp - 1][0] = qjx; H.extra[ksp - 1][1] = qjy; H.extra[ksp - 1][2] = qjz; H.extra[ksp - 1][3] = q;
  return MRG_OK;
}
void mrg_host_set_sort_interval(int32_t n) { H.sort_interval = n < 0 ? 0 : n; }
void mrg_host_set_exit_on_error(int32_t on) { H.exit_on_error = on != 0; }
void mrg_host_set_abort(void (*fn)(int)) { g_abort = fn; }
int mrg_host_status(void) { return H.status; }
void* mrg_host_context(void) { return H.ctx; }
void mrg_host_particles_changed(int32_t ksp) {
  if (ksp >= 1 && ksp <= MRG_MAX_SPECIES) H.resident[ksp - 1] = false;
}

int mrg_host_unique_id(unsigned char id[128]) { return mrg_comm_unique_id(id); }
int mrg_host_set_unique_id(const unsigned char id[128]) {
  std::memcpy(H.id, id, MRG_UNIQUE_ID_BYTES);
  H.have_id = true;
  return MRG_OK;
}

void mrg_host_fulmov(double* x, double* y, double* z, double* vx, double* vy, double* vz, double* qmult, double* wmult,
            int32_t* npr, int32_t* ipc, int32_t* ksp, int32_t* ipar, int32_t* size) {
  H.status = 0;
  if (!H.bound) { std::fprintf(stderr, "fulmov(gpu): mrg_host_bind was not called\n"); H.status = MRG_ERR_STATE; if (H.exit_on_error) std::exit(1); return; }
  const mrg_common_view& v = H.v;
  const int k = *ksp;
  if (k < 1 || k > H.nspecies) {   // the reference handles exactly two species (F:1321-1327, 1377-1386); see mrg_host_set_nspecies
    std::fprintf(stderr, "fulmov(gpu): ksp must be 1..%d\n", H.nspecies);
    H.status = MRG_ERR_ARG;
    if (H.exit_on_error) std::exit(1);
    return;
  }
  int rc;
  if (!H.ctx) {
    rc = mrg_create(&H.ctx, v.mx, v.my, v.mz, *v.xmax, *v.ymax, *v.zmax, H.nspecies, *ipar - 1, *size, H.device);
    if (rc) return die("mrg_create", rc);
    if (*size > 1) {
      if (!H.have_id) { std::fprintf(stderr, "fulmov(gpu): size > 1 needs mrg_host_set_unique_id\n"); H.status = MRG_ERR_STATE; if (H.exit_on_error) std::exit(1); return; }
      rc = mrg_comm_init(H.ctx, H.id);
      if (rc) return die("mrg_comm_init", rc);
    }
  }
  if (!H.resident[k - 1]) {   // first call (or after restrt): take the owned subset, l = ipar, ipar+size, ...
    rc = mrg_upload_particles(H.ctx, k, x, y, z, vx, vy, vz, *npr, *ipar, *size);
    if (rc) return die("mrg_upload_particles", rc);
    H.resident[k - 1] = true;
  }
  if (H.auto_fields && k == 1) H.dirty = 0xFFFu;
  // it = 0 (F:664-706): trans calls the pair with dt = 0, then emfld0 rewrites ALL of COMMON /fields/ on the host
  // (F:691, 3384-3703) and the renewal loop runs -- none of which the three optional marks describe.  So the first
  // call after the it = 0 pair uploads everything and drops the pending device renewal (it would copy the pre-emfld0
  // ex..bz into ex0..bz0).
  if (*v.it == 0) H.it0 = true;
  else if (H.it0) { H.it0 = false; H.dirty = 0xFFFu; H.renew = false; }   // a pending B update stays: E, e0, b0 arrive with this upload
  if (H.renew) {                                          // F:796-807 on the device copies
    rc = mrg_renew_fields(H.ctx);
    if (rc) return die("mrg_renew_fields", rc);
    H.renew = false;
    H.dirty &= ~0xFC0u;
  }
  if (H.b_pending >= 0) H.dirty &= ~0x038u;               // bx,by,bz are computed below from what the device holds
  if (H.dirty) {
    const double* f12[12] = {v.ex, v.ey, v.ez, v.bx, v.by, v.bz, v.ex0, v.ey0, v.ez0, v.bx0, v.by0, v.bz0};
    rc = mrg_set_fields(H.ctx, H.dirty, f12);
    if (rc) return die("mrg_set_fields", rc);
    H.dirty = 0;
  }
  if (H.b_pending >= 0) {                                 // prefld (F:3820-3873) / emfild's B update (F:4238-4302) on the device copies
    rc = mrg_update_b(H.ctx, *v.dt, *v.aimpl, H.b_pending);
    if (rc) return die("mrg_update_b", rc);
    H.b_pending = -1;
  }
  mrg_step_params p;
  p.dt = *v.dt; p.adt = *v.adt; p.hdt = *v.hdt; p.aimpl = *v.aimpl;
  p.bxc = *v.bxc; p.byc = *v.byc; p.bzc = *v.bzc;
  p.ifilx = *v.ifilx; p.ifily = *v.ifily; p.ifilz = *v.ifilz;
  p.drive_on = 1;
  p.Ez00 = *v.Ez00; p.zcent = *v.zcent; p.ycent1 = *v.ycent1; p.ycent2 = *v.ycent2;
  double wkix = 0.0, wkih = 0.0;
  rc = mrg_fulmov(H.ctx, k, *qmult, *wmult, *ipc, &p, v.ranfb, &wkix, &wkih);
  if (rc) return die("mrg_fulmov", rc);
  *v.wkix = wkix;                                         // F:1316-1317
  *v.wkih = wkih;
  if ((*v.it % *v.nha) == 0 && *v.io_pe == 1 && k <= 2) {  // F:1320-1328
    const long row = *v.ldec - 1;
    const int col = (k == 1) ? 5 : 7;
    v.edec[row + 3000L * (col - 1)] = wkix;
    v.edec[row + 3000L * col] = wkih;
  }
  if (*ipc >= 1) {                                        // F:1377-1386
    rc = (k == 1) ? mrg_get_moments(H.ctx, 1, v.qix, v.qiy, v.qiz, v.qi, 1)
       : (k == 2) ? mrg_get_moments(H.ctx, 2, v.qex, v.qey, v.qez, v.qe, 1)
                  : mrg_get_moments(H.ctx, k, H.extra[k - 1][0], H.extra[k - 1][1], H.extra[k - 1][2], H.extra[k - 1][3], 1);
    if (rc) return die("mrg_get_moments", rc);
  } else {
    H.corrector_calls[k - 1]++;
    if (H.sort_interval > 0 && H.corrector_calls[k - 1] % H.sort_interval == 0) {
      rc = mrg_sort(H.ctx, k, *v.hdt);   // key = cell of the next gather position x + hdt*v
      if (rc) return die("mrg_sort", rc);
    }
  }
}

int mrg_host_pull_particles(int32_t ksp, double* x, double* y, double* z, double* vx, double* vy, double* vz,
                            int32_t npr, int32_t ipar, int32_t size) {
  if (!H.ctx || ksp < 1 || ksp > MRG_MAX_SPECIES || !H.resident[ksp - 1]) return MRG_ERR_STATE;
  return mrg_download_particles(H.ctx, ksp, x, y, z, vx, vy, vz, npr, ipar, size);
}

}  // extern "C"
