// A recording stand-in for libmrg_fulmov.so: the entry points csrc/mrg_host.cpp calls, each appending one line to a
// trace.  Linked with mrg_host.cpp into one test library so that the host mirror's state machine -- which members of
// COMMON /fields/ it uploads when, the device renewal, the it = 0 sequence, the prefld / emfild marks, the sort cadence --
// can be checked on a machine without a GPU (tests/test_host_mirror_logic.py).  TEST INFRASTRUCTURE, not a product path.
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/mrg_fulmov.h"

struct mrg_ctx { int dummy; };
static mrg_ctx g_ctx;
static std::string g_trace;
static char g_line[256];
#define TR(...) do { std::snprintf(g_line, sizeof g_line, __VA_ARGS__); g_trace += g_line; g_trace += "\n"; } while (0)

extern "C" {
const char* stub_trace(void) { return g_trace.c_str(); }
void stub_reset(void) { g_trace.clear(); }

int mrg_create(mrg_ctx** ctx, int32_t mx, int32_t my, int32_t mz, double, double, double, int32_t nspecies, int32_t rank,
               int32_t nranks, int32_t device) {
  TR("create %d %d %d nspecies=%d rank=%d nranks=%d device=%d", mx, my, mz, nspecies, rank, nranks, device);
  *ctx = &g_ctx;
  return 0;
}
int mrg_destroy(mrg_ctx*) { TR("destroy"); return 0; }
const char* mrg_last_error(void) { return "stub"; }
int mrg_comm_unique_id(unsigned char id[MRG_UNIQUE_ID_BYTES]) { std::memset(id, 7, MRG_UNIQUE_ID_BYTES); return 0; }
int mrg_comm_init(mrg_ctx*, const unsigned char*) { TR("comm_init"); return 0; }
int mrg_upload_particles(mrg_ctx*, int32_t ksp, const double*, const double*, const double*, const double*, const double*,
                         const double*, int64_t npr, int64_t first, int64_t stride) {
  TR("upload ksp=%d npr=%lld first=%lld stride=%lld", ksp, (long long)npr, (long long)first, (long long)stride);
  return 0;
}
int mrg_download_particles(mrg_ctx*, int32_t ksp, double*, double*, double*, double*, double*, double*, int64_t, int64_t, int64_t) {
  TR("download ksp=%d", ksp);
  return 0;
}
int mrg_set_fields(mrg_ctx*, uint32_t mask, const double* const f12[12]) {
  TR("set_fields mask=0x%03x ex=%g bx=%g", mask, (mask & 1u) ? f12[0][0] : -1.0, (mask & 8u) ? f12[3][0] : -1.0);
  return 0;
}
int mrg_renew_fields(mrg_ctx*) { TR("renew"); return 0; }
int mrg_update_b(mrg_ctx*, double dt, double aimpl, int32_t smooth) { TR("update_b dt=%g aimpl=%g smooth=%d", dt, aimpl, smooth); return 0; }
int mrg_fulmov(mrg_ctx*, int32_t ksp, double qmult, double wmult, int32_t ipc, const mrg_step_params* p, int32_t* ranfb,
               double* wkix, double* wkih) {
  TR("fulmov ksp=%d ipc=%d dt=%g hdt=%g q=%g w=%g", ksp, ipc, p->dt, p->hdt, qmult, wmult);
  *wkix = 10.0 * ksp + ipc; *wkih = 0.5;
  if (ipc == 0) *ranfb += 1;
  return 0;
}
int mrg_get_moments(mrg_ctx*, int32_t ksp, double* qjx, double*, double*, double*, int32_t folded) {
  TR("get_moments ksp=%d folded=%d", ksp, folded);
  qjx[0] = 100.0 + ksp;
  return 0;
}
int mrg_sort(mrg_ctx*, int32_t ksp, double lookahead) { TR("sort ksp=%d lookahead=%g", ksp, lookahead); return 0; }
}
