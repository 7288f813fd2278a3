// mrg_api.cu -- context management and the C ABI of include/mrg_fulmov.h.
// Host-side orchestration of the /fulmov/ path: field upload + preparation
// cache, particle residency, the two particle passes, the NCCL moment sum,
// the fold, the drive kick and the cell sort.  F:n = @mrg37-080A.f03 line n.
#include "../../include/mrg_fulmov.h"
#include "mrg_kernels.cuh"
#include "mrg_tile.cuh"

#include <cuda.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

using namespace mrg;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(MRG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
  } while (0)

#define CKL(ctx)                                                                          \
  do {                                                                                    \
    (ctx)->launches++;                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess)                                                                \
      return fail(MRG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
  } while (0)

// ---- NCCL through dlopen: the library loads without NCCL (CPU symbol check,
// single-GPU runs) and binds to whatever libnccl.so.2 the process already has.
struct Uid { char internal[MRG_UNIQUE_ID_BYTES]; };   // ncclUniqueId: 128 bytes, passed by value
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8, kNcclSum = 0;

int nccl_load() {
  if (g_nccl.h) return MRG_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(MRG_ERR_NCCL, std::string("dlopen libnccl.so.2 failed: ") + dlerror());
  g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(void**, int, Uid, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclBroadcast");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclRecv");
  g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
  g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
    return fail(MRG_ERR_NCCL, "libnccl is missing a required symbol");
  g_nccl.h = h;
  return MRG_OK;
}
std::string nccl_err(int rc) {
  return g_nccl.GetErrorString ? std::string(g_nccl.GetErrorString(rc)) : ("nccl error " + std::to_string(rc));
}

struct Species {
  double* d[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* id = nullptr;       // nullptr = identity order
  long long n = 0, cap = 0;
  double* M4 = nullptr;    // raw moments [ntot][4] + 2 (wkix, wkih)
  double* peerM4[8] = {};  // the other ranks' M4 of this species, mapped through cudaIpc (mrg_peer_import); nullptr = not mapped
  int npeer = 0;
  int early0 = 0, early_n = 0;   // extended planes of the own block already pushed to the peers by the split launch of this call
  double* out4[4] = {nullptr, nullptr, nullptr, nullptr};  // folded, reference layout
  bool have_moments = false;
  // cell index of the current slot order (built by mrg_sort): cell_end[c] = end slot of cell c
  int* cell_end = nullptr;
  bool index_valid = false;
  // fused sort: keys of the NEXT order written by the tiled predictor (+ their histogram in hist); the tiled
  // corrector scatters by them into the spare buffers, so the order is fresh without a sort pass
  bool prekeys_valid = false;
  bool prescan_valid = false;   // cell_end2 / kocc of the next order were already scanned from the predictor's histogram (under the moment exchange)
  bool fresh = false; double fresh_lookahead = 0.0;
  int* cell_end2 = nullptr;
  // next-sort keys emitted by the corrector (fused_keys) + their histogram
  int* key = nullptr; long long key_cap = 0;
  int* hist = nullptr;
  bool keys_valid = false;
  bool hist_valid = false;     // s.hist matches s.key
  double keys_lookahead = 0.0;
  // z planes kp of the gather cells of the next pass (bit kp, kp = 0..mz), tracked by the tiled corrector and the
  // sort; valid for passes whose hdt equals zocc_lookahead.  Lets ensure_prep prepare only the planes in use.
  unsigned* zocc = nullptr;
  std::vector<unsigned> zocc_host;
  bool zocc_valid = false;
  bool zocc_approx = false;    // recorded by the corrector with contracted arithmetic: good to +-1 plane, widened by ensure_prep
  double zocc_lookahead = 0.0;
  // z planes of the current sort order that own slots (exact, from the cell index): tiled launches cover
  // only the arc kz0 .. kz0+nkz-1 (mod mz) instead of every pencil of the replicated grid
  unsigned* kocc = nullptr;
  std::vector<unsigned> kocc_host;
  bool hull_valid = false, kocc_pending = false;
  // every rank found (and the ranks agreed, in the last corrector call) that this species deposits only within
  // HALO_H planes of the rank's own z block: the next ipc>=1 call may exchange slabs instead of allreducing the grid
  bool compact_ok = false;
  int kz0 = 0, nkz = 0;
  // deferred completion (option "defer"): the moment sum, fold and wkix/wkih of the last ipc>=1 call run on the
  // communication stream; done marks their end, wk_user receives wkix/wkih when the host next waits for them
  cudaEvent_t done = nullptr;
  bool pending = false;
  double* wk_user[2] = {nullptr, nullptr};
  double* sink[4] = {nullptr, nullptr, nullptr, nullptr};   // host arrays the folded moments are copied to as soon as they exist
  bool sink_filled = false;
};

struct PrepKey {
  double aimpl, bxc, byc, bzc;
  int ifilx, ifily, ifilz;
  unsigned long long version;
  bool operator==(const PrepKey& o) const {
    return aimpl == o.aimpl && bxc == o.bxc && byc == o.byc && bzc == o.bzc && ifilx == o.ifilx &&
           ifily == o.ifily && ifilz == o.ifilz && version == o.version;
  }
};

}  // namespace

struct mrg_ctx {
  int device = 0, rank = 0, nranks = 1, nspecies = 2;
  GP g;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // particle kernel of the call in flight (= pass_ev of its species / phase)
  cudaEvent_t pass_ev[MRG_MAX_SPECIES][2][2] = {};   // [species][ipc != 0][begin, end]
  bool pass_timed[MRG_MAX_SPECIES][2] = {};
  cudaEvent_t user_ev[8] = {};
  // fields
  double* f12[12] = {};
  const double* fcur[12] = {};   // what k_blend reads: f12[k], or the caller's device array after mrg_bind_fields_device
  // lazily uploaded host fields (mrg_set_fields_lazy): the caller's host array and which interior z planes of it the
  // device copy f12[k] already holds; ensure_prep fetches the planes a preparation reads and nothing else
  const double* fhost[12] = {};
  bool flazy[12] = {};
  double b_dt = 0.0, b_aimpl = 0.0;   // lazily held fields with fhost[3..5] == nullptr: bx,by,bz planes are COMPUTED (k_prefld) when a preparation needs them
  std::vector<char> fplane[12];
  double* T1[6] = {};
  double* T2[6] = {};
  double* F6 = nullptr;
  double* tmp6[6] = {};    // unpack scratch for mrg_get_prepared_fields (lazy)
  unsigned long long field_version = 0;
  bool fields_set = false, prep_valid = false;
  PrepKey prep_key{};
  // planes of F6 (extended index k+2) the cached preparation covers; prep_full = all of them
  bool prep_full = true;
  std::vector<char> prep_planes;
  int* plane_lists = nullptr;    // device copy of the three plane lists of a restricted preparation
  long long prep_count = 0, prep_restricted_count = 0, prep_planes_sum = 0, compact_count = 0;
  // species
  Species sp[MRG_MAX_SPECIES];
  double* alt[6] = {};     // shared spare particle buffer (sort / download)
  int* alt_id = nullptr;
  long long alt_cap = 0;
  // scratch
  double* wk_partial = nullptr; long long wk_partial_cap = 0;
  double* wk2 = nullptr;
  int* sort_key = nullptr; long long sort_key_cap = 0;
  int* hist = nullptr; int* cursor = nullptr; long long ncell = 0;
  int* scan_tiles = nullptr; long long scan_tiles_cap = 0;
  unsigned* lcg_tab = nullptr;   // lambda^(j * 2048^t), t = 0..2, j < 2048 (lcg_pow_tab)
  unsigned* slab_bits = nullptr; int* slab_words = nullptr; long long slab_words_cap = 0;
  int* slab_list = nullptr; long long slab_list_cap = 0;
  int* slab_count = nullptr;
  int* slab_n_host = nullptr;    // pinned landing word of slab_count
  int num_sms = 148;
  double* h_pinned = nullptr; size_t h_pinned_bytes = 0;
  // nccl
  void* comm = nullptr;
  cudaStream_t cstream = nullptr;   // communication stream of the deferred mode (allreduce + fold overlap the next kernel)
  cudaEvent_t ev_kernel = nullptr;
  double* wk_pinned = nullptr;      // [MRG_MAX_SPECIES][2] wkix/wkih landing zone of the deferred mode
  // options / counters
  int opt_deposit = 2, opt_iters = 8, opt_group_min = 2, opt_tile = 1, opt_fused_keys = 1, opt_fused_sort = 1, opt_shard = 0;
  int opt_planes = -1;   // restricted field preparation: -1 = when nranks > 1, 0 = never, 1 = always
  int opt_defer = 0;
  int opt_split_push = 1;       // 0 = off, 1 = last species of the step, 2 = every species: particle kernel in two launches, the planes final after the first are pushed under the second
  cudaEvent_t ev_split = nullptr, ev_split_pre = nullptr, ev_split_b = nullptr;
  cudaStream_t sstream = nullptr;   // second launch of a split predictor
  int opt_peer_push_last = 296; // CTAs for the LAST species of a step (nothing overlaps its exchange: emfild reads all moments next, F:762-771); 0 = same as peer_push
  int opt_peer_push = 64;   // CTAs of the fused add+push kernel (0 = off): slab-wise exchange pushes the finished block into the peers' arrays over NVLink (when mapped) instead of ncclAllGather
  long long push_count = 0;
  long long split_count = 0;
  int opt_sink_share = 0;   // deferred D2H of the folded moments copies only this rank's z block (ranks of a node share the host arrays)
  int opt_kick = -1;     // drive-kick draws: -1 = by particle index when "shard" = 1 (no reference stream exists), else the reference's serial order; 0 / 1 force
  int opt_compact = -1;  // slab-wise moment exchange instead of the whole-grid allreduce: -1 = when possible, 0 = never
  double* halo_rx[2] = {nullptr, nullptr};   // staging of the two neighbour halos of the slab-wise exchange
  int opt_slab_n = 0, opt_slab_i = 0;   // "slab_of"/"slab_index": mrg_loadpt loads slab i of n whatever nranks is (sizing aid)
  long long launches = 0, h2d = 0, d2h = 0;
  double last_kernel_ms = 0.0;
  // per-phase device time of the mrg_fulmov calls (option "phases"): event pairs on the stream each phase runs on,
  // read back lazily (a pair is read right before it is recorded again, and by mrg_phase_ms)
  int opt_phases = 0;
  cudaEvent_t ph_ev[MRG_MAX_SPECIES][2][MRG_NPHASE_DETAIL][2] = {};
  bool ph_rec[MRG_MAX_SPECIES][2][MRG_NPHASE_DETAIL] = {};
  double ph_ms[MRG_NPHASE] = {};
  double ph_detail[MRG_MAX_SPECIES][2][MRG_NPHASE_DETAIL] = {};
  long long ph_calls = 0;
};

namespace {

int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

// ---- phase timers (option "phases") ------------------------------------------------------------
int phase_collect(mrg_ctx* c, int k, int ipc, int ph) {
  if (!c->ph_rec[k][ipc][ph]) return MRG_OK;
  c->ph_rec[k][ipc][ph] = false;
  CK(cudaEventSynchronize(c->ph_ev[k][ipc][ph][1]));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ph_ev[k][ipc][ph][0], c->ph_ev[k][ipc][ph][1]));
  if (ph < MRG_NPHASE) c->ph_ms[ph] += ms;
  c->ph_detail[k][ipc][ph] += ms;
  return MRG_OK;
}
struct PhaseScope {      // records begin on construction, end on done(); no-op unless the option is on
  mrg_ctx* c; int k, ipc, ph; cudaStream_t st; bool on;
  PhaseScope(mrg_ctx* c_, int k_, int ipc_, int ph_, cudaStream_t st_) : c(c_), k(k_), ipc(ipc_ != 0), ph(ph_), st(st_), on(c_->opt_phases != 0) {
    if (!on) return;
    phase_collect(c, k, ipc, ph);
    for (int e = 0; e < 2; e++)
      if (!c->ph_ev[k][ipc][ph][e]) cudaEventCreate(&c->ph_ev[k][ipc][ph][e]);
    cudaEventRecord(c->ph_ev[k][ipc][ph][0], st);
  }
  void done() {
    if (!on) return;
    cudaEventRecord(c->ph_ev[k][ipc][ph][1], st);
    c->ph_rec[k][ipc][ph] = true;
    on = false;
  }
  ~PhaseScope() { done(); }
};

// ---- TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link) ----
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn g_encode = nullptr;
int tensor_map_init() {
  if (g_encode) return MRG_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) return fail(MRG_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  g_encode = (TensorMapEncodeFn)fn;
  return MRG_OK;
}
// the six SoA rows of a particle set as a [6][cap] fp64 tensor; box = 32 slots x 6 rows
int particle_map(const double* base, long long cap, CUtensorMap* tm) {
  int rc = tensor_map_init();
  if (rc) return rc;
  const cuuint64_t gdim[2] = {(cuuint64_t)cap, 6};
  const cuuint64_t gstr[1] = {(cuuint64_t)cap * sizeof(double)};
  const cuuint32_t box[2] = {32, 6}, estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MRG_ERR_CUDA, "cuTensorMapEncodeTiled(particles) failed: " + std::to_string((int)r));
  return MRG_OK;
}
// an int32 array (ids, sort keys) as a 1-D tensor; box = 32 elements
int int_map(const int* base, long long n, CUtensorMap* tm) {
  int rc = tensor_map_init();
  if (rc) return rc;
  const cuuint64_t gdim[1] = {(cuuint64_t)std::max<long long>(n, 32)};
  const cuuint64_t gstr[1] = {0};          // rank 1 has no strides; the encoder still wants a pointer
  const cuuint32_t box[1] = {32}, estr[1] = {1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 1, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MRG_ERR_CUDA, "cuTensorMapEncodeTiled(int32) failed: " + std::to_string((int)r));
  return MRG_OK;
}

int ensure(mrg_ctx* c, void** p, long long* cap, long long need, size_t elem) {
  if (*cap >= need && *p) return MRG_OK;
  if (*p) CK(cudaFree(*p));
  *p = nullptr;
  long long ncap = std::max<long long>(need, 64);
  CK(cudaMalloc(p, (size_t)ncap * elem));
  *cap = ncap;
  (void)c;
  return MRG_OK;
}

int ensure_pinned(mrg_ctx* c, size_t bytes) {
  if (c->h_pinned_bytes >= bytes) return MRG_OK;
  if (c->h_pinned) CK(cudaFreeHost(c->h_pinned));
  c->h_pinned = nullptr;
  CK(cudaMallocHost((void**)&c->h_pinned, bytes));
  c->h_pinned_bytes = bytes;
  return MRG_OK;
}

int check_species(mrg_ctx* c, int ksp) {
  if (!c) return fail(MRG_ERR_ARG, "null context");
  if (ksp < 1 || ksp > c->nspecies) return fail(MRG_ERR_ARG, "ksp out of range (1-based species index)");
  return MRG_OK;
}

long long owned_count(long long npr, long long first, long long stride) {
  if (npr < first) return 0;
  return (npr - first) / stride + 1;   // l = first, first+stride, ... <= npr
}

int alloc_species(mrg_ctx* c, Species& s, long long n) {
  long long cap = ((n + 63) / 64) * 64 + 64;
  if (s.cap < cap) {
    // the six arrays of a set are rows of ONE block (row stride = cap), so that a 2-D TMA box can fetch the
    // same 32 slots of all six with one instruction (particle_map)
    if (s.d[0]) CK(cudaFree(s.d[0]));
    for (int k = 0; k < 6; k++) s.d[k] = nullptr;
    CK(cudaMalloc((void**)&s.d[0], (size_t)cap * 6 * sizeof(double)));
    for (int k = 1; k < 6; k++) s.d[k] = s.d[0] + (size_t)k * cap;
    s.cap = cap;
  }
  if (s.id) { CK(cudaFree(s.id)); s.id = nullptr; }
  s.n = n;
  s.index_valid = false;
  s.keys_valid = false;
  s.prekeys_valid = false;
  s.prescan_valid = false;
  s.fresh = false;
  s.zocc_valid = false;
  s.hull_valid = false;
  s.compact_ok = false;
  if (!s.cell_end) {
    CK(cudaMalloc((void**)&s.cell_end, (size_t)(c->ncell + 1) * sizeof(int)));
    CK(cudaMalloc((void**)&s.cell_end2, (size_t)(c->ncell + 1) * sizeof(int)));
    CK(cudaMalloc((void**)&s.hist, (size_t)(c->ncell + 1) * sizeof(int)));
  }
  if (!s.M4) {
    CK(cudaMalloc((void**)&s.M4, ((size_t)c->g.ntot * 4 + 2) * sizeof(double)));
    for (int k = 0; k < 4; k++) CK(cudaMalloc((void**)&s.out4[k], (size_t)c->g.ntot * sizeof(double)));
  }
  s.have_moments = false;
  return MRG_OK;
}

// spare particle buffer; exact=true keeps its capacity equal to the species'
// so that buffers (and their id arrays) can be swapped by mrg_sort
int ensure_alt(mrg_ctx* c, long long cap, bool exact) {
  if (exact ? (c->alt_cap == cap) : (c->alt_cap >= cap)) return MRG_OK;
  if (c->alt[0]) CK(cudaFree(c->alt[0]));
  for (int k = 0; k < 6; k++) c->alt[k] = nullptr;
  CK(cudaMalloc((void**)&c->alt[0], (size_t)cap * 6 * sizeof(double)));
  for (int k = 1; k < 6; k++) c->alt[k] = c->alt[0] + (size_t)k * cap;
  if (c->alt_id) CK(cudaFree(c->alt_id));
  c->alt_id = nullptr;
  CK(cudaMalloc((void**)&c->alt_id, (size_t)cap * sizeof(int)));
  c->alt_cap = cap;
  return MRG_OK;
}

ParticleSoA soa(const Species& s) {
  ParticleSoA P;
  P.x = s.d[0]; P.y = s.d[1]; P.z = s.d[2]; P.vx = s.d[3]; P.vy = s.d[4]; P.vz = s.d[5];
  P.id = s.id; P.n = s.n;
  return P;
}

// exclusive scan of n ints (in -> out, may alias), optional grand total (device int)
int scan_excl(mrg_ctx* c, const int* in, int* out, long long n, int* total_dev) {
  const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  int rc = ensure(c, (void**)&c->scan_tiles, &c->scan_tiles_cap, ntiles, sizeof(int));
  if (rc) return rc;
  k_scan_reduce<<<ntiles, SCAN_BLOCK, 0, c->stream>>>(in, n, c->scan_tiles); CKL(c);
  k_scan_tiles<<<1, SCAN_BLOCK, 0, c->stream>>>(c->scan_tiles, ntiles, total_dev); CKL(c);
  k_scan_apply<<<ntiles, SCAN_BLOCK, 0, c->stream>>>(in, n, c->scan_tiles, out); CKL(c);
  return MRG_OK;
}

// host side of a deferred ipc>=1 call: wait for the species' moment sum and hand out wkix/wkih
int complete_moments(mrg_ctx* c, int k) {
  Species& s = c->sp[k];
  if (!s.pending) return MRG_OK;
  CK(cudaEventSynchronize(s.done));
  if (s.wk_user[0]) *s.wk_user[0] = c->wk_pinned[2 * k + 0];
  if (s.wk_user[1]) *s.wk_user[1] = c->wk_pinned[2 * k + 1];
  s.wk_user[0] = s.wk_user[1] = nullptr;
  s.pending = false;
  return MRG_OK;
}

// ---- plane tracking -----------------------------------------------------------
bool tracking(const mrg_ctx* c) { return c->opt_planes == 1 || (c->opt_planes < 0 && c->nranks > 1); }
int zocc_words(const mrg_ctx* c) { return (c->g.mz + 1 + 31) / 32 + 1; }   // plane bitmap + one word of flags (bit 0: |vz| dt >= hz seen)

// zero the species' plane bitmap before a kernel that marks it (nullptr when tracking is off)
int zocc_begin(mrg_ctx* c, Species& s, unsigned** out) {
  *out = nullptr;
  s.zocc_valid = false;
  if (!tracking(c)) return MRG_OK;
  const int nw = zocc_words(c);
  if (!s.zocc) CK(cudaMalloc((void**)&s.zocc, (size_t)nw * sizeof(unsigned)));
  CK(cudaMemsetAsync(s.zocc, 0, (size_t)nw * sizeof(unsigned), c->stream));
  *out = s.zocc;
  return MRG_OK;
}
// queue the copy of the bitmap to the host; the caller synchronises the stream afterwards
int zocc_fetch(mrg_ctx* c, Species& s, double lookahead, bool approx) {
  const int nw = zocc_words(c);
  s.zocc_host.assign(nw, 0u);
  CK(cudaMemcpyAsync(s.zocc_host.data(), s.zocc, (size_t)nw * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  s.zocc_valid = true;
  s.zocc_approx = approx;
  s.zocc_lookahead = lookahead;
  return MRG_OK;
}
bool occ_bit(const Species& s, int kp) { return (s.zocc_host[kp >> 5] >> (kp & 31)) & 1u; }
// the corrector saw a particle that moves a whole plane or more per step: the +-1 plane bound of the recorded planes and
// the strips of the slab-wise exchange no longer hold, so the record is not used (full preparation, whole-grid allreduce)
bool occ_violated(const Species& s) { return !s.zocc_host.empty() && (s.zocc_host.back() & 1u); }
// OR the species' gather planes into occ[0..mz]
void add_occupancy(const Species& s, int mz, std::vector<char>& occ) {
  if (s.n == 0) return;
  for (int kp = 0; kp <= mz; kp++) {
    if (!occ_bit(s, kp)) continue;
    occ[kp] = 1;
    if (s.zocc_approx && kp < mz) {        // +-1 plane, and the clamp plane kp = mz next to the seam (F:1177 with z within 1e-9 hz of zmax - hz/2)
      occ[(kp + mz - 1) % mz] = 1;
      occ[(kp + 1) % mz] = 1;
      if (kp <= 1 || kp >= mz - 2) occ[mz] = 1;
    }
  }
}
// key-plane occupancy of the order being built -> bitmap on the host after the caller's synchronize
int kocc_begin(mrg_ctx* c, Species& s, const int* cell_start) {
  if (!tracking(c)) return MRG_OK;
  const int nw = zocc_words(c);
  if (!s.kocc) CK(cudaMalloc((void**)&s.kocc, (size_t)nw * sizeof(unsigned)));
  CK(cudaMemsetAsync(s.kocc, 0, (size_t)nw * sizeof(unsigned), c->stream));
  k_plane_occupancy<<<grid_for(c->g.mz, 128), 128, 0, c->stream>>>(cell_start, c->g.mx * c->g.my, c->g.mz, s.kocc); CKL(c);
  s.kocc_host.assign(nw, 0u);
  CK(cudaMemcpyAsync(s.kocc_host.data(), s.kocc, (size_t)nw * sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  s.kocc_pending = true;
  return MRG_OK;
}
// after the synchronize: smallest arc of planes that holds every particle of the new order
void kocc_finish(mrg_ctx* c, Species& s) {
  s.hull_valid = false;
  if (!s.kocc_pending) return;
  s.kocc_pending = false;
  const int mz = c->g.mz;
  auto bit = [&](int k) { return (s.kocc_host[k >> 5] >> (k & 31)) & 1u; };
  int best_len = 0, best_start = 0;        // longest circular run of empty planes
  for (int k0 = 0; k0 < mz; k0++) {
    if (bit(k0) || !bit((k0 + mz - 1) % mz)) continue;   // runs start right after an occupied plane
    int len = 0;
    while (len < mz && !bit((k0 + len) % mz)) len++;
    if (len > best_len) { best_len = len; best_start = k0; }
  }
  bool any = false;
  for (int k = 0; k < mz; k++) any = any || bit(k);
  if (!any) return;
  s.kz0 = (best_start + best_len) % mz;
  s.nkz = mz - best_len;
  s.hull_valid = true;
}

// Plane sets of a restricted preparation from the occupancy occ[kp], kp = 0..mz (see ensure_prep):
//   G      extended planes (k+2) of F6 to finalize: k = kp-1..kp+1
//   listGI interior planes of G: where the filter sweeps run
//   listB  planes to blend: listGI widened by 2 either side (periodic z sweep) + periodic images of G's ghosts
//   listG  G as a list, plus the guard planes (| PLANE_GUARD) just outside it
struct PlaneSets {
  std::vector<char> G;
  std::vector<int> listB, listGI, listG;
};
void plane_sets(int mz, const std::vector<char>& occ, PlaneSets& ps) {
  const int nz = mz + 4;
  ps.G.assign(nz, 0);
  ps.listB.clear(); ps.listGI.clear(); ps.listG.clear();
  for (int kp = 0; kp <= mz; kp++)
    if (occ[kp]) ps.G[kp + 1] = ps.G[kp + 2] = ps.G[kp + 3] = 1;
  std::vector<char> GI(mz, 0), B(mz, 0);
  for (int k = 0; k < mz; k++) GI[k] = ps.G[k + 2];
  for (int k = 0; k < mz; k++)
    for (int d = -2; d <= 2; d++) B[k] |= GI[((k + d) % mz + mz) % mz];
  if (ps.G[1]) B[mz - 1] = 1;              // ghost plane k = -1 copies the blend of plane mz-1
  if (ps.G[mz + 2]) B[0] = 1;              // k = mz   <- plane 0
  if (ps.G[mz + 3]) B[1] = 1;              // k = mz+1 <- plane 1
  for (int k = 0; k < mz; k++) { if (B[k]) ps.listB.push_back(k); if (GI[k]) ps.listGI.push_back(k); }
  for (int e = 0; e < nz; e++) {
    if (ps.G[e]) ps.listG.push_back(e);
    else if ((e > 0 && ps.G[e - 1]) || (e + 1 < nz && ps.G[e + 1])) ps.listG.push_back(e | PLANE_GUARD);
  }
}

// ---- slab-wise moment exchange ---------------------------------------------------
// With z-slab ownership a rank deposits only near its own block of L = mz/N planes, so the rank sum of the raw
// moments (F:2379-2384, 2533) needs no whole-grid allreduce: every rank adds the two HALO-wide strips its ring
// neighbours deposited into its block (ncclSend/ncclRecv), then the complete blocks are all-gathered in place and
// the four z ghost planes broadcast by the ranks that own them -- about half the bytes of the allreduce.  Ghost
// planes are ordinary planes of the raw arrays (deposits next to the periodic seam go to k = -1 or k = mz, the
// fold maps them later), but a particle that crossed the seam deposits next to plane 0 although its rank owns the
// last block, so the "upper" strip of rank N-1 is the bottom of the array and the "lower" strip of rank 0 its top.
// Eligibility is decided from the recorded gather planes (a deposit lies within 2 planes of the gather plane for
// |vz| dt < hz) and agreed between the ranks in the preceding corrector call.
constexpr int HALO_H = 6;                 // planes beyond the own block a rank may deposit into
constexpr int HALO_PLANES = HALO_H + 2;   // strip width: + the two ghost planes at the seam
// first extended plane (k+2) of the four HALO_PLANES-wide strips of rank r: sent to the upper / lower ring neighbour,
// and where the strips received from the lower / upper neighbour are added
void compact_layout(int mz, int N, int r, int lay[4]) {
  const int L = mz / N;
  lay[0] = (r < N - 1) ? 2 + (r + 1) * L : 0;
  lay[1] = (r > 0) ? 2 + r * L - HALO_PLANES : mz + 4 - HALO_PLANES;
  lay[2] = (r > 0) ? 2 + r * L : 0;
  lay[3] = (r < N - 1) ? 2 + (r + 1) * L - HALO_PLANES : mz + 4 - HALO_PLANES;
}
// may rank r of N, whose (widened) gather planes are occ[0..mz], take part in the slab-wise exchange?
bool compact_planes_ok(int mz, int N, int r, const std::vector<char>& occ) {
  const int L = mz / N, lo = r * L, hi = lo + L - 1;
  for (int kp = 0; kp < mz; kp++) {
    if (!occ[kp] || (kp >= lo && kp <= hi)) continue;
    const int d = std::min(((lo - kp) % mz + mz) % mz, ((kp - hi) % mz + mz) % mz);
    if (d > HALO_H - 2) return false;
  }
  if (occ[mz] && r != 0 && r != N - 1) return false;
  return true;
}
bool compact_possible(const mrg_ctx* c) {
  const int N = c->nranks, mz = c->g.mz;
  return N > 1 && c->opt_compact != 0 && tracking(c) && mz % N == 0 && mz / N >= 2 * HALO_PLANES && g_nccl.AllGather &&
         g_nccl.Broadcast && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd;
}
bool compact_eligible(const mrg_ctx* c, const Species& s, double hdt) {
  if (s.n == 0) return true;
  if (!(s.zocc_valid && s.zocc_lookahead == hdt) || occ_violated(s)) return false;
  const int mz = c->g.mz;
  std::vector<char> occ(mz + 1, 0);
  add_occupancy(s, mz, occ);
  return compact_planes_ok(mz, c->nranks, c->rank, occ);
}
// the exchange itself, on stream ms; M4 = raw moments [nz planes][nxy][4] + (wkix, wkih)
int compact_sum(mrg_ctx* c, Species& s, cudaStream_t ms) {
  double* M4 = s.M4;
  const int ks = (int)(&s - c->sp);
  const GP& g = c->g;
  const int N = c->nranks, r = c->rank, L = g.mz / N;
  const size_t P = (size_t)g.nxy * 4;                       // doubles per plane
  const size_t cnt = (size_t)HALO_PLANES * P;
  const int up = (r + 1) % N, dn = (r + N - 1) % N;
  for (int k = 0; k < 2; k++)
    if (!c->halo_rx[k]) CK(cudaMalloc((void**)&c->halo_rx[k], cnt * sizeof(double)));
  int lay[4];
  compact_layout(g.mz, N, r, lay);
  const size_t send_up = lay[0], send_dn = lay[1], add_lo = lay[2], add_hi = lay[3];   // first plane of each strip
  PhaseScope ph_strips(c, ks, 1, MRG_PH_STRIPS, ms);
  int n = g_nccl.GroupStart();
  if (!n) n = g_nccl.Send(M4 + send_up * P, cnt, kNcclFloat64, up, c->comm, ms);
  if (!n) n = g_nccl.Send(M4 + send_dn * P, cnt, kNcclFloat64, dn, c->comm, ms);
  if (!n) n = g_nccl.Recv(c->halo_rx[0], cnt, kNcclFloat64, dn, c->comm, ms);   // the lower neighbour's upward strip
  if (!n) n = g_nccl.Recv(c->halo_rx[1], cnt, kNcclFloat64, up, c->comm, ms);   // the upper neighbour's downward strip
  const int e = g_nccl.GroupEnd();
  if (n || e) return fail(MRG_ERR_NCCL, "halo exchange: " + nccl_err(n ? n : e));
  ph_strips.done();
  PhaseScope ph_push(c, ks, 1, MRG_PH_PUSH, ms);
  if (s.npeer == N - 1 && c->opt_peer_push) {
    // fused add + push over NVLink peer memory (k_add_push); the allreduce below is the completion barrier
    const size_t e0 = (r == 0) ? 0 : (size_t)(2 + r * L), e1 = (r == N - 1) ? (size_t)g.nz : (size_t)(2 + (r + 1) * L);
    PeerPtrs pp;
    pp.n = 0;
    for (int q = 0; q < N; q++)
      if (q != r) pp.p[pp.n++] = s.peerM4[q];
    for (int q = pp.n; q < 8; q++) pp.p[q] = nullptr;
    // a small grid: 64 CTAs keep NVLink busy (fire-and-forget 128-bit peer stores) and leave the SMs to the other species'
    // particle kernel that runs next to this exchange in deferred mode (1184 CTAs cost that kernel 0.6 ms at 8 GPUs, measured)
    const int ctas = (ks == c->nspecies - 1 && c->opt_peer_push_last > 0) ? c->opt_peer_push_last : c->opt_peer_push;
    k_add_push<<<std::max(1, ctas), 256, 0, ms>>>(M4, e0 * P, (e1 - e0) * P, add_lo * P, c->halo_rx[0], add_hi * P, c->halo_rx[1], cnt, pp,
                                                  (size_t)s.early0 * P, (size_t)s.early_n * P); CKL(c);
    c->push_count++;
    n = 0;
  } else {
    k_add_strips<<<grid_for((long long)cnt, 256), 256, 0, ms>>>(M4 + add_lo * P, c->halo_rx[0], M4 + add_hi * P, c->halo_rx[1], (long long)cnt); CKL(c);
    n = g_nccl.AllGather(M4 + (size_t)(2 + r * L) * P, M4 + 2 * P, (size_t)L * P, kNcclFloat64, c->comm, ms);
    if (!n) n = g_nccl.Broadcast(M4, M4, 2 * P, kNcclFloat64, 0, c->comm, ms);
    if (!n) n = g_nccl.Broadcast(M4 + (size_t)(g.mz + 2) * P, M4 + (size_t)(g.mz + 2) * P, 2 * P, kNcclFloat64, N - 1, c->comm, ms);
  }
  ph_push.done();
  PhaseScope ph_barrier(c, ks, 1, MRG_PH_BARRIER, ms);
  if (!n) n = g_nccl.AllReduce(M4 + (size_t)g.ntot * 4, M4 + (size_t)g.ntot * 4, 2, kNcclFloat64, kNcclSum, c->comm, ms);
  if (n) return fail(MRG_ERR_NCCL, "slab exchange: " + nccl_err(n));
  ph_barrier.done();
  return MRG_OK;
}

// Option "sink_share": the ranks of a node share the host arrays the folded moments land in, so each rank copies only
// its own block of extended z planes [e0, e1) -- rank 0 takes the two ghost planes below, the last rank the two above --
// and the blocks tile the array exactly once.
void sink_block(const mrg_ctx* c, size_t* off, size_t* cnt) {
  const GP& g = c->g;
  if (!c->opt_sink_share || c->nranks == 1) { *off = 0; *cnt = (size_t)g.ntot; return; }
  const int N = c->nranks, r = c->rank;
  const int k0 = (int)((long long)g.mz * r / N), k1 = (int)((long long)g.mz * (r + 1) / N);
  const int e0 = (r == 0) ? 0 : k0 + 2, e1 = (r == N - 1) ? g.nz : k1 + 2;
  *off = (size_t)e0 * g.nxy;
  *cnt = (size_t)(e1 - e0) * g.nxy;
}

// F:1127-1148 on the device, cached on (fields version, aimpl, dc, ifil*).
//
// Restricted preparation: a particle whose gather cell lies on plane kp reads F6 on planes kp-1..kp+1 only
// (F:1217-1270).  When every species' plane bitmap is valid for this hdt, only those planes are finalized
// (set G, extended index k+2), the filters run on the interior planes of G (x and y sweeps stay inside a
// plane), the z sweep needs the blend two planes either side (periodic, F:7365-7395), and the ghost planes
// k = -1, mz, mz+1 of G need the blend of their periodic images (F:3088-3148).  The planes just outside G are
// filled with NaN.  The recorded planes are those of the gather positions themselves, whatever the speeds; only
// the drive kick's nearest-node look-up at the NEW position (F:1347-1354) relies on |vz| hdt < hz to stay
// inside the widened set -- true for |v| < c when c dt < hz (the reference runs hz = 7.5, dt = 1.2).
int ensure_prep(mrg_ctx* c, const mrg_step_params* p, int ksp) {
  if (!c->fields_set) return fail(MRG_ERR_STATE, "mrg_set_fields has not been called");
  if (p->ifilx < 0 || p->ifily < 0 || p->ifilz < 0) return fail(MRG_ERR_ARG, "negative filter count");
  PrepKey key{p->aimpl, p->bxc, p->byc, p->bzc, p->ifilx, p->ifily, p->ifilz, c->field_version};
  const GP& g = c->g;
  const int mz = g.mz, nz = g.nz;
  auto usable = [&](const Species& s) { return s.n == 0 || (s.zocc_valid && s.zocc_lookahead == p->hdt && !occ_violated(s)); };
  if (c->prep_valid && key == c->prep_key) {
    if (c->prep_full) return MRG_OK;
    bool covered = ksp >= 1;          // ksp = 0: the caller wants every plane
    if (covered) {
      const Species& s = c->sp[ksp - 1];
      covered = usable(s);
      if (covered && s.n > 0) {
        std::vector<char> occ(mz + 1, 0);
        add_occupancy(s, mz, occ);
        for (int kp = 0; kp <= mz && covered; kp++)
          if (occ[kp]) covered = c->prep_planes[kp + 1] && c->prep_planes[kp + 2] && c->prep_planes[kp + 3];
      }
    }
    if (covered) return MRG_OK;
  }
  // which planes?
  bool restricted = ksp >= 1 && tracking(c) && p->ifilz <= 1;
  for (int k = 0; k < c->nspecies && restricted; k++) restricted = usable(c->sp[k]);
  PlaneSets ps;
  ps.G.assign(nz, 0);
  if (restricted) {
    std::vector<char> occ(mz + 1, 0);
    for (int k = 0; k < c->nspecies; k++) add_occupancy(c->sp[k], mz, occ);
    plane_sets(mz, occ, ps);
    if (ps.listB.empty() || (long long)ps.listB.size() * 4 > (long long)mz * 3) restricted = false;   // not worth it
  }
  const std::vector<int>&listB = ps.listB, &listGI = ps.listGI, &listG = ps.listG;
  const std::vector<char>& G = ps.G;
  const int B = 256;
  const int *dB = nullptr, *dGI = nullptr, *dG = nullptr;
  int nB = mz, nGI = mz, nG = nz;
  auto in_b = [&](int k) { return restricted ? std::binary_search(listB.begin(), listB.end(), k) : true; };
  // bx,by,bz held as "to be computed" (mrg_update_b on lazily held fields): the planes the blend needs and the device has not made yet
  const bool bcomp = c->flazy[3] && !c->fhost[3];
  std::vector<int> bmiss;
  std::vector<char> bm(mz, 0);
  if (bcomp)
    for (int k = 0; k < mz; k++)
      if (in_b(k) && !c->fplane[3][k]) { bmiss.push_back(k); bm[k] = 1; }
  const int* dBm = nullptr;
  if (restricted || !bmiss.empty()) {
    if (!c->plane_lists) CK(cudaMalloc((void**)&c->plane_lists, (size_t)(4 * (nz + 4)) * sizeof(int)));
    std::vector<int> all(4 * (nz + 4), 0);
    std::copy(listB.begin(), listB.end(), all.begin());
    std::copy(listGI.begin(), listGI.end(), all.begin() + (nz + 4));
    std::copy(listG.begin(), listG.end(), all.begin() + 2 * (nz + 4));
    std::copy(bmiss.begin(), bmiss.end(), all.begin() + 3 * (nz + 4));
    CK(cudaMemcpyAsync(c->plane_lists, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));   // `all` is pageable and goes out of scope
    dBm = c->plane_lists + 3 * (nz + 4);
  }
  if (restricted) {
    dB = c->plane_lists; dGI = c->plane_lists + (nz + 4); dG = c->plane_lists + 2 * (nz + 4);
    nB = (int)listB.size(); nGI = (int)listGI.size(); nG = (int)listG.size();
  }
  // lazily held host fields: fetch the interior planes the blend is about to read (F:1127-1139 reads k = 0..mz-1 only) --
  // and, for ex..ez / ex0..ez0, the planes k, k+-1 (periodic) the B update of the missing bx,by,bz planes reads (F:3834-3880)
  for (int a = 0; a < 12; a++) {
    if (!c->flazy[a] || !c->fhost[a]) continue;
    std::vector<char>& have = c->fplane[a];
    const bool for_b = !bmiss.empty() && (a <= 2 || (a >= 6 && a <= 8));
    auto need = [&](int k) { return in_b(k) || (for_b && (bm[k] || bm[k + 1 < mz ? k + 1 : 0] || bm[k > 0 ? k - 1 : mz - 1])); };
    for (int k = 0; k < mz;) {
      if (have[k] || !need(k)) { k++; continue; }
      int k1 = k;
      while (k1 < mz && !have[k1] && need(k1)) { have[k1] = 1; k1++; }
      const size_t off = (size_t)(k + 2) * g.nxy, cnt = (size_t)(k1 - k) * g.nxy;
      CK(cudaMemcpyAsync(c->f12[a] + off, c->fhost[a] + off, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      c->h2d += (long long)(cnt * sizeof(double));
      k = k1;
    }
  }
  const long long per = (long long)g.mx * (g.my + 1);
  CPtr12 f; for (int k = 0; k < 12; k++) f.p[k] = c->fcur[k];
  if (!bmiss.empty()) {      // prefld / emfild's B update on exactly those planes (bit-identical to the host's arrays)
    k_prefld<<<grid_for(per * (long long)bmiss.size(), B), B, 0, c->stream>>>(g, f, c->f12[3], c->f12[4], c->f12[5], c->b_aimpl, 1.0 - c->b_aimpl,
                                                                       c->b_dt, 2.0 * g.hx, 2.0 * g.hy, 2.0 * g.hz, dBm, (int)bmiss.size()); CKL(c);
    for (int k : bmiss) c->fplane[3][k] = c->fplane[4][k] = c->fplane[5][k] = 1;
  }
  Ptr6 T1, T2; for (int k = 0; k < 6; k++) { T1.p[k] = c->T1[k]; T2.p[k] = c->T2[k]; }
  const double om = 1.0 - p->aimpl;
  // blend (+ first z sweep) -> remaining sweeps -> finalize (+ last y sweep): three kernels for the usual ifil* = 1
  Ptr6 src = T1, dst = T2;
  auto as_const = [](const Ptr6& q) { CPtr6 r; for (int k = 0; k < 6; k++) r.p[k] = q.p[k]; return r; };
  if (p->ifilz > 0) { k_blend_filter_z<<<grid_for(per * nGI, B), B, 0, c->stream>>>(g, f, src, p->aimpl, om, p->bxc, p->byc, p->bzc, dGI, nGI); CKL(c); }
  else { k_blend<<<grid_for(per * nB, B), B, 0, c->stream>>>(g, f, src, p->aimpl, om, p->bxc, p->byc, p->bzc, dB, nB); CKL(c); }
  for (int n = 1; n < p->ifilz; n++) { k_filter<2><<<grid_for(per * nGI, B), B, 0, c->stream>>>(g, as_const(src), dst, dGI, nGI); CKL(c); std::swap(src, dst); }
  for (int n = 0; n < p->ifilx; n++) { k_filter<0><<<grid_for(per * nGI, B), B, 0, c->stream>>>(g, as_const(src), dst, dGI, nGI); CKL(c); std::swap(src, dst); }
  for (int n = 1; n < p->ifily; n++) { k_filter<1><<<grid_for(per * nGI, B), B, 0, c->stream>>>(g, as_const(src), dst, dGI, nGI); CKL(c); std::swap(src, dst); }
  if (p->ifily > 0) k_finalize<true><<<grid_for((long long)g.nxy * nG, B), B, 0, c->stream>>>(g, f, as_const(src), c->F6, p->aimpl, om, p->bxc, p->byc, p->bzc, dG, nG);
  else k_finalize<false><<<grid_for((long long)g.nxy * nG, B), B, 0, c->stream>>>(g, f, as_const(src), c->F6, p->aimpl, om, p->bxc, p->byc, p->bzc, dG, nG);
  CKL(c);
  c->prep_key = key;
  c->prep_valid = true;
  c->prep_full = !restricted;
  if (restricted) c->prep_planes.assign(G.begin(), G.end());
  else c->prep_planes.assign(nz, 1);
  c->prep_count++;
  if (restricted) { c->prep_restricted_count++; c->prep_planes_sum += nG; }
  return MRG_OK;
}

// Simpson table fv2 of loadpt, F:8885-8909 (host side of mrg_loadpt)
void loadpt_table(double vth, double vdr, double fv2[101], double* v2, double* dv2) {
  const double vrg1 = vdr / vth;
  double vv = std::max(-3.0, -vrg1);
  const double dv = (3.0 - vv) / 100.0;
  *v2 = vv * vth;
  *dv2 = dv * vth;
  auto fun2 = [vrg1](double v) { return exp(-(v * v)) * (v + vrg1); };   // F:9195
  fv2[0] = 0.0;
  for (int j = 1; j <= 100; j++) {
    double s = 0.0;
    const double sdv = dv / 1000.0;
    for (int k = 1; k <= 500; k++) {
      vv = vv + 2.0 * sdv;
      s = s + 4.0 * fun2(vv - sdv) + 2.0 * fun2(vv);
    }
    s = (s + 4.0 * fun2(vv + sdv) + fun2(vv + 2.0 * sdv)) * sdv / 3.0;
    fv2[j] = fv2[j - 1] + s;
  }
  const double norm = fv2[100];
  for (int j = 0; j <= 100; j++) fv2[j] = fv2[j] / norm;
}

}  // namespace

// ===========================================================================
extern "C" {

const char* mrg_last_error(void) { return g_err.c_str(); }

const char* mrg_build_info(void) {
  return "mrg_fulmov sm_100a; nvcc " __DATE__ "; fp64; deposit modes 0/1/2";
}

int mrg_create(mrg_ctx** out, int32_t mx, int32_t my, int32_t mz, double xmax, double ymax, double zmax,
               int32_t nspecies, int32_t rank, int32_t nranks, int32_t device) {
  if (!out) return fail(MRG_ERR_ARG, "null output pointer");
  *out = nullptr;
  if (mx < 4 || my < 2 || mz < 4) return fail(MRG_ERR_ARG, "grid too small (need mx,mz >= 4, my >= 2)");
  if (nspecies < 1 || nspecies > MRG_MAX_SPECIES) return fail(MRG_ERR_ARG, "nspecies out of range");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(MRG_ERR_ARG, "bad rank/nranks");
  if (!(xmax > 0 && ymax > 0 && zmax > 0)) return fail(MRG_ERR_ARG, "box sizes must be positive");
  const long long ntot = (long long)(mx + 4) * (my + 3) * (mz + 4);
  if (ntot * 6 >= (1LL << 31)) return fail(MRG_ERR_ARG, "grid too large for 32-bit node indexing");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(MRG_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(MRG_ERR_ARG, "device ordinal out of range");
  CK(cudaSetDevice(device));
  mrg_ctx* c = new mrg_ctx();
  c->device = device; c->rank = rank; c->nranks = nranks; c->nspecies = nspecies;
  GP& g = c->g;
  g.mx = mx; g.my = my; g.mz = mz;
  g.nx = mx + 4; g.ny = my + 3; g.nz = mz + 4; g.nxy = g.nx * g.ny; g.ntot = ntot;
  g.xmax = xmax; g.ymax = ymax; g.zmax = zmax;
  g.hx = xmax / mx; g.hy = ymax / my; g.hz = zmax / mz;                 // F:8454,8467,8484
  g.hxi = 0.9999999999999 / g.hx; g.hyi = 0.9999999999999 / g.hy; g.hzi = 0.9999999999999 / g.hz;  // F:8567-8569
  g.xmaxe = 0.9999999999999 * xmax; g.zmaxe = 0.9999999999999 * zmax;    // F:8575,8577
  g.xlo = -(g.hx / 2); g.xhi = xmax - g.hx / 2;                          // F:1856-1862
  g.zlo = -(g.hz / 2); g.zhi = zmax - g.hz / 2;
  g.ymax2 = 2.0 * ymax;
  auto hi32 = [](double v) { long long b; memcpy(&b, &v, 8); return (int)(b >> 32); };
  g.xhi_h = hi32(g.xhi); g.xlo_h = hi32(g.xlo); g.ymax_h = hi32(g.ymax); g.zhi_h = hi32(g.zhi); g.zlo_h = hi32(g.zlo);
  g.kz0 = 0; g.nkz = mz;
  c->ncell = (long long)mx * my * mz;
  CK(cudaFuncSetAttribute(k_predict_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM_BYTES));
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&c->cstream, cudaStreamNonBlocking, hi));   // the moment sum must not queue behind particle CTAs
  }
  CK(cudaEventCreateWithFlags(&c->ev_kernel, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_split, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_split_pre, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_split_b, cudaEventDisableTiming));
  CK(cudaStreamCreateWithFlags(&c->sstream, cudaStreamNonBlocking));
  for (auto& sp : c->sp) CK(cudaEventCreateWithFlags(&sp.done, cudaEventDisableTiming));
  CK(cudaMallocHost((void**)&c->wk_pinned, MRG_MAX_SPECIES * 2 * sizeof(double)));
  CK(cudaMallocHost((void**)&c->slab_n_host, sizeof(int)));
  *c->slab_n_host = 0;
  CK(cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device));
  for (int k = 0; k < MRG_MAX_SPECIES; k++)
    for (int q = 0; q < 4; q++) CK(cudaEventCreate(&c->pass_ev[k][q >> 1][q & 1]));
  for (int k = 0; k < 8; k++) CK(cudaEventCreate(&c->user_ev[k]));
  const size_t gb = (size_t)ntot * sizeof(double);
  for (int k = 0; k < 12; k++) { CK(cudaMalloc((void**)&c->f12[k], gb)); CK(cudaMemsetAsync(c->f12[k], 0, gb, c->stream)); c->fcur[k] = c->f12[k]; }
  for (int k = 0; k < 6; k++) {
    CK(cudaMalloc((void**)&c->T1[k], gb)); CK(cudaMalloc((void**)&c->T2[k], gb));
    CK(cudaMemsetAsync(c->T1[k], 0, gb, c->stream)); CK(cudaMemsetAsync(c->T2[k], 0, gb, c->stream));
  }
  CK(cudaMalloc((void**)&c->F6, gb * 6));
  CK(cudaMalloc((void**)&c->wk2, 3 * sizeof(double)));   // wkix, wkih + the ranks' vote on the slab-wise exchange
  {
    std::vector<unsigned> tab(3 * 2048);
    unsigned step = 48828125u;                     // lambda^(2048^t)
    for (int t = 0; t < 3; t++) {
      unsigned v = 1u;
      for (int j = 0; j < 2048; j++) { tab[t * 2048 + j] = v; v *= step; }
      step = v;                                    // = step^2048
    }
    CK(cudaMalloc((void**)&c->lcg_tab, tab.size() * sizeof(unsigned)));
    CK(cudaMemcpy(c->lcg_tab, tab.data(), tab.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc((void**)&c->slab_count, sizeof(int)));
  CK(cudaMalloc((void**)&c->hist, (size_t)(c->ncell + 1) * sizeof(int)));
  CK(cudaMalloc((void**)&c->cursor, (size_t)(c->ncell + 1) * sizeof(int)));
  CK(cudaStreamSynchronize(c->stream));
  *out = c;
  return MRG_OK;
}

int mrg_destroy(mrg_ctx* c) {
  if (!c) return MRG_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->cstream) cudaStreamSynchronize(c->cstream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  for (int k = 0; k < 12; k++) cudaFree(c->f12[k]);
  for (int k = 0; k < 6; k++) { cudaFree(c->T1[k]); cudaFree(c->T2[k]); cudaFree(c->tmp6[k]); }
  cudaFree(c->alt[0]);
  cudaFree(c->F6); cudaFree(c->alt_id);
  for (auto& s : c->sp) {
    for (int q = 0; q < 8; q++) if (s.peerM4[q]) cudaIpcCloseMemHandle(s.peerM4[q]);
    cudaFree(s.d[0]);
    cudaFree(s.id); cudaFree(s.M4); cudaFree(s.cell_end); cudaFree(s.cell_end2); cudaFree(s.key); cudaFree(s.hist);
    cudaFree(s.zocc); cudaFree(s.kocc);
    if (s.done) cudaEventDestroy(s.done);
    for (int k = 0; k < 4; k++) cudaFree(s.out4[k]);
  }
  cudaFree(c->wk_partial); cudaFree(c->wk2); cudaFree(c->sort_key); cudaFree(c->hist); cudaFree(c->cursor);
  cudaFree(c->lcg_tab); cudaFree(c->halo_rx[0]); cudaFree(c->halo_rx[1]);
  cudaFree(c->scan_tiles); cudaFree(c->slab_bits); cudaFree(c->slab_words); cudaFree(c->slab_list); cudaFree(c->slab_count);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->wk_pinned) cudaFreeHost(c->wk_pinned);
  if (c->slab_n_host) cudaFreeHost(c->slab_n_host);
  cudaFree(c->plane_lists);
  if (c->ev_kernel) cudaEventDestroy(c->ev_kernel);
  if (c->ev_split) cudaEventDestroy(c->ev_split);
  if (c->ev_split_pre) cudaEventDestroy(c->ev_split_pre);
  if (c->ev_split_b) cudaEventDestroy(c->ev_split_b);
  if (c->sstream) cudaStreamDestroy(c->sstream);
  if (c->cstream) cudaStreamDestroy(c->cstream);
  for (int k = 0; k < MRG_MAX_SPECIES; k++)
    for (int q = 0; q < 4; q++) if (c->pass_ev[k][q >> 1][q & 1]) cudaEventDestroy(c->pass_ev[k][q >> 1][q & 1]);
  for (int k = 0; k < 8; k++) cudaEventDestroy(c->user_ev[k]);
  for (int k = 0; k < MRG_MAX_SPECIES; k++)
    for (int i = 0; i < 2; i++)
      for (int ph = 0; ph < MRG_NPHASE_DETAIL; ph++)
        for (int e = 0; e < 2; e++) if (c->ph_ev[k][i][ph][e]) cudaEventDestroy(c->ph_ev[k][i][ph][e]);
  cudaStreamDestroy(c->stream);
  delete c;
  return MRG_OK;
}

int mrg_comm_unique_id(unsigned char id[MRG_UNIQUE_ID_BYTES]) {
  if (!id) return fail(MRG_ERR_ARG, "null id");
  int rc = nccl_load();
  if (rc) return rc;
  Uid u;
  int n = g_nccl.GetUniqueId(&u);
  if (n != 0) return fail(MRG_ERR_NCCL, "ncclGetUniqueId: " + nccl_err(n));
  memcpy(id, u.internal, MRG_UNIQUE_ID_BYTES);
  return MRG_OK;
}

int mrg_comm_init(mrg_ctx* c, const unsigned char id[MRG_UNIQUE_ID_BYTES]) {
  if (!c || !id) return fail(MRG_ERR_ARG, "null argument");
  if (c->nranks == 1) return MRG_OK;
  int rc = nccl_load();
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  Uid u;
  memcpy(u.internal, id, MRG_UNIQUE_ID_BYTES);
  int n = g_nccl.CommInitRank(&c->comm, c->nranks, u, c->rank);
  if (n != 0) return fail(MRG_ERR_NCCL, "ncclCommInitRank: " + nccl_err(n));
  return MRG_OK;
}

int64_t mrg_num_local(mrg_ctx* c, int32_t ksp) {
  if (check_species(c, ksp)) return -1;
  return c->sp[ksp - 1].n;
}

int mrg_upload_particles(mrg_ctx* c, int32_t ksp, const double* x, const double* y, const double* z,
                         const double* vx, const double* vy, const double* vz, int64_t npr, int64_t first,
                         int64_t stride) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (npr < 0 || first < 1 || stride < 1) return fail(MRG_ERR_ARG, "need npr >= 0, first >= 1 (1-based), stride >= 1");
  const double* h[6] = {x, y, z, vx, vy, vz};
  for (int k = 0; k < 6; k++) if (!h[k] && npr > 0) return fail(MRG_ERR_ARG, "null particle array");
  CK(cudaSetDevice(c->device));
  const long long n = owned_count(npr, first, stride);
  if (n >= (1LL << 31) - 64) return fail(MRG_ERR_ARG, "more than 2^31 particles of one species on one GPU");
  Species& s = c->sp[ksp - 1];
  rc = alloc_species(c, s, n);
  if (rc) return rc;
  if (n == 0) return MRG_OK;
  if (stride == 1) {
    for (int k = 0; k < 6; k++)
      CK(cudaMemcpyAsync(s.d[k], h[k] + (first - 1), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  } else {
    const long long chunk = 1 << 20;
    rc = ensure_pinned(c, (size_t)chunk * sizeof(double));
    if (rc) return rc;
    for (int k = 0; k < 6; k++)
      for (long long m0 = 0; m0 < n; m0 += chunk) {
        const long long cnt = std::min(chunk, n - m0);
        for (long long m = 0; m < cnt; m++) c->h_pinned[m] = h[k][(first - 1) + (m0 + m) * stride];
        CK(cudaMemcpyAsync(s.d[k] + m0, c->h_pinned, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
      }
  }
  CK(cudaStreamSynchronize(c->stream));
  c->h2d += 6 * n * (long long)sizeof(double);
  return MRG_OK;
}

int mrg_download_particles(mrg_ctx* c, int32_t ksp, double* x, double* y, double* z, double* vx, double* vy,
                           double* vz, int64_t npr, int64_t first, int64_t stride) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (npr < 0 || first < 1 || stride < 1) return fail(MRG_ERR_ARG, "need npr >= 0, first >= 1 (1-based), stride >= 1");
  Species& s = c->sp[ksp - 1];
  const long long n = owned_count(npr, first, stride);
  if (n != s.n) return fail(MRG_ERR_ARG, "npr/first/stride do not match the resident particle count");
  double* h[6] = {x, y, z, vx, vy, vz};
  CK(cudaSetDevice(c->device));
  if (n == 0) return MRG_OK;
  if (s.id) { rc = ensure_alt(c, s.cap, false); if (rc) return rc; }
  const long long chunk = 1 << 20;
  if (stride != 1) { rc = ensure_pinned(c, (size_t)chunk * sizeof(double)); if (rc) return rc; }
  for (int k = 0; k < 6; k++) {
    if (!h[k]) continue;
    const double* src = s.d[k];
    if (s.id) {   // back to original local order
      k_unpermute<<<grid_for(n, 256), 256, 0, c->stream>>>(n, s.id, s.d[k], c->alt[0]); CKL(c);
      src = c->alt[0];
    }
    if (stride == 1) {
      CK(cudaMemcpyAsync(h[k] + (first - 1), src, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    } else {
      for (long long m0 = 0; m0 < n; m0 += chunk) {
        const long long cnt = std::min(chunk, n - m0);
        CK(cudaMemcpyAsync(c->h_pinned, src + m0, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (long long m = 0; m < cnt; m++) h[k][(first - 1) + (m0 + m) * stride] = c->h_pinned[m];
      }
    }
    c->d2h += n * (long long)sizeof(double);
  }
  return MRG_OK;
}

int mrg_loadpt(mrg_ctx* c, int32_t ksp, int32_t ppc, double vth, double vdr, double vbeam, int32_t* ranfa,
               int32_t* ranfb) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (ppc < 1 || !(vth > 0) || !ranfa || !ranfb) return fail(MRG_ERR_ARG, "bad loadpt arguments");
  CK(cudaSetDevice(c->device));
  const GP& g = c->g;
  const long long npr = (long long)g.mx * g.my * g.mz * ppc;       // F:8937-8957
  const long long first = c->rank + 1, stride = c->nranks;         // F:219, F:1162
  long long n = owned_count(npr, first, stride);
  Species& s = c->sp[ksp - 1];
  const int nslab = c->opt_slab_n > 0 ? c->opt_slab_n : c->nranks;
  const int islab = c->opt_slab_n > 0 ? c->opt_slab_i : c->rank;
  const bool slab = c->opt_shard == 1 && nslab > 1;
  if (!slab && n >= (1LL << 31) - 64) return fail(MRG_ERR_ARG, "more than 2^31 particles of one species on one GPU");
  if (!slab) {
    rc = alloc_species(c, s, n);
    if (rc) return rc;
  }
  LoadParams L;
  loadpt_table(vth, vdr, L.fv2, &L.v2, &L.dv2);
  L.vdr = vdr; L.vbeam = vbeam;
  L.half_hx = g.hx / 2; L.half_hz = g.hz / 2;
  L.zcent = 0.50 * g.zmax; L.dzcent = 0.125 * g.zmax; L.dzsmt = 0.15 * g.zmax;      // F:9001-9003
  L.ycent1 = 0.30 * g.ymax; L.ycent2 = 0.70 * g.ymax; L.dycent = 0.05 * g.ymax;     // F:9005-9009
  L.rrz = 0.25 * g.zmax; L.rry = 0.075 * g.ymax;                                    // F:9026-9027
  L.sa = (unsigned)*ranfa; L.sb = (unsigned)*ranfb;
  L.first = first; L.stride = stride;
  if (slab) {   // z-slab ownership: count, scan, fill (local order = increasing l)
    // npr itself may exceed 2^31 (BASELINE configs[3] at 8 GPUs: 3.36 G per species): the candidate index l0 and the LCG
    // skip-ahead are 64-bit, the scanned block offsets count OWNED particles and stay below the per-GPU limit checked below
    if ((npr + 255) / 256 >= (1LL << 31)) return fail(MRG_ERR_ARG, "slab loading needs npr < 2^39");
    const long long nb = (npr + 255) / 256;
    int* bc = nullptr;
    CK(cudaMalloc((void**)&bc, (size_t)(nb + 1) * sizeof(int)));
    CK(cudaMemsetAsync(bc + nb, 0, sizeof(int), c->stream));
    k_loadpt_slab_count<<<(unsigned)nb, 256, 0, c->stream>>>(g, L, npr, nslab, islab, bc); CKL(c);
    rc = scan_excl(c, bc, bc, nb + 1, nullptr);
    if (rc) { cudaFree(bc); return rc; }
    int total = 0;
    CK(cudaMemcpyAsync(&total, bc + nb, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    n = total;
    if (n < 0 || n >= (1LL << 31) - 64) { cudaFree(bc); return fail(MRG_ERR_ARG, "more than 2^31 particles of one species on one GPU"); }
    rc = alloc_species(c, s, n);
    if (rc) { cudaFree(bc); return rc; }
    if (n > 0) { k_loadpt_slab_fill<<<(unsigned)nb, 256, 0, c->stream>>>(g, L, soa(s), npr, nslab, islab, bc); CKL(c); }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(bc));
  } else if (n > 0) { k_loadpt<<<grid_for(n, 256), 256, 0, c->stream>>>(g, L, soa(s)); CKL(c); }
  CK(cudaStreamSynchronize(c->stream));
  // every rank of the reference runs the whole serial loader: 3 ranfp + 4 ranf draws per particle
  *ranfb = (int32_t)lcg_skip((unsigned)*ranfb, 3ull * (unsigned long long)npr);
  *ranfa = (int32_t)lcg_skip((unsigned)*ranfa, 4ull * (unsigned long long)npr);
  return MRG_OK;
}

// kind: cudaMemcpyHostToDevice / cudaMemcpyDeviceToDevice copy into the context's own arrays;
// cudaMemcpyDefault = bind (no copy: k_blend reads the caller's device arrays)
static int set_fields_impl(mrg_ctx* c, uint32_t mask, const double* const f12[12], cudaMemcpyKind kind) {
  if (!c) return fail(MRG_ERR_ARG, "null context");
  if (mask >> 12) return fail(MRG_ERR_ARG, "mask has bits above 11");
  CK(cudaSetDevice(c->device));
  // new fields exist only after the moments they were solved from: order the particle stream behind any
  // moment sum still running on the communication stream (deferred mode)
  for (int k = 0; k < c->nspecies; k++)
    if (c->sp[k].pending) CK(cudaStreamWaitEvent(c->stream, c->sp[k].done, 0));
  const size_t gb = (size_t)c->g.ntot * sizeof(double);
  for (int k = 0; k < 12; k++) {
    if (!((mask >> k) & 1u)) continue;
    if (!f12 || !f12[k]) return fail(MRG_ERR_ARG, "selected field pointer is null");
    c->flazy[k] = false;
    if (kind == cudaMemcpyDefault) { c->fcur[k] = f12[k]; continue; }
    CK(cudaMemcpyAsync(c->f12[k], f12[k], gb, kind, c->stream));
    c->fcur[k] = c->f12[k];
    if (kind == cudaMemcpyHostToDevice) c->h2d += (long long)gb;
  }
  if (kind == cudaMemcpyHostToDevice) CK(cudaStreamSynchronize(c->stream));
  c->field_version++;
  c->fields_set = true;
  return MRG_OK;
}
int mrg_set_fields(mrg_ctx* c, uint32_t mask, const double* const f12[12]) {
  return set_fields_impl(c, mask, f12, cudaMemcpyHostToDevice);
}
int mrg_set_fields_device(mrg_ctx* c, uint32_t mask, const double* const f12[12]) {
  return set_fields_impl(c, mask, f12, cudaMemcpyDeviceToDevice);
}
int mrg_bind_fields_device(mrg_ctx* c, uint32_t mask, const double* const f12[12]) {
  return set_fields_impl(c, mask, f12, cudaMemcpyDefault);
}
// Host fields held lazily: nothing is copied now; every preparation fetches the planes it reads (ensure_prep).
int mrg_set_fields_lazy(mrg_ctx* c, uint32_t mask, const double* const f12[12]) {
  if (!c) return fail(MRG_ERR_ARG, "null context");
  if (mask >> 12) return fail(MRG_ERR_ARG, "mask has bits above 11");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < c->nspecies; k++)
    if (c->sp[k].pending) CK(cudaStreamWaitEvent(c->stream, c->sp[k].done, 0));
  for (int k = 0; k < 12; k++) {
    if (!((mask >> k) & 1u)) continue;
    if (!f12 || !f12[k]) return fail(MRG_ERR_ARG, "selected field pointer is null");
    c->fhost[k] = f12[k];
    c->flazy[k] = true;
    c->fplane[k].assign(c->g.mz, 0);
    c->fcur[k] = c->f12[k];
  }
  c->field_version++;
  c->fields_set = true;
  return MRG_OK;
}

// entry prefld of emfild (F:3820-3873) on the device copies: bx, by, bz from ex..ez, ex0..ez0, bx0..bz0 -- bit-identical to
// the host's prefld, which therefore need not upload the three arrays in front of the predictor pass -- and
// the same update as emfild performs it behind its solve (F:4238-4302): from the new E, and on the steps with
// mod(it,5) = 1 followed by outmesh3 + filt3e(sym = +1, no dc), i.e. one z, x and y sweep of the preparation's filter
// kernels on the three arrays (channels 3..5 carry the B signs of the wall mirror rows).
// With lazily held fields (each rank has only the planes its preparations fetched) nothing is computed here: bx,by,bz
// become "to be computed", and every preparation makes exactly the planes it reads after fetching the planes of ex..ez,
// ex0..ez0 around them (ensure_prep).  The smoothing needs whole arrays and is refused in that mode (the host uploads).
int mrg_update_b(mrg_ctx* c, double dt, double aimpl, int32_t smooth) {
  if (!c) return fail(MRG_ERR_ARG, "null context");
  if (!c->fields_set) return fail(MRG_ERR_STATE, "mrg_set_fields has not been called");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < c->nspecies; k++)
    if (c->sp[k].pending) CK(cudaStreamWaitEvent(c->stream, c->sp[k].done, 0));
  const GP& g = c->g;
  bool lazy = false;
  for (int k = 0; k < 12; k++)
    if (c->flazy[k] && !(k >= 3 && k <= 5)) lazy = true;
  if (lazy) {
    if (smooth) return fail(MRG_ERR_STATE, "fields are held lazily: the smoothing of bx,by,bz needs whole arrays (upload them instead)");
    for (int k = 3; k <= 5; k++) {
      c->flazy[k] = true; c->fhost[k] = nullptr; c->fplane[k].assign(g.mz, 0); c->fcur[k] = c->f12[k];
    }
    c->b_dt = dt; c->b_aimpl = aimpl;
    c->field_version++;
    return MRG_OK;
  }
  CPtr12 f; for (int k = 0; k < 12; k++) f.p[k] = c->fcur[k];
  const long long n = (long long)g.mx * (g.my + 1) * g.mz;
  k_prefld<<<grid_for(n, 256), 256, 0, c->stream>>>(g, f, c->f12[3], c->f12[4], c->f12[5], aimpl, 1.0 - aimpl, dt,
                                                     2.0 * g.hx, 2.0 * g.hy, 2.0 * g.hz, nullptr, g.mz); CKL(c);
  for (int k = 3; k <= 5; k++) { c->fcur[k] = c->f12[k]; c->flazy[k] = false; }
  c->field_version++;
  if (!smooth) return MRG_OK;
  // channels 0..2 ride along on scratch (their results are not used): the sweeps are the 6-channel kernels of ensure_prep
  CPtr6 s0; Ptr6 d1, d2, d3; CPtr6 c1, c2;
  for (int k = 0; k < 3; k++) {
    s0.p[k] = c->T2[k]; d1.p[k] = c->T1[k]; c1.p[k] = c->T1[k]; d2.p[k] = c->T2[k]; c2.p[k] = c->T2[k]; d3.p[k] = c->T1[k];
    s0.p[k + 3] = c->f12[k + 3]; d1.p[k + 3] = c->T1[k + 3]; c1.p[k + 3] = c->T1[k + 3];
    d2.p[k + 3] = c->T2[k + 3]; c2.p[k + 3] = c->T2[k + 3]; d3.p[k + 3] = c->f12[k + 3];
  }
  k_filter<2><<<grid_for(n, 256), 256, 0, c->stream>>>(g, s0, d1, nullptr, g.mz); CKL(c);
  k_filter<0><<<grid_for(n, 256), 256, 0, c->stream>>>(g, c1, d2, nullptr, g.mz); CKL(c);
  k_filter<1><<<grid_for(n, 256), 256, 0, c->stream>>>(g, c2, d3, nullptr, g.mz); CKL(c);
  c->prep_valid = false;      // T1 / T2 are the preparation's scratch
  return MRG_OK;
}
int mrg_prefld(mrg_ctx* c, double dt, double aimpl) { return mrg_update_b(c, dt, aimpl, 0); }

// F:796-807 ("Renewal: ex0 <- ex") on the device copies of the fields.
int mrg_renew_fields_host(mrg_ctx* c, const double* const old6[6]) {
  if (!c) return fail(MRG_ERR_ARG, "null context");
  if (!c->fields_set) return fail(MRG_ERR_STATE, "mrg_set_fields has not been called");
  CK(cudaSetDevice(c->device));
  const size_t gb = (size_t)c->g.ntot * sizeof(double);
  for (int k = 0; k < 6; k++) {
    if (c->flazy[k] && !(old6 && old6[k]))
      return fail(MRG_ERR_ARG, "lazily held fields need the host's ex0..bz0 arrays for the renewal (mrg_renew_fields_host)");
    CK(cudaMemcpyAsync(c->f12[k + 6], c->fcur[k], gb, cudaMemcpyDeviceToDevice, c->stream));
    c->fcur[k + 6] = c->f12[k + 6];
    c->flazy[k + 6] = c->flazy[k];
    if (c->flazy[k]) { c->fhost[k + 6] = old6[k]; c->fplane[k + 6] = c->fplane[k]; }   // same planes, same values (the host copied them too)
  }
  c->field_version++;
  return MRG_OK;
}
int mrg_renew_fields(mrg_ctx* c) {
  if (c) for (int k = 0; k < 6; k++)
    if (c->flazy[k]) return fail(MRG_ERR_STATE, "fields are held lazily: use mrg_renew_fields_host");
  return mrg_renew_fields_host(c, nullptr);
}
int mrg_fulmov(mrg_ctx* c, int32_t ksp, double qmult, double wmult, int32_t ipc, const mrg_step_params* p,
               int32_t* ranfb, double* wkix, double* wkih) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (!p) return fail(MRG_ERR_ARG, "null step parameters");
  if (ipc < 0) return fail(MRG_ERR_ARG, "ipc must be 0 (update) or >= 1 (predict + deposit)");
  if (wmult == 0.0) return fail(MRG_ERR_ARG, "wmult must be non-zero");
  CK(cudaSetDevice(c->device));
  Species& s = c->sp[ksp - 1];
  if (!s.M4) { rc = alloc_species(c, s, 0); if (rc) return rc; }
  if (ipc >= 1) { rc = complete_moments(c, ksp - 1); if (rc) return rc; }   // M4 / out4 are about to be reused
  c->ev0 = c->pass_ev[ksp - 1][ipc != 0][0];
  c->ev1 = c->pass_ev[ksp - 1][ipc != 0][1];
  c->pass_timed[ksp - 1][ipc != 0] = false;
  if (c->opt_phases) c->ph_calls++;
  {
    PhaseScope ps(c, ksp - 1, ipc, MRG_PH_PREP, c->stream);
    rc = ensure_prep(c, p, ksp);
    if (rc) return rc;
  }
  const GP& g = c->g;
  PushParams pp;
  pp.dt = p->dt; pp.adt = p->adt; pp.hdt = p->hdt; pp.aimpl = p->aimpl;
  pp.hh = p->dt * qmult / wmult;                                   // F:1150-1152
  pp.ht = 0.5 * pp.hh;
  pp.ht2 = pp.ht * pp.ht;
  pp.qmult = qmult;
  pp.zcent = p->zcent; pp.ycent1 = p->ycent1; pp.ycent2 = p->ycent2;
  pp.zw = 0.15 * g.zmax; pp.yw = 0.025 * g.ymax;                   // F:1343-1345
  pp.drive_on = (ipc == 0 && p->drive_on) ? 1 : 0;
  pp.kick_inline = 0; pp.kick_state = 0u; pp.Ez00 = p->Ez00; pp.yw2 = 0.05 * g.ymax;
  const ParticleSoA P = soa(s);
  double wk_host[2] = {0.0, 0.0};

  if (ipc >= 1) {
    PhaseScope ph_setup(c, ksp - 1, 1, MRG_PH_SETUP, c->stream);
    s.early0 = 0; s.early_n = 0;
    if (c->nranks > 1 && s.compact_ok && compact_possible(c)) {
      // slab-wise exchange ahead: this rank deposits only into its own block and the two strips it sends, every other
      // plane is overwritten by the peers' blocks -- and must not be touched here once peers push into it
      const int N = c->nranks, r = c->rank, L = g.mz / N;
      const size_t P = (size_t)g.nxy * 4;
      int lay[4];
      compact_layout(g.mz, N, r, lay);
      const size_t e0 = (r == 0) ? 0 : (size_t)(2 + r * L), e1 = (r == N - 1) ? (size_t)g.nz : (size_t)(2 + (r + 1) * L);
      CK(cudaMemsetAsync(s.M4 + e0 * P, 0, (e1 - e0) * P * sizeof(double), c->stream));
      CK(cudaMemsetAsync(s.M4 + (size_t)lay[0] * P, 0, (size_t)HALO_PLANES * P * sizeof(double), c->stream));
      CK(cudaMemsetAsync(s.M4 + (size_t)lay[1] * P, 0, (size_t)HALO_PLANES * P * sizeof(double), c->stream));
      CK(cudaMemsetAsync(s.M4 + (size_t)g.ntot * 4, 0, 2 * sizeof(double), c->stream));
    } else {
      CK(cudaMemsetAsync(s.M4, 0, ((size_t)g.ntot * 4 + 2) * sizeof(double), c->stream));
    }
    int blocks = 1;
    const int B = 128;
    if (s.n > 0) {
      const int iters = (c->opt_deposit == 2) ? c->opt_iters : 1;
      const bool tiled = c->opt_tile == 1 && c->opt_deposit == 2 && s.index_valid;
      s.prekeys_valid = false;
      s.prescan_valid = false;
      const long long per_block = (long long)(B / 32) * 32 * iters;
      blocks = (int)((s.n + per_block - 1) / per_block);
      GP gl = g;                               // tiled launches cover only the z planes of the order that own slots
      if (tiled && s.hull_valid) { gl.kz0 = s.kz0; gl.nkz = s.nkz; }
      if (tiled) blocks = ((g.mx + TILE_CELLS - 1) / TILE_CELLS) * g.my * gl.nkz;
      const int nparts = tiled ? 1 : blocks;   // tiled: two atomic accumulators
      rc = ensure(c, (void**)&c->wk_partial, &c->wk_partial_cap, 2LL * nparts, sizeof(double));
      if (rc) return rc;
      if (tiled) CK(cudaMemsetAsync(c->wk_partial, 0, 2 * sizeof(double), c->stream));
      ph_setup.done();
      PhaseScope ph_kernel(c, ksp - 1, 1, MRG_PH_KERNEL, c->stream);
      CK(cudaEventRecord(c->ev0, c->stream));
      const int gm = c->opt_group_min * 4;   // option counts particles; a particle is a quad of lanes
      if (tiled) {
        int* prekey = nullptr;
        if (c->opt_fused_sort) {   // keys + histogram of the next order; the corrector scatters by them
          rc = ensure(c, (void**)&s.key, &s.key_cap, s.n + 2, sizeof(int));
          if (rc) return rc;
          CK(cudaMemsetAsync(s.hist, 0, (size_t)(c->ncell + 1) * sizeof(int), c->stream));
          prekey = s.key;
        }
        CUtensorMap tmP;
        rc = particle_map(s.d[0], s.cap, &tmP);
        if (rc) return rc;
        // Split launch: with the slab-wise exchange ahead and nothing else to hide it under (last species of the step), the
        // pencils of the first three quarters of the block run first; the planes of the block that are final after them --
        // no pencil of the rest deposits within 2 planes, no neighbour strip lands there -- are pushed to the peers on the
        // communication stream while the last quarter runs (a pencil is a tile of GATHER cells, so this needs the fresh
        // order of the fused sort; |vz| dt < hz is the precondition of the exchange itself)
        int nA = 0;
        const bool exch = c->nranks > 1 && s.compact_ok && compact_possible(c);
        if (exch && c->opt_defer && c->opt_peer_push && s.npeer == c->nranks - 1 && s.hull_valid && s.fresh && s.fresh_lookahead == p->hdt &&
            (c->opt_split_push == 2 || (c->opt_split_push == 1 && ksp == c->nspecies))) {
          const int L = g.mz / c->nranks, lo = c->rank * L, split = lo + (3 * L) / 4;
          const int na = ((split - gl.kz0) % g.mz + g.mz) % g.mz;
          const int e_lo = lo + HALO_PLANES, e_hi = std::min(split - 2, lo + L - HALO_PLANES);
          if (na > 0 && na < gl.nkz && e_hi > e_lo) { nA = na; s.early0 = e_lo + 2; s.early_n = e_hi - e_lo; }
        }
        if (nA) {
          GP ga = gl, gb = gl;
          ga.nkz = nA;
          gb.kz0 = (gl.kz0 + nA) % g.mz; gb.nkz = gl.nkz - nA;
          const int per_plane = ((g.mx + TILE_CELLS - 1) / TILE_CELLS) * g.my;
          // the second launch goes to a side stream right behind the first: no dependency between them, so its CTAs fill the
          // SMs as the first launch drains (in one stream the drain + ramp-up bubble cost more than the early push saved)
          CK(cudaEventRecord(c->ev_split_pre, c->stream));
          CK(cudaStreamWaitEvent(c->sstream, c->ev_split_pre, 0));
          k_predict_tile<<<per_plane * ga.nkz, B, PRED_SMEM_BYTES, c->stream>>>(ga, pp, tmP, c->F6, s.M4, s.cell_end, c->wk_partial, gm, prekey, s.hist); CKL(c);
          CK(cudaEventRecord(c->ev_split, c->stream));
          k_predict_tile<<<per_plane * gb.nkz, B, PRED_SMEM_BYTES, c->sstream>>>(gb, pp, tmP, c->F6, s.M4, s.cell_end, c->wk_partial, gm, prekey, s.hist);
          CK(cudaEventRecord(c->ev_split_b, c->sstream));
          CK(cudaStreamWaitEvent(c->stream, c->ev_split_b, 0));
          CK(cudaStreamWaitEvent(c->cstream, c->ev_split, 0));
          PeerPtrs peers;
          peers.n = 0;
          for (int q = 0; q < c->nranks; q++)
            if (q != c->rank) peers.p[peers.n++] = s.peerM4[q];
          for (int q = peers.n; q < 8; q++) peers.p[q] = nullptr;
          const size_t PL = (size_t)g.nxy * 4;
          k_add_push<<<std::max(1, c->opt_peer_push), 256, 0, c->cstream>>>(s.M4, (size_t)s.early0 * PL, (size_t)s.early_n * PL, 0, nullptr, 0, nullptr, 0, peers, 0, 0); CKL(c);
          c->split_count++;
        } else {
          k_predict_tile<<<blocks, B, PRED_SMEM_BYTES, c->stream>>>(gl, pp, tmP, c->F6, s.M4, s.cell_end, c->wk_partial, gm, prekey, s.hist);
        }
        s.prekeys_valid = prekey != nullptr;
        s.keys_valid = false;
        s.prescan_valid = false;
        if (prekey) {
          // cell starts + occupied planes of the next order, needed by the corrector's scatter: scanned here, on the particle
          // stream, so that they run under this species' moment exchange instead of in front of the corrector
          rc = scan_excl(c, s.hist, s.cell_end2, c->ncell + 1, nullptr);
          if (rc) return rc;
          rc = kocc_begin(c, s, s.cell_end2);
          if (rc) return rc;
          s.prescan_valid = true;
        }
      }
      else if (c->opt_deposit == 0) k_predict_direct<<<blocks, B, 0, c->stream>>>(g, pp, P, c->F6, s.M4, c->wk_partial);
      else if (iters == 1) k_predict_run<1><<<blocks, B, 0, c->stream>>>(g, pp, P, c->F6, s.M4, c->wk_partial, gm);
      else if (iters == 4) k_predict_run<4><<<blocks, B, 0, c->stream>>>(g, pp, P, c->F6, s.M4, c->wk_partial, gm);
      else if (iters == 8) k_predict_run<8><<<blocks, B, 0, c->stream>>>(g, pp, P, c->F6, s.M4, c->wk_partial, gm);
      else if (iters == 16) k_predict_run<16><<<blocks, B, 0, c->stream>>>(g, pp, P, c->F6, s.M4, c->wk_partial, gm);
      else if (iters == 32) k_predict_run<32><<<blocks, B, 0, c->stream>>>(g, pp, P, c->F6, s.M4, c->wk_partial, gm);
      else return fail(MRG_ERR_ARG, "option iters must be 4, 8, 16 or 32");
      CKL(c);
      CK(cudaEventRecord(c->ev1, c->stream));
      ph_kernel.done();
      k_wk_final<<<1, 256, 0, c->stream>>>(c->wk_partial, nparts, s.M4 + (size_t)g.ntot * 4); CKL(c);
    }
    ph_setup.done();
    // moment sum + fold; in deferred mode they run on the communication stream, so the next call's particle
    // kernel overlaps them and the host does not wait here
    cudaStream_t ms = c->stream;
    if (c->opt_defer) {
      CK(cudaEventRecord(c->ev_kernel, c->stream));
      CK(cudaStreamWaitEvent(c->cstream, c->ev_kernel, 0));
      ms = c->cstream;
    }
    if (c->nranks > 1) {                                           // F:2379-2384, 2533, 1312-1315
      PhaseScope ph_sum(c, ksp - 1, 1, MRG_PH_SUM, ms);
      if (!c->comm) return fail(MRG_ERR_STATE, "nranks > 1 but mrg_comm_init was not called");
      // Which collective runs is decided ONLY by state every rank agrees on: the vote taken in the preceding corrector
      // call (compact_ok, summed over the ranks there) and the job-wide configuration.  A rank-local condition here could
      // make ranks enqueue different collectives and hang; calls that reset compact_ok (upload, loadpt,
      // options "planes" / "compact") are therefore collective: every rank makes them between the same two steps.
      if (s.compact_ok && compact_possible(c)) {
        rc = compact_sum(c, s, ms);
        if (rc) return rc;
        c->compact_count++;
      } else {
        int n = g_nccl.AllReduce(s.M4, s.M4, (size_t)g.ntot * 4 + 2, kNcclFloat64, kNcclSum, c->comm, ms);
        if (n != 0) return fail(MRG_ERR_NCCL, "ncclAllReduce: " + nccl_err(n));
      }
    }
    Ptr4 o; for (int k = 0; k < 4; k++) o.p[k] = s.out4[k];
    {
      PhaseScope ph_fold(c, ksp - 1, 1, MRG_PH_FOLD, ms);
      k_fold_unpack<<<grid_for(g.ntot, 256), 256, 0, ms>>>(g, s.M4, o, 1); CKL(c);   // F:2398, 2544
    }
    s.have_moments = true;
    if (c->opt_defer) {
      CK(cudaMemcpyAsync(c->wk_pinned + 2 * (ksp - 1), s.M4 + (size_t)g.ntot * 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, ms));
      s.sink_filled = false;
      if (s.sink[0] || s.sink[1] || s.sink[2] || s.sink[3]) {   // D2H of the moments overlaps the next particle kernel too
        size_t off, cnt;
        sink_block(c, &off, &cnt);
        for (int k = 0; k < 4; k++) {
          if (!s.sink[k]) continue;
          CK(cudaMemcpyAsync(s.sink[k] + off, s.out4[k] + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, ms));
          c->d2h += (long long)(cnt * sizeof(double));
        }
        s.sink_filled = true;
      }
      CK(cudaEventRecord(s.done, ms));
      s.pending = true;
      s.wk_user[0] = wkix; s.wk_user[1] = wkih;
      c->pass_timed[ksp - 1][1] = s.n > 0;
      c->d2h += 2 * (long long)sizeof(double);
      return MRG_OK;
    }
    s.sink_filled = false;
    CK(cudaMemcpyAsync(wk_host, s.M4 + (size_t)g.ntot * 4, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  } else {
    bool slab_pending = false;
    if (pp.drive_on && !ranfb) return fail(MRG_ERR_ARG, "ranfb state pointer is required when the drive kick is on");
    {
      const bool want = c->opt_kick == 1 || (c->opt_kick < 0 && c->opt_shard == 1 && (c->nranks > 1 || c->opt_slab_n > 1));
      const bool tile1 = c->opt_tile == 1 && s.index_valid;   // the kernel that can kick on its own
      if (pp.drive_on && want && tile1 && s.n > 0) { pp.kick_inline = 1; pp.kick_state = (unsigned)*ranfb; }
    }
    if (pp.drive_on && !pp.kick_inline) {
      const long long nwords = (s.n + 31) / 32 + 1;
      if (c->slab_words_cap < nwords) {
        if (c->slab_bits) CK(cudaFree(c->slab_bits));
        if (c->slab_words) CK(cudaFree(c->slab_words));
        c->slab_bits = nullptr; c->slab_words = nullptr;
        CK(cudaMalloc((void**)&c->slab_bits, (size_t)nwords * sizeof(unsigned)));
        CK(cudaMalloc((void**)&c->slab_words, (size_t)nwords * sizeof(int)));
        c->slab_words_cap = nwords;
      }
      rc = ensure(c, (void**)&c->slab_list, &c->slab_list_cap, s.n + 1, sizeof(int));
      if (rc) return rc;
      CK(cudaMemsetAsync(c->slab_bits, 0, (size_t)nwords * sizeof(unsigned), c->stream));
      CK(cudaMemsetAsync(c->slab_count, 0, sizeof(int), c->stream));
    }
    PhaseScope ph_setup(c, ksp - 1, 0, MRG_PH_SETUP, c->stream);
    CK(cudaMemsetAsync(c->wk2, 0, 3 * sizeof(double), c->stream));
    s.keys_valid = false;
    s.fresh = false;
    bool fused_scatter = false;
    if (s.n > 0) {
      const bool tiled = c->opt_tile == 1 && s.index_valid;
      const int B = tiled ? 128 : 256;
      GP gl = g;
      if (tiled && s.hull_valid) { gl.kz0 = s.kz0; gl.nkz = s.nkz; }
      const int blocks = tiled ? ((g.mx + TILE_CELLS - 1) / TILE_CELLS) * g.my * gl.nkz : grid_for(s.n, B);
      const int nparts = tiled ? 1 : blocks;
      rc = ensure(c, (void**)&c->wk_partial, &c->wk_partial_cap, 2LL * nparts, sizeof(double));
      if (rc) return rc;
      if (tiled) CK(cudaMemsetAsync(c->wk_partial, 0, 2 * sizeof(double), c->stream));
      int* key_out = nullptr;
      // fused sort: the tiled predictor left the keys of the next order in s.key and their histogram in s.hist
      const bool scatter = tiled && s.prekeys_valid && c->opt_fused_sort;
      SortArrays D{};
      if (scatter) {   // cell starts of the next order; the kernel advances them to the ends
        rc = ensure_alt(c, s.cap, true);
        if (rc) return rc;
        if (!s.prescan_valid) {
          rc = scan_excl(c, s.hist, s.cell_end2, c->ncell + 1, nullptr);
          if (rc) return rc;
          rc = kocc_begin(c, s, s.cell_end2);
          if (rc) return rc;
        }
        for (int k = 0; k < 6; k++) { D.src[k] = s.d[k]; D.dst[k] = c->alt[k]; }
        D.id_src = s.id; D.id_dst = c->alt_id;
      }
      if (tiled && c->opt_fused_keys && !scatter) {   // emit next step's sort keys (cell of x + hdt*v); the old kernel also builds their histogram
        rc = ensure(c, (void**)&s.key, &s.key_cap, s.n + 2, sizeof(int));
        if (rc) return rc;
        CK(cudaMemsetAsync(s.hist, 0, (size_t)(c->ncell + 1) * sizeof(int), c->stream));
        key_out = s.key;
      }
      unsigned* zocc = nullptr;
      if (tiled) { rc = zocc_begin(c, s, &zocc); if (rc) return rc; }
      else s.zocc_valid = false;
      ph_setup.done();
      PhaseScope ph_kernel(c, ksp - 1, 0, MRG_PH_KERNEL, c->stream);
      CK(cudaEventRecord(c->ev0, c->stream));
      if (tiled) {
        CUtensorMap tmP, tmId, tmKey;
        memset(&tmId, 0, sizeof(tmId));
        memset(&tmKey, 0, sizeof(tmKey));
        rc = particle_map(s.d[0], s.cap, &tmP);
        if (!rc && s.id) rc = int_map(s.id, s.cap, &tmId);
        if (!rc && scatter) rc = int_map(s.key, s.key_cap, &tmKey);
        if (rc) return rc;
        for (int k = 0; k < 6; k++) D.src_rw[k] = s.d[k];
        k_correct_tile<<<blocks, B, 0, c->stream>>>(gl, pp, tmP, tmId, tmKey, s.id ? 1 : 0, c->F6, s.cell_end, c->wk_partial,
                                                    c->slab_bits, c->slab_list, c->slab_count, key_out, s.hist, p->hdt,
                                                    scatter ? 1 : 0, s.cell_end2, D, zocc, c->lcg_tab);
        s.hist_valid = !scatter;
        fused_scatter = scatter;
      } else {
        k_correct<<<blocks, B, 0, c->stream>>>(g, pp, P, c->F6, c->wk_partial, c->slab_bits, c->slab_list, c->slab_count);
      }
      CKL(c);
      CK(cudaEventRecord(c->ev1, c->stream));
      ph_kernel.done();
      if (zocc) { rc = zocc_fetch(c, s, p->hdt, true); if (rc) return rc; }   // completed by the synchronize below
      if (key_out) { s.keys_valid = true; s.keys_lookahead = p->hdt; }
      if (fused_scatter) {   // the spare buffers now hold the updated particles in the next order
        for (int k = 0; k < 6; k++) std::swap(s.d[k], c->alt[k]);
        int* old_id = s.id;
        s.id = c->alt_id;
        std::swap(s.cap, c->alt_cap);
        if (old_id) c->alt_id = old_id;
        else { c->alt_id = nullptr; CK(cudaMalloc((void**)&c->alt_id, (size_t)c->alt_cap * sizeof(int))); }
        std::swap(s.cell_end, s.cell_end2);
        s.fresh = true;
        s.fresh_lookahead = p->hdt;
      }
      s.prekeys_valid = false;
      s.prescan_valid = false;
      k_wk_final<<<1, 256, 0, c->stream>>>(c->wk_partial, nparts, c->wk2); CKL(c);
    }
    ph_setup.done();
    double wk3[3] = {0.0, 0.0, 0.0};
    PhaseScope ph_kick(c, ksp - 1, 0, MRG_PH_KICK, c->stream);
    if (pp.kick_inline) {
      *ranfb = (int32_t)lcg_skip((unsigned)*ranfb, (unsigned long long)s.n);   // every particle owns one draw of the call
    } else if (pp.drive_on && s.n > 0) {
      // the chain runs without knowing the slab count on the host (no synchronize between the corrector and the
      // kick): k_kick reads it on the device and strides over the list; the host reads it with wk below
      const long long nwords = (s.n + 31) / 32 + 1;
      k_popc<<<grid_for(nwords, 256), 256, 0, c->stream>>>(c->slab_bits, nwords, c->slab_words); CKL(c);
      rc = scan_excl(c, c->slab_words, c->slab_words, nwords, nullptr);
      if (rc) return rc;
      const int kick_blocks = (int)std::min<long long>(grid_for(s.n, 256), 8LL * c->num_sms);
      k_kick<<<kick_blocks, 256, 0, c->stream>>>(g, soa(s), c->F6, c->slab_bits, c->slab_words, c->slab_list,
                                                 c->slab_count, (unsigned)*ranfb, p->Ez00, p->ycent1,
                                                 p->ycent2, 0.05 * g.ymax, c->lcg_tab); CKL(c);
      CK(cudaMemcpyAsync(c->slab_n_host, c->slab_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      slab_pending = true;
    }
    ph_kick.done();
    PhaseScope ph_sum(c, ksp - 1, 0, MRG_PH_SUM, c->stream);
    if (c->nranks > 1) {                                           // F:1312-1315
      if (!c->comm) return fail(MRG_ERR_STATE, "nranks > 1 but mrg_comm_init was not called");
      // third word: this rank's vote on exchanging slabs instead of allreducing the grid in the next ipc>=1 call of
      // the species (0 = my deposits will stay near my block); the planes were recorded by the kernel above
      const bool vote = compact_possible(c);
      if (vote) {
        CK(cudaStreamSynchronize(c->stream));                      // the recorded planes are on the host now
        const double veto = compact_eligible(c, s, p->hdt) ? 0.0 : 1.0;
        CK(cudaMemcpyAsync(c->wk2 + 2, &veto, sizeof(double), cudaMemcpyHostToDevice, c->stream));
      }
      int n = g_nccl.AllReduce(c->wk2, c->wk2, 3, kNcclFloat64, kNcclSum, c->comm, c->stream);
      if (n != 0) return fail(MRG_ERR_NCCL, "ncclAllReduce: " + nccl_err(n));
      CK(cudaMemcpyAsync(wk3, c->wk2, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      s.compact_ok = vote && wk3[2] == 0.0;
    } else {
      CK(cudaMemcpyAsync(wk3, c->wk2, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    wk_host[0] = wk3[0]; wk_host[1] = wk3[1];
    if (slab_pending) *ranfb = (int32_t)lcg_skip((unsigned)*ranfb, (unsigned long long)*c->slab_n_host);
    if (fused_scatter) kocc_finish(c, s);   // planes of the order the particles are in now
  }
  c->pass_timed[ksp - 1][ipc != 0] = s.n > 0;
  if (s.n > 0) {
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->last_kernel_ms = ms;
  } else {
    c->last_kernel_ms = 0.0;
  }
  c->d2h += 2 * (long long)sizeof(double);
  if (wkix) *wkix = wk_host[0];
  if (wkih) *wkih = wk_host[1];
  return MRG_OK;
}

int mrg_get_moments(mrg_ctx* c, int32_t ksp, double* qjx, double* qjy, double* qjz, double* q, int32_t folded) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  Species& s = c->sp[ksp - 1];
  if (!s.have_moments) return fail(MRG_ERR_STATE, "no ipc>=1 call has produced moments for this species yet");
  CK(cudaSetDevice(c->device));
  rc = complete_moments(c, ksp - 1);
  if (rc) return rc;
  double* h[4] = {qjx, qjy, qjz, q};
  const size_t gb = (size_t)c->g.ntot * sizeof(double);
  double* const* src = s.out4;
  if (!folded) {   // unpack the rank-summed raw arrays into T1[0..3] (prep scratch is not live here)
    Ptr4 o; for (int k = 0; k < 4; k++) o.p[k] = c->T2[k];
    k_fold_unpack<<<grid_for(c->g.ntot, 256), 256, 0, c->stream>>>(c->g, s.M4, o, 0); CKL(c);
    src = c->T2;
  }
  size_t off = 0, cnt = (size_t)c->g.ntot;
  if (folded) sink_block(c, &off, &cnt);      // shared host arrays: this rank's block only
  for (int k = 0; k < 4; k++) {
    if (!h[k]) continue;
    if (folded && s.sink_filled && h[k] == s.sink[k]) continue;   // already delivered by the deferred call
    CK(cudaMemcpyAsync(h[k] + off, src[k] + off, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    c->d2h += (long long)(cnt * sizeof(double));
  }
  (void)gb;
  CK(cudaStreamSynchronize(c->stream));
  return MRG_OK;
}

int mrg_set_moment_sink(mrg_ctx* c, int32_t ksp, double* qjx, double* qjy, double* qjz, double* q) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  rc = complete_moments(c, ksp - 1);
  if (rc) return rc;
  Species& s = c->sp[ksp - 1];
  s.sink[0] = qjx; s.sink[1] = qjy; s.sink[2] = qjz; s.sink[3] = q;
  s.sink_filled = false;
  return MRG_OK;
}

int mrg_get_moments_device(mrg_ctx* c, int32_t ksp, const double* dev4[4]) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  Species& s = c->sp[ksp - 1];
  if (!s.have_moments) return fail(MRG_ERR_STATE, "no ipc>=1 call has produced moments for this species yet");
  rc = complete_moments(c, ksp - 1);
  if (rc) return rc;
  for (int k = 0; k < 4; k++) dev4[k] = s.out4[k];
  return MRG_OK;
}

int mrg_get_prepared_fields(mrg_ctx* c, const mrg_step_params* p, double* const a6[6]) {
  if (!c || !p || !a6) return fail(MRG_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  int rc = ensure_prep(c, p, 0);   // every plane
  if (rc) return rc;
  const size_t gb = (size_t)c->g.ntot * sizeof(double);
  Ptr6 o;
  for (int k = 0; k < 6; k++) {
    if (!c->tmp6[k]) CK(cudaMalloc((void**)&c->tmp6[k], gb));
    o.p[k] = c->tmp6[k];
  }
  k_unpack6<<<grid_for(c->g.ntot, 256), 256, 0, c->stream>>>(c->g, c->F6, o); CKL(c);
  for (int k = 0; k < 6; k++) {
    if (!a6[k]) continue;
    CK(cudaMemcpyAsync(a6[k], c->tmp6[k], gb, cudaMemcpyDeviceToHost, c->stream));
    c->d2h += (long long)gb;
  }
  CK(cudaStreamSynchronize(c->stream));
  return MRG_OK;
}

// the device copies of COMMON /fields/ (what mrg_set_fields, mrg_renew_fields and mrg_prefld left there)
int mrg_get_fields(mrg_ctx* c, uint32_t mask, double* const f12[12]) {
  if (!c || !f12) return fail(MRG_ERR_ARG, "null argument");
  if (mask >> 12) return fail(MRG_ERR_ARG, "mask has bits above 11");
  if (!c->fields_set) return fail(MRG_ERR_STATE, "mrg_set_fields has not been called");
  CK(cudaSetDevice(c->device));
  const size_t gb = (size_t)c->g.ntot * sizeof(double);
  for (int k = 0; k < 12; k++) {
    if (!((mask >> k) & 1u)) continue;
    if (!f12[k]) return fail(MRG_ERR_ARG, "selected field pointer is null");
    if (c->flazy[k]) return fail(MRG_ERR_STATE, "this field is held lazily: the device has only some of its planes");
    CK(cudaMemcpyAsync(f12[k], c->fcur[k], gb, cudaMemcpyDeviceToHost, c->stream));
    c->d2h += (long long)gb;
  }
  CK(cudaStreamSynchronize(c->stream));
  return MRG_OK;
}

int mrg_sort(mrg_ctx* c, int32_t ksp, double lookahead) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  Species& s = c->sp[ksp - 1];
  if (s.n == 0) return MRG_OK;
  if (s.fresh && s.index_valid && s.fresh_lookahead == lookahead) return MRG_OK;   // the fused sort of the last corrector already produced this order
  s.prekeys_valid = false;
  s.prescan_valid = false;
  s.fresh = false;
  rc = ensure_alt(c, s.cap, true);
  if (rc) return rc;
  rc = ensure(c, (void**)&s.key, &s.key_cap, s.n + 2, sizeof(int));
  if (rc) return rc;
  const int B = 256;
  // planes the next pass gathers from: kept when the last corrector tracked them for this look-ahead
  unsigned* zocc = nullptr;
  if (!(s.zocc_valid && s.zocc_lookahead == lookahead)) { rc = zocc_begin(c, s, &zocc); if (rc) return rc; }
  if (!(s.keys_valid && s.keys_lookahead == lookahead)) {   // the corrector may already have emitted the keys
    CK(cudaMemsetAsync(s.hist, 0, (size_t)(c->ncell + 1) * sizeof(int), c->stream));
    k_sort_keys<<<grid_for(s.n, B), B, 0, c->stream>>>(c->g, soa(s), lookahead, s.key, s.hist, zocc); CKL(c);
  } else {
    if (!s.hist_valid) {
      CK(cudaMemsetAsync(s.hist, 0, (size_t)(c->ncell + 1) * sizeof(int), c->stream));
      k_key_hist<<<grid_for(s.n, B), B, 0, c->stream>>>(s.n, s.key, s.hist); CKL(c);
    }
    if (zocc) { k_mark_planes<<<grid_for(s.n, B), B, 0, c->stream>>>(c->g, soa(s), lookahead, zocc); CKL(c); }
  }
  if (zocc) { rc = zocc_fetch(c, s, lookahead, false); if (rc) return rc; }   // completed by the synchronize below
  s.keys_valid = false;
  rc = scan_excl(c, s.hist, s.cell_end, c->ncell + 1, nullptr);
  if (rc) return rc;
  s.kocc_pending = false;
  rc = kocc_begin(c, s, s.cell_end);
  if (rc) return rc;
  SortArrays A;
  for (int k = 0; k < 6; k++) { A.src[k] = s.d[k]; A.dst[k] = c->alt[k]; }
  A.id_src = s.id; A.id_dst = c->alt_id;
  // the scatter advances cell_end[c] from the start to the end slot of cell c
  k_sort_scatter<<<grid_for(s.n, B), B, 0, c->stream>>>(s.n, s.key, s.cell_end, A); CKL(c);
  CK(cudaStreamSynchronize(c->stream));
  kocc_finish(c, s);
  s.index_valid = true;
  // the spare buffer becomes the species' storage and vice versa
  for (int k = 0; k < 6; k++) std::swap(s.d[k], c->alt[k]);
  int* old_id = s.id;
  s.id = c->alt_id;
  const long long old_cap = s.cap;
  s.cap = c->alt_cap;
  c->alt_cap = old_cap;
  if (old_id) {
    c->alt_id = old_id;
  } else {
    c->alt_id = nullptr;
    CK(cudaMalloc((void**)&c->alt_id, (size_t)c->alt_cap * sizeof(int)));
  }
  return MRG_OK;
}

int mrg_set_option(mrg_ctx* c, const char* name, int64_t value) {
  if (!c || !name) return fail(MRG_ERR_ARG, "null argument");
  const std::string n(name);
  if (n == "deposit") {
    if (value < 0 || value > 2) return fail(MRG_ERR_ARG, "deposit must be 0, 1 or 2");
    c->opt_deposit = (int)value;
  } else if (n == "iters") {
    if (value != 4 && value != 8 && value != 16 && value != 32) return fail(MRG_ERR_ARG, "iters must be 4, 8, 16 or 32");
    c->opt_iters = (int)value;
  } else if (n == "tile") {
    if (value < 0 || value > 1) return fail(MRG_ERR_ARG, "tile must be 0 (gather through L1, any particle order) or 1 (TMA-staged tiles)");
    c->opt_tile = (int)value;
  } else if (n == "fused_keys") {
    c->opt_fused_keys = value != 0;
  } else if (n == "fused_sort") {
    c->opt_fused_sort = value != 0;
  } else if (n == "shard") {
    if (value < 0 || value > 1) return fail(MRG_ERR_ARG, "shard must be 0 (round-robin, the reference) or 1 (z slabs)");
    c->opt_shard = (int)value;
  } else if (n == "slab_of") {
    if (value < 0 || value > 4096) return fail(MRG_ERR_ARG, "slab_of must be 0 (= nranks) or a slab count");
    c->opt_slab_n = (int)value;
    if (c->opt_slab_i >= std::max(c->opt_slab_n, 1)) c->opt_slab_i = 0;
  } else if (n == "slab_index") {
    if (value < 0 || value >= std::max(c->opt_slab_n, 1)) return fail(MRG_ERR_ARG, "slab_index must be in [0, slab_of)");
    c->opt_slab_i = (int)value;
  } else if (n == "planes") {
    if (value < -1 || value > 1) return fail(MRG_ERR_ARG, "planes must be -1 (when nranks > 1), 0 (never) or 1 (always)");
    c->opt_planes = (int)value;
    for (auto& sp : c->sp) { sp.zocc_valid = false; sp.compact_ok = false; }
  } else if (n == "kick") {
    if (value < -1 || value > 1) return fail(MRG_ERR_ARG, "kick must be -1, 0 (the reference's serial draw order) or 1 (draw by particle index)");
    c->opt_kick = (int)value;
  } else if (n == "compact") {
    if (value < -1 || value > 0) return fail(MRG_ERR_ARG, "compact must be -1 (slab-wise moment exchange when the ranks agree it is possible) or 0 (always allreduce)");
    c->opt_compact = (int)value;
    for (auto& sp : c->sp) sp.compact_ok = false;
  } else if (n == "peer_push") {
    if (value < 0 || value > 4096) return fail(MRG_ERR_ARG, "peer_push must be 0 (ncclAllGather) or the number of CTAs of the push kernel");
    c->opt_peer_push = (int)value;
  } else if (n == "split_push") {
    if (value < 0 || value > 2) return fail(MRG_ERR_ARG, "split_push must be 0 (off), 1 (last species of the step) or 2 (every species)");
    c->opt_split_push = (int)value;
  } else if (n == "peer_push_last") {
    if (value < 0 || value > 4096) return fail(MRG_ERR_ARG, "peer_push_last must be 0 (= peer_push) or the number of CTAs of the push kernel of the last species");
    c->opt_peer_push_last = (int)value;
  } else if (n == "phases") {
    c->opt_phases = value != 0;
  } else if (n == "sink_share") {
    c->opt_sink_share = value != 0;
  } else if (n == "defer") {
    if (!value) { for (int k = 0; k < c->nspecies; k++) { int rc = complete_moments(c, k); if (rc) return rc; } }
    c->opt_defer = value != 0;
  } else if (n == "group_min") {
    if (value < 1 || value > 9) return fail(MRG_ERR_ARG, "group_min must be in 1..9 (particles per sub-iteration group)");
    c->opt_group_min = (int)value;
  } else {
    return fail(MRG_ERR_ARG, "unknown option: " + n);
  }
  return MRG_OK;
}

int mrg_get_counters(mrg_ctx* c, int64_t out[3], int32_t reset) {
  if (!c || !out) return fail(MRG_ERR_ARG, "null argument");
  out[0] = c->launches; out[1] = c->h2d; out[2] = c->d2h;
  if (reset) { c->launches = 0; c->h2d = 0; c->d2h = 0; }
  return MRG_OK;
}

int mrg_last_kernel_ms(mrg_ctx* c, double* ms) {
  if (!c || !ms) return fail(MRG_ERR_ARG, "null argument");
  *ms = c->last_kernel_ms;
  return MRG_OK;
}

int mrg_event_record(mrg_ctx* c, int32_t slot) {
  if (!c || slot < 0 || slot >= 8) return fail(MRG_ERR_ARG, "bad event slot");
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->user_ev[slot], c->stream));
  return MRG_OK;
}

int mrg_event_elapsed_ms(mrg_ctx* c, int32_t a, int32_t b, double* ms) {
  if (!c || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) return fail(MRG_ERR_ARG, "bad event slot");
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(c->user_ev[b]));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, c->user_ev[a], c->user_ev[b]));
  *ms = f;
  return MRG_OK;
}

int mrg_synchronize(mrg_ctx* c) {
  if (!c) return fail(MRG_ERR_ARG, "null context");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->cstream));
  for (int k = 0; k < c->nspecies; k++) { int rc = complete_moments(c, k); if (rc) return rc; }
  return MRG_OK;
}

int mrg_pass_ms(mrg_ctx* c, int32_t ksp, int32_t ipc, double* ms) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (!ms) return fail(MRG_ERR_ARG, "null argument");
  *ms = 0.0;
  if (!c->pass_timed[ksp - 1][ipc != 0]) return MRG_OK;
  CK(cudaSetDevice(c->device));
  cudaEvent_t* e = c->pass_ev[ksp - 1][ipc != 0];
  CK(cudaEventSynchronize(e[1]));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, e[0], e[1]));
  *ms = f;
  return MRG_OK;
}

int mrg_self_check(mrg_ctx* c, int32_t ksp, double sums[4], int64_t counts[4]) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (!sums || !counts) return fail(MRG_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  Species& s = c->sp[ksp - 1];
  rc = complete_moments(c, ksp - 1);
  if (rc) return rc;
  double* d4 = nullptr;
  CK(cudaMalloc((void**)&d4, 4 * sizeof(double) + 2 * sizeof(unsigned long long)));
  CK(cudaMemsetAsync(d4, 0, 4 * sizeof(double) + 2 * sizeof(unsigned long long), c->stream));
  unsigned long long* d2 = reinterpret_cast<unsigned long long*>(d4 + 4);
  for (int k = 0; k < 4; k++) sums[k] = 0.0;
  if (s.have_moments) { k_moment_sums<<<592, 256, 0, c->stream>>>(s.M4, c->g.ntot, d4); CKL(c); }
  unsigned long long h2[2] = {0ull, 0ull};
  if (s.id && s.n > 0) { k_id_sums<<<592, 256, 0, c->stream>>>(s.id, s.n, d2); CKL(c); }
  int last = -1;
  if (s.index_valid && s.cell_end) CK(cudaMemcpyAsync(&last, s.cell_end + (c->ncell - 1), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(sums, d4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(h2, d2, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaFree(d4));
  const unsigned long long n = (unsigned long long)s.n;
  if (!s.id) {      // identity order
    h2[0] = n ? n * (n - 1) / 2 : 0ull;
    const unsigned __int128 t = n ? (unsigned __int128)(n - 1) * n * (2 * n - 1) / 6 : 0;
    h2[1] = (unsigned long long)t;
  }
  counts[0] = s.n;
  counts[1] = s.index_valid ? (int64_t)last : (int64_t)s.n;
  counts[2] = (int64_t)h2[0];
  counts[3] = (int64_t)h2[1];
  return MRG_OK;
}

int mrg_dfma_peak(mrg_ctx* c, double* dfma_per_s) {
  if (!c || !dfma_per_s) return fail(MRG_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount * 8, iters = 1 << 14;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(e0, c->stream));
    k_dfma_peak<<<blocks, 256, 0, c->stream>>>(c->wk2, iters, 1.0000001, 1e-9); CKL(c);
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double rate = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  CK(cudaEventDestroy(e0));
  CK(cudaEventDestroy(e1));
  *dfma_per_s = best;
  return MRG_OK;
}

int mrg_peer_export(mrg_ctx* c, int32_t ksp, unsigned char handle[MRG_IPC_HANDLE_BYTES]) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (!handle) return fail(MRG_ERR_ARG, "null handle");
  CK(cudaSetDevice(c->device));
  Species& s = c->sp[ksp - 1];
  if (!s.M4) { rc = alloc_species(c, s, 0); if (rc) return rc; }
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, s.M4));
  static_assert(sizeof(h) == MRG_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(handle, &h, sizeof(h));
  return MRG_OK;
}

int mrg_peer_import(mrg_ctx* c, int32_t ksp, int32_t rank, const unsigned char handle[MRG_IPC_HANDLE_BYTES]) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (!handle || rank < 0 || rank >= c->nranks || rank >= 8 || rank == c->rank) return fail(MRG_ERR_ARG, "bad peer rank / handle");
  CK(cudaSetDevice(c->device));
  Species& s = c->sp[ksp - 1];
  if (s.peerM4[rank]) return MRG_OK;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* ptr = nullptr;
  CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  s.peerM4[rank] = (double*)ptr;
  s.npeer++;
  return MRG_OK;
}

int mrg_phase_ms(mrg_ctx* c, double out[MRG_NPHASE], int64_t* calls, int32_t reset) {
  if (!c || !out) return fail(MRG_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  for (int k = 0; k < MRG_MAX_SPECIES; k++)
    for (int i = 0; i < 2; i++)
      for (int ph = 0; ph < MRG_NPHASE_DETAIL; ph++) { int rc = phase_collect(c, k, i, ph); if (rc) return rc; }
  for (int ph = 0; ph < MRG_NPHASE; ph++) out[ph] = c->ph_ms[ph];
  if (calls) *calls = c->ph_calls;
  if (reset) {
    for (int ph = 0; ph < MRG_NPHASE; ph++) c->ph_ms[ph] = 0.0;
    memset(c->ph_detail, 0, sizeof(c->ph_detail));
    c->ph_calls = 0;
  }
  return MRG_OK;
}

int mrg_phase_detail(mrg_ctx* c, int32_t ksp, int32_t ipc, double out[MRG_NPHASE_DETAIL]) {
  int rc = check_species(c, ksp);
  if (rc) return rc;
  if (!out) return fail(MRG_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  const int i = ipc != 0;
  for (int ph = 0; ph < MRG_NPHASE_DETAIL; ph++) { rc = phase_collect(c, ksp - 1, i, ph); if (rc) return rc; }
  for (int ph = 0; ph < MRG_NPHASE_DETAIL; ph++) out[ph] = c->ph_detail[ksp - 1][i][ph];
  return MRG_OK;
}

int mrg_plane_sets(int32_t mz, const uint8_t* occ, int32_t* listB, int32_t* listGI, int32_t* listG, int32_t n[3]) {
  if (mz < 4 || !occ || !listB || !listGI || !listG || !n) return fail(MRG_ERR_ARG, "bad argument");
  PlaneSets ps;
  plane_sets(mz, std::vector<char>(occ, occ + mz + 1), ps);
  std::copy(ps.listB.begin(), ps.listB.end(), listB);
  std::copy(ps.listGI.begin(), ps.listGI.end(), listGI);
  std::copy(ps.listG.begin(), ps.listG.end(), listG);
  n[0] = (int32_t)ps.listB.size(); n[1] = (int32_t)ps.listGI.size(); n[2] = (int32_t)ps.listG.size();
  return MRG_OK;
}

int mrg_compact_layout(int32_t mz, int32_t nranks, int32_t rank, const uint8_t* occ, int32_t out[7]) {
  if (mz < 4 || nranks < 2 || rank < 0 || rank >= nranks || !out) return fail(MRG_ERR_ARG, "bad argument");
  out[0] = (mz % nranks == 0 && mz / nranks >= 2 * HALO_PLANES) ? 1 : 0;     // can this grid be exchanged slab-wise at all
  int lay[4] = {0, 0, 0, 0};
  if (out[0]) compact_layout(mz, nranks, rank, lay);
  for (int k = 0; k < 4; k++) out[1 + k] = lay[k];
  out[5] = HALO_PLANES;
  out[6] = (out[0] && occ) ? (compact_planes_ok(mz, nranks, rank, std::vector<char>(occ, occ + mz + 1)) ? 1 : 0) : 0;
  return MRG_OK;
}

int mrg_get_prep_stats(mrg_ctx* c, int64_t out[4], int32_t reset) {
  if (!c || !out) return fail(MRG_ERR_ARG, "null argument");
  out[0] = c->prep_count; out[1] = c->prep_restricted_count; out[2] = c->prep_planes_sum; out[3] = c->compact_count;
  if (reset) { c->prep_count = 0; c->prep_restricted_count = 0; c->prep_planes_sum = 0; c->compact_count = 0; }
  return MRG_OK;
}

int64_t mrg_peer_pushes(mrg_ctx* c, int32_t reset) {
  if (!c) return -1;
  const long long v = c->push_count;
  if (reset) c->push_count = 0;
  return v;
}

int64_t mrg_split_pushes(mrg_ctx* c, int32_t reset) {
  if (!c) return -1;
  const long long v = c->split_count;
  if (reset) c->split_count = 0;
  return v;
}

}  // extern "C"
