/*
 * mrg_fulmov.h -- C ABI of the B200 (sm_100a) replacement for the /fulmov/
 * particle hot path of @mrg37-080A.f03 (F:n = that file, line n).
 *
 * The reference has no FFI: the seam is the Fortran subroutine
 *   fulmov(x,y,z,vx,vy,vz,qmult,wmult,npr,ipc,ksp,ipar,size)      F:1044
 * plus the COMMON blocks it reads and writes (F:1061-1120).  Every entry
 * point below names the reference lines it replaces; the ISO_C_BINDING shim
 * a maintainer adds on the Fortran side is in INTEGRATION.md and
 * fortran/mrg_gpu.f03.
 *
 * Conventions: every function returns 0 on success and a non-zero code on
 * failure (mrg_last_error() gives the text); plain pointers and sizes only;
 * all grids use the reference layout real(C_DOUBLE)(-2:mx+1,-1:my+1,-2:mz+1),
 * i fastest, mxyzA = (mx+4)(my+3)(mz+4) doubles (param_080A.h:33); particle
 * counts are int64_t; one context per rank / GPU, driven by one host thread.
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef MRG_FULMOV_H
#define MRG_FULMOV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mrg_ctx mrg_ctx;

enum {
  MRG_OK = 0,
  MRG_ERR_ARG = 1,      /* bad argument                                    */
  MRG_ERR_CUDA = 2,     /* CUDA runtime error (no device, OOM, launch)      */
  MRG_ERR_NCCL = 3,     /* NCCL missing or failed                           */
  MRG_ERR_STATE = 4     /* call order (fields not set, species empty, ...)  */
};

#define MRG_MAX_SPECIES 4      /* qspec(4), wspec(4): F:1100               */
#define MRG_UNIQUE_ID_BYTES 128

/* Per-call scalars that fulmov reads from COMMON /parm1/,/parm2/,/profl/.  */
typedef struct mrg_step_params {
  double dt, adt, hdt, aimpl;        /* F:1094; adt,hdt are passed, not
                                        derived: trans zeroes them at it=0
                                        (F:678-680)                         */
  double bxc, byc, bzc;              /* F:1097, set at F:8601-8603          */
  int32_t ifilx, ifily, ifilz;       /* F:1085 (forced to 1 at F:368-370)   */
  int32_t drive_on;                  /* 0 skips the ipc=0 kick loop entirely
                                        (for tests); 1 = reference          */
  double Ez00, zcent, ycent1, ycent2;/* /profl/ F:1109-1110                 */
} mrg_step_params;

/* Sizes / grid constants.  Derives hx..,hxi..,xmaxe.. exactly as
 * F:8454-8484 and F:8567-8577.  `device` is the CUDA ordinal; rank/nranks
 * mirror ipar-1 and size of F:215-219.                                      */
int mrg_create(mrg_ctx** ctx, int32_t mx, int32_t my, int32_t mz, double xmax,
               double ymax, double zmax, int32_t nspecies, int32_t rank,
               int32_t nranks, int32_t device);
int mrg_destroy(mrg_ctx* ctx);
const char* mrg_last_error(void);
/* "sm_100a;<build flags>" -- lets a caller check what was loaded.           */
const char* mrg_build_info(void);

/* NCCL communicator for the moment sums that replace mpi_allreduce at
 * F:1312-1315, F:2379-2384 and F:2533.  Rank 0 calls mrg_comm_unique_id and
 * the host broadcasts the 128 bytes (MPI_Bcast in the Fortran host,
 * torch.distributed in bench.py); every rank then calls mrg_comm_init.
 * With nranks == 1 neither call is needed.                                  */
int mrg_comm_unique_id(unsigned char id[MRG_UNIQUE_ID_BYTES]);
int mrg_comm_init(mrg_ctx* ctx, const unsigned char id[MRG_UNIQUE_ID_BYTES]);

/* Particles.  The host arrays are the reference's x(np0).. (F:121-122,
 * F:1056); the rank owns l = first, first+stride, ... <= npr (1-based, as in
 * `do l= ipar,npr,size`, F:1162) and only those are copied.  Download writes
 * them back in the original l order whatever the device order is (needed by
 * restrt, F:9622-9668, and diag1).                                          */
int mrg_upload_particles(mrg_ctx* ctx, int32_t ksp, const double* x,
                         const double* y, const double* z, const double* vx,
                         const double* vy, const double* vz, int64_t npr,
                         int64_t first, int64_t stride);
int mrg_download_particles(mrg_ctx* ctx, int32_t ksp, double* x, double* y,
                           double* z, double* vx, double* vy, double* vz,
                           int64_t npr, int64_t first, int64_t stride);
/* Number of particles of species ksp resident on this GPU.                  */
int64_t mrg_num_local(mrg_ctx* ctx, int32_t ksp);

/* Synthetic two-flux-bundle load generated on the device with the exact
 * LCG skip-ahead: same values as loadpt, F:8937-9040, for the owned l with
 * `ppc` particles per cell (the source hard-codes 32, F:8941).  ranfa/ranfb
 * (COMMON /ranfa/,/ranfb/ after rantbl, F:9255-9256) are advanced as the
 * serial loader would.                                                      */
int mrg_loadpt(mrg_ctx* ctx, int32_t ksp, int32_t ppc, double vth, double vdr,
               double vbeam, int32_t* ranfa, int32_t* ranfb);

/* COMMON /fields/ (F:1066): f12 = ex,ey,ez,bx,by,bz,ex0,ey0,ez0,bx0,by0,bz0,
 * each mxyzA doubles on the HOST; bit i of mask selects f12[i] (unselected
 * pointers may be NULL).  Invalidates the cached prepared fields.           */
int mrg_set_fields(mrg_ctx* ctx, uint32_t mask, const double* const f12[12]);
/* Same, from DEVICE pointers (resident field solve, bench `value` leg).     */
int mrg_set_fields_device(mrg_ctx* ctx, uint32_t mask,
                          const double* const f12_dev[12]);

/* Zero-copy variant: the context reads the caller's DEVICE arrays in place
 * (no copy is made).  The selected arrays must stay valid and unchanged until
 * they are replaced by another set/bind call.  For a device-resident field
 * solve that already owns ex..bz0 in HBM.                                   */
int mrg_bind_fields_device(mrg_ctx* ctx, uint32_t mask,
                           const double* const f12_dev[12]);
/* Lazy variant for HOST arrays: nothing is copied now.  The context keeps the
 * pointers and, when a field preparation runs (F:1127-1148 inside the next
 * mrg_fulmov / mrg_get_prepared_fields), fetches exactly the interior z planes
 * that preparation reads and does not hold yet -- with option "planes" a rank
 * that owns a z slab uploads its slab (+ filter halo) of the replicated
 * COMMON /fields/ arrays instead of all of them.  The selected host arrays
 * must stay valid and unchanged until the caller replaces them (another set /
 * bind call for the same array) -- true for the reference, whose fields change
 * only in prefld, emfild and the renewal loop.  With lazily held ex..bz the
 * renewal must be mrg_renew_fields_host.                                     */
int mrg_set_fields_lazy(mrg_ctx* ctx, uint32_t mask, const double* const f12[12]);
/* mrg_renew_fields for lazily held fields: old6 = the host's ex0..bz0 arrays
 * (which the host's own renewal loop has just filled with ex..bz); planes not
 * yet on the device are later fetched from them.  old6 may be NULL when no
 * field is held lazily.                                                      */
int mrg_renew_fields_host(mrg_ctx* ctx, const double* const old6[6]);

/* entry prefld of emfild (F:3820-3873) on the device copies of COMMON /fields/:
 * bx,by,bz = b0 - dt*curl(aimpl*e + (1-aimpl)*e0) on the interior nodes, with
 * the reference's wall rows (j = 0, my: one-sided, by = 0).  Bit-identical to
 * the host's prefld (tests), so a host that keeps calling its own prefld need
 * not upload bx,by,bz before the ipc >= 1 calls: SURVEY 8(f1), the first piece
 * of the field-side assembly that reads only what the particle path already
 * holds on the device.  Needs whole (not lazily held) ex..ez, ex0..bz0.      */
int mrg_prefld(mrg_ctx* ctx, double dt, double aimpl);
/* bx,by,bz as emfild leaves them behind its solve (F:4238-4302): the same
 * update from the NEW ex,ey,ez (uploaded first), and with smooth != 0 -- the
 * steps with mod(it,5) = 1 -- outmesh3 + filt3e(sym=+1) of the three arrays
 * (F:4298-4302).  Bit-identical to the host's emfild, so only ex,ey,ez cross
 * PCIe after the field solve.  Needs ifilx = ifily = ifilz = 1 (F:368-370). */
int mrg_update_b(mrg_ctx* ctx, double dt, double aimpl, int32_t smooth);
/* The device copies of the selected members of COMMON /fields/, to the host.  */
int mrg_get_fields(mrg_ctx* ctx, uint32_t mask, double* const f12[12]);

/* "Renewal: ex0 <- ex" of trans, F:796-807, on the device copies: after the
 * host has done that loop on its own arrays it calls this instead of
 * uploading ex0..bz0 again (they equal the ex..bz the device already has).  */
int mrg_renew_fields(mrg_ctx* ctx);

/* fulmov, F:1044-1390, for this rank's particles of species ksp (1-based):
 *   section 0 (F:1127-1152)  blend + outmesh3 + filt3e, cached while neither
 *                            the fields nor aimpl/bxc../ifil* change;
 *   ipc >= 1  (F:1162-1309, 1375-1386) half-step gather, implicit rotation,
 *             predicted position/velocity, partbc, srimp1 + srimp2 scatter,
 *             NCCL sum over ranks, vmesh3/vmesh1 fold;
 *   ipc == 0  (F:1162-1365) the same gather/rotation, in-place update,
 *             partbc, E x B drive kick with the rank's ranfp stream.
 * wkix/wkih receive the rank-summed values of F:1312-1317 (the tiled kernels
 * add their warp partials with fp64 atomics, so the last bits of these two
 * diagnostics depend on the order of the adds, like any parallel sum).  ranfb is the
 * rank's COMMON /ranfb/ state (in/out; only ipc==0 advances it).            */
int mrg_fulmov(mrg_ctx* ctx, int32_t ksp, double qmult, double wmult,
               int32_t ipc, const mrg_step_params* p, int32_t* ranfb,
               double* wkix, double* wkih);

/* Moments of the last ipc>=1 call for species ksp into the caller's
 * COMMON /srimp7/ arrays qix|qex, qiy|qey, qiz|qez, qi|qe (F:1067), each
 * mxyzA doubles on the host.  folded=1: after vmesh3/vmesh1 (what srimp1/2
 * return); folded=0: the rank-summed arrays before the fold.  NULL pointers
 * are skipped.                                                              */
int mrg_get_moments(mrg_ctx* ctx, int32_t ksp, double* qjx, double* qjy,
                    double* qjz, double* q, int32_t folded);
/* Deferred mode only: host arrays (the caller's COMMON /srimp7/ members,
 * ideally page-locked) that every later mrg_fulmov(ipc >= 1) of species ksp
 * copies its folded moments into right after the fold, on the communication
 * stream, so the transfer overlaps the next species' particle kernel.  A
 * following mrg_get_moments with the same pointers only waits.  NULL = none. */
int mrg_set_moment_sink(mrg_ctx* ctx, int32_t ksp, double* qjx, double* qjy,
                        double* qjz, double* q);
/* Device pointers of the same four arrays (folded), valid until the next
 * mrg_fulmov(ipc>=1) of that species; for a device-resident field solve.    */
int mrg_get_moments_device(mrg_ctx* ctx, int32_t ksp, const double* dev4[4]);

/* The prepared fields exa,eya,eza,bxa,bya,bza of F:1127-1148 (host, mxyzA
 * doubles each; NULL skipped).  Runs the preparation if it is stale.        */
int mrg_get_prepared_fields(mrg_ctx* ctx, const mrg_step_params* p,
                            double* const a6[6]);

/* Re-order the species' particles in HBM by cell of x + lookahead*v (after
 * the periodic/wall wrap).  Pure maintenance: results of every other call
 * are independent of the device order up to fp64 summation order.           */
int mrg_sort(mrg_ctx* ctx, int32_t ksp, double lookahead);

/* Kernel selection and tuning knobs (name/value); unknown names fail.
 *   "deposit"    0 = per-particle global atomics, 1 = warp pre-reduction
 *                (ballot/shuffle) per 32 particles then atomics, 2 = cell-run
 *                register accumulation + warp pre-reduction (default)
 *   "tile"       1 (default) = after mrg_sort, particle passes run on
 *                TMA-staged shared-memory field tiles, particles stream
 *                through tensor-TMA stages, cell-run totals of the moments
 *                leave with red.global.add.f64; 0 = gather through L1 only
 *                (any particle order)
 *   "fused_sort" 1 (default) = the tiled predictor emits the cell keys of the
 *                next order and the tiled corrector writes the updated
 *                particles straight into that order, so mrg_sort(ksp, hdt)
 *                after a predictor+corrector pair returns at once
 *   "fused_keys" 1 (default) = a tiled corrector that does not scatter emits
 *                the next sort keys (cell of x + hdt*v), so mrg_sort(ksp, hdt)
 *                skips its key pass
 *   "shard"      ownership used by mrg_loadpt when nranks > 1: 0 (default) = the
 *                reference's round-robin l = rank+1 (mod nranks), F:1162;
 *                1 = the particles whose INITIAL z lies in z slab `rank` of
 *                nranks equal slabs (local order = increasing l).  Moments
 *                summed over ranks do not depend on the ownership; with
 *                replicated grids the slab choice keeps every GPU's particles
 *                dense in the cells it touches
 *   "slab_of", "slab_index"  sizing aid for "shard" = 1: mrg_loadpt loads z
 *                slab slab_index of slab_of whatever nranks is, so that one
 *                GPU can hold exactly what one rank of a larger job holds
 *                (slab_of = 0, the default, means nranks / rank)
 *   "planes"     restricted field preparation: the tiled corrector and
 *                mrg_sort record which z planes the next pass gathers from,
 *                and section 0 (blend/filter/ghost fill) then runs on those
 *                planes and their stencil neighbours only; a rank that owns a
 *                z slab prepares its slab instead of the whole replicated
 *                grid.  -1 (default) = on when nranks > 1, 0 = off, 1 = on.
 *                Needs |vz|*dt < hz (true for |v| < c when c*dt < hz; the
 *                reference runs hz = 7.5, dt = 1.2)
 *   "kick"       random numbers of the E x B drive kick (F:1342-1364).  0 = the
 *                reference's: every rank draws ranfp once per owned particle
 *                inside the slab, in l order (bit-exact with the reference for
 *                its round-robin ownership, any nranks).  1 = the particle with
 *                original local index id takes draw id+1 of the stream starting
 *                at *ranfb, and *ranfb advances by the number of owned
 *                particles per call: same LCG, same kick probability, no serial
 *                order -- the kick then happens inside the tiled corrector
 *                instead of through an index bitmap, a scan and a second
 *                kernel.  -1 (default) = 1 when "shard" = 1 (z-slab ownership
 *                has no reference stream to reproduce), else 0
 *   "compact"    rank sum of the moments for ranks that own z slabs: -1
 *                (default) = when every rank's deposits stay within 6 planes
 *                of its own block of mz/nranks planes (decided from the
 *                recorded planes and agreed between the ranks in the
 *                preceding ipc == 0 call), each rank adds its two neighbours'
 *                boundary strips (ncclSend/Recv) and the complete blocks are
 *                all-gathered in place -- half the bytes of the whole-grid
 *                allreduce, same sums; 0 = always ncclAllReduce.  Every rank
 *                must use the same setting, and calls that invalidate the
 *                agreement (mrg_upload_particles, mrg_loadpt, options "planes",
 *                "compact") are collective: every rank makes them between the
 *                same two steps.  A violated precondition (|vz| dt >= hz) would
 *                drop charge: mrg_self_check's sums[3] = qmult * N detects it
 *   "defer"      1 = mrg_fulmov(ipc >= 1) returns once its work is queued: the
 *                NCCL moment sum and the fold run on a second stream and
 *                overlap the next species' particle kernel.  *wkix, *wkih are
 *                then written when the host next waits for that species:
 *                mrg_get_moments, mrg_get_moments_device, the next
 *                mrg_fulmov(ipc >= 1) of the species, or mrg_synchronize (the
 *                pointers must stay valid until then).  0 (default) = every
 *                call completes before it returns
 *   "sink_share" 1 = the host arrays given to mrg_set_moment_sink /
 *                mrg_get_moments(folded = 1) are SHARED by the ranks of the
 *                node (POSIX shared memory): every rank copies only its own
 *                block of z planes (rank 0 and the last rank include the ghost
 *                planes), the blocks tile each array exactly once, and the
 *                caller synchronises the ranks (MPI_Barrier) before reading.
 *                Removes the N-fold D2H of the replicated moments.  0 (default)
 *                = every rank receives the whole arrays
 *   "iters"      particles per warp / 32 of the untiled predictor (4..32)
 *   "group_min"  smallest stray group (particles) that is pre-reduced        */
int mrg_set_option(mrg_ctx* ctx, const char* name, int64_t value);

/* Counters since the last reset: [0] kernels launched by this library,
 * [1] bytes copied host->device, [2] bytes copied device->host.             */
int mrg_get_counters(mrg_ctx* ctx, int64_t out[3], int32_t reset);

/* Device-time of the particle kernel of the last mrg_fulmov call (CUDA
 * events on the library's stream), in milliseconds.                         */
int mrg_last_kernel_ms(mrg_ctx* ctx, double* ms);

/* Same for the last call of species ksp with ipc == 0 / ipc >= 1 (waits for
 * that kernel only); usable in deferred mode.                               */
int mrg_pass_ms(mrg_ctx* ctx, int32_t ksp, int32_t ipc, double* ms);

/* Since the last reset: [0] runs of section 0 (field preparation), [1] of
 * which restricted to a plane set, [2] planes finalized by those, [3] rank
 * sums of the moments done slab-wise instead of by a whole-grid allreduce.  */
int mrg_get_prep_stats(mrg_ctx* ctx, int64_t out[4], int32_t reset);

/* Host-only helper behind option "planes" (no GPU needed; exported so the
 * dependency analysis can be tested on its own): occ[kp], kp = 0..mz, marks
 * the z planes of the particles' gather cells; out come the planes to blend
 * (listB, k values), to filter (listGI, k values) and to finalize (listG,
 * extended index k+2; bit 30 marks a NaN guard plane), each with room for
 * mz+4 entries, and their lengths n[0..2].                                  */
int mrg_plane_sets(int32_t mz, const uint8_t* occ, int32_t* listB,
                   int32_t* listGI, int32_t* listG, int32_t n[3]);

/* Host-only helper behind option "compact" (no GPU needed; exported so the
 * exchange pattern can be tested on its own).  out[0] = 1 when a grid of mz
 * planes can be exchanged slab-wise by nranks ranks; out[1..4] = first
 * extended plane (k+2) of the strips rank sends to its upper / lower ring
 * neighbour and of the regions where it adds what it receives from the lower
 * / upper neighbour; out[5] = strip width in planes; out[6] = 1 when a rank
 * whose gather cells lie on the planes occ[kp], kp = 0..mz (may be NULL), is
 * eligible.                                                                 */
int mrg_compact_layout(int32_t mz, int32_t nranks, int32_t rank,
                       const uint8_t* occ, int32_t out[7]);

/* Device-side stopwatch on the library's stream (where every kernel of this
 * context is launched): record marks slot 0..7, elapsed gives the CUDA-event
 * time between two recorded slots in milliseconds (synchronises on b).       */
int mrg_event_record(mrg_ctx* ctx, int32_t slot);
int mrg_event_elapsed_ms(mrg_ctx* ctx, int32_t a, int32_t b, double* ms);

/* NVLink peer memory for the slab-wise moment exchange (option "compact"): every
 * rank exports a cudaIpc handle of its raw-moment array of species ksp, the
 * host carries the 64 bytes to the other ranks (MPI_Allgather / torch
 * all_gather) and every rank imports its peers' handles.  Once all nranks-1
 * peers are mapped, a rank finishes the exchange with ONE kernel that adds the
 * neighbour strips into its own block and stores the block into every peer's
 * array over NVLink (instead of ncclAllGather + two ncclBroadcast); option
 * "peer_push" = number of CTAs of that kernel (default 64; 0 keeps the NCCL path); "peer_push_last" = CTAs when the
 * species is the last of the step (ksp = nspecies), whose exchange no particle kernel overlaps (default 296; 0 = same as
 * "peer_push").  Ranks must be processes of one node with peer access (NVSwitch); at most 8 ranks.  mrg_peer_pushes counts the
 * exchanges finished that way.                                               */
#define MRG_IPC_HANDLE_BYTES 64
int mrg_peer_export(mrg_ctx* ctx, int32_t ksp, unsigned char handle[MRG_IPC_HANDLE_BYTES]);
int mrg_peer_import(mrg_ctx* ctx, int32_t ksp, int32_t rank, const unsigned char handle[MRG_IPC_HANDLE_BYTES]);
int64_t mrg_peer_pushes(mrg_ctx* ctx, int32_t reset);
/* Option "split_push" (default 1 = last species of the step, 2 = every species, 0 = off): in deferred mode with mapped
 * peers the tiled predictor of such a call runs as two launches (the pencils of the first three quarters of the rank's z
 * block, then the rest); the planes of the block no later pencil and no neighbour strip can touch are pushed to the peers
 * while the second launch runs, the fused add+push kernel after it sends the remainder.  Needs the order of the fused
 * sort (pencil = gather cell).  mrg_split_pushes counts the calls that did it.                                       */
int64_t mrg_split_pushes(mrg_ctx* ctx, int32_t reset);

/* Device time of the phases of the mrg_fulmov calls since the last reset, in
 * milliseconds, measured with CUDA events on the stream each phase runs on
 * (option "phases" = 1 turns the recording on; it costs a few event records
 * per call).  out[MRG_PH_PREP] field preparation (F:1127-1148, incl. lazy
 * plane uploads), [MRG_PH_SETUP] memsets, histogram scans and other work
 * around the particle kernel, [MRG_PH_KERNEL] the particle kernel,
 * [MRG_PH_SUM] the rank sum of the moments / of wkix,wkih (NCCL), [MRG_PH_FOLD]
 * vmesh fold + unpack, [MRG_PH_KICK] the serial-order drive kick chain.
 * Phases on the two streams overlap in deferred mode, so the sum can exceed
 * the step time; *calls = number of mrg_fulmov calls covered.                */
#define MRG_PH_PREP 0
#define MRG_PH_SETUP 1
#define MRG_PH_KERNEL 2
#define MRG_PH_SUM 3
#define MRG_PH_FOLD 4
#define MRG_PH_KICK 5
#define MRG_NPHASE 6
int mrg_phase_ms(mrg_ctx* ctx, double out[MRG_NPHASE], int64_t* calls, int32_t reset);
/* The same timers per species and kind of call (ipc = 0 | 1), with the rank sum of an ipc >= 1 call split further:
 * [MRG_PH_STRIPS] the strips exchanged with the two ring neighbours (ncclSend/ncclRecv; includes waiting for the
 * neighbours' particle kernels), [MRG_PH_PUSH] the add + push / all-gather of the blocks, [MRG_PH_BARRIER] the
 * 2-double all-reduce that completes it (waits for the slowest rank).  Sums since the last reset of mrg_phase_ms. */
#define MRG_PH_STRIPS 6
#define MRG_PH_PUSH 7
#define MRG_PH_BARRIER 8
#define MRG_NPHASE_DETAIL 9
int mrg_phase_detail(mrg_ctx* ctx, int32_t ksp, int32_t ipc, double out[MRG_NPHASE_DETAIL]);

/* Cheap invariants of the resident state, for callers that want to check a
 * run without a CPU reference (bench.py prints them with every line):
 *   sums[0..3]  sums over the extended grid of the RAW (rank-summed, unfolded)
 *               qjx,qjy,qjz,q of the last ipc>=1 call.  The scatter weights of
 *               srimp1/srimp2 are a partition of unity (F:2296-2308), so
 *               sums[3] = qmult * (particles of all ranks) up to rounding and
 *               sums[0..2] = qmult * sum of the predicted velocities;
 *   counts[0]   resident particles; counts[1] end slot of the last cell of
 *               the cell index (= counts[0] when every particle is owned by a
 *               cell); counts[2], counts[3] sum and sum of squares (mod 2^64)
 *               of the slots' original local indices -- n(n-1)/2 and
 *               (n-1)n(2n-1)/6 mod 2^64 while the slots hold a permutation,
 *               i.e. no particle was lost or duplicated by the fused sort.    */
int mrg_self_check(mrg_ctx* ctx, int32_t ksp, double sums[4], int64_t counts[4]);

/* Measured fp64 FMA rate of this GPU (thread-level DFMA per second, 8
 * independent chains per thread, all SMs): the denominator of the second
 * roofline bench.py reports next to the HBM one.                             */
int mrg_dfma_peak(mrg_ctx* ctx, double* dfma_per_s);

/* Block the host until all work queued by this context has finished.        */
int mrg_synchronize(mrg_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* MRG_FULMOV_H */
