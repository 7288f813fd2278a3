/* ref_runtime.c -- run-time support for the reference's own routines translated by oracle/f03c.py.
 *
 * TEST INFRASTRUCTURE (see the header of oracle/f03c.py).  This file is hand-written and contains no reference
 * code; it #includes the generated translation (oracle/_ref/mrgref_gen.c, never committed) and provides
 *   - the arithmetic helpers the generated code calls (integer powers in the __powidf2 multiplication order
 *     gfortran uses, min/max/sign);
 *   - simulated MPI: one rank = one persistent thread with private COMMON storage; mpi_allreduce sums the
 *     ranks' buffers in rank order (MPI leaves the order open), mpi_allgather concatenates them;
 *   - a small C API for the tests: set the compile-time sizes of param_080A.h, start a pool of ranks, look up
 *     COMMON members (as any translated unit declares them), call a translated unit on all ranks at once.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define REF_MPI_REAL8 1
#define REF_MPI_DOUBLE_PRECISION 1
#define REF_MPI_INTEGER 2
#define REF_MPI_SUM 1
#define REF_MPI_COMM_WORLD 0
#define REF_MPI_ANY_TAG (-1)
#define REF_MPI_STATUS_SIZE 6
#define REF_MAX_RANKS 64

static inline double ref_powi(double x, int n) {
  /* libgcc __powidf2: what gfortran emits for real**integer (and folds for constant exponents) */
  unsigned m = n < 0 ? -(unsigned)n : (unsigned)n;
  double y = (m & 1u) ? x : 1.0;
  while (m >>= 1) {
    x = x * x;
    if (m & 1u) y = y * x;
  }
  return n < 0 ? 1.0 / y : y;
}
static inline float ref_powif(float x, int n) {
  unsigned m = n < 0 ? -(unsigned)n : (unsigned)n;
  float y = (m & 1u) ? x : 1.0f;
  while (m >>= 1) {
    x = x * x;
    if (m & 1u) y = y * x;
  }
  return n < 0 ? 1.0f / y : y;
}
static inline int ref_ipow(int b, int n) {
  int y = 1;
  if (n < 0) return (b == 1) ? 1 : ((b == -1) ? ((n & 1) ? -1 : 1) : 0);
  while (n-- > 0) y *= b;
  return y;
}
static inline int ref_imin(int a, int b) { return a < b ? a : b; }
static inline int ref_imax(int a, int b) { return a > b ? a : b; }
static inline float ref_fmin(float a, float b) { return a < b ? a : b; }
static inline float ref_fmax(float a, float b) { return a > b ? a : b; }
static inline double ref_dmin(double a, double b) { return a < b ? a : b; }
static inline double ref_dmax(double a, double b) { return a > b ? a : b; }
static inline int ref_isign(int a, int b) { a = a < 0 ? -a : a; return b >= 0 ? a : -a; }
static inline float ref_fsign(float a, float b) { return copysignf(a, b); }
static inline double ref_dsign(double a, double b) { return copysign(a, b); }

static void ref_stop(int line) {
  fprintf(stderr, "reference STOP at F:%d\n", line);
  abort();
}

/* ---- simulated MPI --------------------------------------------------------------------------------- */
static int g_nranks = 1;
static __thread int t_rank = 0;
static pthread_barrier_t g_bar;
static void *g_slot[REF_MAX_RANKS];
static double g_t_allreduce[REF_MAX_RANKS];     /* seconds spent inside the collectives, per rank */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void ref_mpi_barrier(void) {
  if (g_nranks > 1) pthread_barrier_wait(&g_bar);
}

static void ref_mpi_allreduce(void *s, void *r, int count, int dtype) {
  const double t0 = now_s();
  const size_t esz = (dtype == REF_MPI_INTEGER) ? 4 : 8;
  if (g_nranks == 1) {
    if (s != r) memmove(r, s, esz * (size_t)count);
    return;
  }
  g_slot[t_rank] = s;
  pthread_barrier_wait(&g_bar);
  void *tmp = malloc(esz * (size_t)count);
  if (dtype == REF_MPI_INTEGER) {
    int *o = (int *)tmp;
    for (int i = 0; i < count; i++) {
      int a = ((int *)g_slot[0])[i];
      for (int k = 1; k < g_nranks; k++) a += ((int *)g_slot[k])[i];
      o[i] = a;
    }
  } else {
    double *o = (double *)tmp;
    for (int i = 0; i < count; i++) {
      double a = ((double *)g_slot[0])[i];
      for (int k = 1; k < g_nranks; k++) a += ((double *)g_slot[k])[i];
      o[i] = a;
    }
  }
  pthread_barrier_wait(&g_bar);
  memcpy(r, tmp, esz * (size_t)count);
  free(tmp);
  g_t_allreduce[t_rank] += now_s() - t0;
}

static void ref_mpi_allgather(void *s, int scount, void *r, int rcount, int dtype) {
  const size_t esz = (dtype == REF_MPI_INTEGER) ? 4 : 8;
  (void)rcount;
  if (g_nranks == 1) {
    if (s != r) memmove(r, s, esz * (size_t)scount);
    return;
  }
  g_slot[t_rank] = s;
  pthread_barrier_wait(&g_bar);
  char *tmp = (char *)malloc(esz * (size_t)scount * g_nranks);
  for (int k = 0; k < g_nranks; k++) memcpy(tmp + esz * (size_t)scount * k, g_slot[k], esz * (size_t)scount);
  pthread_barrier_wait(&g_bar);
  memcpy(r, tmp, esz * (size_t)scount * g_nranks);
  free(tmp);
}

/* ---- point to point (mpi_isend / mpi_irecv / mpi_wait of the solver's halo exchange, F:6411-6501) -------------------
 * A send is buffered at once (a copy queued for the destination, FIFO per sender), so it is complete when it returns and
 * its request is null; a receive is recorded in a request and satisfied by mpi_wait, which blocks until the sender's next
 * message has arrived.  Tags are not matched: every exchange of the reference uses tag 0 / mpi_any_tag. */
typedef struct ref_msg { struct ref_msg *next; size_t bytes; char data[]; } ref_msg_t;
static ref_msg_t *g_q_head[REF_MAX_RANKS][REF_MAX_RANKS], *g_q_tail[REF_MAX_RANKS][REF_MAX_RANKS];   /* [dst][src] */
static pthread_mutex_t g_q_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_q_cv = PTHREAD_COND_INITIALIZER;
typedef struct { void *buf; size_t bytes; int src; int used; } ref_req_t;
#define REF_MAX_REQ 16
static __thread ref_req_t t_req[REF_MAX_REQ];

static void ref_mpi_isend(const void *buf, int count, int dtype, int dst, int *request) {
  const size_t bytes = ((dtype == REF_MPI_INTEGER) ? 4 : 8) * (size_t)count;
  *request = 0;
  if (dst < 0 || dst >= g_nranks) { fprintf(stderr, "ref_mpi_isend: rank %d sends to %d of %d ranks\n", t_rank, dst, g_nranks); abort(); }
  ref_msg_t *m = (ref_msg_t *)malloc(sizeof(ref_msg_t) + bytes);
  m->next = NULL; m->bytes = bytes;
  memcpy(m->data, buf, bytes);
  pthread_mutex_lock(&g_q_mu);
  if (g_q_tail[dst][t_rank]) g_q_tail[dst][t_rank]->next = m; else g_q_head[dst][t_rank] = m;
  g_q_tail[dst][t_rank] = m;
  pthread_cond_broadcast(&g_q_cv);
  pthread_mutex_unlock(&g_q_mu);
}
static void ref_mpi_irecv(void *buf, int count, int dtype, int src, int *request) {
  if (src < 0 || src >= g_nranks) { fprintf(stderr, "ref_mpi_irecv: rank %d receives from %d of %d ranks\n", t_rank, src, g_nranks); abort(); }
  for (int k = 0; k < REF_MAX_REQ; k++)
    if (!t_req[k].used) {
      t_req[k].used = 1; t_req[k].buf = buf; t_req[k].src = src;
      t_req[k].bytes = ((dtype == REF_MPI_INTEGER) ? 4 : 8) * (size_t)count;
      *request = k + 1;
      return;
    }
  fprintf(stderr, "ref_mpi_irecv: too many pending requests\n"); abort();
}
static void ref_mpi_wait(int *request) {
  if (*request <= 0) return;                 /* a completed (buffered) send */
  ref_req_t *r = &t_req[*request - 1];
  const double t0 = now_s();
  pthread_mutex_lock(&g_q_mu);
  while (!g_q_head[t_rank][r->src]) pthread_cond_wait(&g_q_cv, &g_q_mu);
  ref_msg_t *m = g_q_head[t_rank][r->src];
  g_q_head[t_rank][r->src] = m->next;
  if (!m->next) g_q_tail[t_rank][r->src] = NULL;
  pthread_mutex_unlock(&g_q_mu);
  memcpy(r->buf, m->data, m->bytes < r->bytes ? m->bytes : r->bytes);
  free(m);
  r->used = 0;
  *request = 0;
  g_t_allreduce[t_rank] += now_s() - t0;
}

/* ---- unformatted sequential files (write(u) list / read(u) list of restrt, F:9696-9751) ----------------------------
 * What a Fortran run-time library does for form='unformatted': the items of one statement form one record, framed by two
 * 4-byte length markers; a record longer than the sub-record limit (gfortran: 2^31 - 9 bytes) is split, a negative leading
 * marker saying "continued", a negative trailing marker "has a predecessor".  File names are character expressions the
 * translator does not evaluate: the test sets the path of a unit with ref_set_unit_path.  One open file per unit and rank. */
#define REF_MAX_UNIT 100
static char g_unit_path[REF_MAX_UNIT][512];
static __thread FILE *t_unit[REF_MAX_UNIT];
static __thread struct { int unit, writing; char *buf; size_t len, cap, pos; } t_rec;
static unsigned long long g_max_sub = 2147483639ull;
void ref_set_unit_path(int unit, const char *path) { if (unit >= 0 && unit < REF_MAX_UNIT) snprintf(g_unit_path[unit], sizeof g_unit_path[unit], "%s", path); }
void ref_set_max_subrecord(unsigned long long bytes) { g_max_sub = bytes ? bytes : 2147483639ull; }
static void ref_unit_open(int unit, int replace) {
  if (t_unit[unit]) fclose(t_unit[unit]);
  t_unit[unit] = fopen(g_unit_path[unit], replace ? "wb" : "rb");
  if (!t_unit[unit]) { fprintf(stderr, "ref_unit_open: cannot open unit %d (%s)\n", unit, g_unit_path[unit]); abort(); }
}
static void ref_unit_close(int unit) { if (t_unit[unit]) { fclose(t_unit[unit]); t_unit[unit] = NULL; } }
static void ref_rec_begin(int unit, int writing) {
  t_rec.unit = unit; t_rec.writing = writing; t_rec.len = 0; t_rec.pos = 0;
  if (!t_unit[unit]) { fprintf(stderr, "ref_rec_begin: unit %d is not open\n", unit); abort(); }
  if (writing) return;
  for (;;) {                                    /* gather the sub-records of one record */
    int m0 = 0, m1 = 0;
    if (fread(&m0, 4, 1, t_unit[unit]) != 1) { fprintf(stderr, "ref read: end of file on unit %d\n", unit); abort(); }
    const size_t n = (size_t)(m0 < 0 ? -(long)m0 : m0);
    if (t_rec.len + n > t_rec.cap) { t_rec.cap = (t_rec.len + n) * 2 + 64; t_rec.buf = (char *)realloc(t_rec.buf, t_rec.cap); }
    if (n && fread(t_rec.buf + t_rec.len, 1, n, t_unit[unit]) != n) { fprintf(stderr, "ref read: short record\n"); abort(); }
    t_rec.len += n;
    if (fread(&m1, 4, 1, t_unit[unit]) != 1) { fprintf(stderr, "ref read: missing trailing marker\n"); abort(); }
    if (m0 >= 0) break;
  }
}
static void ref_rec_item(void *p, long bytes) {
  if (t_rec.writing) {
    if (t_rec.len + (size_t)bytes > t_rec.cap) { t_rec.cap = (t_rec.len + (size_t)bytes) * 2 + 64; t_rec.buf = (char *)realloc(t_rec.buf, t_rec.cap); }
    memcpy(t_rec.buf + t_rec.len, p, (size_t)bytes);
    t_rec.len += (size_t)bytes;
  } else {
    if (t_rec.pos + (size_t)bytes > t_rec.len) { fprintf(stderr, "ref read: io-list longer than the record\n"); abort(); }
    memcpy(p, t_rec.buf + t_rec.pos, (size_t)bytes);
    t_rec.pos += (size_t)bytes;
  }
}
static void ref_rec_end(void) {
  if (!t_rec.writing) return;
  FILE *f = t_unit[t_rec.unit];
  size_t off = 0;
  int first = 1;
  do {
    const size_t n = (t_rec.len - off > g_max_sub) ? (size_t)g_max_sub : t_rec.len - off;
    const int more = off + n < t_rec.len;
    const int lead = more ? -(int)n : (int)n, trail = first ? (int)n : -(int)n;
    fwrite(&lead, 4, 1, f);
    if (n) fwrite(t_rec.buf + off, 1, n, f);
    fwrite(&trail, 4, 1, f);
    off += n; first = 0;
  } while (off < t_rec.len);
}

#include "mrgref_gen.c"

/* ---- pool of ranks --------------------------------------------------------------------------------- */
typedef struct {
  const ref_unit_t *unit;
  void **args;          /* [nranks][nargs] */
  double ret[REF_MAX_RANKS];
  double secs[REF_MAX_RANKS];
} job_t;

static pthread_t g_thr[REF_MAX_RANKS];
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv_job = PTHREAD_COND_INITIALIZER, g_cv_done = PTHREAD_COND_INITIALIZER;
static job_t *g_job = NULL;
static long g_job_seq = 0, g_start_seq = 0;
static int g_done = 0, g_quit = 0, g_running = 0;
static char *g_cm[REF_MAX_RANKS][CM_COUNT];
static long g_cm_bytes[CM_COUNT];

typedef void (*fn0)(void);
static double call_unit(const ref_unit_t *u, void **a) {
  void *p[20] = {0};
  for (int i = 0; i < u->nargs && i < 20; i++) p[i] = a[i];
#define ARGS p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9], p[10], p[11], p[12], p[13], p[14], p[15], p[16], p[17], p[18], p[19]
  typedef void (*vf)(void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *,
                     void *, void *, void *, void *, void *);
  typedef int (*jf)(void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *,
                    void *, void *, void *, void *, void *);
  typedef double (*df)(void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *, void *,
                       void *, void *, void *, void *, void *, void *);
  /* every dummy is a pointer, so surplus pointer arguments in registers / on the stack are harmless on x86-64 */
  if (u->rtype == 0) { ((vf)u->fn)(ARGS); return 0.0; }
  if (u->rtype == 3) return ((df)u->fn)(ARGS);
  return (double)((jf)u->fn)(ARGS);
#undef ARGS
}

static void *worker(void *arg) {
  t_rank = (int)(long)arg;
  for (int b = 0; b < CM_COUNT; b++) ref_cm[b] = g_cm[t_rank][b];
  long seen = g_start_seq;      /* jobs posted before this pool started are not ours */
  for (;;) {
    pthread_mutex_lock(&g_mu);
    while (!g_quit && g_job_seq == seen) pthread_cond_wait(&g_cv_job, &g_mu);
    if (g_quit) { pthread_mutex_unlock(&g_mu); break; }
    seen = g_job_seq;
    job_t *j = g_job;
    pthread_mutex_unlock(&g_mu);
    const double t0 = now_s();
    j->ret[t_rank] = call_unit(j->unit, j->args + (size_t)t_rank * j->unit->nargs);
    j->secs[t_rank] = now_s() - t0;
    pthread_mutex_lock(&g_mu);
    if (++g_done == g_nranks) pthread_cond_signal(&g_cv_done);
    pthread_mutex_unlock(&g_mu);
  }
  return NULL;
}

/* ---- exported API ---------------------------------------------------------------------------------- */
int ref_pool_stop(void);

/* sizes of param_080A.h (P:12-16); everything derived from them is recomputed */
int ref_set_params(int npc, int mx, int my, int mz, long np0) {
  if (g_running) return 1;
  ref_default_params();
  int rc = 0;
  rc |= ref_set_param("npc", npc);
  rc |= ref_set_param("mx", mx);
  rc |= ref_set_param("my", my);
  rc |= ref_set_param("mz", mz);
  rc |= ref_set_param("np0", np0);
  ref_derive_params();
  return rc;
}
long ref_param(const char *name) { return ref_get_param(name); }

int ref_pool_start(int nranks) {
  if (g_running || nranks < 1 || nranks > REF_MAX_RANKS) return 1;
  g_nranks = nranks;
  g_quit = 0;
  g_done = 0;
  pthread_barrier_init(&g_bar, NULL, (unsigned)nranks);
  for (int b = 0; b < CM_COUNT; b++) {
    long total = 0;
    ref_member(b, NULL, NULL, NULL, NULL, &total);
    g_cm_bytes[b] = total;
    for (int r = 0; r < nranks; r++) g_cm[r][b] = (char *)calloc((size_t)total + 64, 1);
  }
  for (int r = 0; r < nranks; r++) g_t_allreduce[r] = 0.0;
  g_start_seq = g_job_seq;
  g_job = NULL;
  for (int r = 0; r < nranks; r++) pthread_create(&g_thr[r], NULL, worker, (void *)(long)r);
  g_running = 1;
  return 0;
}

int ref_pool_stop(void) {
  if (!g_running) return 0;
  pthread_mutex_lock(&g_mu);
  g_quit = 1;
  pthread_cond_broadcast(&g_cv_job);
  pthread_mutex_unlock(&g_mu);
  for (int r = 0; r < g_nranks; r++) pthread_join(g_thr[r], NULL);
  for (int b = 0; b < CM_COUNT; b++)
    for (int r = 0; r < g_nranks; r++) { free(g_cm[r][b]); g_cm[r][b] = NULL; }
  pthread_barrier_destroy(&g_bar);
  g_running = 0;
  return 0;
}

/* pointer to a COMMON member of one rank, named as `unit` declares it (unit = NULL: the first unit that knows the name);
 * type: 1 = int32, 2 = float, 3 = double */
void *ref_common(int rank, const char *block, const char *unit, const char *name, long *count, int *type) {
  if (!g_running || rank < 0 || rank >= g_nranks) return NULL;
  for (int b = 0; ref_block_names[b]; b++) {
    if (strcmp(ref_block_names[b], block)) continue;
    long off = ref_member(b, unit, name, count, type, NULL);
    if (off < 0) return NULL;
    return g_cm[rank][b] + off;
  }
  return NULL;
}

int ref_has_unit(const char *name) {
  for (const ref_unit_t *u = ref_units; u->name; u++)
    if (!strcmp(u->name, name)) return u->nargs;
  return -1;
}

/* call unit `name` on every rank at once; args = [nranks][nargs] pointers (every Fortran argument is by reference).
 * ret[nranks] receives function results, secs[nranks] the wall time of each rank.  Returns 0 when the unit exists. */
int ref_call(const char *name, void **args, double *ret, double *secs) {
  if (!g_running) return 2;
  const ref_unit_t *u = ref_units;
  for (; u->name; u++)
    if (!strcmp(u->name, name)) break;
  if (!u->name) return 1;
  job_t j;
  memset(&j, 0, sizeof(j));
  j.unit = u;
  j.args = args;
  pthread_mutex_lock(&g_mu);
  g_job = &j;
  g_done = 0;
  g_job_seq++;
  pthread_cond_broadcast(&g_cv_job);
  while (g_done < g_nranks) pthread_cond_wait(&g_cv_done, &g_mu);
  pthread_mutex_unlock(&g_mu);
  for (int r = 0; r < g_nranks; r++) {
    if (ret) ret[r] = j.ret[r];
    if (secs) secs[r] = j.secs[r];
  }
  return 0;
}

double ref_collective_seconds(int rank, int reset) {
  double v = g_t_allreduce[rank];
  if (reset) g_t_allreduce[rank] = 0.0;
  return v;
}
