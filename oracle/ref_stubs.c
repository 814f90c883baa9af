/* ref_stubs.c -- TEST INFRASTRUCTURE ONLY: what the transpiled reference (oracle/_ref/nekref_*.c, produced by
 * oracle/f77c.py from /root/reference) links against in place of the third-party / system pieces that are absent here.
 *
 *  - gslib v1.0.9 (3rd_party/gslib/install:5; not vendored, no network): serial (np = 1) restatement of the published
 *    semantics of gs_setup / gs_op / gs_op_many / gs_op_fields / gs_free as the reference calls them
 *    (core/dssum.f:20,79,198,277; core/hsmg.f:335,350,362; core/navier8.f:177-180): entries of equal non-zero id are
 *    replaced by their +,*,min,max.  Within a group the local combine runs in ascending index order starting from the
 *    first member (gslib's gs_gather over its sorted map).  Fortran codes: dom 1=double 2=float 3=int 4=long,
 *    op 1=+ 2=* 3=min 4=max.
 *  - crs_setup / crs_solve (core/fcrs.c:45-96 over core/crs_xxt.c, which needs gslib): dense Cholesky of the assembled
 *    coarse matrix with XXT's single-rank null-space rule (crs_xxt.c:893-956: last degree of freedom pinned to 0, then the
 *    mean removed).
 *  - crystal router tuple transfer (np = 1: nothing moves), timers, exit hooks, user hooks, file I/O (abort if reached).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;

#ifndef NEKHYB   /* the drop-in variant takes gslib's and crs' Fortran API from libnekb200.so */
/* ---------------------------------------------------------------- gather-scatter ---------------------------------- */
typedef struct {
  int n;        /* length of the id vector */
  int ngroups;
  int *off;     /* CSR over groups with >= 2 members */
  int *idx;
} gs_t;

#define GS_MAX 64
static gs_t *gs_tab[GS_MAX];
static int gs_n = 0;

typedef struct { i64 id; int i; } pair_t;
static int pair_cmp(const void *a, const void *b)
{
  const pair_t *x = a, *y = b;
  if (x->id != y->id) return x->id < y->id ? -1 : 1;
  return x->i < y->i ? -1 : (x->i > y->i);
}

void fgslib_gs_setup_(int *handle, const i64 *id, const int *n, const int *comm, const int *np)
{
  int m = 0, i;
  pair_t *p = malloc(sizeof(pair_t) * (size_t)(*n > 0 ? *n : 1));
  for (i = 0; i < *n; i++)
    if (id[i] != 0) { p[m].id = id[i] < 0 ? -id[i] : id[i]; p[m].i = i; m++; }
  qsort(p, (size_t)m, sizeof(pair_t), pair_cmp);
  gs_t *g = calloc(1, sizeof(gs_t));
  g->n = *n;
  g->off = malloc(sizeof(int) * (size_t)(m + 2));
  g->idx = malloc(sizeof(int) * (size_t)(m + 1));
  int ng = 0, k = 0;
  g->off[0] = 0;
  for (i = 0; i < m;) {
    int j = i;
    while (j < m && p[j].id == p[i].id) j++;
    if (j - i >= 2) {
      for (int q = i; q < j; q++) g->idx[k++] = p[q].i;
      g->off[++ng] = k;
    }
    i = j;
  }
  g->ngroups = ng;
  free(p);
  if (gs_n >= GS_MAX) { fprintf(stderr, "ref_stubs: too many gs handles\n"); abort(); }
  gs_tab[gs_n] = g;
  *handle = gs_n++;
}

void fgslib_gs_free_(const int *handle)
{
  gs_t *g = gs_tab[*handle];
  if (g) { free(g->off); free(g->idx); free(g); gs_tab[*handle] = 0; }
}

#define GS_LOOP(T)                                                                                     \
  for (int q = 0; q < g->ngroups; q++) {                                                               \
    int a = g->off[q], b = g->off[q + 1];                                                              \
    T s = u[(size_t)g->idx[a] * stride];                                                               \
    for (int k = a + 1; k < b; k++) {                                                                  \
      T v = u[(size_t)g->idx[k] * stride];                                                             \
      switch (op) { case 1: s = s + v; break; case 2: s = s * v; break;                                \
                    case 3: s = v < s ? v : s; break; case 4: s = v > s ? v : s; break; }              \
    }                                                                                                  \
    for (int k = a; k < b; k++) u[(size_t)g->idx[k] * stride] = s;                                     \
  }

static void gs_apply(gs_t *g, void *uv, int dom, int op, size_t stride)
{
  if (dom == 1) { double *u = uv; GS_LOOP(double) }
  else if (dom == 2) { float *u = uv; GS_LOOP(float) }
  else if (dom == 3) { int *u = uv; GS_LOOP(int) }
  else if (dom == 4) { i64 *u = uv; GS_LOOP(i64) }
  else { fprintf(stderr, "ref_stubs: gs dom %d\n", dom); abort(); }
}

static gs_t *gs_get(const int *handle)
{
  if (*handle < 0 || *handle >= gs_n || !gs_tab[*handle]) { fprintf(stderr, "ref_stubs: bad gs handle %d\n", *handle); abort(); }
  return gs_tab[*handle];
}

void fgslib_gs_op_(const int *handle, void *u, const int *dom, const int *op, const int *transpose)
{
  gs_apply(gs_get(handle), u, *dom, *op, 1);
}

void fgslib_gs_op_many_(const int *handle, void *u1, void *u2, void *u3, void *u4, void *u5, void *u6,
                        const int *n, const int *dom, const int *op, const int *transpose)
{
  void *u[6] = {u1, u2, u3, u4, u5, u6};
  for (int k = 0; k < *n; k++) gs_apply(gs_get(handle), u[k], *dom, *op, 1);
}

void fgslib_gs_op_fields_(const int *handle, void *u, const int *stride, const int *n, const int *dom, const int *op,
                          const int *transpose)
{
  size_t sz = (*dom == 1 || *dom == 4) ? 8 : 4;
  for (int k = 0; k < *n; k++) gs_apply(gs_get(handle), (char *)u + sz * (size_t)k * (size_t)*stride, *dom, *op, 1);
}

/* np = 1: every tuple already lives on its target rank */
void fgslib_crystal_ituple_transfer_() {}
void fgslib_crystal_tuple_transfer_() {}
void fgslib_crystal_setup_(int *h) { *h = 0; }
void fgslib_crystal_free_() {}

/* ---------------------------------------------------------------- coarse solve ------------------------------------ */
typedef struct {
  int un, cn, null_space;
  int *perm;      /* user index -> compressed dof (-1: id 0) */
  double *L;      /* dense Cholesky factor, cn x cn (row major), of the first m rows/cols */
  int m;
} crs_t;
static crs_t *crs_tab[16];
static int crs_n = 0;

void crs_setup_(int *handle, const int *sid, const int *comm, const int *np, const int *n, const i64 *id, const int *nz,
                const int *Ai, const int *Aj, const double *A, const int *null_space, const double *param,
                const char *datafname, int *ierr)
{
  crs_t *c = calloc(1, sizeof(crs_t));
  int un = *n, i, m = 0;
  pair_t *p = malloc(sizeof(pair_t) * (size_t)(un + 1));
  c->un = un;
  c->perm = malloc(sizeof(int) * (size_t)(un + 1));
  for (i = 0; i < un; i++) { c->perm[i] = -1; if (id[i] != 0) { p[m].id = id[i]; p[m].i = i; m++; } }
  qsort(p, (size_t)m, sizeof(pair_t), pair_cmp);
  int cn = 0;
  for (i = 0; i < m; i++) {
    if (i == 0 || p[i].id != p[i - 1].id) cn++;
    c->perm[p[i].i] = cn - 1;
  }
  free(p);
  c->cn = cn;
  c->null_space = *null_space;
  double *M = calloc((size_t)cn * (size_t)cn + 1, sizeof(double));
  for (i = 0; i < *nz; i++) {
    int r = c->perm[Ai[i]], q = c->perm[Aj[i]];      /* 0-based local indices (fcrs passes them through as uint) */
    if (r >= 0 && q >= 0) M[(size_t)r * cn + q] += A[i];
  }
  c->m = c->null_space ? cn - 1 : cn;
  int mm = c->m;
  /* in-place dense Cholesky of the leading mm x mm block, lower triangle */
  for (int j = 0; j < mm; j++) {
    double d = M[(size_t)j * cn + j];
    for (int k = 0; k < j; k++) d -= M[(size_t)j * cn + k] * M[(size_t)j * cn + k];
    if (!(d > 0)) { fprintf(stderr, "ref_stubs: coarse matrix not SPD at %d (%g)\n", j, d); abort(); }
    d = sqrt(d);
    M[(size_t)j * cn + j] = d;
    for (int r = j + 1; r < mm; r++) {
      double s = M[(size_t)r * cn + j];
      for (int k = 0; k < j; k++) s -= M[(size_t)r * cn + k] * M[(size_t)j * cn + k];
      M[(size_t)r * cn + j] = s / d;
    }
  }
  c->L = M;
  crs_tab[crs_n] = c;
  *handle = crs_n++;
  if (ierr) *ierr = 0;
}

void crs_solve_(const int *handle, double *x, const double *b)
{
  crs_t *c = crs_tab[*handle];
  int cn = c->cn, mm = c->m, i;
  double *v = calloc((size_t)cn + 1, sizeof(double));
  for (i = 0; i < c->un; i++) if (c->perm[i] >= 0) v[c->perm[i]] += b[i];
  for (int r = 0; r < mm; r++) {
    double s = v[r];
    for (int k = 0; k < r; k++) s -= c->L[(size_t)r * cn + k] * v[k];
    v[r] = s / c->L[(size_t)r * cn + r];
  }
  for (int r = mm - 1; r >= 0; r--) {
    double s = v[r];
    for (int k = r + 1; k < mm; k++) s -= c->L[(size_t)k * cn + r] * v[k];
    v[r] = s / c->L[(size_t)r * cn + r];
  }
  if (c->null_space) {
    v[cn - 1] = 0;
    double s = 0;
    for (i = 0; i < cn; i++) s += v[i] / cn;
    for (i = 0; i < cn; i++) v[i] -= s;
  }
  for (i = 0; i < c->un; i++) x[i] = c->perm[i] >= 0 ? v[c->perm[i]] : 0;
  free(v);
}

void crs_free_(const int *handle) {}

/* ---------------------------------------------------------------- system / hooks ---------------------------------- */
#else
void fgslib_crystal_ituple_transfer_() {}
void fgslib_crystal_tuple_transfer_() {}
void fgslib_crystal_setup_(int *h) { *h = 0; }
void fgslib_crystal_free_() {}
#endif
double etime_(float *t) { return 0.0; }
#include <time.h>
double dnekclock_(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
double dnekclock_sync_(void) { return dnekclock_(); }
void cexit_(void) { fprintf(stderr, "ref_stubs: cexit\n"); abort(); }
#include <execinfo.h>
static void where(void)
{ /* which translated routine gave up (the reference's own message went to a dropped WRITE) */
  void *bt[16];
  int n = backtrace(bt, 16);
  backtrace_symbols_fd(bt, n, 2);
}
void exitt_(void) { fprintf(stderr, "ref_stubs: the reference called exitt\n"); where(); abort(); }
void exitti_(const char *msg, const int *i, long len)
{
  fprintf(stderr, "ref_stubs: the reference called exitti: %.*s %d\n", (int)len, msg, *i);
  abort();
}
void exittr_(const char *msg, const double *r, long len)
{
  fprintf(stderr, "ref_stubs: the reference called exittr: %.*s %g\n", (int)len, msg, *r);
  abort();
}
void usrsetvert_() {}
void printpartstat_() {}
void nekgsync_() {}
#define DEAD(name) void name() { fprintf(stderr, "ref_stubs: " #name " reached\n"); abort(); }
DEAD(byte_open_) DEAD(byte_close_) DEAD(byte_write_) DEAD(byte_read_) DEAD(mpi_file_open_) DEAD(mpi_file_close_)
DEAD(mpi_file_set_view_) DEAD(mpi_file_write_all_) DEAD(fem_amg_solve_) DEAD(fem_amg_setup_) DEAD(outpost_)
DEAD(outpost2_) DEAD(fgslib_gs_unique_)

/* ---------------------------------------------------------------- write(6,...) trace (oracle/f77c.py trace_write) ------------- */
/* The reference logs some quantities and keeps them nowhere else (cggo's residual per iteration, core/hmholtz.f:770-773).
 * In routines named in ref_build.TRACE_UNITS the translator turns write(6,...) into f77_trace(unit, n, numeric items...);
 * records are kept here while tracing is on and read back by oracle/ref.py (Ref.trace). */
#include <stdarg.h>
int f77_trace_on = 0;
#define TRACE_MAX 65536
#define TRACE_VALS 12
typedef struct { char unit[24]; int n; double v[TRACE_VALS]; } trace_rec_t;
static trace_rec_t *trace_buf = 0;
static int trace_n = 0;
void f77_trace(const char *unit, int n, ...)
{
  if (!trace_buf) trace_buf = (trace_rec_t *)calloc(TRACE_MAX, sizeof(trace_rec_t));
  if (trace_n >= TRACE_MAX || !trace_buf) return;
  trace_rec_t *r = &trace_buf[trace_n++];
  strncpy(r->unit, unit, sizeof r->unit - 1);
  r->n = n > TRACE_VALS ? TRACE_VALS : n;
  va_list ap;
  va_start(ap, n);
  for (int i = 0; i < r->n; i++) r->v[i] = va_arg(ap, double);
  va_end(ap);
}
void nekref_trace_enable(int on) { f77_trace_on = on; trace_n = 0; }
int nekref_trace_count(void) { return trace_n; }
int nekref_trace_get(int i, char *unit24, double *vals12)
{
  if (i < 0 || i >= trace_n) return -1;
  memcpy(unit24, trace_buf[i].unit, 24);
  memcpy(vals12, trace_buf[i].v, sizeof(double) * TRACE_VALS);
  return trace_buf[i].n;
}
