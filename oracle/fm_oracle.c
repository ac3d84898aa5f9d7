/* TEST INFRASTRUCTURE ONLY -- not part of the shipped engine.
 *
 * Plain-C restatement of the FMwR hot path (forward / predict and the
 * SGD, FTRL-Proximal, TDAP, ALS and MCMC training loops) used as the parity
 * oracle for the CUDA engine.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose arithmetic and operation ORDER it follows; the restatement is pinned
 * against the reference's own headers (oracle/_ref, built from the reference
 * tree by oracle/Makefile) in tests/test_oracle_pinning.py and against the
 * golden vectors under tests/golden/.
 *
 * Arithmetic outside the reference tree:
 *   - R's Rf_rnorm / Rf_rgamma (libR) and glibc rand(): replaced on both sides by
 *     injected streams (standard normals, unit-scale gammas, raw rand() ints).
 *     Native-RNG streams are therefore "parity unpinned"; injected-stream runs are pinned.
 *   - The two lookup tables (src/util/RandomData.h, RandomData_.h) are NOT copied:
 *     they are regenerated from their defining formulas (Phi on a 1/549.9667 grid,
 *     phi/(1-Phi) on a 2e-4 grid) and checked against the reference tables to 1e-9.
 *
 * Layout convention at this ABI: v is [p][k] (feature-major == R's k x p column-major
 * NumericMatrix memory order); the reference's internal DMatrix is [k][p].
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_abi.h"

typedef unsigned int uint;

/* ------------------------------------------------------------------ streams */
static struct {
  const double* normals; long n_normals, i_normal;
  const double* gammas;  long n_gammas, i_gamma;
  const int*    rands;   long n_rands, i_rand;
  long overrun;
} S;

void fmwr_orc_set_streams(const double* normals, long n_normals, const double* gammas, long n_gammas,
                          const int* rands, long n_rands)
{
  S.normals = normals; S.n_normals = n_normals; S.i_normal = 0;
  S.gammas = gammas; S.n_gammas = n_gammas; S.i_gamma = 0;
  S.rands = rands; S.n_rands = n_rands; S.i_rand = 0;
  S.overrun = 0;
}

void fmwr_orc_stream_pos(long* out4)
{
  out4[0] = S.i_normal; out4[1] = S.i_gamma; out4[2] = S.i_rand; out4[3] = S.overrun;
}

static double draw_norm(double mean, double sd)   /* stands in for Rf_rnorm */
{
  double z = 0.0;
  if (S.normals && S.i_normal < S.n_normals) z = S.normals[S.i_normal]; else S.overrun++;
  S.i_normal++;
  return mean + sd * z;
}

static double draw_gamma(double shape, double scale)   /* stands in for Rf_rgamma */
{
  (void)shape;
  double g = 1.0;
  if (S.gammas && S.i_gamma < S.n_gammas) g = S.gammas[S.i_gamma]; else S.overrun++;
  S.i_gamma++;
  return scale * g;
}

static int draw_rand(void)   /* stands in for glibc rand() */
{
  int r = 0;
  if (S.rands && S.i_rand < S.n_rands) r = S.rands[S.i_rand];
  else if (!S.rands) r = rand();
  else S.overrun++;
  S.i_rand++;
  return r;
}

static char g_err[256];
const char* fmwr_orc_last_error(void) { return g_err; }

/* ------------------------------------------------------------------ tables */
/* src/util/RandomData.h: 2861 points of Phi(x), x_i = i / 549.966731401936, max 5.20031455849973.
 * src/util/RandomData_.h: 40001 points of phi(x)/(1-Phi(x)) on x_i = -3 + 2e-4 i. */
#define PN_N 2861
#define PN_HINV 549.966731401936
#define PN_MAX 5.20031455849973
#define DP_N 40001
#define DP_MIN (-3.0)
#define DP_MAX 5.0
static double pn_x[PN_N + 1], pn_y[PN_N + 1];
static double dp_x[DP_N + 1], dp_y[DP_N + 1];
static int tables_ready = 0;

static void build_tables(void)
{
  if (tables_ready) return;
#pragma omp critical(fmwr_orc_tables)
  {
    if (!tables_ready) {
      for (int i = 0; i <= PN_N; ++i) {
        double x = (double)i / PN_HINV;
        pn_x[i] = x;
        pn_y[i] = 0.5 * erfc(-x / sqrt(2.0));
      }
      for (int i = 0; i <= DP_N; ++i) {
        double x = DP_MIN + 2e-4 * (double)i;
        x = round(x * 1e4) / 1e4;                   /* the reference grid is printed to 4 decimals */
        dp_x[i] = x;
        double tail = 0.5 * erfc(x / sqrt(2.0));    /* 1 - Phi(x), accurate in the right tail */
        dp_y[i] = exp(-0.5 * x * x) / 2.5066282746310002 / tail;
      }
      tables_ready = 1;
    }
  }
}

/* src/util/Random.h:95-111 */
double fmwr_orc_pnorm(double x)
{
  build_tables();
  double ax = x < 0 ? -x : x;
  double res;
  if (ax > PN_MAX) {
    res = 0.999999900524235;
  } else {
    int i = (int)(ax * PN_HINV);
    double w = (ax - pn_x[i]) * PN_HINV;
    res = w * pn_y[i + 1] + (1.0 - w) * pn_y[i];
  }
  if (ax == x) return res;
  return 1.0 - res;
}

/* src/util/Random.h:114-124 */
double fmwr_orc_dpnorm(double x)
{
  build_tables();
  double ax = x < 0 ? -x : x;
  if (x < DP_MIN) return 0.0;
  if (x > DP_MAX) return 0.1943369 + 0.9754752 * x + 0.4136861 * sqrt(ax) - 0.5034295 * log(ax + 1e-07);
  int i = (int)((x - DP_MIN) * 5000);
  double w = (x - dp_x[i]) * 5000;
  return w * dp_y[i + 1] + (1.0 - w) * dp_y[i];
}

void fmwr_orc_tables(double* pnx, double* pny, double* dpx, double* dpy)
{
  build_tables();
  if (pnx) memcpy(pnx, pn_x, sizeof(double) * PN_N);
  if (pny) memcpy(pny, pn_y, sizeof(double) * PN_N);
  if (dpx) memcpy(dpx, dp_x, sizeof(double) * DP_N);
  if (dpy) memcpy(dpy, dp_y, sizeof(double) * DP_N);
}

/* ------------------------------------------------------------------ samplers */
/* src/util/Random.h:20-24 */
static double runif(void) { return draw_rand() / ((double)RAND_MAX + 1); }
/* :26-29 */
static double rexp1(void) { return -log(1 - runif()); }
/* :31-48  Leva's ratio-of-uniforms normal */
static double rnorm_leva(void)
{
  double u, v, av, x, y, Q;
  do {
    do { u = runif(); } while (u == 0.0);
    v = 1.7156 * (runif() - 0.5);
    av = v < 0 ? -v : v;
    x = u - 0.449871;
    y = av + 0.386595;
    Q = x * x + y * (0.19600 * y - 0.25472 * x);
    if (Q < 0.27597) break;
  } while ((Q > 0.27846) || ((v * v) > (-4.0 * u * u * log(u))));
  return v / u;
}
/* :51-76 */
static double trnorm_left_std(double left)
{
  if (left < 0.0) {
    for (;;) { double r = rnorm_leva(); if (r >= left) return r; }
  } else {
    double a = 0.5 * (left + sqrt(left * left + 4.0));
    for (;;) {
      double z = rexp1() / a + left;
      double d = z - a;
      d = exp(-(d * d) / 2);
      double u = runif();
      if (u < d) return z;
    }
  }
}
/* :79-93 */
double fmwr_orc_trnorm_left(double left, double mean, double sd) { return mean + sd * trnorm_left_std((left - mean) / sd); }
double fmwr_orc_trnorm_right(double right, double mean, double sd) { return mean + sd * (-trnorm_left_std(-((right - mean) / sd))); }
/* :126-132 */
unsigned fmwr_orc_random_select(int n) { if (n == 1) return 1; return (uint)(runif() * n + 1); }

/* ------------------------------------------------------------------ forward */
/* Model::predict_batch, src/core/Model.h:106-161.  v is [p][k]. */
static void predict_batch(const fmwr_oracle_cfg* c, uint n, const uint* rowptr, const uint* col, const float* val,
                          double w0, const double* w, const double* v, double* out, int nthreads)
{
  const int k = c->k;
  (void)nthreads;
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (uint i = 0; i < n; ++i) {
    double o = c->k0 ? w0 : 0.0;                      /* :115-116 */
    const uint b = rowptr[i], e = rowptr[i + 1];
    if (c->k1) {                                      /* :121-134 */
      double ws = 0.0;
      for (uint j = b; j < e; ++j) ws += w[col[j]] * val[j];
      o += ws;
    }
    if (k > 0) {                                      /* :136-160: f outer, nnz inner */
      double vres = 0.0;
      for (int f = 0; f < k; ++f) {
        double s = 0, q = 0;
        for (uint j = b; j < e; ++j) {
          double t = val[j] * v[(size_t)col[j] * k + f];
          s += t; q += t * t;
        }
        vres += (0.5 * s * s - 0.5 * q);
      }
      o += vres;
    }
    out[i] = o;
  }
}

/* Model::predict (single row; nnz outer, f inner; leaves S_f in sum[]), src/core/Model.h:75-103 */
static double predict_row(const fmwr_oracle_cfg* c, uint b, uint e, const uint* col, const float* val,
                          double w0, const double* w, const double* v, double* sum, double* sum_sqr)
{
  const int k = c->k;
  double pred = 0.0;
  if (c->k0) pred += w0;
  for (int f = 0; f < k; ++f) { sum[f] = 0.0; sum_sqr[f] = 0.0; }
  for (uint j = b; j < e; ++j) {
    double x = val[j];
    uint idx = col[j];
    if (c->k1) pred += w[idx] * x;
    const double* vr = v + (size_t)idx * k;
    for (int f = 0; f < k; ++f) {
      double t = vr[f] * x;
      sum[f] += t;
      sum_sqr[f] += t * t;
    }
  }
  for (int f = 0; f < k; ++f) pred += 0.5 * (sum[f] * sum[f] - sum_sqr[f]);
  return pred;
}

/* Model::predict_prob link, src/core/Model.h:163-180 */
static double link_prob(int solver, double x)
{
  if (solver == ORC_MCMC || solver == ORC_ALS) return fmwr_orc_pnorm(x);
  return 1.0 / (1.0 + exp(-x));
}

int fmwr_orc_predict(const fmwr_oracle_cfg* c, uint n, uint p, uint nnz, const uint* rowptr, const uint* col,
                     const float* val, double w0, const double* w, const double* v, int link, double* out)
{
  (void)p; (void)nnz;
  build_tables();
  int nt = c->nthreads > 0 ? c->nthreads : 1;
  predict_batch(c, n, rowptr, col, val, w0, w, v, out, nt);
  if (link) {
#pragma omp parallel for num_threads(nt)
    for (uint i = 0; i < n; ++i) out[i] = link_prob(c->solver, out[i]);
  }
  return 0;
}

int fmwr_orc_predict_rows(const fmwr_oracle_cfg* c, uint n, uint p, uint nnz, const uint* rowptr, const uint* col,
                          const float* val, double w0, const double* w, const double* v, double* out, double* sums)
{
  (void)p; (void)nnz;
  double* s = (double*)malloc(sizeof(double) * (c->k + 1) * 2);
  double* q = s + c->k + 1;
  for (uint i = 0; i < n; ++i) {
    if (rowptr[i + 1] == rowptr[i]) { out[i] = NAN; continue; }
    out[i] = predict_row(c, rowptr[i], rowptr[i + 1], col, val, w0, w, v, s, q);
    if (sums) for (int f = 0; f < c->k; ++f) sums[(size_t)i * c->k + f] = s[f];
  }
  free(s);
  return 0;
}

/* ------------------------------------------------------------------ CSR -> CSC */
/* Stable counting sort; equals SMatrix::transpose (src/util/Smatrix.h:155-185) output on
 * data with no empty rows (rows ascending inside each column). */
int fmwr_orc_transpose(uint n, uint p, uint nnz, const uint* rowptr, const uint* col, const float* val, int unused,
                       uint* t_ptr, uint* t_idx, float* t_val)
{
  (void)unused;
  for (uint c = 0; c <= p; ++c) t_ptr[c] = 0;
  for (uint e = 0; e < nnz; ++e) t_ptr[col[e] + 1]++;
  for (uint c = 0; c < p; ++c) t_ptr[c + 1] += t_ptr[c];
  uint* cur = (uint*)malloc(sizeof(uint) * (p + 1));
  memcpy(cur, t_ptr, sizeof(uint) * (p + 1));
  for (uint r = 0; r < n; ++r)
    for (uint e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      uint d = cur[col[e]]++;
      t_idx[d] = r; t_val[d] = val[e];
    }
  free(cur);
  return 0;
}

/* ------------------------------------------------------------------ metrics */
/* src/core/Evaluation.h:20-115 */
static int cmp_abs(const void* a, const void* b)
{
  double x = fabs(*(const double*)a), y = fabs(*(const double*)b);
  return (x < y) ? -1 : (x > y) ? 1 : 0;
}

double fmwr_orc_evaluate(int task, int metric, uint n, const double* y_hat, const float* y_true)
{
  if (task == ORC_REGRESSION) {
    if (metric <= ORC_RMSE) {                         /* :24-26, rmse :91-102 */
      double s = 0.0;
      for (uint i = 0; i < n; ++i) { double e = y_hat[i] - y_true[i]; s += e * e; }
      return sqrt(s / n);
    } else {                                          /* mae, with the reference's stray sqrt :104-115 */
      double s = 0.0;
      for (uint i = 0; i < n; ++i) { double e = y_hat[i] - y_true[i]; s += fabs(e); }
      return sqrt(s / n);
    }
  }
  if (metric >= ORC_ACC) {                            /* accurancy :44-54, cutoff 0.5 */
    uint ok = 0;
    for (uint i = 0; i < n; ++i)
      if (((y_hat[i] >= 0.5) && (y_true[i] > 0)) || ((y_hat[i] < 0.5) && (y_true[i] < 0))) ok++;
    return (double)ok / (double)n;
  }
  if (metric == ORC_LL) {                             /* loglikehood :80-89 */
    double r = 0.0;
    for (uint i = 0; i < n; ++i)
      r += (1 + y_true[i]) * log(y_hat[i] + 1e-20) + (1 - y_true[i]) * log(1 - y_hat[i] - 1e-20);
    return r / 2.0;
  }
  /* auc :56-78: signed scores sorted by |score| (unstable std::sort in the reference; ties only
   * matter between a positive and a negative of identical |score|) */
  double* t = (double*)malloc(sizeof(double) * (n ? n : 1));
  for (uint i = 0; i < n; ++i) t[i] = y_true[i] > 0 ? y_hat[i] : -y_hat[i];
  qsort(t, n, sizeof(double), cmp_abs);
  double area = 0, cum_tp = 0;
  for (uint i = 0; i < n; ++i) { if (t[i] > 0) cum_tp += 1.0; else area += cum_tp; }
  free(t);
  if (cum_tp == 0 || cum_tp == n) return 1.0;
  area /= cum_tp * (n - cum_tp);
  return area < 0.5 ? 1 - area : area;
}

/* ------------------------------------------------------------------ shared learner bits */
/* calculate_grad_mult, identical in src/solver/SGD_Learner.h:180-191, FTRL_Learner.h:204-215, TDAP_Learner.h:235-246 */
static double grad_mult(const fmwr_oracle_cfg* c, double y_hat, float y)
{
  if (c->task == ORC_REGRESSION) {
    y_hat = fmin(c->max_target, y_hat);
    y_hat = fmax(c->min_target, y_hat);
    return -(y - y_hat);
  }
  return -y * (1.0 - 1.0 / (1.0 + exp(-y * y_hat)));
}

typedef struct {
  int max_rec; double* eval_train; int* rec_index; fmwr_oracle_trace_info* info;
  int step_size;          /* possibly re-derived by tracker_init */
  int ii; int n_rec; double old_score; int conv_times; int convergent;
} tracker_t;

/* Tracker::init, src/core/Tracker.h:41-52 (MAX_REC 10000) */
static void tracker_init(tracker_t* t, const fmwr_oracle_cfg* c, int max_rec, double* eval_train, int* rec_index,
                         fmwr_oracle_trace_info* info)
{
  memset(t, 0, sizeof(*t));
  t->max_rec = max_rec; t->eval_train = eval_train; t->rec_index = rec_index; t->info = info;
  t->step_size = c->step_size; t->ii = -1;
  if (t->step_size > 0) {
    int rt = (int)(ceil(((double)c->max_iter - 0.5) / (double)t->step_size)) + 1;
    if (rt > 10000) t->step_size = (int)((double)(c->max_iter + 1) / 10000.0) + 1;
  }
}

static void tracker_record(tracker_t* t, int iter, double score)
{
  if (t->n_rec < t->max_rec) {
    if (t->eval_train) t->eval_train[t->n_rec] = score;
    if (t->rec_index) t->rec_index[t->n_rec] = iter;
  }
  t->n_rec++;
}

/* train-set score used by the trackers: SGD_Learner.h:140-156 (same block in FTRL/TDAP) */
static double train_score(const fmwr_oracle_cfg* c, uint n, const uint* rowptr, const uint* col, const float* val,
                          const float* y, double w0, const double* w, const double* v, double* buf)
{
  predict_batch(c, n, rowptr, col, val, w0, w, v, buf, 1);
  if (c->task == ORC_REGRESSION) {
    for (uint i = 0; i < n; ++i) { if (buf[i] < c->min_target) buf[i] = c->min_target; else if (buf[i] > c->max_target) buf[i] = c->max_target; }
  } else {
    for (uint i = 0; i < n; ++i) buf[i] = link_prob(c->solver, buf[i]);
  }
  return fmwr_orc_evaluate(c->task, c->metric, n, buf, y);
}

/* the per-sample tracker / convergence block shared by SGD/FTRL/TDAP (SGD_Learner.h:140-175).
 * returns 1 when the sample loop must stop.  SGD uses integer abs() on a double in the
 * reference (SGD_Learner.h:157) -- with <cmath> in scope it resolves to the double overload. */
static int tracker_step(tracker_t* t, const fmwr_oracle_cfg* c, int* iter, uint n, const uint* rowptr, const uint* col,
                        const float* val, const float* y, double w0, const double* w, const double* v, double* buf)
{
  if (t->step_size > 0) {
    t->ii++;
    if (t->ii == t->step_size) t->ii = 0;
    if (t->ii == 0 || *iter == c->max_iter - 1) {
      double s = train_score(c, n, rowptr, col, val, y, w0, w, v, buf);
      if (*iter > t->step_size && fabs((s - t->old_score) / (t->old_score + 1e-30)) <= c->convergence) t->conv_times++;
      else t->conv_times = 0;
      t->old_score = s;
      tracker_record(t, *iter, s);
    }
  }
  (*iter)++;
  if (t->conv_times >= 3) { t->convergent = 1; return 1; }
  if (*iter >= c->max_iter) return 1;
  return 0;
}

/* apply_penalty (cumulative L1, Tsuruoka), src/solver/SGD_Learner.h:195-204 */
static void apply_penalty(double* theta, double u, double* q)
{
  double old = *theta;
  if (*theta > 0) *theta = fmax(0.0, old - (u + *q));
  else if (*theta < 0) *theta = fmin(0.0, old + (u - *q));
  *q += *theta - old;
}

/* ------------------------------------------------------------------ SGD */
/* SGD_Learner::init + learn, src/solver/SGD_Learner.h:44-178 */
static void learn_sgd(const fmwr_oracle_cfg* c, uint n, uint p, const uint* rowptr, const uint* col, const float* val,
                      const float* y, double* w0, double* w, double* v, tracker_t* t)
{
  const int k = c->k;
  int l1 = 0; double regw, regv;
  if (c->l1_w > 0 || c->l1_v > 0) { l1 = 1; regw = c->l1_w; regv = c->l1_v; }     /* :46-55 */
  else { regw = c->l2_w; regv = c->l2_v; }
  if (c->task != ORC_CLASSIFICATION) l1 = 0;                                         /* :57-59: L1 rates become L2 rates */
  double *q_w = NULL, *q_v = NULL, u_w = 0.0, u_v = 0.0;
  if (l1) { q_w = (double*)calloc(p, sizeof(double)); q_v = (double*)calloc((size_t)p * (k ? k : 1), sizeof(double)); }
  double* sum = (double*)malloc(sizeof(double) * (2 * k + 2));
  double* sum_sqr = sum + k + 1;
  double* buf = t->step_size > 0 ? (double*)malloc(sizeof(double) * n) : NULL;
  const double lr = c->learn_rate;
  int iter = 0, stop = 0;
  for (;;) {
    for (uint i = fmwr_orc_random_select(c->random_step); i < n; i += fmwr_orc_random_select(c->random_step)) {  /* :88, F5 */
      if (l1) { u_w += lr * regw; u_v += lr * regv; }                                /* :92-97 */
      const uint b = rowptr[i], e = rowptr[i + 1];
      double y_hat = predict_row(c, b, e, col, val, *w0, w, v, sum, sum_sqr);
      double mult = grad_mult(c, y_hat, y[i]);
      if (c->k0) *w0 -= lr * (mult + c->l2_w0 * (*w0));                              /* :106-109 */
      if (c->k1) {                                                                   /* :111-122 */
        for (uint j = b; j < e; ++j) {
          double* wj = &w[col[j]];
          *wj -= lr * mult * val[j];
          if (l1) apply_penalty(wj, u_w, &q_w[col[j]]);
          else *wj -= lr * regw * (*wj);
        }
      }
      for (int f = 0; f < k; ++f) {                                                  /* :124-138: f outer, S_f frozen */
        double sf = sum[f];
        for (uint j = b; j < e; ++j) {
          double* vj = &v[(size_t)col[j] * k + f];
          double x = val[j];
          double grad = sf * x - (*vj) * x * x;
          *vj -= lr * mult * grad;
          if (l1) apply_penalty(vj, u_v, &q_v[(size_t)col[j] * k + f]);
          else *vj -= lr * regv * (*vj);
        }
      }
      if (tracker_step(t, c, &iter, n, rowptr, col, val, y, *w0, w, v, buf)) { stop = 1; break; }
    }
    if (stop) break;
    if (n <= 1 && c->random_step == 1) break;   /* guard: the reference would spin forever on n<=1 */
  }
  if (t->info) t->info->iters_done = iter;
  free(sum); free(buf); free(q_w); free(q_v);
}

/* ------------------------------------------------------------------ FTRL-Proximal */
/* FTRL_Learner::init + learn + calculate_param, src/solver/FTRL_Learner.h:48-202 */
static void learn_ftrl(const fmwr_oracle_cfg* c, uint n, uint p, const uint* rowptr, const uint* col, const float* val,
                       const float* y, double* w0, double* w, double* v, tracker_t* t)
{
  const int k = c->k;
  double z_w0 = 0.0, n_w0 = 0.0;
  double* z_w = (double*)calloc(p, sizeof(double));
  double* n_w = (double*)calloc(p, sizeof(double));
  double* z_v = (double*)calloc((size_t)p * (k ? k : 1), sizeof(double));
  double* n_v = (double*)calloc((size_t)p * (k ? k : 1), sizeof(double));
  double* sum = (double*)malloc(sizeof(double) * (2 * k + 2));
  double* sum_sqr = sum + k + 1;
  double* buf = t->step_size > 0 ? (double*)malloc(sizeof(double) * n) : NULL;
  int iter = 0, stop = 0;
  for (;;) {
    for (uint i = fmwr_orc_random_select(c->random_step); i < n; i += fmwr_orc_random_select(c->random_step)) {  /* :74 */
      const uint b = rowptr[i], e = rowptr[i + 1];
      double y_hat = predict_row(c, b, e, col, val, *w0, w, v, sum, sum_sqr);
      double mult = grad_mult(c, y_hat, y[i]);
      double g, delta;
      if (c->k0) {                                                                   /* :80-86 */
        g = mult;
        double old = n_w0;
        n_w0 += g * g;
        delta = (sqrt(n_w0) - sqrt(old)) / c->alpha_w;
        z_w0 += g - delta * (*w0);
      }
      if (c->k1) {                                                                   /* :88-98 */
        for (uint j = b; j < e; ++j) {
          uint id = col[j];
          g = mult * val[j];
          double old = n_w[id];
          n_w[id] += g * g;
          delta = (sqrt(n_w[id]) - sqrt(old)) / c->alpha_w;
          z_w[id] += g - delta * w[id];
        }
      }
      for (int f = 0; f < k; ++f) {                                                  /* :100-113 */
        double sf = sum[f];
        for (uint j = b; j < e; ++j) {
          size_t id = (size_t)col[j] * k + f;
          double x = val[j];
          g = mult * (sf * x - v[id] * x * x);
          double old = n_v[id];
          n_v[id] += g * g;
          delta = (sqrt(n_v[id]) - sqrt(old)) / c->alpha_v;
          z_v[id] += g - delta * v[id];
        }
      }
      /* calculate_param :158-202 (row's coordinates only) */
      *w0 = -z_w0 * c->alpha_w / (c->beta_w + sqrt(n_w0));                           /* :161 -- even when keep.w0 is false */
      for (uint j = b; j < e; ++j) {
        uint id = col[j];
        double z = z_w[id];
        if (fabs(z) <= c->l1_w) w[id] = 0.0;
        else {
          double sign = z < 0.0 ? -1.0 : 1.0;
          w[id] = -(z - sign * c->l1_w) / ((c->beta_w + sqrt(n_w[id])) / c->alpha_w + c->l2_w);
        }
      }
      for (int f = 0; f < k; ++f)
        for (uint j = b; j < e; ++j) {
          size_t id = (size_t)col[j] * k + f;
          double z = z_v[id];
          if (fabs(z) <= c->l1_v) v[id] = 0.0;
          else {
            double sign = z < 0.0 ? -1.0 : 1.0;
            v[id] = -(z - sign * c->l1_v) / ((c->beta_v + sqrt(n_v[id])) / c->alpha_v + c->l2_v);
          }
        }
      if (tracker_step(t, c, &iter, n, rowptr, col, val, y, *w0, w, v, buf)) { stop = 1; break; }
    }
    if (stop) break;
    if (n <= 1 && c->random_step == 1) break;
  }
  if (t->info) t->info->iters_done = iter;
  free(z_w); free(n_w); free(z_v); free(n_v); free(sum); free(buf);
}

/* ------------------------------------------------------------------ TDAP */
/* TDAP_Learner::init + learn + calculate_param, src/solver/TDAP_Learner.h:55-233.
 * Reproduces F6: the linear refresh reads z_w[POSITION IN ROW] (:207), not z_w[col]. */
static void learn_tdap(const fmwr_oracle_cfg* c, uint n, uint p, const uint* rowptr, const uint* col, const float* val,
                       const float* y, double* w0, double* w, double* v, tracker_t* t)
{
  const int k = c->k;
  const size_t pk = (size_t)p * (k ? k : 1);
  double u_w0 = 0, nu_w0 = 0, delta_w0 = 0, h_w0 = 0, z_w0 = 0;
  double* u_w = (double*)calloc(p, sizeof(double)); double* nu_w = (double*)calloc(p, sizeof(double));
  double* d_w = (double*)calloc(p, sizeof(double)); double* h_w = (double*)calloc(p, sizeof(double));
  double* z_w = (double*)calloc(p, sizeof(double));
  double* u_v = (double*)calloc(pk, sizeof(double)); double* nu_v = (double*)calloc(pk, sizeof(double));
  double* d_v = (double*)calloc(pk, sizeof(double)); double* h_v = (double*)calloc(pk, sizeof(double));
  double* z_v = (double*)calloc(pk, sizeof(double));
  double* sum = (double*)malloc(sizeof(double) * (2 * k + 2));
  double* sum_sqr = sum + k + 1;
  double* buf = t->step_size > 0 ? (double*)malloc(sizeof(double) * n) : NULL;
  const double egamma = exp(-c->gamma);                                              /* :83 */
  int iter = 0, stop = 0;
  for (;;) {
    for (uint i = fmwr_orc_random_select(c->random_step); i < n; i += fmwr_orc_random_select(c->random_step)) {  /* :90 */
      const uint b = rowptr[i], e = rowptr[i + 1];
      double y_hat = predict_row(c, b, e, col, val, *w0, w, v, sum, sum_sqr);
      double mult = grad_mult(c, y_hat, y[i]);
      double g, sigma;
      if (c->k0) {                                                                   /* :96-105 */
        g = mult;
        double old = u_w0;
        u_w0 += g * g; nu_w0 += g;
        sigma = (sqrt(u_w0) - sqrt(old)) / c->alpha_w;
        delta_w0 = egamma * (delta_w0 + sigma);
        h_w0 = egamma * (h_w0 + sigma * (*w0));
        z_w0 = nu_w0 - h_w0;
      }
      if (c->k1) {                                                                   /* :107-123 */
        for (uint j = b; j < e; ++j) {
          uint id = col[j];
          g = mult * val[j];
          double old = u_w[id];
          u_w[id] += g * g; nu_w[id] += g;
          sigma = (sqrt(u_w[id]) - sqrt(old)) / c->alpha_w;
          d_w[id] = egamma * (d_w[id] + sigma);
          h_w[id] = egamma * (h_w[id] + sigma * w[id]);
          z_w[id] = nu_w[id] - h_w[id];
        }
      }
      for (int f = 0; f < k; ++f) {                                                  /* :125-143 */
        double sf = sum[f];
        for (uint j = b; j < e; ++j) {
          size_t id = (size_t)col[j] * k + f;
          double x = val[j];
          g = mult * (sf * x - v[id] * x * x);
          double old = u_v[id];
          u_v[id] += g * g; nu_v[id] += g;
          sigma = (sqrt(u_v[id]) - sqrt(old)) / c->alpha_v;
          d_v[id] = egamma * (d_v[id] + sigma);
          h_v[id] = egamma * (h_v[id] + sigma * v[id]);
          z_v[id] = nu_v[id] - h_v[id];
        }
      }
      /* calculate_param :189-233 */
      *w0 = -z_w0 / delta_w0;                                                        /* :192 (0/0 = NaN when keep.w0 is false) */
      for (uint j = b; j < e; ++j) {
        uint id = col[j];
        double z = z_w[j - b];                                                       /* :207, F6: position, not column */
        if (fabs(z) <= c->l1_w) w[id] = 0.0;
        else {
          double sign = z < 0.0 ? -1.0 : 1.0;
          w[id] = -(z - sign * c->l1_w) / (d_w[id] + c->l2_w);
        }
      }
      for (int f = 0; f < k; ++f)
        for (uint j = b; j < e; ++j) {
          size_t id = (size_t)col[j] * k + f;
          double z = z_v[id];
          if (fabs(z) <= c->l1_v) v[id] = 0.0;
          else {
            double sign = z < 0.0 ? -1.0 : 1.0;
            v[id] = -(z - sign * c->l1_v) / (d_v[id] + c->l2_v);
          }
        }
      if (tracker_step(t, c, &iter, n, rowptr, col, val, y, *w0, w, v, buf)) { stop = 1; break; }
    }
    if (stop) break;
    if (n <= 1 && c->random_step == 1) break;
  }
  if (t->info) t->info->iters_done = iter;
  free(u_w); free(nu_w); free(d_w); free(h_w); free(z_w);
  free(u_v); free(nu_v); free(d_v); free(h_v); free(z_v); free(sum); free(buf);
}

/* ------------------------------------------------------------------ ALS / MCMC */
static int bad(double x) { return isnan(x) || isinf(x); }   /* CHECK_PARAM predicate, src/util/Macros.h:36-41 */

/* MCMC_ALS_Learner::init/learn/update_*, src/solver/MCMC_ALS_Learner.h:59-562, nthreads == 1
 * (exact Gauss-Seidel).  One attribute group (src/FM.cpp:75).  enable_v restores :151-155 (F1). */
static void learn_mcmc_als(const fmwr_oracle_cfg* c, uint n, uint p, uint nnz, const uint* rowptr, const uint* col,
                           const float* val, const float* y, double* w0p, double* w, double* v, tracker_t* t)
{
  const int k = c->k;
  const int do_sample = (c->solver == ORC_MCMC), do_multilevel = (c->solver == ORC_MCMC);   /* :567-587 */
  /* init() overwrites whatever FM.cpp set (F2) :64-71 */
  const double alpha_0 = 1.0, gamma_0 = 1.0, beta_0 = 1.0, mu_0 = 0.0, w0_mean_0 = 0.0;
  double alpha = 1.0;
  double w_mu = 0.0, w_lambda = 0.0;
  double* v_mu = (double*)calloc(k ? k : 1, sizeof(double));
  double* v_lambda = (double*)calloc(k ? k : 1, sizeof(double));
  /* CSC twin */
  uint* tp = (uint*)malloc(sizeof(uint) * (p + 1));
  uint* ti = (uint*)malloc(sizeof(uint) * (nnz ? nnz : 1));
  float* tv = (float*)malloc(sizeof(float) * (nnz ? nnz : 1));
  fmwr_orc_transpose(n, p, nnz, rowptr, col, val, 0, tp, ti, tv);
  double* err = (double*)malloc(sizeof(double) * (n ? n : 1));
  double* e2 = (double*)malloc(sizeof(double) * (n ? n : 1));     /* update_v's local copies :276-277 */
  double* vq = (double*)malloc(sizeof(double) * (n ? n : 1));
  double* yh = (double*)malloc(sizeof(double) * (n ? n : 1));
  int ii = -1;
  int sweep;
  for (sweep = 0; sweep < c->max_iter; ++sweep) {                                    /* :98 */
    predict_batch(c, n, rowptr, col, val, *w0p, w, v, err, 1);                       /* :100 */
    if (t->step_size > 0) {                                                          /* :101-124 */
      ii++;
      if (ii == t->step_size) ii = 0;
      if (ii == 0 || sweep == c->max_iter - 1) {
        if (c->task == ORC_REGRESSION) {
          for (uint i = 0; i < n; ++i) { yh[i] = err[i]; if (yh[i] < c->min_target) yh[i] = c->min_target; else if (yh[i] > c->max_target) yh[i] = c->max_target; }
        } else {
          for (uint i = 0; i < n; ++i) yh[i] = fmwr_orc_pnorm(err[i]);
        }
        tracker_record(t, sweep, fmwr_orc_evaluate(c->task, c->metric, n, yh, y));
      }
    }
    /* calculate_error :520-562 */
    if (c->task == ORC_REGRESSION) {
      for (uint i = 0; i < n; ++i) err[i] -= y[i];
    } else if (do_sample) {
      for (uint i = 0; i < n; ++i) {
        double e = err[i];
        if (y[i] >= 0.0) err[i] -= fmwr_orc_trnorm_left(e, 0.0, 1.0);               /* :536 -- N(0,1) truncated at y_hat, as shipped */
        else err[i] -= fmwr_orc_trnorm_right(e, 0.0, 1.0);
      }
    } else {
      for (uint i = 0; i < n; ++i) {
        double e = err[i];
        if (y[i] >= 0.0) err[i] = -fmwr_orc_dpnorm(-e);
        else err[i] = fmwr_orc_dpnorm(e);
      }
    }
    /* update_alpha :360-380 */
    if (!do_multilevel) alpha = alpha_0;
    else {
      double alpha_n = alpha_0 + n, gamma_n = gamma_0;
      for (uint i = 0; i < n; ++i) gamma_n += err[i] * err[i];
      double a = draw_gamma(alpha_n / 2.0, 2.0 / gamma_n);
      if (!bad(a)) alpha = a;
    }
    /* update_w0 :162-188 */
    if (c->k0) {
      double w0 = *w0p, e = 0;
      for (uint i = 0; i < n; ++i) e += err[i] - w0;
      double var = 1.0 / (c->l2_w0 + alpha * n);
      double mean = -(alpha * e - w0_mean_0 * c->l2_w0) * var;
      double nw = do_sample ? draw_norm(mean, sqrt(var)) : mean;
      if (bad(nw)) nw = w0;
      *w0p = nw;
      double diff = w0 - nw;
      for (uint i = 0; i < n; ++i) err[i] -= diff;
    }
    if (c->k1) {
      /* update_w_lambda :415-445 */
      if (do_multilevel) {
        double g = 0.0;
        for (uint i = 0; i < p; ++i) g += (w[i] - w_mu) * (w[i] - w_mu);
        g += beta_0 * (w_mu - mu_0) * (w_mu - mu_0) + gamma_0;
        double la = alpha_0 + p + 1;
        double nl = do_sample ? draw_gamma(la / 2.0, 2.0 / g) : la / g;
        if (!bad(nl)) w_lambda = nl;
      }
      /* update_w_mu :383-412 */
      if (!do_multilevel) w_mu = mu_0;
      else {
        double m = 0.0;
        for (uint i = 0; i < p; ++i) m += w[i];
        m = (m + beta_0 * mu_0) / (p + beta_0);
        double var = 1.0 / ((p + beta_0) * w_lambda);
        double nm = do_sample ? draw_norm(m, sqrt(var)) : m;
        if (!bad(nm)) w_mu = nm;
      }
      /* update_w :190-270, single thread => Gauss-Seidel in feature order */
      for (uint i = 0; i < p; ++i) {
        double mean = 0.0, var = 0.0, old = w[i], nw;
        int upd = 1;
        for (uint j = tp[i]; j < tp[i + 1]; ++j) {
          double x = tv[j];
          mean += err[ti[j]] * x - old * x * x;
          var += x * x;
        }
        var = 1.0 / (w_lambda + alpha * var);
        mean = -var * (alpha * mean - w_mu * w_lambda);
        if (bad(var)) nw = 0.0;
        else nw = do_sample ? draw_norm(mean, var) : mean;                            /* :239, F7: variance passed as s.d. */
        if (bad(nw)) { nw = old; upd = 0; }
        w[i] = nw;
        if (upd) {
          double d = old - nw;
          for (uint j = tp[i]; j < tp[i + 1]; ++j) err[ti[j]] -= tv[j] * d;
        }
      }
    }
    if (c->enable_v && k > 0) {
      /* update_v_lambda :486-517 */
      if (do_multilevel) {
        for (int f = 0; f < k; ++f) {
          double g = 0.0;
          for (uint i = 0; i < p; ++i) { double d = v[(size_t)i * k + f] - v_mu[f]; g += d * d; }
          g += beta_0 * (v_mu[f] - mu_0) * (v_mu[f] - mu_0) + gamma_0;
          double la = alpha_0 + p + 1;
          double nl = do_sample ? draw_gamma(la / 2.0, 2.0 / g) : la / g;
          if (!bad(nl)) v_lambda[f] = nl;
        }
      }
      /* update_v_mu :448-483 -- F7: sums v(f, attr_group[i]) == v(f, 0) p times */
      if (!do_multilevel) { for (int f = 0; f < k; ++f) v_mu[f] = mu_0; }
      else {
        for (int f = 0; f < k; ++f) {
          double m = 0.0;
          for (uint i = 0; i < p; ++i) m += v[(size_t)0 * k + f];
          m = (m + beta_0 * mu_0) / (p + beta_0);
          double var = 1.0 / ((p + beta_0) * v_lambda[f]);
          double nm = do_sample ? draw_norm(m, sqrt(var)) : m;
          if (!bad(nm)) v_mu[f] = nm;
        }
      }
      /* update_v :272-354 on local copies of e and q */
      memcpy(e2, err, sizeof(double) * n);
      for (int f = 0; f < k; ++f) {
        for (uint r = 0; r < n; ++r) vq[r] = 0.0;
        for (uint i = 0; i < p; ++i) {                                               /* :289-299 */
          double vv = v[(size_t)i * k + f];
          for (uint j = tp[i]; j < tp[i + 1]; ++j) vq[ti[j]] += tv[j] * vv;
        }
        for (uint i = 0; i < p; ++i) {                                               /* :303-351 */
          double mean = 0, var = 0, old = v[(size_t)i * k + f], nv;
          int upd = 1;
          for (uint m = tp[i]; m < tp[i + 1]; ++m) {
            float x = tv[m];
            uint r = ti[m];
            double h = x * vq[r] - x * x * old;
            mean += h * e2[r];
            var += h * h;
          }
          mean -= old * var;
          var = 1.0 / (v_lambda[f] + alpha * var);
          mean = -var * (alpha * mean - v_mu[f] * v_lambda[f]);
          if (bad(var)) nv = 0.0;
          else nv = do_sample ? draw_norm(mean, sqrt(var)) : mean;
          if (bad(nv)) { nv = old; upd = 0; }
          v[(size_t)i * k + f] = nv;
          double d = old - nv;
          if (upd) {
            for (uint m = tp[i]; m < tp[i + 1]; ++m) {
              float x = tv[m];
              uint r = ti[m];
              double h = x * vq[r] - x * x * old;
              vq[r] -= x * d;
              e2[r] -= h * d;
            }
          }
        }
      }
    }
  }
  if (t->info) t->info->iters_done = sweep;
  free(v_mu); free(v_lambda); free(tp); free(ti); free(tv); free(err); free(e2); free(vq); free(yh);
}

/* ------------------------------------------------------------------ entry */
int fmwr_orc_train(const fmwr_oracle_cfg* c, uint n, uint p, uint nnz, const uint* rowptr, const uint* col,
                   const float* val, const float* y, int unused,
                   double* w0, double* w, double* v,
                   int max_rec, double* eval_train, int* rec_index, fmwr_oracle_trace_info* info)
{
  (void)unused;
  build_tables();
  tracker_t t;
  tracker_init(&t, c, max_rec, eval_train, rec_index, info);
  if (info) { info->n_rec = 0; info->convergent = 0; info->iters_done = 0; }
  switch (c->solver) {
    case ORC_SGD:  learn_sgd(c, n, p, rowptr, col, val, y, w0, w, v, &t); break;
    case ORC_FTRL: learn_ftrl(c, n, p, rowptr, col, val, y, w0, w, v, &t); break;
    case ORC_TDAP: learn_tdap(c, n, p, rowptr, col, val, y, w0, w, v, &t); break;
    case ORC_ALS: case ORC_MCMC: learn_mcmc_als(c, n, p, nnz, rowptr, col, val, y, w0, w, v, &t); break;
    default: snprintf(g_err, sizeof g_err, "Unknown solver..."); return 1;
  }
  if (info) { info->n_rec = t.n_rec; info->convergent = t.convergent; }
  return 0;
}

/* SMatrix::scales, src/util/Smatrix.h:98-131 (z-score of the NON-ZEROS of the listed columns, in place) */
int fmwr_orc_scales(uint n, uint p, uint nnz, const uint* rowptr, const uint* col, float* val,
                    const int* norm_cols, int n_norm, double* mean, double* sd)
{
  (void)rowptr;
  for (uint c = 0; c < p; ++c) { mean[c] = 0.0; sd[c] = 0.0; }
  for (uint e = 0; e < nnz; ++e) { double x = val[e]; mean[col[e]] += x; sd[col[e]] += x * x; }
  double mult_dim = (double)n * ((double)n - 1);
  int i = 0;
  for (uint c = 0; c < p; ++c) {
    if (i < n_norm && c == (uint)norm_cols[i]) {
      sd[c] = sqrt(sd[c] / (n - 1) - mean[c] * mean[c] / mult_dim);
      mean[c] /= n;
      i++;
    } else { sd[c] = 1.0; mean[c] = 0.0; }
  }
  for (uint e = 0; e < nnz; ++e) {
    val[e] -= mean[col[e]];             /* float -= double, then float /= double: as the reference does on float storage */
    val[e] /= (sd[col[e]] + 1e-30);
  }
  return 0;
}

int fmwr_orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
