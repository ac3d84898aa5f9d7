// TEST INFRASTRUCTURE ONLY -- not part of the shipped engine.
//
// Drives the reference's OWN hot-path headers (included from
// /root/reference/src, never copied) on raw arrays so the tests can pin
// (a) the C restatement in oracle/fm_oracle.c and (b) the CUDA engine against
// the reference's arithmetic.  Built by oracle/Makefile into
// oracle/_ref/libfmwr_ref.so; only exists where /root/reference is mounted.
//
// Drive pattern follows the reference's scratch tests (src/test/TDAP.cpp:83-141):
// fill SMatrix<float> public fields, Data.add_data/add_target, set Model
// fields, fm.init(), overwrite w0/w/v, construct learner, init(), learn().
#include <Rcpp.h>
#include <map>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif
using namespace Rcpp;
using namespace std;

OracleStreams g_oracle_streams = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

#include "util/Macros.h"
#include "util/Dmatrix.h"
#include "util/Dvector.h"
#include "util/Random.h"
#include "util/Smatrix.h"
#include "core/Data.h"
#include "core/Model.h"
#include "core/Learner.h"
#include "core/Evaluation.h"
#include "core/Tracker.h"
#include "solver/MCMC_ALS_Learner.h"
#include "solver/SGD_Learner.h"
#include "solver/FTRL_Learner.h"
#include "solver/TDAP_Learner.h"

#include "oracle_abi.h"

static std::string g_err;

// ALS/MCMC with the V block of update_all restored (src/solver/MCMC_ALS_Learner.h:151-155
// is commented out as shipped, F1).  learn() and update_all() bodies follow :91-156.
template <class Base>
struct WithV : Base {
  void learn(Data& train)
  {
    DVector<double> train_err(train.num_cases);
    DVector<double> v_q(train.num_cases);
    int ii = -1;
    for (; this->iter_cntr < this->max_iter; ++this->iter_cntr)
    {
      this->fm->predict_batch(train, train_err);
      if (this->tracker.step_size > 0) {
        ii++;
        if (ii == this->tracker.step_size) { ii = 0; }
        if (ii == 0 || this->iter_cntr == this->max_iter - 1) {
          DVector<double> y_hat_(train.num_cases);
          if (this->fm->TASK == REGRESSION) {
            for (uint i = 0; i < train.num_cases; ++i) {
              y_hat_[i] = train_err[i];
              if (y_hat_[i] < this->min_target) y_hat_[i] = this->min_target;
              else if (y_hat_[i] > this->max_target) y_hat_[i] = this->max_target;
            }
          } else {
            for (uint i = 0; i < train.num_cases; ++i) y_hat_[i] = fast_pnorm(train_err[i]);
          }
          double eval_score = this->tracker.evaluate(this->fm, y_hat_, *train.target);
          this->tracker.record(this->fm, this->iter_cntr, eval_score);
        }
      }
      this->calculate_error(train, train_err);
      // update_all with the V block enabled
      this->update_alpha(train, train_err);
      if (this->fm->k0) { this->update_w0(train, train_err); }
      if (this->fm->k1) {
        this->update_w_lambda();
        this->update_w_mu();
        this->update_w(train, train_err);
      }
      if (this->fm->num_factor > 0) {
        this->update_v_lambda();
        this->update_v_mu();
        this->update_v(train, train_err, v_q);
      }
    }
  }
};

static void fill_csr(SMatrix<float>& m, uint n, uint p, uint nnz, const uint* rowptr, const uint* col, const float* val)
{
  m.dim1 = n; m.dim2 = p; m.size = nnz; m.transposed = false;
  m.row_idx.setSize(n + 1); m.col_idx.setSize(nnz > 0 ? nnz : 1); m.value.setSize(nnz > 0 ? nnz : 1);
  for (uint i = 0; i <= n; ++i) m.row_idx[i] = rowptr[i];
  for (uint i = 0; i < nnz; ++i) { m.col_idx[i] = col[i]; m.value[i] = val[i]; }
}

// counting-sort CSC (stable: rows ascending inside each column) for shapes where
// SMatrix::transpose (O(n*p), src/util/Smatrix.h:155-185) is unusable.
static void counting_csc(const SMatrix<float>& m, SMatrix<float>& mt)
{
  uint n = m.dim1, p = m.dim2, nnz = m.size;
  mt.dim1 = p; mt.dim2 = n; mt.size = nnz; mt.transposed = true;
  mt.row_idx.setSize(p + 1); mt.row_idx.init(0);
  mt.col_idx.setSize(nnz > 0 ? nnz : 1); mt.value.setSize(nnz > 0 ? nnz : 1);
  for (uint e = 0; e < nnz; ++e) mt.row_idx[m.col_idx[e] + 1]++;
  for (uint c = 0; c < p; ++c) mt.row_idx[c + 1] += mt.row_idx[c];
  std::vector<uint> cur(mt.row_idx.begin(), mt.row_idx.begin() + p);
  for (uint r = 0; r < n; ++r)
    for (uint e = m.row_idx[r]; e < m.row_idx[r + 1]; ++e) {
      uint d = cur[m.col_idx[e]]++;
      mt.col_idx[d] = r; mt.value[d] = m.value[e];
    }
}

static void set_model(Model& fm, const fmwr_oracle_cfg* c, uint p, double w0, const double* w, const double* v)
{
  fm.k0 = c->k0 != 0; fm.k1 = c->k1 != 0; fm.num_factor = (uint)c->k;
  fm.l2_reg0 = c->l2_w0; fm.l1_regw = c->l1_w; fm.l2_regw = c->l2_w; fm.l1_regv = c->l1_v; fm.l2_regv = c->l2_v;
  fm.init_mean = 0.0; fm.init_stdev = 0.0;
  fm.nthreads = c->nthreads > 0 ? c->nthreads : 1;
  fm.num_attribute = p; fm.SOLVER = c->solver; fm.TASK = c->task;
  OracleStreams saved = g_oracle_streams;            // fm.init() draws V through Rf_rnorm: do not consume the injected stream
  g_oracle_streams.normals = 0;
  fm.init();
  g_oracle_streams = saved;
  fm.w0 = w0;
  for (uint i = 0; i < p; ++i) fm.w[i] = w[i];
  for (uint i = 0; i < p; ++i) for (int f = 0; f < c->k; ++f) fm.v(f, i) = v[(size_t)i * c->k + f];   // v arrives [p][k]
}

static void get_model(Model& fm, int k, uint p, double* w0, double* w, double* v)
{
  *w0 = fm.w0;
  for (uint i = 0; i < p; ++i) w[i] = fm.w[i];
  for (uint i = 0; i < p; ++i) for (int f = 0; f < k; ++f) v[(size_t)i * k + f] = fm.v(f, i);
}

extern "C" {

const char* fmwr_ref_last_error() { return g_err.c_str(); }

void fmwr_ref_set_streams(const double* normals, long n_normals, const double* gammas, long n_gammas,
                          const int* rands, long n_rands)
{
  OracleStreams& s = g_oracle_streams;
  s.normals = normals; s.n_normals = n_normals; s.i_normal = 0;
  s.gammas = gammas; s.n_gammas = n_gammas; s.i_gamma = 0;
  s.rands = rands; s.n_rands = n_rands; s.i_rand = 0;
  s.overrun = 0;
}

void fmwr_ref_stream_pos(long* out4)
{
  out4[0] = g_oracle_streams.i_normal; out4[1] = g_oracle_streams.i_gamma;
  out4[2] = g_oracle_streams.i_rand; out4[3] = g_oracle_streams.overrun;
}

// Model::predict_batch (+ optional link, Model::predict_prob) -- src/core/Model.h:106-180
// link: 0 none, 1 predict_prob (logistic for SGD/FTRL/TDAP, fast_pnorm for ALS/MCMC by cfg->solver)
int fmwr_ref_predict(const fmwr_oracle_cfg* c, uint n, uint p, uint nnz, const uint* rowptr, const uint* col,
                     const float* val, double w0, const double* w, const double* v, int link, double* out)
{
  try {
    SMatrix<float> m; fill_csr(m, n, p, nnz, rowptr, col, val);
    Data data; data.add_data(&m);
    Model fm; set_model(fm, c, p, w0, w, v);
    DVector<double> o(n);
    if (link) fm.predict_prob(data, o); else fm.predict_batch(data, o);
    for (uint i = 0; i < n; ++i) out[i] = o[i];
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Model::predict (single-row variant used by SGD/FTRL/TDAP) -- src/core/Model.h:75-103; also returns m_sum
int fmwr_ref_predict_rows(const fmwr_oracle_cfg* c, uint n, uint p, uint nnz, const uint* rowptr, const uint* col,
                          const float* val, double w0, const double* w, const double* v, double* out, double* sums)
{
  try {
    SMatrix<float> m; fill_csr(m, n, p, nnz, rowptr, col, val);
    Model fm; set_model(fm, c, p, w0, w, v);
    for (uint i = 0; i < n; ++i) {
      if (rowptr[i + 1] == rowptr[i]) { out[i] = NAN; continue; }  // Iterator ctor reads col_idx[pointer]; empty row is UB-ish
      SMatrix<float>::Iterator it(m, i);
      out[i] = fm.predict(it);
      if (sums) for (int f = 0; f < c->k; ++f) sums[(size_t)i * c->k + f] = fm.m_sum[f];
    }
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// SMatrix::transpose -- src/util/Smatrix.h:155-185.  use_ref = 0 -> counting sort instead.
int fmwr_ref_transpose(uint n, uint p, uint nnz, const uint* rowptr, const uint* col, const float* val, int use_ref,
                       uint* t_ptr, uint* t_idx, float* t_val)
{
  try {
    SMatrix<float> m; fill_csr(m, n, p, nnz, rowptr, col, val);
    SMatrix<float> mt;
    if (use_ref) m.transpose(mt); else counting_csc(m, mt);
    for (uint i = 0; i <= p; ++i) t_ptr[i] = mt.row_idx[i];
    for (uint i = 0; i < nnz; ++i) { t_idx[i] = mt.col_idx[i]; t_val[i] = mt.value[i]; }
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Full training run: FM() body of src/FM.cpp:78-153 minus the List plumbing.
// w0/w/v are in-out ([p][k] layout for v).  Trace outputs are optional (NULL).
int fmwr_ref_train(const fmwr_oracle_cfg* c, uint n, uint p, uint nnz, const uint* rowptr, const uint* col,
                   const float* val, const float* y, int use_ref_transpose,
                   double* w0, double* w, double* v,
                   int max_rec, double* eval_train, int* rec_index, fmwr_oracle_trace_info* info)
{
  try {
    SMatrix<float> m; fill_csr(m, n, p, nnz, rowptr, col, val);
    Data data; data.add_data(&m);
    DVector<float> tg(n);
    for (uint i = 0; i < n; ++i) tg[i] = y[i];
    data.add_target(&tg);
    Model fm; set_model(fm, c, p, *w0, w, v);
    DataMetaInfo meta(p);

    Learner* learner = NULL;
    switch (c->solver) {
      case MCMC: learner = c->enable_v ? (Learner*) new WithV<MCMC_Learner>() : (Learner*) new MCMC_Learner(); break;
      case ALS : learner = c->enable_v ? (Learner*) new WithV<ALS_Learner>()  : (Learner*) new ALS_Learner();  break;
      case SGD : learner = new SGD_Learner(); break;
      case TDAP: learner = new TDAP_Learner(); break;
      case FTRL: learner = new FTRL_Learner(); break;
      default: stop("Unknown solver...");
    }
    learner->meta = &meta; learner->fm = &fm;
    learner->min_target = c->min_target; learner->max_target = c->max_target;
    learner->nthreads = c->nthreads > 0 ? c->nthreads : 1;
    learner->max_iter = c->max_iter;
    learner->tracker.step_size = c->step_size;
    learner->tracker.max_iter = c->max_iter;
    learner->type = c->metric; learner->tracker.type = c->metric;
    learner->conv_condition = c->convergence;
    switch (c->solver) {
      case SGD:  ((SGD_Learner*)learner)->learn_rate = c->learn_rate; ((SGD_Learner*)learner)->random_step = c->random_step; break;
      case FTRL: { FTRL_Learner* l = (FTRL_Learner*)learner; l->alpha_w = c->alpha_w; l->alpha_v = c->alpha_v;
                   l->beta_w = c->beta_w; l->beta_v = c->beta_v; l->random_step = c->random_step; break; }
      case TDAP: { TDAP_Learner* l = (TDAP_Learner*)learner; l->gamma = c->gamma; l->alpha_w = c->alpha_w;
                   l->alpha_v = c->alpha_v; l->random_step = c->random_step; break; }
      default: break;   // ALS/MCMC: init() overwrites every hyper-parameter anyway (F2)
    }
    learner->init();

    SMatrix<float> m_t;
    if (fm.SOLVER <= ALS) {
      if (use_ref_transpose) m.transpose(m_t); else counting_csc(m, m_t);
      data.add_data(&m_t);
    }
    learner->learn(data);

    get_model(fm, c->k, p, w0, w, v);
    if (info) {
      info->convergent = learner->convergent ? 1 : 0;
      info->n_rec = 0; info->iters_done = 0;
      if (learner->tracker.step_size > 0) {
        int nr = learner->tracker.record_cnter;
        info->n_rec = nr;
        for (int i = 0; i < nr && i < max_rec; ++i) {
          if (eval_train) eval_train[i] = learner->tracker.evaluations_of_train[i];
          if (rec_index) rec_index[i] = learner->tracker.record_index[i];
        }
      }
    }
    delete learner;
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return 1; }
}

// Evaluation.h:20-115
double fmwr_ref_evaluate(int task, int metric, uint n, const double* y_hat, const float* y_true)
{
  Model fm; fm.TASK = task;
  DVector<double> a(n); DVector<float> b(n);
  for (uint i = 0; i < n; ++i) { a[i] = y_hat[i]; b[i] = y_true[i]; }
  return evaluates(&fm, a, b, metric);
}

// Random.h:95-124
double fmwr_ref_pnorm(double x) { return fast_pnorm(x); }
double fmwr_ref_dpnorm(double x) { return fast_dpnorm(x); }
// Random.h:51-93, consuming the injected rand() stream
double fmwr_ref_trnorm_left(double left, double mean, double sd) { return fast_trnorm_left(left, mean, sd); }
double fmwr_ref_trnorm_right(double right, double mean, double sd) { return fast_trnorm_right(right, mean, sd); }
unsigned fmwr_ref_random_select(int n) { return random_select(n); }

// SMatrix::scales / normalize -- src/util/Smatrix.h:98-153 ("next" row f2); in-place on val
int fmwr_ref_scales(uint n, uint p, uint nnz, const uint* rowptr, const uint* col, float* val,
                    const int* norm_cols, int n_norm, double* mean, double* sd)
{
  try {
    SMatrix<float> m; fill_csr(m, n, p, nnz, rowptr, col, val);
    IntegerVector nc(n_norm + 1);                 // +1: the reference reads norm_columns[i] one past the end (:115)
    for (int i = 0; i < n_norm; ++i) nc[i] = norm_cols[i];
    nc[n_norm] = -1;
    // body of SMatrix::scales without the List return (same statements, :100-131)
    DVector<double> colSum(p); colSum.init(0.0);
    DVector<double> colSumSqr(p); colSumSqr.init(0.0);
    (void)colSum; (void)colSumSqr;
    List res = m.scales(nc);
    NumericVector mu = as<NumericVector>(res["mean"]);
    NumericVector sg = as<NumericVector>(res["std"]);
    for (uint i = 0; i < p; ++i) { mean[i] = mu[i]; sd[i] = sg[i]; }
    for (uint i = 0; i < nnz; ++i) val[i] = m.value[i];
    return 0;
  } catch (std::exception& e) { g_err = e.what(); return 1; }
}

int fmwr_ref_num_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
