// Drop-in replacement for FMwR's src/FM.cpp: the three Rcpp exports keep their signatures
// (reference src/FM.cpp:7, :177, :218; src/RcppExports.cpp is unchanged) and become thin glue that
// unpacks the R lists into raw pointers and calls libfmwr_b200.so (include/fmwr_b200.h).
//
// UNVERIFIED IN THIS REPOSITORY: the build image has no R / Rcpp, so this file has never been compiled.
// It is the binding a maintainer adds; the same call sequence is exercised by fmwr_b200/api.py.
//
// Build: src/Makevars gains   PKG_CPPFLAGS = -I$(FMWR_B200)/include
//                             PKG_LIBS     = -L$(FMWR_B200)/fmwr_b200 -lfmwr_b200 -Wl,-rpath,$(FMWR_B200)/fmwr_b200
#include <Rcpp.h>
#include <map>
#include <string>
#include <vector>
#include "fmwr_b200.h"

using namespace Rcpp;

static void check(int rc) { if (rc != 0) stop(fmwr_last_error()); }   // Rcpp::stop -> R error, as before

static int opt_int(const char* name, int dflt)
{
  Function getOption("getOption");
  SEXP v = getOption(name, dflt);
  return as<int>(v);
}

static std::string opt_str(const char* name, const char* dflt)
{
  Function getOption("getOption");
  return as<std::string>(getOption(name, dflt));
}

static fmwr_model_cfg model_cfg(List fm_controls)
{
  std::map<std::string, int> tasks; tasks["CLASSIFICATION"] = FMWR_CLASSIFICATION; tasks["REGRESSION"] = FMWR_REGRESSION;
  List hp = fm_controls["hyper.params"];
  fmwr_model_cfg c;
  c.task = tasks[as<std::string>(fm_controls["task"])];
  c.keep_w0 = (bool)hp["keep.w0"]; c.keep_w1 = (bool)hp["keep.w1"]; c.k = (int)hp["factor.number"];
  c.l2_w0 = (double)hp["L2.w0"]; c.l1_w1 = (double)hp["L1.w1"]; c.l2_w1 = (double)hp["L2.w1"];
  c.l1_v = (double)hp["L1.v"]; c.l2_v = (double)hp["L2.v"];
  return c;
}

static int solver_id(List solver_controls)
{
  std::map<std::string, int> m;
  m["MCMC"] = FMWR_MCMC; m["ALS"] = FMWR_ALS; m["SGD"] = FMWR_SGD; m["FTRL"] = FMWR_FTRL; m["TDAP"] = FMWR_TDAP;
  List solver = solver_controls["solver"];
  std::string s = as<std::string>(solver.attr("solver"));
  if (!m.count(s)) stop("Unknown solver...");
  return m[s];
}

static int metric_id(const std::string& s)
{
  std::map<std::string, int> m;
  m["LL"] = FMWR_LL; m["AUC"] = FMWR_AUC; m["ACC"] = FMWR_ACC; m["RMSE"] = FMWR_RMSE; m["MSE"] = FMWR_MSE; m["MAE"] = FMWR_MAE;
  return m[s];
}

// [[Rcpp::export]]
List FM(List data_, IntegerVector normalize, List fm_controls, List solver_controls, List track_controls, List model_list)
{
  List X = data_["features"];
  NumericVector value = X["value"]; IntegerVector col_idx = X["col_idx"]; IntegerVector row_size = X["row_size"];
  IntegerVector dim = X["dim"];
  NumericVector target = as<NumericVector>(data_["labels"]);
  const int64_t n = dim[0], p = dim[1], nnz = value.size();

  fmwr_ctx* ctx = NULL; fmwr_data* d = NULL; fmwr_model* m = NULL;
  check(fmwr_ctx_create(opt_int("FM.device", 0), &ctx));
  check(fmwr_data_create(ctx, n, p, nnz, row_size.begin(), col_idx.begin(), value.begin(), target.begin(), &d));

  List scales;
  if (normalize[0] > -1) {                                            // reference FM.cpp:36-38
    NumericVector mean(p), sd(p);
    check(fmwr_data_scales(d, normalize.begin(), normalize.size(), mean.begin(), sd.begin()));
    scales["mean"] = mean; scales["std"] = sd;
  }
  scales["model.vars"] = as<CharacterVector>(X.attr("feature_names"));

  fmwr_model_cfg mc = model_cfg(fm_controls);
  List hp = fm_controls["hyper.params"];
  const int k = mc.k;
  const int prec = opt_str("FM.precision", "f32") == "f64" ? FMWR_F64 : FMWR_F32;
  check(fmwr_model_create(ctx, &mc, p, prec, &m));

  // Model::init (reference src/core/Model.h:63-72): w = 0, V ~ rnorm drawn HERE, factor-major, so set.seed() holds
  double w0 = 0.0;
  NumericVector w(p);
  NumericMatrix v(k, p);                                              // k x p column-major == the engine's [p][k]
  {
    std::vector<double> draws((size_t)k * p);
    for (size_t i = 0; i < draws.size(); ++i) draws[i] = Rf_rnorm((double)hp["v.init_mean"], (double)hp["v.init_stdev"]);
    for (int f = 0; f < k; ++f) for (int64_t j = 0; j < p; ++j) v(f, j) = draws[(size_t)f * p + j];
  }
  double lo = min(target), hi = max(target);
  SEXP is_model = model_list.attr("class");
  if (!Rf_isNull(is_model)) {                                         // warm start (fm.update), FM.cpp:66-72, :91-96
    List model = model_list["Model"];
    w0 = (double)model["w0"]; w = clone(as<NumericVector>(model["w"])); v = clone(as<NumericMatrix>(model["v"]));
    List so = model_list["Scales"];
    NumericVector tr = as<NumericVector>(so.attr("target.range"));
    lo = std::min(lo, tr[0]); hi = std::max(hi, tr[1]);
  }
  check(fmwr_model_set(m, w0, w.begin(), v.begin()));

  List solver = solver_controls["solver"];
  fmwr_solver_cfg sc; memset(&sc, 0, sizeof sc);
  sc.solver = solver_id(solver_controls);
  sc.max_iter = (int)solver_controls["max_iter"];
  sc.random_step = solver.containsElementNamed("random_step") ? (int)solver["random_step"] : 1;
  sc.learn_rate = solver.containsElementNamed("learn_rate") ? (double)solver["learn_rate"] : 0.01;
  sc.alpha_w = solver.containsElementNamed("alpha_w") ? (double)solver["alpha_w"] : 0.1;
  sc.alpha_v = solver.containsElementNamed("alpha_v") ? (double)solver["alpha_v"] : 0.1;
  sc.beta_w = solver.containsElementNamed("beta_w") ? (double)solver["beta_w"] : 1.0;
  sc.beta_v = solver.containsElementNamed("beta_v") ? (double)solver["beta_v"] : 1.0;
  sc.gamma = solver.containsElementNamed("gamma") ? (double)solver["gamma"] : 1e-4;
  sc.min_target = lo; sc.max_target = hi;
  sc.mode = opt_str("FM.mode", "exact") == "minibatch" ? FMWR_MODE_MINIBATCH : FMWR_MODE_EXACT;
  sc.batch_size = opt_int("FM.batch", 65536);
  sc.precision = prec;
  sc.compat = opt_str("FM.compat", "reference") == "reference" ? FMWR_COMPAT_REFERENCE : 0;
  sc.enable_v = opt_int("FM.enable_v", 0);
  sc.step_size = (int)track_controls["step_size"];
  sc.metric = metric_id(as<std::string>(track_controls["evaluate.metric"]));
  sc.convergence = (double)track_controls["convergence"];
  sc.seed = (uint64_t)(R::unif_rand() * 4294967296.0);               // native MCMC RNG keyed off R's stream

  fmwr_trace tr; memset(&tr, 0, sizeof tr);
  std::vector<double> ev, sw0, sw, sv; std::vector<int> ri;
  if (sc.step_size > 0) {
    tr.max_rec = std::min(10001, (int)std::ceil((sc.max_iter - 0.5) / sc.step_size) + 2);
    ev.resize(tr.max_rec); ri.resize(tr.max_rec); sw0.resize(tr.max_rec);
    sw.resize((size_t)tr.max_rec * p); sv.resize((size_t)tr.max_rec * p * std::max(k, 1));
    tr.eval_train = ev.data(); tr.rec_index = ri.data(); tr.snap_w0 = sw0.data(); tr.snap_w = sw.data(); tr.snap_v = sv.data();
  }
  int rc = fmwr_train_dev(ctx, m, d, &sc, sc.step_size > 0 ? &tr : NULL);
  if (rc == 0) rc = fmwr_model_get(m, &w0, w.begin(), v.begin());
  fmwr_model_destroy(m); fmwr_data_destroy(d); fmwr_ctx_destroy(ctx);
  check(rc);

  List md = List::create(_["w0"] = w0, _["w"] = w, _["v"] = v);      // Model::save_model
  md.attr("model.control") = fm_controls;
  md.attr("solver.control") = solver_controls;
  md.attr("track.control") = track_controls;
  md.attr("convergence") = (bool)tr.convergent;
  List res;
  res["Model"] = md;
  scales.attr("target.range") = NumericVector::create(lo, hi);
  res["Scales"] = scales;
  if (sc.step_size > 0) {                                             // Tracker::save, src/core/Tracker.h:96-119
    List valid(tr.n_rec + 1);
    NumericVector idx(tr.n_rec);
    for (int i = 0; i < tr.n_rec; ++i) idx[i] = ri[i];
    valid[0] = idx;
    for (int i = 0; i < tr.n_rec; ++i) {
      NumericVector wi(sw.begin() + (size_t)i * p, sw.begin() + (size_t)(i + 1) * p);
      NumericMatrix vi(k, p);
      std::copy(sv.begin() + (size_t)i * p * k, sv.begin() + (size_t)(i + 1) * p * k, vi.begin());
      valid[i + 1] = List::create(_["w0"] = sw0[i], _["w"] = wi, _["v"] = vi);
    }
    res["Trace"] = List::create(_["trace"] = valid, _["evaluation.train"] = NumericVector(ev.begin(), ev.begin() + tr.n_rec));
  }
  res.attr("class") = "FM";
  return res;
}

// [[Rcpp::export]]
NumericVector FMPredict(List newdata, bool normalize, List model_list, int max_threads)
{
  List X = newdata["features"];
  NumericVector value = clone(as<NumericVector>(X["value"]));
  IntegerVector col_idx = X["col_idx"]; IntegerVector row_size = X["row_size"]; IntegerVector dim = X["dim"];
  const int64_t n = dim[0], p = dim[1], nnz = value.size();
  List model = model_list["Model"];
  List scales = model_list["Scales"];                                 // read unconditionally (latent crash in the reference)
  fmwr_model_cfg mc = model_cfg(as<List>(model.attr("model.control")));
  const int solver = solver_id(as<List>(model.attr("solver.control")));
  const int prec = opt_str("FM.precision", "f32") == "f64" ? FMWR_F64 : FMWR_F32;
  NumericVector w = model["w"]; NumericMatrix v = model["v"];
  NumericVector tr = as<NumericVector>(scales.attr("target.range"));

  fmwr_ctx* ctx = NULL; fmwr_data* d = NULL; fmwr_model* m = NULL;
  check(fmwr_ctx_create(opt_int("FM.device", 0), &ctx));
  check(fmwr_data_create(ctx, n, p, nnz, row_size.begin(), col_idx.begin(), value.begin(), NULL, &d));
  if (normalize) {
    NumericVector mean = scales["mean"], sd = scales["std"];
    check(fmwr_data_normalize(d, mean.begin(), sd.begin()));
  }
  check(fmwr_model_create(ctx, &mc, p, prec, &m));
  check(fmwr_model_set(m, (double)model["w0"], w.begin(), v.begin()));
  int link = mc.task == FMWR_CLASSIFICATION
                 ? ((solver == FMWR_MCMC || solver == FMWR_ALS) ? FMWR_LINK_PROBIT_TABLE : FMWR_LINK_LOGISTIC)
                 : FMWR_LINK_CLAMP;                                   // Model::predict_prob / FM.cpp:204-210
  NumericVector pred(n);
  int rc = fmwr_predict_dev(ctx, m, d, link, tr[0], tr[1]);
  if (rc == 0) rc = fmwr_predict_fetch(ctx, d, pred.begin());
  fmwr_model_destroy(m); fmwr_data_destroy(d); fmwr_ctx_destroy(ctx);
  check(rc);
  return pred;
}

// [[Rcpp::export]]
NumericVector FMTrack(List newdata, List model_list, bool normalize, String type, int max_threads)
{
  List X = newdata["features"];
  NumericVector value = clone(as<NumericVector>(X["value"]));
  IntegerVector col_idx = X["col_idx"]; IntegerVector row_size = X["row_size"]; IntegerVector dim = X["dim"];
  NumericVector labels = as<NumericVector>(newdata["labels"]);
  const int64_t n = dim[0], p = dim[1], nnz = value.size();
  List model = model_list["Model"]; List scales = model_list["Scales"];
  fmwr_model_cfg mc = model_cfg(as<List>(model.attr("model.control")));
  const int solver = solver_id(as<List>(model.attr("solver.control")));
  const int prec = opt_str("FM.precision", "f32") == "f64" ? FMWR_F64 : FMWR_F32;
  if (normalize) {                                                    // SMatrix::normalize on the host copy, Smatrix.h:144-150
    NumericVector mean = scales["mean"], sd = scales["std"];
    for (int64_t e = 0; e < nnz; ++e) { int c = col_idx[e]; if (sd[c] != 0) value[e] = (float)(((float)value[e] - mean[c]) / sd[c]); }
  }
  List trace = model_list["Trace"]; List snaps = trace["trace"];
  const int ns = snaps.size() - 1, k = mc.k;
  std::vector<double> sw0(ns), sw((size_t)ns * p), sv((size_t)ns * p * std::max(k, 1));
  for (int i = 0; i < ns; ++i) {
    List s = snaps[i + 1];
    sw0[i] = (double)s["w0"];
    NumericVector wi = s["w"]; std::copy(wi.begin(), wi.end(), sw.begin() + (size_t)i * p);
    NumericMatrix vi = s["v"]; std::copy(vi.begin(), vi.end(), sv.begin() + (size_t)i * p * k);
  }
  NumericVector tr = as<NumericVector>(scales.attr("target.range"));
  NumericVector out(ns);
  check(fmwr_track(&mc, solver, prec, n, p, nnz, row_size.begin(), col_idx.begin(), value.begin(), labels.begin(), ns,
                   sw0.data(), sw.data(), sv.data(), metric_id(type), tr[0], tr[1], out.begin()));
  return out;
}
