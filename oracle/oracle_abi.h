/* TEST INFRASTRUCTURE ONLY -- shared C ABI of the two checkers:
 *   oracle/_ref/libfmwr_ref.so   the reference's own headers, compiled unmodified (ref_driver.cpp)
 *   oracle/libfm_oracle.so       the plain-C restatement (fm_oracle.c)
 * Both export the same entry points with prefix fmwr_ref_ / fmwr_orc_ so tests
 * can pin the restatement against the reference function by function.
 * Enum values are the reference's (src/util/Macros.h:10-30).
 */
#ifndef FMWR_ORACLE_ABI_H_
#define FMWR_ORACLE_ABI_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CLASSIFICATION = 10, ORC_REGRESSION = 20 };
enum { ORC_MCMC = 100, ORC_ALS = 200, ORC_SGD = 300, ORC_FTRL = 500, ORC_TDAP = 600 };
enum { ORC_LL = 0, ORC_AUC = 111, ORC_ACC = 222, ORC_RMSE = 333, ORC_MSE = 444, ORC_MAE = 555 };

typedef struct {
  int task;            /* ORC_CLASSIFICATION / ORC_REGRESSION */
  int solver;          /* ORC_* */
  int k0, k1, k;       /* keep.w0, keep.w1, factor.number */
  double l2_w0, l1_w, l2_w, l1_v, l2_v;
  int max_iter;        /* SGD/FTRL/TDAP: sample updates; ALS/MCMC: sweeps (F4) */
  int random_step;
  int nthreads;
  double learn_rate;                         /* SGD */
  double alpha_w, alpha_v, beta_w, beta_v;   /* FTRL (TDAP uses alpha_*) */
  double gamma;                              /* TDAP */
  int enable_v;        /* ALS/MCMC: 1 = run the update_v block that the shipped update_all comments out (F1) */
  double min_target, max_target;
  int step_size;       /* tracker: <=0 off */
  int metric;          /* ORC_LL ... */
  double convergence;
} fmwr_oracle_cfg;

typedef struct {
  int n_rec;           /* records written */
  int convergent;
  int iters_done;
} fmwr_oracle_trace_info;

#ifdef __cplusplus
}
#endif
#endif
