// Exact (batch = 1) SGD / FTRL-Proximal / TDAP in the reference's visit order.
//
// Replaces SGD_Learner::learn, FTRL_Learner::learn + calculate_param, TDAP_Learner::learn +
// calculate_param (reference src/solver/SGD_Learner.h:79-178, FTRL_Learner.h:64-202,
// TDAP_Learner.h:79-233).  Samples are strictly serial (w0 is read by every forward and written
// by every update), so the parallelism is INSIDE a sample: one persistent CTA walks the visit
// sequence; thread (slot, l) owns non-zero `slot` of the row and the 16-byte vector l of that
// feature's factor row.  Within one sample every coordinate update is independent given `mult`
// and the frozen S_f (each (f, j) is touched once), which is exactly what the reference's
// f-outer / nnz-inner loops compute.
#include "forward.cuh"
#include "coord.cuh"

#include <cmath>
#include <cstdlib>

namespace fmwr {

constexpr int EX_THREADS = 512;
constexpr int EX_WARPS = EX_THREADS / 32;

template <class T>
struct ExactArgs {
  const uint32_t* rowptr; const uint32_t* col; const float* val; const float* y;
  T* w; T* v; double* scal;
  T* sw[4]; T* sv[4];
  int kp, k0, k1, task;
  int64_t n;
  const uint32_t* order;     // explicit visit order (NULL: scan)
  int64_t order_len;
  int skip_row0;             // F5: scan rows 1..n-1
  int64_t t_begin, t_end;    // sample counters of this launch
  int tdap_zw_index;         // F6
  SolverParams<T> sp;
  double lo, hi;
};

template <class T, int LPR, int CH>
__global__ void __launch_bounds__(EX_THREADS, 1) exact_kernel(ExactArgs<T> a)
{
  typedef typename Vec<T>::type V16;
  constexpr int VN = Vec<T>::N;
  constexpr int KP = LPR * CH * VN;
  constexpr int SLOTS = EX_THREADS / LPR;
  constexpr int SPW = 32 / LPR;            // slots per warp
  __shared__ T sS[EX_WARPS][KP];
  __shared__ T sPart[EX_WARPS];
  __shared__ T sFin[KP];
  __shared__ T sScore;
  __shared__ double sc[8];                  // w0 and the scalar optimizer state

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = tid / LPR, l = tid % LPR;
  const int kp = a.kp;
  const SolverParams<T> sp = a.sp;
  if (tid < 8) sc[tid] = a.scal[tid];
  __syncthreads();

  for (int64_t t = a.t_begin; t < a.t_end; ++t) {
    int64_t row;
    if (a.order) row = a.order[t % a.order_len];
    else if (a.skip_row0) row = 1 + (t % (a.n - 1));
    else row = t % a.n;
    const uint32_t b = a.rowptr[row], e = a.rowptr[row + 1];

    // ---- pass A: S_f, sum Q, linear term (Model::predict, reference src/core/Model.h:75-103)
    T S[CH][VN];
    T qsum = T(0), lin = T(0);
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VN; ++i) S[ch][i] = T(0);
    for (uint32_t j0 = b; j0 < e; j0 += SLOTS) {
      const uint32_t j = j0 + slot;
      if (j < e) {
        const uint32_t c = a.col[j];
        const T x = T(a.val[j]);
        const V16* vr = reinterpret_cast<const V16*>(a.v + (size_t)c * kp);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          T arr[VN];
          vec_to_arr(vr[ch * LPR + l], arr);
#pragma unroll
          for (int i = 0; i < VN; ++i) {
            const T tt = arr[i] * x;
            S[ch][i] += tt;
            qsum += tt * tt;
          }
        }
        if (a.k1 && l == 0) lin += a.w[c] * x;
      }
    }
    // slots of this warp -> lanes < LPR hold the warp's S
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VN; ++i) S[ch][i] += __shfl_xor_sync(0xffffffffu, S[ch][i], o);
    const T part = warp_sum(lin - T(0.5) * qsum);
    if (lane < LPR) {
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VN; ++i) sS[warp][(ch * LPR + l) * VN + i] = S[ch][i];
    }
    if (lane == 0) sPart[warp] = part;
    __syncthreads();
    if (warp == 0) {
      T acc = T(0);
      for (int f = lane; f < KP; f += 32) {
        T sf = T(0);
#pragma unroll
        for (int w2 = 0; w2 < EX_WARPS; ++w2) sf += sS[w2][f];
        sFin[f] = sf;
        acc += T(0.5) * sf * sf;
      }
      if (lane < EX_WARPS) acc += sPart[lane];
      acc = warp_sum(acc);
      if (lane == 0) {
        sScore = (a.k0 ? T(sc[0]) : T(0)) + acc;
        // SGD cumulative-L1 totals advance once per sample, before the updates (SGD_Learner.h:92-97)
        if (sp.solver == FMWR_SGD && sp.l1) { sc[1] += (double)sp.lr * (double)sp.reg_w; sc[2] += (double)sp.lr * (double)sp.reg_v; }
      }
    }
    __syncthreads();

    // ---- multiplier (calculate_grad_mult)
    const T yv = T(a.y[row]);
    const T mult = grad_mult<T>(a.task, sScore, yv, T(a.lo), T(a.hi));
    T Sf[CH][VN];
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VN; ++i) Sf[ch][i] = sFin[(ch * LPR + l) * VN + i];

    const bool sgd_l1 = (sp.solver == FMWR_SGD) && sp.l1;
    const T u_w = sgd_l1 ? T(sc[1]) : T(0), u_v = sgd_l1 ? T(sc[2]) : T(0);

    // ---- pass B: coordinate updates
    for (uint32_t j0 = b; j0 < e; j0 += SLOTS) {
      const uint32_t j = j0 + slot;
      if (j < e) {
        const uint32_t c = a.col[j];
        const T x = T(a.val[j]);
        // linear weight
        if (a.k1 && l == 0) {
          const T g = mult * x;
          T th = a.w[c];
          if (sp.solver == FMWR_SGD) {
            T q = sp.l1 ? a.sw[0][c] : T(0);
            th = sgd_step(th, g, sp.lr, sp.reg_w, sp.l1, u_w, q);
            if (sp.l1) a.sw[0][c] = q;
            a.w[c] = th;
          } else if (sp.solver == FMWR_FTRL) {
            T z = a.sw[0][c], nn = a.sw[1][c];
            th = ftrl_step(th, g, z, nn, sp.alpha_w, sp.beta_w, sp.l1_w, sp.l2_w);
            a.sw[0][c] = z; a.sw[1][c] = nn;
            a.w[c] = th;
          } else {
            T u = a.sw[0][c], nu = a.sw[1][c], dl = a.sw[2][c], h = a.sw[3][c];
            const T z = tdap_state(th, g, u, nu, dl, h, sp.alpha_w, sp.egamma);
            a.sw[0][c] = u; a.sw[1][c] = nu; a.sw[2][c] = dl; a.sw[3][c] = h;
            if (!a.tdap_zw_index) a.w[c] = tdap_refresh(z, dl, sp.l1_w, sp.l2_w);
          }
        }
        // factors
        V16* vr = reinterpret_cast<V16*>(a.v + (size_t)c * kp);
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          const int vi = ch * LPR + l;
          T th[VN];
          vec_to_arr(vr[vi], th);
          if (sp.solver == FMWR_SGD) {
            T q[VN];
            V16* qr = sp.l1 ? reinterpret_cast<V16*>(a.sv[0] + (size_t)c * kp) : nullptr;
            if (sp.l1) vec_to_arr(qr[vi], q);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
              const T grad = fm_grad(Sf[ch][i], th[i], x);
              T qq = sp.l1 ? q[i] : T(0);
              th[i] = sgd_step(th[i], mult * grad, sp.lr, sp.reg_v, sp.l1, u_v, qq);
              if (sp.l1) q[i] = qq;
            }
            if (sp.l1) qr[vi] = arr_to_vec(q);
          } else if (sp.solver == FMWR_FTRL) {
            V16* zr = reinterpret_cast<V16*>(a.sv[0] + (size_t)c * kp);
            V16* nr = reinterpret_cast<V16*>(a.sv[1] + (size_t)c * kp);
            T z[VN], nn[VN];
            vec_to_arr(zr[vi], z); vec_to_arr(nr[vi], nn);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
              const T g = mult * fm_grad(Sf[ch][i], th[i], x);
              th[i] = ftrl_step(th[i], g, z[i], nn[i], sp.alpha_v, sp.beta_v, sp.l1_v, sp.l2_v);
            }
            zr[vi] = arr_to_vec(z); nr[vi] = arr_to_vec(nn);
          } else {
            V16* ur = reinterpret_cast<V16*>(a.sv[0] + (size_t)c * kp);
            V16* nur = reinterpret_cast<V16*>(a.sv[1] + (size_t)c * kp);
            V16* dr = reinterpret_cast<V16*>(a.sv[2] + (size_t)c * kp);
            V16* hr = reinterpret_cast<V16*>(a.sv[3] + (size_t)c * kp);
            T u[VN], nu[VN], dl[VN], h[VN];
            vec_to_arr(ur[vi], u); vec_to_arr(nur[vi], nu); vec_to_arr(dr[vi], dl); vec_to_arr(hr[vi], h);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
              const T g = mult * fm_grad(Sf[ch][i], th[i], x);
              const T z = tdap_state(th[i], g, u[i], nu[i], dl[i], h[i], sp.alpha_v, sp.egamma);
              th[i] = tdap_refresh(z, dl[i], sp.l1_v, sp.l2_v);
            }
            ur[vi] = arr_to_vec(u); nur[vi] = arr_to_vec(nu); dr[vi] = arr_to_vec(dl); hr[vi] = arr_to_vec(h);
          }
          vr[vi] = arr_to_vec(th);
        }
      }
    }

    // ---- scalars: w0 and its optimizer state (thread 0)
    if (tid == 0) {
      const double g = (double)mult;
      if (sp.solver == FMWR_SGD) {
        if (a.k0) sc[0] -= (double)sp.lr * (g + (double)sp.reg_w0 * sc[0]);       // SGD_Learner.h:106-109
      } else if (sp.solver == FMWR_FTRL) {
        if (a.k0) {                                                              // FTRL_Learner.h:80-86
          const double old = sc[2];
          sc[2] += g * g;
          const double delta = (sqrt(sc[2]) - sqrt(old)) / (double)sp.alpha_w;
          sc[1] += g - delta * sc[0];
        }
        sc[0] = -sc[1] * (double)sp.alpha_w / ((double)sp.beta_w + sqrt(sc[2]));  // :161, unconditional
      } else {
        if (a.k0) {                                                              // TDAP_Learner.h:96-105
          const double old = sc[1];
          sc[1] += g * g; sc[2] += g;
          const double sigma = (sqrt(sc[1]) - sqrt(old)) / (double)sp.alpha_w;
          sc[3] = (double)sp.egamma * (sc[3] + sigma);
          sc[4] = (double)sp.egamma * (sc[4] + sigma * sc[0]);
          sc[5] = sc[2] - sc[4];
        }
        sc[0] = -sc[5] / sc[3];                                                  // :192 (0/0 = NaN when keep.w0 is false)
      }
    }
    __syncthreads();

    // ---- F6: TDAP linear refresh reads z_w[position in row], not z_w[column] (TDAP_Learner.h:207)
    if (sp.solver == FMWR_TDAP && a.tdap_zw_index && a.k1) {
      for (uint32_t j0 = b; j0 < e; j0 += SLOTS) {
        const uint32_t j = j0 + slot;
        if (j < e && l == 0) {
          const uint32_t c = a.col[j];
          const uint32_t pos = j - b;                       // pos < nnz(row) <= p
          const T z = a.sw[1][pos] - a.sw[3][pos];           // z_w[pos] = nu_w[pos] - h_w[pos]
          a.w[c] = tdap_refresh(z, a.sw[2][c], sp.l1_w, sp.l2_w);
        }
      }
      __syncthreads();
    }
  }
  if (tid < 8) a.scal[tid] = sc[tid];
  (void)SPW;
}

template <class T>
struct ExactLaunch {
  fmwr_ctx* ctx; ExactArgs<T> args;
  template <class TT, int LPR, int CH>
  void run() { FMWR_LAUNCH(ctx, (exact_kernel<TT, LPR, CH>), 1, EX_THREADS, 0, args); }
};

template <class T>
static SolverParams<T> make_params(const fmwr_model* m, const fmwr_solver_cfg* s)
{
  SolverParams<T> sp;
  memset(&sp, 0, sizeof sp);
  sp.solver = s->solver;
  const fmwr_model_cfg& c = m->cfg;
  if (s->solver == FMWR_SGD) {
    // SGD_Learner::init (reference src/solver/SGD_Learner.h:44-59): any L1 > 0 selects the L1 rates;
    // L1 proper is only used for classification, in regression the L1 rates act as L2 rates.
    double regw, regv; int l1 = 0;
    if (c.l1_w1 > 0 || c.l1_v > 0) { l1 = 1; regw = c.l1_w1; regv = c.l1_v; }
    else { regw = c.l2_w1; regv = c.l2_v; }
    if (c.task != FMWR_CLASSIFICATION) l1 = 0;
    sp.l1 = l1; sp.lr = T(s->learn_rate); sp.reg_w = T(regw); sp.reg_v = T(regv); sp.reg_w0 = T(c.l2_w0);
  } else {
    sp.alpha_w = T(s->alpha_w); sp.alpha_v = T(s->alpha_v); sp.beta_w = T(s->beta_w); sp.beta_v = T(s->beta_v);
    sp.l1_w = T(c.l1_w1); sp.l1_v = T(c.l1_v); sp.l2_w = T(c.l2_w1); sp.l2_v = T(c.l2_v);
    sp.egamma = T(std::exp(-s->gamma));
  }
  return sp;
}

SolverParams<double> make_params_f64(const fmwr_model* m, const fmwr_solver_cfg* s) { return make_params<double>(m, s); }

int solver_state_count(const SolverParams<double>& sp)
{
  if (sp.solver == FMWR_SGD) return sp.l1 ? 1 : 0;
  if (sp.solver == FMWR_FTRL) return 2;
  return 4;
}

// train-set score for the tracker (reference src/solver/SGD_Learner.h:140-156)
double tracker_score(fmwr_ctx* ctx, fmwr_model* m, fmwr_data* d, const fmwr_solver_cfg* s)
{
  int link;
  if (m->cfg.task == FMWR_REGRESSION) link = FMWR_LINK_CLAMP;
  else link = (s->solver == FMWR_MCMC || s->solver == FMWR_ALS) ? FMWR_LINK_PROBIT_TABLE : FMWR_LINK_LOGISTIC;
  forward_launch(ctx, m, d, link, s->min_target, s->max_target);
  return evaluate_dev(ctx, d, m->cfg.task, s->metric);
}

void tracker_snapshot(fmwr_model* m, fmwr_trace* tr, int idx, int iter, double score)
{
  if (!tr || idx >= tr->max_rec) return;
  if (tr->eval_train) tr->eval_train[idx] = score;
  if (tr->rec_index) tr->rec_index[idx] = iter;
  if (tr->snap_w0 || tr->snap_w || tr->snap_v) {
    double w0 = 0;
    model_get_host(m, &w0, tr->snap_w ? tr->snap_w + (size_t)idx * m->p : nullptr,
                   tr->snap_v ? tr->snap_v + (size_t)idx * m->p * m->k : nullptr);
    if (tr->snap_w0) tr->snap_w0[idx] = w0;
  }
}

// Tracker::init step-size rule (reference src/core/Tracker.h:41-52, MAX_REC = 10000)
int tracker_step_size(int step_size, int max_iter)
{
  if (step_size <= 0) return step_size;
  const int rt = (int)(std::ceil(((double)max_iter - 0.5) / (double)step_size)) + 1;
  if (rt > 10000) step_size = (int)((double)(max_iter + 1) / 10000.0) + 1;
  return step_size;
}

template <class T>
static void train_exact_t(fmwr_ctx* ctx, fmwr_model* m, fmwr_data* d, const fmwr_solver_cfg* s, fmwr_trace* tr)
{
  const SolverParams<double> spd = make_params<double>(m, s);
  model_alloc_state(m, solver_state_count(spd), s->solver, s->warm_state != 0);   // Learner::init(): the state restarts at zero unless warm

  ExactArgs<T> a;
  memset(&a, 0, sizeof a);
  a.rowptr = d->rowptr.p; a.col = d->col.p; a.val = d->val.p; a.y = d->y.p;
  a.w = (T*)m->w.p; a.v = (T*)m->v.p; a.scal = (double*)m->scal.p;
  for (int i = 0; i < 4; ++i) { a.sw[i] = (T*)m->sw[i].p; a.sv[i] = (T*)m->sv[i].p; }
  a.kp = m->kp; a.k0 = m->cfg.keep_w0; a.k1 = m->cfg.keep_w1; a.task = m->cfg.task;
  a.n = d->n;
  a.skip_row0 = (s->compat & FMWR_COMPAT_SKIP_ROW0) ? 1 : 0;
  a.tdap_zw_index = (s->compat & FMWR_COMPAT_TDAP_ZW_INDEX) ? 1 : 0;
  a.sp = make_params<T>(m, s);
  a.lo = s->min_target; a.hi = s->max_target;

  // visit order.  random_step == 1: the reference scans i = 1 .. n-1 (F5).  random_step > 1: strides of
  // 1 + floor(U * random_step) from glibc rand() (reference src/util/Random.h:126-132); the caller may pass
  // the sequence explicitly (visit_order) so both sides of a comparison use the same one.
  DBuf<uint32_t> order_dev;
  std::vector<uint32_t> order_host;
  const int64_t max_iter = s->max_iter;
  if (s->visit_order && s->n_visit > 0) {
    order_host.assign(s->visit_order, s->visit_order + s->n_visit);
  } else if (s->random_step > 1) {
    auto draw = [&]() -> uint32_t { return (uint32_t)((std::rand() / ((double)RAND_MAX + 1)) * s->random_step + 1); };
    while ((int64_t)order_host.size() < max_iter) {
      const size_t before = order_host.size();
      for (uint32_t i = draw(); i < (uint32_t)d->n && (int64_t)order_host.size() < max_iter; i += draw()) order_host.push_back(i);
      if (order_host.size() == before && d->n <= 1) break;
    }
  }
  if (!order_host.empty()) {
    for (uint32_t r : order_host) FMWR_REQUIRE((int64_t)r < d->n, FMWR_ERR_ARG, "visit_order entry out of range");
    order_dev.alloc(order_host.size());
    FMWR_CUDA(cudaMemcpyAsync(order_dev.p, order_host.data(), 4 * order_host.size(), cudaMemcpyHostToDevice, ctx->stream));
    a.order = order_dev.p; a.order_len = (int64_t)order_host.size();
  } else {
    if (a.skip_row0 && d->n <= 1) { if (tr) tr->iters_done = 0; return; }   // the reference would never visit a row
    if (d->n == 0) { if (tr) tr->iters_done = 0; return; }
  }

  const int step = tracker_step_size(s->step_size, s->max_iter);
  int64_t iter = 0;
  int n_rec = 0, conv_times = 0, convergent = 0;
  double old_score = 0.0;
  while (iter < max_iter) {
    // run up to the next tracker point: the reference evaluates after the update of samples
    // 0, step, 2*step, ... and of sample max_iter-1 (SGD_Learner.h:140-143)
    int64_t next = max_iter;
    if (step > 0) {
      const int64_t k = iter / step;
      const int64_t cand = (iter % step == 0) ? iter + 1 : (k + 1) * step + 1;
      next = std::min<int64_t>(cand, max_iter);
    }
    a.t_begin = iter; a.t_end = next;
    ExactLaunch<T> L{ctx, a};
    dispatch_layout<T>(m->kp, L);
    iter = next;
    if (step > 0) {
      const int64_t last = iter - 1;     // index of the sample just processed
      if (last % step == 0 || last == max_iter - 1) {
        const double score = tracker_score(ctx, m, d, s);
        if (last > step && std::fabs((score - old_score) / (old_score + 1e-30)) <= s->convergence) conv_times++;
        else conv_times = 0;
        old_score = score;
        tracker_snapshot(m, tr, n_rec, (int)last, score);
        n_rec++;
        if (conv_times >= 3) { convergent = 1; break; }
      }
    }
  }
  FMWR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (tr) { tr->n_rec = n_rec; tr->convergent = convergent; tr->iters_done = (int)iter; }
}

void train_exact(fmwr_ctx* ctx, fmwr_model* m, fmwr_data* d, const fmwr_solver_cfg* s, fmwr_trace* tr)
{
  if (m->prec == FMWR_F64) train_exact_t<double>(ctx, m, d, s, tr);
  else train_exact_t<float>(ctx, m, d, s, tr);
}

}  // namespace fmwr
