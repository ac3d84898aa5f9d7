// Train/test metrics on the device: LL, AUC, ACC, RMSE, MAE
// (reference src/core/Evaluation.h:20-115, used by Tracker::evaluate / report).
#include "common.cuh"


namespace fmwr {

constexpr int RED_BLOCKS = 1024;

// mode: 0 rmse sum-of-squares, 1 mae sum-of-abs, 2 accuracy count, 3 log-likelihood
template <class T>
__global__ void metric_partial(const T* __restrict__ yh, const float* __restrict__ y, int64_t n, int mode,
                               double* __restrict__ part)
{
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = (double)yh[i];
    const double t = (double)y[i];
    if (mode == 0) { const double e = a - t; acc += e * e; }
    else if (mode == 1) { acc += fabs(a - t); }
    else if (mode == 2) { acc += (((a >= 0.5) && (t > 0)) || ((a < 0.5) && (t < 0))) ? 1.0 : 0.0; }   // Evaluation.h:44-54
    else { acc += (1 + t) * log(a + 1e-20) + (1 - t) * log(1 - a - 1e-20); }                         // Evaluation.h:80-89
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    part[blockIdx.x] = s;
  }
}

template <class T>
__global__ void auc_keys(const T* __restrict__ yh, const float* __restrict__ y, int64_t n, uint64_t* __restrict__ key,
                         uint32_t* __restrict__ pos)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = y[i] > 0 ? (double)yh[i] : -(double)yh[i];     // Evaluation.h:60-63
  key[i] = (uint64_t)__double_as_longlong(fabs(t));   // non-negative doubles order like their bit patterns
  pos[i] = t > 0 ? 1u : 0u;
}

__global__ void auc_area_partial(const uint32_t* __restrict__ pos_sorted, const uint32_t* __restrict__ cum, int64_t n,
                                 double* __restrict__ part)
{
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!pos_sorted[i]) acc += (double)cum[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    part[blockIdx.x] = s;
  }
}

// tiny single-block exclusive scan helpers are in data.cu; metrics use a serial-per-block variant for clarity
__global__ void scan_u32_serial_blocks(const uint32_t* __restrict__ in, int64_t n, int64_t per, uint32_t* __restrict__ out,
                                       uint32_t* __restrict__ block_tot)
{
  // phase A: each thread scans a contiguous chunk of `per` items (exclusive) and records its total
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = t * per, e = (b + per < n) ? b + per : n;
  uint32_t run = 0;
  for (int64_t i = b; i < e; ++i) { const uint32_t x = in[i]; out[i] = run; run += x; }
  block_tot[t] = run;
}

__global__ void scan_u32_add_offsets(uint32_t* __restrict__ out, int64_t n, int64_t per, const uint32_t* __restrict__ off)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] += off[i / per];
}

static double sum_partials(fmwr_ctx* ctx, const double* dpart, int nblk)
{
  std::vector<double> h(nblk);
  FMWR_CUDA(cudaMemcpyAsync(h.data(), dpart, sizeof(double) * nblk, cudaMemcpyDeviceToHost, ctx->stream));
  FMWR_CUDA(cudaStreamSynchronize(ctx->stream));
  double s = 0;
  for (int i = 0; i < nblk; ++i) s += h[i];
  return s;
}

template <class T>
static double evaluate_t(fmwr_ctx* ctx, fmwr_data* d, const T* yh, int task, int metric)
{
  const int64_t n = d->n;
  const int nblk = (int)std::min<int64_t>(RED_BLOCKS, std::max<int64_t>(1, ceil_div64(n, 256)));
  ctx->red_scratch.ensure(RED_BLOCKS);
  if (task == FMWR_REGRESSION) {
    // evaluates(): type <= RMSE -> rmse, else "mae" with the reference's stray sqrt (Evaluation.h:24-29, :104-115)
    const int mode = metric <= FMWR_RMSE ? 0 : 1;
    FMWR_LAUNCH(ctx, metric_partial<T>, nblk, 256, 0, yh, d->y.p, n, mode, ctx->red_scratch.p);
    return std::sqrt(sum_partials(ctx, ctx->red_scratch.p, nblk) / (double)n);
  }
  if (metric >= FMWR_ACC) {
    FMWR_LAUNCH(ctx, metric_partial<T>, nblk, 256, 0, yh, d->y.p, n, 2, ctx->red_scratch.p);
    return sum_partials(ctx, ctx->red_scratch.p, nblk) / (double)n;
  }
  if (metric == FMWR_LL) {
    FMWR_LAUNCH(ctx, metric_partial<T>, nblk, 256, 0, yh, d->y.p, n, 3, ctx->red_scratch.p);
    return sum_partials(ctx, ctx->red_scratch.p, nblk) / 2.0;
  }
  // AUC (Evaluation.h:56-78): sort signed scores by |score|, count positives below each negative
  DBuf<uint64_t> key_in, key_out;
  DBuf<uint32_t> pos_in, pos_out, cum, tot;
  key_in.alloc(n); key_out.alloc(n); pos_in.alloc(n); pos_out.alloc(n); cum.alloc(n);
  FMWR_LAUNCH(ctx, auc_keys<T>, ceil_div(n, 256), 256, 0, yh, d->y.p, n, key_in.p, pos_in.p);
  sort_pairs_u64(ctx, key_in.p, key_out.p, pos_in.p, pos_out.p, n, 64);
  // two-level scan of the positive flags
  const int64_t per = 1024;
  const int64_t nthreads = ceil_div64(n, per);
  tot.alloc(nthreads + 1);
  FMWR_LAUNCH(ctx, scan_u32_serial_blocks, ceil_div(nthreads, 128), 128, 0, pos_out.p, n, per, cum.p, tot.p);
  std::vector<uint32_t> ht(nthreads);
  FMWR_CUDA(cudaMemcpyAsync(ht.data(), tot.p, 4 * nthreads, cudaMemcpyDeviceToHost, ctx->stream));
  FMWR_CUDA(cudaStreamSynchronize(ctx->stream));
  uint64_t run = 0;
  for (int64_t i = 0; i < nthreads; ++i) { const uint32_t x = ht[i]; ht[i] = (uint32_t)run; run += x; }
  const double cum_tp = (double)run;
  FMWR_CUDA(cudaMemcpyAsync(tot.p, ht.data(), 4 * nthreads, cudaMemcpyHostToDevice, ctx->stream));
  FMWR_LAUNCH(ctx, scan_u32_add_offsets, ceil_div(n, 256), 256, 0, cum.p, n, per, tot.p);
  FMWR_LAUNCH(ctx, auc_area_partial, nblk, 256, 0, pos_out.p, cum.p, n, ctx->red_scratch.p);
  double area = sum_partials(ctx, ctx->red_scratch.p, nblk);
  if (cum_tp == 0 || cum_tp == (double)n) return 1.0;
  area /= cum_tp * ((double)n - cum_tp);
  return area < 0.5 ? 1 - area : area;
}

double evaluate_dev(fmwr_ctx* ctx, fmwr_data* d, int task, int metric)
{
  if (d->n == 0) return 0.0;
  if (d->pred_prec == FMWR_F64) {
    FMWR_REQUIRE(d->pred64.p, FMWR_ERR_ARG, "no forward result on the device");
    return evaluate_t<double>(ctx, d, d->pred64.p, task, metric);
  }
  FMWR_REQUIRE(d->pred32.p, FMWR_ERR_ARG, "no forward result on the device");
  return evaluate_t<float>(ctx, d, d->pred32.p, task, metric);
}

}  // namespace fmwr
