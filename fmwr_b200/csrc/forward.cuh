// Warp-per-row FM forward over device CSR.
//
// Replaces Model::predict / Model::predict_batch (reference src/core/Model.h:75-161):
//   y = [k0] w0 + [k1] sum_j w_j x_j + sum_f 1/2 (S_f^2 - Q_f),  S_f = sum_j v_jf x_j,  Q_f = sum_j (v_jf x_j)^2
//
// Layout: V is [p][kp] (feature-major, kp = padded k), so one feature's factors are one or a
// few contiguous 16-byte vectors.  A warp owns a row; it is split into G = 32/LPR groups of LPR
// lanes; a group fetches one V row per step with ONE 128-bit load per lane (k=32 fp32: LPR=8, a
// 128-byte line per group, 4 non-zeros per warp-wide load instruction), U steps are kept in flight.
// Lane l of every group accumulates S and Q for factors [l*VN, l*VN+VN) (+ chunk offsets); groups
// are combined with xor-shuffles at the end of the row.
#pragma once
#include "common.cuh"

namespace fmwr {

template <class T>
struct RowAcc {
  // filled by row_forward: identical in every lane of the warp
  T score;   // w0*k0 + linear + pairwise (no link)
};

// LPR: lanes per V row (power of two), CH: 16-byte chunks per lane, U: non-zeros in flight per group.
// On return S[ch][i] holds the complete S_f for factor (ch*LPR + l)*VN + i in EVERY group.
template <class T, int LPR, int CH, int U>
__device__ __forceinline__ void row_gather(const uint32_t* __restrict__ col, const float* __restrict__ val,
                                           uint32_t b, uint32_t e, const T* __restrict__ w, const T* __restrict__ v,
                                           int kp, int k1, T (&S)[CH][Vec<T>::N], T& lin_out, T& q_out)
{
  typedef typename Vec<T>::type V16;
  constexpr int VN = Vec<T>::N;
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int g = lane / LPR, l = lane % LPR;

  T Q[CH][VN];
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int i = 0; i < VN; ++i) { S[ch][i] = T(0); Q[ch][i] = T(0); }
  T lin = T(0);

  for (uint32_t base = b; base < e; base += G * U) {
    uint32_t c[U];
    T x[U];
    V16 vv[U][CH];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t j = base + u * G + g;
      const bool ok = j < e;
      c[u] = ok ? __ldg(col + j) : 0u;
      x[u] = ok ? T(__ldg(val + j)) : T(0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t j = base + u * G + g;
      const V16* vr = reinterpret_cast<const V16*>(v + (size_t)c[u] * kp);
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        if (j < e) vv[u][ch] = vr[ch * LPR + l];
        else memset(&vv[u][ch], 0, sizeof(V16));
      }
    }
    if (k1 && l == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = base + u * G + g;
        if (j < e) lin += w[c[u]] * x[u];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        T a[VN];
        vec_to_arr(vv[u][ch], a);
#pragma unroll
        for (int i = 0; i < VN; ++i) {
          const T t = a[i] * x[u];
          S[ch][i] += t;
          Q[ch][i] += t * t;
        }
      }
  }

  // combine the G groups
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VN; ++i) {
        S[ch][i] += __shfl_xor_sync(0xffffffffu, S[ch][i], o);
        Q[ch][i] += __shfl_xor_sync(0xffffffffu, Q[ch][i], o);
      }
  // lin lives in lane 0 of every group; sum Q over this group's factors (identical in every group after the combine)
  T qs = T(0);
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
#pragma unroll
    for (int i = 0; i < VN; ++i) qs += Q[ch][i];
  lin_out = lin;
  q_out = (g == 0) ? qs : T(0);
}

template <class T, int LPR, int CH, int U>
__device__ __forceinline__ T row_forward(const uint32_t* __restrict__ col, const float* __restrict__ val,
                                         uint32_t b, uint32_t e, const T* __restrict__ w, const T* __restrict__ v,
                                         int kp, T w0, int k0, int k1, T (&S)[CH][Vec<T>::N])
{
  constexpr int VN = Vec<T>::N;
  const int g = (threadIdx.x & 31) / LPR;
  T lin, qs;
  row_gather<T, LPR, CH, U>(col, val, b, e, w, v, kp, k1, S, lin, qs);
  T r = lin - T(0.5) * qs;
  if (g == 0) {
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int i = 0; i < VN; ++i) r += T(0.5) * S[ch][i] * S[ch][i];
  }
  r = warp_sum(r);
  return (k0 ? w0 : T(0)) + r;
}

// column-slice variant: S_f of the slice plus the ADDITIVE part of the score (linear term - 1/2 sum Q);
// 1/2 sum_f S_f^2 is formed after the partial S_f have been summed over the shards
template <class T, int LPR, int CH, int U>
__device__ __forceinline__ void row_forward_partial(const uint32_t* __restrict__ col, const float* __restrict__ val,
                                                    uint32_t b, uint32_t e, const T* __restrict__ w, const T* __restrict__ v,
                                                    int kp, int k1, T (&S)[CH][Vec<T>::N], T& addend)
{
  T lin, qs;
  row_gather<T, LPR, CH, U>(col, val, b, e, w, v, kp, k1, S, lin, qs);
  addend = warp_sum(lin - T(0.5) * qs);
}

// table-exact fast_pnorm (reference src/util/Random.h:95-111; Y table regenerated, see link_tables.cu)
__device__ __forceinline__ double dev_fast_pnorm(const double* __restrict__ Y, double x)
{
  const double HINV = 549.966731401936, XMAX = 5.20031455849973;
  const double ax = x < 0 ? -x : x;
  double res;
  if (ax > XMAX) {
    res = 0.999999900524235;
  } else {
    const int i = (int)(ax * HINV);
    const double wgt = (ax - (double)i / HINV) * HINV;
    res = wgt * Y[i + 1] + (1.0 - wgt) * Y[i];
  }
  return (ax == x) ? res : 1.0 - res;
}

// table-exact fast_dpnorm = phi(x)/(1-Phi(x)) (reference src/util/Random.h:114-124)
__device__ __forceinline__ double dev_fast_dpnorm(const double* __restrict__ Y, double x)
{
  const double ax = x < 0 ? -x : x;
  if (x < -3.0) return 0.0;
  if (x > 5.0) return 0.1943369 + 0.9754752 * x + 0.4136861 * sqrt(ax) - 0.5034295 * log(ax + 1e-07);
  const int i = (int)((x + 3.0) * 5000);
  const double xi = rint((-3.0 + 2e-4 * (double)i) * 1e4) / 1e4;
  const double wgt = (x - xi) * 5000;
  return wgt * Y[i + 1] + (1.0 - wgt) * Y[i];
}

__device__ __forceinline__ double apply_link(int link, double s, double lo, double hi, const double* __restrict__ pnY)
{
  switch (link) {
    case FMWR_LINK_LOGISTIC: return 1.0 / (1.0 + exp(-s));
    case FMWR_LINK_PROBIT_TABLE: return dev_fast_pnorm(pnY, s);
    case FMWR_LINK_CLAMP: return s < lo ? lo : (s > hi ? hi : s);
    default: return s;
  }
}

// calculate_grad_mult (reference src/solver/SGD_Learner.h:180-191 and twins)
template <class T>
__device__ __forceinline__ T grad_mult(int task, T y_hat, T y, T lo, T hi)
{
  if (task == FMWR_REGRESSION) {
    y_hat = fmin(hi, y_hat);
    y_hat = fmax(lo, y_hat);
    return -(y - y_hat);
  }
  return -y * (T(1) - T(1) / (T(1) + exp(-y * y_hat)));
}

// dispatch helper: calls F.template run<T, LPR, CH>() for the model's layout
template <class T, class F>
static inline void dispatch_layout(int kp, F&& f)
{
  const int units = kp / Vec<T>::N;
  switch (units) {
    case 1: f.template run<T, 1, 1>(); break;
    case 2: f.template run<T, 2, 1>(); break;
    case 4: f.template run<T, 4, 1>(); break;
    case 8: f.template run<T, 8, 1>(); break;
    case 16: f.template run<T, 16, 1>(); break;
    case 32: f.template run<T, 32, 1>(); break;
    case 64: f.template run<T, 32, 2>(); break;
    case 128: f.template run<T, 32, 4>(); break;
    default: throw Error(FMWR_ERR_UNSUPPORTED, "unsupported padded factor count");
  }
}

}  // namespace fmwr
