// Per-coordinate optimizer steps shared by the exact (batch=1) and minibatch kernels.
//
// Every function takes the coordinate's gradient g (for the exact mode g is one sample's
// gradient, for the minibatch mode the sum over the batch rows that touch the coordinate),
// the current parameter value and its optimizer state, and returns the new value.
//   SGD   reference src/solver/SGD_Learner.h:106-138 (+ apply_penalty :195-204)
//   FTRL  reference src/solver/FTRL_Learner.h:80-113 (state) and :177-199 (closed-form refresh)
//   TDAP  reference src/solver/TDAP_Learner.h:96-143 (state) and :203-231 (refresh)
#pragma once
#include "common.cuh"

namespace fmwr {

template <class T>
struct SolverParams {
  int solver;        // FMWR_SGD / FMWR_FTRL / FMWR_TDAP
  int l1;            // SGD: cumulative-penalty L1 active
  T lr;              // SGD
  T reg_w, reg_v;    // SGD: L2 rate (or L1 rate when l1)
  T reg_w0;          // SGD: L2.w0
  T alpha_w, alpha_v, beta_w, beta_v;   // FTRL / TDAP
  T l1_w, l1_v, l2_w, l2_v;             // FTRL / TDAP refresh
  T egamma;          // TDAP exp(-gamma)
};

// d/dv_f of the pairwise term for one non-zero: S_f x - v x^2, with the reference's operation order and
// NO fused multiply-add (reference src/solver/SGD_Learner.h:129: `sum * x - v * x * x`).  On a row whose
// only non-zero is this feature S_f == v*x, so the expression is exactly 0 in the reference; a contracted
// FMA would leave a rounding residue of random sign, and TDAP (no beta in its denominator) turns the sign of
// an infinitesimal gradient into a +-alpha jump of the parameter.
__device__ __forceinline__ float fm_grad(float S, float v, float x) { return __fsub_rn(__fmul_rn(S, x), __fmul_rn(__fmul_rn(v, x), x)); }
__device__ __forceinline__ double fm_grad(double S, double v, double x) { return __dsub_rn(__dmul_rn(S, x), __dmul_rn(__dmul_rn(v, x), x)); }

// cumulative-penalty clip (Tsuruoka et al.), reference src/solver/SGD_Learner.h:195-204
template <class T>
__device__ __forceinline__ void sgd_apply_penalty(T& theta, T u, T& q)
{
  // both clips computed, one selected: the same values as the reference's if / else-if, without a divergent branch per element
  const T old = theta;
  const T pos = fmax(T(0), old - (u + q)), neg = fmin(T(0), old + (u - q));
  theta = old > T(0) ? pos : (old < T(0) ? neg : old);
  q += theta - old;
}

// SGD step for one coordinate.  g excludes the learning rate.  q is touched only when l1.
template <class T>
__device__ __forceinline__ T sgd_step(T theta, T g, T lr, T reg, int l1, T u, T& q)
{
  theta -= lr * g;
  if (l1) sgd_apply_penalty(theta, u, q);
  else theta -= lr * reg * theta;
  return theta;
}

// Throughput-mode arithmetic: MUFU approximations (sqrt.approx / rcp.approx, <= 2 ulp) instead of the IEEE
// sequences (~10 instructions each, and each ends a basic block) -- the minibatch update kernel is issue-bound on
// them and the exact (batch = 1) kernel serialises a vector's elements behind them.  FAST is only ever set for
// fp32 FTRL / TDAP kernels; SGD has no sqrt / div, and every fp64 instantiation keeps IEEE operations.
template <bool FAST> __device__ __forceinline__ float m_sqrt(float x)
{
  if (FAST) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
  return sqrtf(x);
}
template <bool FAST> __device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
template <bool FAST> __device__ __forceinline__ float m_div(float a, float b) { return FAST ? __fdividef(a, b) : a / b; }
template <bool FAST> __device__ __forceinline__ double m_div(double a, double b) { return a / b; }

// FTRL-Proximal: state (z, n) in/out, returns refreshed theta
// SEL: compute the closed form unconditionally and select (same value; no branch, so the elements of a vector
// interleave in the latency-bound exact kernel)
template <class T, bool FAST = false, bool SEL = false>
__device__ __forceinline__ T ftrl_step(T theta, T g, T& z, T& n, T alpha, T beta, T l1, T l2)
{
  const T n_old = n;
  n = n_old + g * g;
  const T sq = m_sqrt<FAST>(n);
  const T sigma = m_div<FAST>(sq - m_sqrt<FAST>(n_old), alpha);
  z += g - sigma * theta;
  if (!SEL && fabs(z) <= l1) return T(0);
  const T sign = z < T(0) ? T(-1) : T(1);
  const T r = -m_div<FAST>(z - sign * l1, m_div<FAST>(beta + sq, alpha) + l2);
  return (SEL && fabs(z) <= l1) ? T(0) : r;
}

// TDAP state update: (u, nu, delta, h) in/out; returns z = nu - h
template <class T, bool FAST = false>
__device__ __forceinline__ T tdap_state(T theta, T g, T& u, T& nu, T& delta, T& h, T alpha, T egamma)
{
  const T u_old = u;
  u = u_old + g * g;
  nu += g;
  const T sigma = m_div<FAST>(m_sqrt<FAST>(u) - m_sqrt<FAST>(u_old), alpha);
  delta = egamma * (delta + sigma);
  h = egamma * (h + sigma * theta);
  return nu - h;
}

template <class T, bool FAST = false, bool SEL = false>
__device__ __forceinline__ T tdap_refresh(T z, T delta, T l1, T l2)
{
  if (!SEL && fabs(z) <= l1) return T(0);
  const T sign = z < T(0) ? T(-1) : T(1);
  const T r = -m_div<FAST>(z - sign * l1, delta + l2);
  return (SEL && fabs(z) <= l1) ? T(0) : r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

}  // namespace fmwr
