// Optimizer step on the flat parameter / gradient buffers:
// train.py:138 (clip_grad_norm_, used with 1e9 = "report the norm") and train.py:140 (AdamW.step, torch
// defaults) for all 108 tensors in two launches.  The learning rate is a host scalar per step
// (util.LinearWarmupCosineDecay sets it, train.py:139), so nothing here synchronises.
//
// Arithmetic follows torch.optim.AdamW's multi-tensor (foreach) path op by op in fp32 (decay, lerp, addcmul, sqrt,
// divide, add eps, addcdiv) with the host-side scalars computed in double exactly as torch does; only the
// norm differs in form (one fp64 sum of squares instead of a norm of per-tensor fp32 norms).
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/tru_b200.h"
#include "tru_common.cuh"

namespace tru {
namespace {

constexpr int ONT = 256;
constexpr int OMAXB = 148;          // one CTA per SM at most; the buffers are ~1.5 MB each and live in L2

__device__ __forceinline__ double block_sum(double s, double* sm) {
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = s;
  __syncthreads();
  if (w == 0) {
    s = l < ONT / 32 ? sm[l] : 0.0;
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) sm[0] = s;
  }
  __syncthreads();
  s = sm[0];
  __syncthreads();
  return s;
}

// partial[b] = sum of g^2 over the float4 elements CTA b owns (fixed ownership: deterministic).
__global__ void __launch_bounds__(ONT) grad_sumsq_kernel(const float4* __restrict__ g, long long n4, double* __restrict__ partial) {
  __shared__ double sm[ONT / 32];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * ONT + threadIdx.x; i < n4; i += (long long)gridDim.x * ONT) {
    const float4 v = g[i];
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  s = block_sum(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

struct AdamWScalars {
  float decay;        // 1 - lr * weight_decay
  float w1;           // 1 - beta1 (lerp weight)
  float beta2, w2;    // beta2, 1 - beta2
  float bc2s;         // sqrt(1 - beta2^t)
  float eps;
  float neg_step;     // -lr / (1 - beta1^t)
  float max_norm;     // <= 0: no clipping
};

__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, const AdamWScalars& c) {
  p = __fmul_rn(p, c.decay);
  m = c.w1 < 0.5f ? fmaf(c.w1, g - m, m) : g - __fmul_rn(g - m, 1.0f - c.w1);     // Tensor.lerp_
  v = __fmul_rn(v, c.beta2);
  v = fmaf(__fmul_rn(c.w2, g), g, v);                                            // addcmul_(g, g, value = 1 - beta2)
  const float den = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2s), c.eps);
  p = fmaf(c.neg_step, __fdiv_rn(m, den), p);                                    // addcdiv_(m, den, value = -step_size)
  return p;
}

__global__ void __launch_bounds__(ONT) adamw_flat_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                         float4* __restrict__ v, long long n4, const double* __restrict__ partial,
                                                         int npartial, float* __restrict__ norm_out, AdamWScalars c) {
  __shared__ double sm[ONT / 32];
  // every CTA adds the partials in the same order, so all of them see the same norm
  double s = 0.0;
  for (int i = threadIdx.x; i < npartial; i += ONT) s += partial[i];
  s = block_sum(s, sm);
  const float norm = (float)sqrt(s);
  if (blockIdx.x == 0 && threadIdx.x == 0 && norm_out) *norm_out = norm;
  float coef = 1.0f;
  if (c.max_norm > 0.0f) coef = fminf(__fdiv_rn(c.max_norm, __fadd_rn(norm, 1e-6f)), 1.0f);   // clip_grad_norm_
  const bool clip = coef < 1.0f;
  for (long long i = (long long)blockIdx.x * ONT + threadIdx.x; i < n4; i += (long long)gridDim.x * ONT) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    if (clip) {
      gg.x = __fmul_rn(gg.x, coef); gg.y = __fmul_rn(gg.y, coef); gg.z = __fmul_rn(gg.z, coef); gg.w = __fmul_rn(gg.w, coef);
      g[i] = gg;                                                                 // clip_grad_norm_ scales the grads in place
    }
    adamw_one(pp.x, gg.x, mm.x, vv.x, c);
    adamw_one(pp.y, gg.y, mm.y, vv.y, c);
    adamw_one(pp.z, gg.z, mm.z, vv.z, c);
    adamw_one(pp.w, gg.w, mm.w, vv.w, c);
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

int grid_for(long long n4) {
  long long b = (n4 + ONT - 1) / ONT;
  return (int)(b < 1 ? 1 : (b > OMAXB ? OMAXB : b));
}

}  // namespace
}  // namespace tru

using namespace tru;

extern "C" size_t tru_flat_adamw_workspace_bytes(const TruAdamWDesc* d) {
  (void)d;
  return OMAXB * sizeof(double);
}

extern "C" int tru_flat_grad_norm(long long n, const float* grads, float* norm_out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  TRU_REQUIRE(n > 0 && n % 4 == 0, TRU_ERR_ARG, "flat_grad_norm: n must be a positive multiple of 4, got %lld", n);
  TRU_REQUIRE(grads && norm_out && workspace, TRU_ERR_ARG, "flat_grad_norm: null pointer");
  TRU_REQUIRE(workspace_bytes >= OMAXB * sizeof(double), TRU_ERR_WORKSPACE, "flat_grad_norm: workspace too small");
  TRU_REQUIRE(((uintptr_t)grads & 15) == 0 && ((uintptr_t)workspace & 7) == 0, TRU_ERR_ALIGN, "flat_grad_norm: misaligned pointer");
  const long long n4 = n / 4;
  const int nb = grid_for(n4);
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof("grad_sumsq", 4.0 * n, 2.0 * n, st);
    grad_sumsq_kernel<<<nb, ONT, 0, st>>>((const float4*)grads, n4, (double*)workspace);
    TRU_LAUNCH_CHECK();
  }
  // a one-element "update" that only publishes the norm: reuse the reduction of the main kernel
  AdamWScalars c{};
  {
    ProfScope prof("grad_norm_out", 8.0 * nb, 0, st);
    adamw_flat_kernel<<<1, ONT, 0, st>>>(nullptr, nullptr, nullptr, nullptr, 0, (const double*)workspace, nb, norm_out, c);
    TRU_LAUNCH_CHECK();
  }
  return TRU_OK;
}

extern "C" int tru_flat_adamw_step(const TruAdamWDesc* d, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                   float* grad_norm_out, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  TRU_REQUIRE(d, TRU_ERR_ARG, "flat_adamw: null descriptor");
  TRU_REQUIRE(d->n > 0 && d->n % 4 == 0, TRU_ERR_ARG, "flat_adamw: n must be a positive multiple of 4, got %lld", d->n);
  TRU_REQUIRE(d->step >= 1, TRU_ERR_ARG, "flat_adamw: step is 1-based, got %lld", d->step);
  TRU_REQUIRE(d->lr >= 0 && d->eps >= 0 && d->weight_decay >= 0 && d->beta1 >= 0 && d->beta1 < 1 && d->beta2 >= 0 && d->beta2 < 1,
              TRU_ERR_ARG, "flat_adamw: hyper-parameter out of range");
  TRU_REQUIRE(params && grads && exp_avg && exp_avg_sq && workspace, TRU_ERR_ARG, "flat_adamw: null pointer");
  TRU_REQUIRE(workspace_bytes >= OMAXB * sizeof(double), TRU_ERR_WORKSPACE, "flat_adamw: workspace too small");
  TRU_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0 &&
                  ((uintptr_t)workspace & 7) == 0,
              TRU_ERR_ALIGN, "flat_adamw: buffers must be 16-byte aligned");
  AdamWScalars c;
  const double bc1 = 1.0 - pow(d->beta1, (double)d->step), bc2 = 1.0 - pow(d->beta2, (double)d->step);
  c.decay = (float)(1.0 - d->lr * d->weight_decay);
  c.w1 = (float)(1.0 - d->beta1);
  c.beta2 = (float)d->beta2;
  c.w2 = (float)(1.0 - d->beta2);
  c.bc2s = (float)sqrt(bc2);
  c.eps = (float)d->eps;
  c.neg_step = (float)(-(d->lr / bc1));
  c.max_norm = (float)d->max_grad_norm;
  const long long n4 = d->n / 4;
  const int nb = grid_for(n4);
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof("grad_sumsq", 4.0 * d->n, 2.0 * d->n, st);
    grad_sumsq_kernel<<<nb, ONT, 0, st>>>((const float4*)grads, n4, (double*)workspace);
    TRU_LAUNCH_CHECK();
  }
  {
    ProfScope prof("adamw_flat", 28.0 * d->n, 12.0 * d->n, st);
    adamw_flat_kernel<<<nb, ONT, 0, st>>>((float4*)params, (float4*)grads, (float4*)exp_avg, (float4*)exp_avg_sq, n4,
                                          (const double*)workspace, nb, grad_norm_out, c);
    TRU_LAUNCH_CHECK();
  }
  return TRU_OK;
}
