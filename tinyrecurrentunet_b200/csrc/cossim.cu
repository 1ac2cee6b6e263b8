// cos_loss.py:41-56 (CosSimLoss.forward): mean over segments [g[i-1], g[i]) of 1 - cosine similarity of the
// prediction and the target slice, per batch row.  The reference builds the result with torch.FloatTensor(list), which only
// works for one row and drops the gradient; here rows are averaged and the loss is differentiable w.r.t. the prediction.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/tru_b200.h"
#include "tru_common.cuh"

namespace tru {
namespace {

constexpr int CNT = 256;

__device__ __forceinline__ double block_sum3(double v, double* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < CNT / 32; ++w) s += sm[w];
  __syncthreads();
  return s;
}

// stats[(b * n_seg + s) * 3 + {0,1,2}] = sum x y, sum x^2, sum y^2 over segment s of row b
__global__ void __launch_bounds__(CNT) cossim_stats_kernel(TruCosSimDesc d, const float* __restrict__ x, const float* __restrict__ y,
                                                           double* __restrict__ stats) {
  __shared__ double sm[CNT / 32];
  const int s = blockIdx.x, b = blockIdx.y;
  const int lo = d.bounds[s], hi = d.bounds[s + 1];
  const float* xr = x + (size_t)b * d.n_samples;
  const float* yr = y + (size_t)b * d.n_samples;
  double xy = 0.0, xx = 0.0, yy = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += CNT) {
    const double a = xr[i], c = yr[i];
    xy += a * c; xx += a * a; yy += c * c;
  }
  xy = block_sum3(xy, sm); xx = block_sum3(xx, sm); yy = block_sum3(yy, sm);
  if (threadIdx.x == 0) {
    double* o = stats + ((size_t)b * d.n_seg + s) * 3;
    o[0] = xy; o[1] = xx; o[2] = yy;
  }
}

__global__ void cossim_loss_kernel(TruCosSimDesc d, const double* __restrict__ stats, float* __restrict__ loss) {
  double acc = 0.0;
  const int n = d.batch * d.n_seg;
  for (int i = threadIdx.x; i < n; i += 32) {
    const double nx = fmax(sqrt(stats[3 * i + 1]), d.eps), ny = fmax(sqrt(stats[3 * i + 2]), d.eps);   // nn.CosineSimilarity clamps each norm
    acc += 1.0 - stats[3 * i] / (nx * ny);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) *loss = (float)(acc / n);
}

__global__ void __launch_bounds__(CNT) cossim_bwd_kernel(TruCosSimDesc d, const float* __restrict__ x, const float* __restrict__ y,
                                                         const double* __restrict__ stats, const float* __restrict__ gl,
                                                         float* __restrict__ gx) {
  const int b = blockIdx.y;
  const float g = __ldg(gl) / (float)(d.batch * d.n_seg);
  for (int i = blockIdx.x * CNT + threadIdx.x; i < d.n_samples; i += gridDim.x * CNT) {
    float v = 0.f;
    int s = -1;
    for (int k = 0; k < d.n_seg; ++k) if (i >= d.bounds[k] && i < d.bounds[k + 1]) s = k;
    if (s >= 0) {
      const double* st = stats + ((size_t)b * d.n_seg + s) * 3;
      const double rx = sqrt(st[1]), ry = sqrt(st[2]);
      const double nx = fmax(rx, d.eps), ny = fmax(ry, d.eps);
      // d/dx [ x.y / (nx ny) ] as torch differentiates it: the clamp is applied to a detached copy of the norm, so the
      // norm's own derivative x / |x| stays in the second term even where the clamp is active
      double t = (double)y[(size_t)b * d.n_samples + i] / (nx * ny);
      if (rx > 0.0) t -= st[0] * (double)x[(size_t)b * d.n_samples + i] / (nx * nx * rx * ny);
      v = -g * (float)t;
    }
    gx[(size_t)b * d.n_samples + i] = v;
  }
}

int check_desc(const TruCosSimDesc* d) {
  TRU_REQUIRE(d, TRU_ERR_ARG, "cossim: null descriptor");
  TRU_REQUIRE(d->batch > 0 && d->n_samples > 0 && d->n_seg > 0 && d->n_seg <= 8, TRU_ERR_ARG, "cossim: batch %d, n_samples %d, n_seg %d",
              d->batch, d->n_samples, d->n_seg);
  for (int s = 0; s < d->n_seg; ++s)
    TRU_REQUIRE(d->bounds[s] >= 0 && d->bounds[s] < d->bounds[s + 1] && d->bounds[s + 1] <= d->n_samples, TRU_ERR_ARG,
                "cossim: segment %d = [%d, %d) does not fit %d samples", s, d->bounds[s], d->bounds[s + 1], d->n_samples);
  return TRU_OK;
}

}  // namespace
}  // namespace tru

using namespace tru;

extern "C" int tru_cossim_fwd(const TruCosSimDesc* d, const float* x, const float* y, double* stats, float* loss, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_desc(d))) return rc;
  TRU_REQUIRE(x && y && stats && loss, TRU_ERR_ARG, "cossim_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  {
    ProfScope prof("cossim_stats", 8.0 * d->batch * (d->bounds[d->n_seg] - d->bounds[0]), 6.0 * d->batch * (d->bounds[d->n_seg] - d->bounds[0]), st);
    cossim_stats_kernel<<<dim3(d->n_seg, d->batch), CNT, 0, st>>>(*d, x, y, stats);
    TRU_LAUNCH_CHECK();
  }
  {
    ProfScope prof("cossim_loss", 24.0 * d->batch * d->n_seg, 0, st);
    cossim_loss_kernel<<<1, 32, 0, st>>>(*d, stats, loss);
    TRU_LAUNCH_CHECK();
  }
  return TRU_OK;
}

extern "C" int tru_cossim_bwd(const TruCosSimDesc* d, const float* x, const float* y, const double* stats, const float* grad_loss,
                              float* grad_x, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  if ((rc = check_desc(d))) return rc;
  TRU_REQUIRE(x && y && stats && grad_loss && grad_x, TRU_ERR_ARG, "cossim_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int gx = (d->n_samples + CNT - 1) / CNT;
  const int cap = (4 * sm_count() + d->batch - 1) / d->batch;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  ProfScope prof("cossim_bwd", 12.0 * d->batch * d->n_samples, 4.0 * d->batch * d->n_samples, st);
  cossim_bwd_kernel<<<dim3(gx, d->batch), CNT, 0, st>>>(*d, x, y, stats, grad_loss, grad_x);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
