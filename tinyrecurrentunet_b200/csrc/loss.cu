// Fused multi-resolution STFT loss (+ L1), forward and backward, one launch each.
//
// Replaces stft_loss.py:9-30 (stft: Hann-windowed torch.stft, magnitude with a
// 1e-7 clamp), :33-50 (spectral convergence, Frobenius over the whole batch),
// :53-69 (log-magnitude L1), :141-166 (average over resolutions x lambda) and
// util.py:239-240 (nn.L1Loss).  Magnitudes are never materialised: each CTA
// frames, windows and transforms (prediction, target) as ONE complex FFT
// (x in the real part, y in the imaginary part), reduces its three partial sums
// and adds them to fp64 accumulators.  The backward recomputes the spectra,
// forms dL/dX, runs one inverse complex FFT per PAIR of frames (Hermitian
// extension of two one-sided gradients) and scatter-adds the windowed result
// (adjoint of framing + reflect padding) into grad_x.
#include "tru_common.cuh"
#include "tru_fft.cuh"

namespace tru {
namespace {

constexpr int NT = 256;
constexpr int FBUF = TRU_FFT_PAD(2048);       // floats per re/im plane (shared by all groups)
constexpr int DBUF = 4 * 4 * 257;             // backward: per group 2 frames x (re,im) x bins
constexpr int L1_ELEMS = 8192;                // elements per L1 block

struct LossParams {
  const float* x; const float* y; const float* win[3]; double* sums; const double* csums;
  const float* gout; float* gx; float* out; const float2* tw;
  int B, N, nres;
  int nfft[3], hop[3], wlen[3], T[3], fpb[3], bpc[3];   // frames, frames/block, blocks/clip
  int seg[5];                                            // block ranges: res 0..2, L1
  float sc_lambda, mag_lambda;
};

__device__ __forceinline__ void block_add(double* dst, float v0, float v1, float v2, int n) {
  __shared__ double red[3][NT / 32];
  double d0 = warp_sum((double)v0), d1 = warp_sum((double)v1), d2 = warp_sum((double)v2);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = d0; red[1][w] = d1; red[2][w] = d2; }
  __syncthreads();
  if (threadIdx.x < 3 && threadIdx.x < n) {
    double s = 0;
    for (int i = 0; i < NT / 32; ++i) s += red[threadIdx.x][i];
    atomicAdd(dst + threadIdx.x, s);
  }
}

__device__ __forceinline__ float sqrta(float x) { float y; asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rsqrta(float x) { float y; asm("rsqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// load frame t of (x,y), windowed, into (re, im)
template <int N>
__device__ __forceinline__ void load_frame(const LossParams& p, int r, int b, int t, bool valid,
                                           float* z, int l) {
  const int left = (N - p.wlen[r]) >> 1;
  const float* x = p.x + (size_t)b * p.N;
  const float* y = p.y + (size_t)b * p.N;
  const float* w = p.win[r];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = l + (N / 8) * j;
    const int wi = n - left;
    float a = 0.f, c = 0.f;
    if (valid && wi >= 0 && wi < p.wlen[r]) {
      const int src = reflect_idx(t * p.hop[r] + n - N / 2, p.N);
      const float wv = __ldg(w + wi);
      a = wv * __ldg(x + src); c = wv * __ldg(y + src);
    }
    TRU_FFT_RE(z, n) = a; TRU_FFT_IM(z, n) = c;
  }
}

__device__ __forceinline__ void split_xy(const float* z, int k, int kn,
                                         float& xr, float& xi, float& yr, float& yi) {
  const float zr = TRU_FFT_RE(z, k), zi = TRU_FFT_IM(z, k);
  const float wr = TRU_FFT_RE(z, kn), wi = TRU_FFT_IM(z, kn);
  xr = 0.5f * (zr + wr); xi = 0.5f * (zi - wi);
  yr = 0.5f * (zi + wi); yi = -0.5f * (zr - wr);
}

template <int N>
__device__ void loss_fwd_block(const LossParams& p, int r, int blk, float* fre, float* fim, const float2* tw) {
  constexpr int GT = N / 8, NG = NT / GT, PADN = TRU_FFT_PAD(N);
  const int tid = threadIdx.x, g = tid / GT, l = tid % GT;
  const int b = blk / p.bpc[r], t0 = (blk % p.bpc[r]) * p.fpb[r];
  float* z = fre + g * (2 * PADN);                      // interleaved (re, im) pairs, tru_fft.cuh
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int round = 0; round * NG < p.fpb[r]; ++round) {
    const int t = t0 + round * NG + g;
    const bool valid = t < p.T[r];
    __syncthreads();
    load_frame<N>(p, r, b, t, valid, z, l);
    fft_smem<N, -1>(z, tw, l);
    if (valid) {
      for (int k = l; k <= N / 2; k += GT) {
        float xr, xi, yr, yi;
        split_xy(z, k, (N - k) & (N - 1), xr, xi, yr, yi);
        // (SFU sqrt / log2, ~2 ulp: the per-bin arithmetic with libm's logf and the IEEE sqrt was a third of this kernel's
        // instructions; |log my - log mx| = ln2 / 2 * |log2 my^2 - log2 mx^2|)
        const float px = fmaxf(xr * xr + xi * xi, 1e-7f), py = fmaxf(yr * yr + yi * yi, 1e-7f);     // stft_loss.py:30 (clamp on the power)
        const float mx = sqrta(px), my = sqrta(py);
        const float d = my - mx;
        a0 += d * d; a1 += py; a2 += 0.34657359027997264f * fabsf(lg2a(py) - lg2a(px));
      }
    }
  }
  block_add(p.sums + 4 * r, a0, a1, a2, 3);
}

__global__ void __launch_bounds__(NT) loss_fwd_kernel(LossParams p) {
  extern __shared__ __align__(16) float smem[];
  float* fre = smem;
  float* fim = fre + FBUF;
  float2* tw = (float2*)(fim + FBUF);
  const int blk = blockIdx.x;
  int r = 0;
  while (r < 3 && blk >= p.seg[r + 1]) ++r;
  if (r < p.nres) {
    const int n = p.nfft[r];
    for (int k = threadIdx.x; k < n; k += NT) tw[k] = p.tw[k * (2048 / n)];
    if (n == 512) loss_fwd_block<512>(p, r, blk - p.seg[r], fre, fim, tw);
    else if (n == 1024) loss_fwd_block<1024>(p, r, blk - p.seg[r], fre, fim, tw);
    else loss_fwd_block<2048>(p, r, blk - p.seg[r], fre, fim, tw);
  } else {                                              // L1 segment (util.py:239)
    const size_t total = (size_t)p.B * p.N;
    const size_t base = (size_t)(blk - p.seg[3]) * L1_ELEMS;
    float a = 0.f;
    for (size_t i = base + threadIdx.x; i < base + L1_ELEMS && i < total; i += NT)
      a += fabsf(__ldg(p.x + i) - __ldg(p.y + i));
    block_add(p.sums + 12, a, 0.f, 0.f, 1);
  }
}

__global__ void loss_finalize_kernel(LossParams p) {
  if (threadIdx.x != 0) return;
  double sc = 0, mg = 0;
  for (int r = 0; r < p.nres; ++r) {
    const double cnt = (double)p.B * p.T[r] * (p.nfft[r] / 2 + 1);
    sc += sqrt(p.sums[4 * r]) / sqrt(p.sums[4 * r + 1]);             // stft_loss.py:50
    mg += p.sums[4 * r + 2] / cnt;                                    // :69
  }
  p.out[0] = (float)(p.sums[12] / ((double)p.B * p.N));
  p.out[1] = (float)(sc * p.sc_lambda / p.nres);                      // :161-162
  p.out[2] = (float)(mg * p.mag_lambda / p.nres);                     // :163-164
}

template <int N>
__device__ void loss_bwd_block(const LossParams& p, int r, int blk, float* fre, float* fim,
                               float* dbuf, const float2* tw) {
  constexpr int GT = N / 8, NG = NT / GT, PADN = TRU_FFT_PAD(N), NBIN = N / 2 + 1;
  const int tid = threadIdx.x, g = tid / GT, l = tid % GT;
  const int b = blk / p.bpc[r], t0 = (blk % p.bpc[r]) * p.fpb[r];
  float* z = fre + g * (2 * PADN);                      // interleaved (re, im) pairs, tru_fft.cuh
  float* dre = dbuf + g * 4 * NBIN;                     // [2][NBIN]
  float* dim = dre + 2 * NBIN;
  const double A = p.csums[4 * r], Bn = p.csums[4 * r + 1];
  const float g_sc = __ldg(p.gout + 1), g_mag = __ldg(p.gout + 2);
  const float csc = (A > 0.0 && Bn > 0.0) ? (float)(-(double)g_sc * p.sc_lambda / p.nres / (sqrt(A) * sqrt(Bn))) : 0.f;
  const float cmg = (float)((double)g_mag * p.mag_lambda / p.nres / ((double)p.B * p.T[r] * NBIN));
  const int left = (N - p.wlen[r]) >> 1;
  float* gx = p.gx + (size_t)b * p.N;

  for (int round = 0; round * 2 * NG < p.fpb[r]; ++round) {
    const int ta = t0 + (round * NG + g) * 2;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int t = ta + h;
      const bool valid = t < p.T[r];
      __syncthreads();
      load_frame<N>(p, r, b, t, valid, z, l);
      fft_smem<N, -1>(z, tw, l);
      for (int k = l; k <= N / 2; k += GT) {
        float dr = 0.f, di = 0.f;
        if (valid) {
          float xr, xi, yr, yi;
          split_xy(z, k, (N - k) & (N - 1), xr, xi, yr, yi);
          const float px = xr * xr + xi * xi;
          if (px >= 1e-7f) {                           // clamp passes gradient only above the floor
            const float inv = rsqrta(px), mx = px * inv;                       // (SFU rsqrt / sqrt instead of IEEE sqrt + two divisions)
            const float my = sqrta(fmaxf(yr * yr + yi * yi, 1e-7f));
            const float dl = my - mx;                                          // sign(log my - log mx): the logarithm is monotone
            const float sg = dl > 0.f ? 1.f : (dl < 0.f ? -1.f : 0.f);
            const float dmx = csc * dl - cmg * sg * inv;
            const float sc = dmx * inv;
            dr = sc * xr; di = sc * xi;
          }
        }
        dre[h * NBIN + k] = dr; dim[h * NBIN + k] = di;
      }
    }
    __syncthreads();
    // Hermitian extension of the two one-sided gradients, packed as Da' + i Db'
    for (int k = l; k <= N / 2; k += GT) {
      const float ar = dre[k], ai = dim[k], br = dre[NBIN + k], bi = dim[NBIN + k];
      if (k == 0 || k == N / 2) {
        TRU_FFT_RE(z, k) = ar; TRU_FFT_IM(z, k) = br;
      } else {
        TRU_FFT_RE(z, k) = 0.5f * (ar - bi); TRU_FFT_IM(z, k) = 0.5f * (ai + br);
        TRU_FFT_RE(z, N - k) = 0.5f * (ar + bi); TRU_FFT_IM(z, N - k) = 0.5f * (br - ai);
      }
    }
    fft_smem<N, 1>(z, tw, l);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = l + GT * j;
      const int wi = n - left;
      if (wi >= 0 && wi < p.wlen[r]) {
        const float wv = __ldg(p.win[r] + wi);
        if (ta < p.T[r])
          atomicAdd(gx + reflect_idx(ta * p.hop[r] + n - N / 2, p.N), wv * TRU_FFT_RE(z, n));
        if (ta + 1 < p.T[r])
          atomicAdd(gx + reflect_idx((ta + 1) * p.hop[r] + n - N / 2, p.N), wv * TRU_FFT_IM(z, n));
      }
    }
  }
}

__global__ void __launch_bounds__(NT) loss_bwd_kernel(LossParams p) {
  extern __shared__ __align__(16) float smem[];
  float* fre = smem;
  float* fim = fre + FBUF;
  float* dbuf = fim + FBUF;
  float2* tw = (float2*)(dbuf + DBUF);
  const int blk = blockIdx.x;
  int r = 0;
  while (r < 3 && blk >= p.seg[r + 1]) ++r;
  if (r < p.nres) {
    const int n = p.nfft[r];
    for (int k = threadIdx.x; k < n; k += NT) tw[k] = p.tw[k * (2048 / n)];
    if (n == 512) loss_bwd_block<512>(p, r, blk - p.seg[r], fre, fim, dbuf, tw);
    else if (n == 1024) loss_bwd_block<1024>(p, r, blk - p.seg[r], fre, fim, dbuf, tw);
    else loss_bwd_block<2048>(p, r, blk - p.seg[r], fre, fim, dbuf, tw);
  } else {
    const size_t total = (size_t)p.B * p.N;
    const size_t base = (size_t)(blk - p.seg[3]) * L1_ELEMS;
    const float c = __ldg(p.gout) / (float)total;
    for (size_t i = base + threadIdx.x; i < base + L1_ELEMS && i < total; i += NT) {
      const float d = __ldg(p.x + i) - __ldg(p.y + i);
      if (d != 0.f) atomicAdd(p.gx + i, d > 0.f ? c : -c);
    }
  }
}

int fill(const TruLossDesc* d, const float* const* windows, LossParams& p) {
  TRU_REQUIRE(d && d->batch > 0 && d->n_samples > 0 && d->n_res >= 1 && d->n_res <= 3, TRU_ERR_ARG,
              "loss: bad descriptor");
  TRU_REQUIRE(windows, TRU_ERR_ARG, "loss: windows is null");
  p.B = d->batch; p.N = d->n_samples; p.nres = d->n_res;
  p.sc_lambda = (float)d->sc_lambda; p.mag_lambda = (float)d->mag_lambda;
  p.tw = twiddle_table();
  int seg = 0;
  for (int r = 0; r < 3; ++r) {
    p.seg[r] = seg;
    if (r >= d->n_res) continue;
    const int n = d->fft_size[r];
    TRU_REQUIRE(n == 512 || n == 1024 || n == 2048, TRU_ERR_ARG, "loss: fft_size %d unsupported (512/1024/2048)", n);
    TRU_REQUIRE(d->win_length[r] > 0 && d->win_length[r] <= n && d->hop_size[r] > 0, TRU_ERR_ARG,
                "loss: bad window/hop for resolution %d", r);
    TRU_REQUIRE(d->n_samples > n / 2, TRU_ERR_ARG, "loss: reflect padding needs N > fft_size/2 (N=%d, fft=%d)",
                d->n_samples, n);
    TRU_REQUIRE(windows[r], TRU_ERR_ARG, "loss: window %d is null", r);
    p.nfft[r] = n; p.hop[r] = d->hop_size[r]; p.wlen[r] = d->win_length[r]; p.win[r] = windows[r];
    p.T[r] = 1 + d->n_samples / d->hop_size[r];
    p.fpb[r] = 4 * (NT / (n / 8));                    // 16 / 8 / 4 frames per block
    p.bpc[r] = (p.T[r] + p.fpb[r] - 1) / p.fpb[r];
    seg += p.B * p.bpc[r];
  }
  p.seg[3] = seg;
  seg += (int)(((size_t)p.B * p.N + L1_ELEMS - 1) / L1_ELEMS);
  p.seg[4] = seg;
  return TRU_OK;
}

constexpr size_t LOSS_FWD_SMEM = (size_t)(2 * FBUF) * 4 + 2048 * 8;
constexpr size_t LOSS_BWD_SMEM = (size_t)(2 * FBUF + DBUF) * 4 + 2048 * 8;

}  // namespace
}  // namespace tru

using namespace tru;

extern "C" int tru_loss_fwd(const TruLossDesc* d, const float* x, const float* y, const float* const* windows,
                            double* sums, float* out, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  LossParams p{};
  if ((rc = fill(d, windows, p))) return rc;
  TRU_REQUIRE(x && y && sums && out, TRU_ERR_ARG, "loss_fwd: null pointer");
  p.x = x; p.y = y; p.sums = sums; p.out = out;
  cudaStream_t st = (cudaStream_t)stream;
  TRU_CUDA(cudaMemsetAsync(sums, 0, 16 * sizeof(double), st));
  TRU_CUDA(cudaFuncSetAttribute(loss_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOSS_FWD_SMEM));
  double fl = 0;
  for (int r = 0; r < p.nres; ++r) fl += (double)p.B * p.T[r] * 5.0 * p.nfft[r] * log2((double)p.nfft[r]);
  {
    ProfScope prof("loss_fwd", 8.0 * p.B * p.N, fl, st);
    loss_fwd_kernel<<<p.seg[4], NT, LOSS_FWD_SMEM, st>>>(p);
  }
  TRU_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 32, 0, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

extern "C" int tru_loss_bwd(const TruLossDesc* d, const float* x, const float* y, const float* const* windows,
                            const double* sums, const float* grad_out, float* grad_x, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  LossParams p{};
  if ((rc = fill(d, windows, p))) return rc;
  TRU_REQUIRE(x && y && sums && grad_out && grad_x, TRU_ERR_ARG, "loss_bwd: null pointer");
  p.x = x; p.y = y; p.csums = sums; p.gout = grad_out; p.gx = grad_x;
  cudaStream_t st = (cudaStream_t)stream;
  TRU_CUDA(cudaMemsetAsync(grad_x, 0, (size_t)p.B * p.N * sizeof(float), st));
  TRU_CUDA(cudaFuncSetAttribute(loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOSS_BWD_SMEM));
  double fl = 0;
  for (int r = 0; r < p.nres; ++r) fl += 1.5 * p.B * p.T[r] * 5.0 * p.nfft[r] * log2((double)p.nfft[r]);
  ProfScope prof("loss_bwd", 12.0 * p.B * p.N, fl, st);
  loss_bwd_kernel<<<p.seg[4], NT, LOSS_BWD_SMEM, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
