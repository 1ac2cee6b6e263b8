// Shared-memory Stockham FFT building blocks (radix 8/4, in-register butterflies).
//
// One complex FFT of size N is carried out by N/8 threads; data lives in ONE padded array of interleaved (re, im) pairs
// (element a at pair index a + a/16), so every butterfly operand is one 8-byte shared-memory access: ncu had the FFT-bound
// kernels (loss_fwd / loss_bwd) stalled on "mio throttle" - the shared-memory instruction rate - with separate re / im
// planes (16 + 16 four-byte accesses per radix-8 pass and thread); the padding keeps the 8-byte loads conflict-free per
// half warp and leaves 2-way conflicts on one store pass only.  The per-thread halves of a pass (load+twiddle+butterfly,
// then store) are plain inline functions so the exact same code is exercised on the host by tests/host/host_fft_test.cpp.
// Buffers are declared as float z[2 * TRU_FFT_PAD(N)] (a little more than the 2 * (N + N/16) floats in use).
//
// Used by: frontend.cu (STFT-512, dataset.py:260-264), backend.cu (irFFT-512,
// dataset.py:293-296), loss.cu (STFT 512/1024/2048, stft_loss.py:21-23).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define TRU_HD __host__ __device__ __forceinline__
#else
#define TRU_HD inline
struct alignas(8) float2 { float x, y; };
#endif

#define TRU_FFT_IDX(a) ((a) + ((a) >> 4))
#define TRU_FFT_PAD(n) ((n) + ((n) >> 3))
#define TRU_FFT_RE(z, a) (z)[2 * TRU_FFT_IDX(a)]
#define TRU_FFT_IM(z, a) (z)[2 * TRU_FFT_IDX(a) + 1]

namespace tru {

// multiply (x,y) by DIR*i
template <int DIR>
TRU_HD void mul_i(float& x, float& y) {
  float t = x;
  if (DIR > 0) { x = -y; y = t; } else { x = y; y = -t; }
}

// 4-point DFT, kernel exp(DIR*2*pi*i*nk/4)
template <int DIR>
TRU_HD void dft4(float& x0r, float& x0i, float& x1r, float& x1i,
                 float& x2r, float& x2i, float& x3r, float& x3i) {
  float s0r = x0r + x2r, s0i = x0i + x2i;
  float s1r = x0r - x2r, s1i = x0i - x2i;
  float s2r = x1r + x3r, s2i = x1i + x3i;
  float s3r = x1r - x3r, s3i = x1i - x3i;
  mul_i<DIR>(s3r, s3i);
  x0r = s0r + s2r; x0i = s0i + s2i;
  x2r = s0r - s2r; x2i = s0i - s2i;
  x1r = s1r + s3r; x1i = s1i + s3i;
  x3r = s1r - s3r; x3i = s1i - s3i;
}

// 8-point DFT in place, natural order in and out.
template <int DIR>
TRU_HD void dft8(float (&r)[8], float (&i)[8]) {
  const float h = 0.70710678118654752440f;
  // decimation in frequency: b = a[n]+a[n+4] (even outputs), d = W^n (a[n]-a[n+4]) (odd outputs)
  float br[4], bi[4], dr[4], di[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    br[n] = r[n] + r[n + 4]; bi[n] = i[n] + i[n + 4];
    dr[n] = r[n] - r[n + 4]; di[n] = i[n] - i[n + 4];
  }
  // W^1 = (h, DIR*h), W^2 = DIR*i, W^3 = (-h, DIR*h)
  {
    float tr = dr[1], ti = di[1];
    dr[1] = h * (tr - DIR * ti); di[1] = h * (ti + DIR * tr);
    mul_i<DIR>(dr[2], di[2]);
    tr = dr[3]; ti = di[3];
    dr[3] = h * (-tr - DIR * ti); di[3] = h * (-ti + DIR * tr);
  }
  dft4<DIR>(br[0], bi[0], br[1], bi[1], br[2], bi[2], br[3], bi[3]);
  dft4<DIR>(dr[0], di[0], dr[1], di[1], dr[2], di[2], dr[3], di[3]);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    r[2 * q] = br[q]; i[2 * q] = bi[q];
    r[2 * q + 1] = dr[q]; i[2 * q + 1] = di[q];
  }
}

// One radix-R butterfly (index i in [0, N/R)) of the Stockham pass with current
// sub-transform length p.  tw[k] = exp(-2*pi*i*k/N), k < N.
template <int N, int R, int DIR>
TRU_HD void fft_butterfly_load(const float* z, const float2* tw,
                               int p, int i, float (&ur)[R], float (&ui)[R]) {
  const int k = i & (p - 1);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int a = i + r * (N / R);
    const float2 v = *(const float2*)(z + 2 * TRU_FFT_IDX(a));
    ur[r] = v.x;
    ui[r] = v.y;
  }
  if (p > 1) {
    const int step = k * (N / (p * R));
#pragma unroll
    for (int r = 1; r < R; ++r) {
      const float2 w = tw[r * step];
      const float wr = w.x, wi = (DIR < 0) ? w.y : -w.y;
      const float xr = ur[r], xi = ui[r];
      ur[r] = xr * wr - xi * wi;
      ui[r] = xr * wi + xi * wr;
    }
  }
}

template <int N, int R>
TRU_HD void fft_butterfly_store(float* z, int p, int i,
                                const float (&ur)[R], const float (&ui)[R]) {
  const int k = i & (p - 1);
  const int j0 = (i - k) * R + k;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int a = j0 + r * p;
    float2 v;
    v.x = ur[r]; v.y = ui[r];
    *(float2*)(z + 2 * TRU_FFT_IDX(a)) = v;
  }
}

// radix schedule: 512 = 8*8*8, 1024 = 8*8*4*4, 2048 = 8*8*8*4
template <int N> struct FftPlan;
template <> struct FftPlan<512>  { static constexpr int n8 = 3, n4 = 0; };
template <> struct FftPlan<1024> { static constexpr int n8 = 2, n4 = 2; };
template <> struct FftPlan<2048> { static constexpr int n8 = 3, n4 = 1; };

#if defined(__CUDACC__)
// Cooperative FFT by N/8 threads (tid in [0, N/8)); every thread of the CTA must
// call it (it contains __syncthreads).  Data must be in place before the call
// (a __syncthreads is issued on entry); result is visible to all on return.
template <int N, int DIR>
__device__ __forceinline__ void fft_smem(float* z, const float2* tw, int tid) {
  int p = 1;
  __syncthreads();
#pragma unroll
  for (int s = 0; s < FftPlan<N>::n8; ++s) {
    float ur[8], ui[8];
    fft_butterfly_load<N, 8, DIR>(z, tw, p, tid, ur, ui);
    dft8<DIR>(ur, ui);
    __syncthreads();
    fft_butterfly_store<N, 8>(z, p, tid, ur, ui);
    __syncthreads();
    p *= 8;
  }
#pragma unroll
  for (int s = 0; s < FftPlan<N>::n4; ++s) {
    float ur[2][4], ui[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      fft_butterfly_load<N, 4, DIR>(z, tw, p, tid + h * (N / 8), ur[h], ui[h]);
      dft4<DIR>(ur[h][0], ui[h][0], ur[h][1], ui[h][1], ur[h][2], ui[h][2], ur[h][3], ui[h][3]);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h)
      fft_butterfly_store<N, 4>(z, p, tid + h * (N / 8), ur[h], ui[h]);
    __syncthreads();
    p *= 4;
  }
}
#endif

}  // namespace tru
