// Batch assembly in front of the hot path: dataset.py:79-126 (DataAugment: gain, low-pass biquad,
// high-pass biquad on the noise clip -- torchaudio.functional, each biquad clamped to [-1,1]) and dataset.py:352-386
// (random crop of the clean clip, noisy = clean + augmented noise), batched for B clips on the device.
//
// A biquad is y[n] = u[n] - a1 y[n-1] - a2 y[n-2] with u = b0 x[n] + b1 x[n-1] + b2 x[n-2]: sequential over 64,000 samples in
// torchaudio.  Here one CTA owns one clip and walks it in tiles of 256 x 63 samples held in shared memory (global traffic
// is one coalesced read and one coalesced write per sample, both filters run on the tile in place).  Inside a tile thread j
// owns 63 consecutive samples (63 is odd: the 32 lanes of a warp hit 32 different banks) and the recurrence is split by
// superposition, which is exact for a linear recurrence:
//   pass A   every thread runs its chunk from a ZERO state and publishes the end state e_j;
//   carry    s_{j+1} = A^63 s_j + e_j over the 256 chunks: warp 0, 8 chunks per lane + a 5-step scan across lanes with
//            A^(63*8*2^k) (all powers come from the host, evaluated in fp64, in the coefficient row);
//   pass B   every thread reruns its chunk from its true start state s_j and stores clamp(y).
// The clamp is applied to the stored output only, the recurrence continues on the unclamped values (torchaudio clamps
// after the whole filter has run).
#include <cuda_runtime.h>

#include "../../include/tru_b200.h"
#include "tru_common.cuh"

namespace tru {
namespace {

constexpr int ANT = 256;                 // threads = chunks per tile
constexpr int AK = TRU_AUGMENT_CHUNK;    // samples per chunk
constexpr int ATS = ANT * AK;            // samples per tile (16,128 = 63 KB of shared memory)
constexpr int ACO = TRU_AUGMENT_NCOEF;   // floats per coefficient row
static_assert(ANT == 256 && ACO == 59, "the carry scan is written for 32 lanes x 8 chunks and 6 matrix powers per filter");
static_assert(AK % 2 == 1 && AK >= 3, "an odd chunk keeps the strided shared-memory walk conflict-free");

struct Biquad {
  float b0, b1, b2, a1, a2;              // already divided by a0 (in fp32, as torchaudio does)
  float m00, m01, m10, m11;              // A^AK, A = [[-a1, -a2], [1, 0]] acting on (y[n-1], y[n-2])
  const float* pw;                       // the row's five further powers A^(AK * {8, 16, 32, 64, 128}), row major (read by the scan)
};

struct StageCarry {                      // lives in thread 0's registers across tiles
  float x1, x2;                          // last two inputs of the previous tile
  float y1, y2;                          // last two (unclamped) outputs of the previous tile
};

// One biquad over the nv valid samples of the tile, in place.  es/ss: shared [ANT][2] end states / start states.
__device__ __forceinline__ void biquad_tile(float* tile, int nv, const Biquad& c, StageCarry& carry, float2* es, float2* ss) {
  const int j = threadIdx.x;
  const int base = j * AK;
  int cnt = nv - base;
  cnt = cnt < 0 ? 0 : (cnt > AK ? AK : cnt);
  float xm1 = 0.f, xm2 = 0.f;
  if (j == 0) { xm1 = carry.x1; xm2 = carry.x2; }
  else if (cnt > 0) { xm1 = tile[base - 1]; xm2 = tile[base - 2]; }
  float nx1 = 0.f, nx2 = 0.f;
  if (j == 0 && nv == ATS) { nx1 = tile[ATS - 1]; nx2 = tile[ATS - 2]; }     // next tile's FIR history, before pass B overwrites it
  // pass A: zero-state response of this chunk
  {
    float y1 = 0.f, y2 = 0.f, a = xm1, b = xm2;
    for (int i = 0; i < cnt; ++i) {
      const float x = tile[base + i];
      const float u = fmaf(c.b2, b, fmaf(c.b1, a, c.b0 * x));
      const float y = fmaf(-c.a1, y1, fmaf(-c.a2, y2, u));
      y2 = y1; y1 = y; b = a; a = x;
    }
    es[j] = make_float2(y1, y2);
  }
  __syncthreads();
  // carry over the 256 chunk ends by warp 0: lane L folds its 8 chunks (Horner), a 5-step scan across the lanes with the
  // precomputed powers A^(63*8*2^k) composes them, then the lane replays its 8 chunks from its true start state.
  if (j < 32) {
    const float m00 = c.m00, m01 = c.m01, m10 = c.m10, m11 = c.m11;
    const float in1 = __shfl_sync(0xffffffffu, carry.y1, 0), in2 = __shfl_sync(0xffffffffu, carry.y2, 0);
    float2 e[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = es[j * 8 + i];
    float p1 = j == 0 ? in1 : 0.f, p2 = j == 0 ? in2 : 0.f;           // lane 0 starts from the tile's incoming state
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t1 = fmaf(m00, p1, fmaf(m01, p2, e[i].x)), t2 = fmaf(m10, p1, fmaf(m11, p2, e[i].y));
      p1 = t1; p2 = t2;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {                                      // inclusive scan: P_L = A^(63*8*d) P_{L-d} + P_L
      const int d = 1 << k;
      const float q1 = __shfl_up_sync(0xffffffffu, p1, d), q2 = __shfl_up_sync(0xffffffffu, p2, d);
      const float w0 = __ldg(c.pw + 4 * k), w1 = __ldg(c.pw + 4 * k + 1), w2 = __ldg(c.pw + 4 * k + 2), w3 = __ldg(c.pw + 4 * k + 3);
      if (j >= d) {
        const float t1 = fmaf(w0, q1, fmaf(w1, q2, p1)), t2 = fmaf(w2, q1, fmaf(w3, q2, p2));
        p1 = t1; p2 = t2;
      }
    }
    float s1 = __shfl_up_sync(0xffffffffu, p1, 1), s2 = __shfl_up_sync(0xffffffffu, p2, 1);   // state after the previous lane's chunks
    if (j == 0) { s1 = in1; s2 = in2; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ss[j * 8 + i] = make_float2(s1, s2);
      const float t1 = fmaf(m00, s1, fmaf(m01, s2, e[i].x)), t2 = fmaf(m10, s1, fmaf(m11, s2, e[i].y));
      s1 = t1; s2 = t2;
    }
    const float o1 = __shfl_sync(0xffffffffu, p1, 31), o2 = __shfl_sync(0xffffffffu, p2, 31);
    if (j == 0) {
      carry.y1 = o1; carry.y2 = o2;        // state after the tile's last chunk (meaningful when the tile is full)
      carry.x1 = nx1; carry.x2 = nx2;
    }
  }
  __syncthreads();
  // pass B: the same chunk from its true start state
  if (cnt > 0) {
    const float2 s = ss[j];
    float y1 = s.x, y2 = s.y, a = xm1, b = xm2;
    for (int i = 0; i < cnt; ++i) {
      const float x = tile[base + i];
      const float u = fmaf(c.b2, b, fmaf(c.b1, a, c.b0 * x));
      const float y = fmaf(-c.a1, y1, fmaf(-c.a2, y2, u));
      tile[base + i] = fminf(fmaxf(y, -1.f), 1.f);
      y2 = y1; y1 = y; b = a; a = x;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ Biquad load_biquad(const float* p) {
  Biquad c;
  c.b0 = p[0]; c.b1 = p[1]; c.b2 = p[2]; c.a1 = p[3]; c.a2 = p[4];
  c.m00 = p[5]; c.m01 = p[6]; c.m10 = p[7]; c.m11 = p[8];
  c.pw = p + 9;
  return c;
}

// noise (B, n) -> out (B, n); coef (B, ACO) = {gain ratio, low-pass row[29], high-pass row[29]}
__global__ void __launch_bounds__(ANT) augment_kernel(const float* __restrict__ noise, const float* __restrict__ coef,
                                                      float* __restrict__ out, int n) {
  extern __shared__ float tile[];
  __shared__ float2 es[ANT], ss[ANT];
  const int b = blockIdx.x;
  const float* src = noise + (size_t)b * n;
  float* dst = out + (size_t)b * n;
  const float* cf = coef + (size_t)b * ACO;
  const float gain = cf[0];
  const Biquad lp = load_biquad(cf + 1), hp = load_biquad(cf + 1 + (ACO - 1) / 2);
  StageCarry c1{0.f, 0.f, 0.f, 0.f}, c2{0.f, 0.f, 0.f, 0.f};
  for (int t0 = 0; t0 < n; t0 += ATS) {
    const int nv = min(ATS, n - t0);
    for (int i = threadIdx.x; i < nv; i += ANT) tile[i] = src[t0 + i] * gain;          // F.gain: waveform * ratio
    __syncthreads();
    biquad_tile(tile, nv, lp, c1, es, ss);                                               // F.lowpass_biquad (+ clamp)
    biquad_tile(tile, nv, hp, c2, es, ss);                                               // F.highpass_biquad (+ clamp)
    for (int i = threadIdx.x; i < nv; i += ANT) dst[t0 + i] = tile[i];
    __syncthreads();
  }
}

// clean_out[b, i] = clean[b, cs_b + i];  noisy_out[b, i] = clean_out[b, i] + aug[b, (ns_b + i) mod n_noise]
__global__ void __launch_bounds__(256) mix_crop_kernel(const float* __restrict__ clean, const float* __restrict__ aug,
                                                       const int* __restrict__ cstart, const int* __restrict__ nstart,
                                                       float* __restrict__ clean_out, float* __restrict__ noisy_out,
                                                       int n_clean, int n_noise, int n_out) {
  const int b = blockIdx.y;
  int cs = cstart ? cstart[b] : 0, ns = nstart ? nstart[b] : 0;
  cs = max(0, min(cs, n_clean - n_out));            // the host wrapper validates the offsets; this only keeps a bad one in bounds
  ns = ((ns % n_noise) + n_noise) % n_noise;
  const float* c = clean + (size_t)b * n_clean + cs;
  const float* a = aug + (size_t)b * n_noise;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
    const float v = c[i];
    int k = ns + i;
    if (k >= n_noise) k %= n_noise;
    clean_out[(size_t)b * n_out + i] = v;
    noisy_out[(size_t)b * n_out + i] = v + a[k];
  }
}

}  // namespace
}  // namespace tru

using namespace tru;

extern "C" int tru_augment_fwd(int batch, int n_samples, const float* noise, const float* coef, float* out, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  TRU_REQUIRE(batch > 0 && n_samples > 0, TRU_ERR_ARG, "augment: batch %d, n_samples %d", batch, n_samples);
  TRU_REQUIRE(noise && coef && out, TRU_ERR_ARG, "augment: null pointer");
  const size_t smem = (size_t)ATS * sizeof(float);
  TRU_CUDA(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof("augment", 8.0 * batch * n_samples, 4.0 * 10.0 * batch * n_samples, (cudaStream_t)stream);
  augment_kernel<<<batch, ANT, smem, (cudaStream_t)stream>>>(noise, coef, out, n_samples);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

extern "C" int tru_mix_crop(int batch, int n_clean, int n_noise, int n_out, const float* clean, const float* aug_noise,
                            const int* clean_start, const int* noise_start, float* clean_out, float* noisy_out, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  TRU_REQUIRE(batch > 0 && n_clean > 0 && n_noise > 0 && n_out > 0 && n_out <= n_clean, TRU_ERR_ARG,
              "mix_crop: batch %d, n_clean %d, n_noise %d, n_out %d", batch, n_clean, n_noise, n_out);
  TRU_REQUIRE(clean && aug_noise && clean_out && noisy_out, TRU_ERR_ARG, "mix_crop: null pointer");
  int gx = (n_out + 255) / 256;
  const int cap = (4 * sm_count() + batch - 1) / batch;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  ProfScope prof("mix_crop", 16.0 * batch * n_out, 1.0 * batch * n_out, (cudaStream_t)stream);
  mix_crop_kernel<<<dim3(gx, batch), 256, 0, (cudaStream_t)stream>>>(clean, aug_noise, clean_start, noise_start, clean_out,
                                                                      noisy_out, n_clean, n_noise, n_out);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
