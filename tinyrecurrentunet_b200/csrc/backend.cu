// Fused back end: mod_phase (de-norm, dB->amplitude, atan2) + phase-aware
// beta-sigmoid mask + complex multiply + inverse rFFT-512 (two frames per complex
// transform) + rectangular overlap-add / envelope division, and its backward.
//
// Replaces dataset.py:182-203 (mod_phase), phm.py:31-45 (PhaseAwareMask),
// util.py:221-234 (call site) and dataset.py:293-296 (torch.istft
// with no window => rectangular, centre => trim n_fft/2 each side).
#include "tru_common.cuh"
#include "tru_fft.cuh"

namespace tru {
namespace {

constexpr int NFFT = TRU_NFFT, HOP = TRU_HOP, NB = TRU_NBINS;
constexpr int NT = 256;
constexpr int FPAD = TRU_FFT_PAD(NFFT);
constexpr int JC = 29;            // output hop-blocks per CTA in the forward
constexpr int NFR = JC + 3;       // frames touched by those blocks (32)
constexpr int TCB = 32;           // frames per CTA in the backward

struct BackParams {
  const float* net; const float* gaudio; float* audio; float* gnet; const float2* tw;
  int B, T, C, nchunks;
  int ch_m, ch_s, ch_c, ch_s1, ch_c1, use_mask;
  float beta;
};

// 10^(de_norm(m)/20), de_norm = (clamp(m,-1,1)+1)/2*100 - 100 + 25  (dataset.py:214-218,238-243)
__device__ __forceinline__ float amp_from_norm(float m) {
  const float mc = fminf(fmaxf(m, -1.0f), 1.0f);
  const float db = ((mc + 1.0f) / 2.0f) * 100.0f + (-100.0f) + 25.0f;
  return exp10f(db / 20.0f);
}

// Spectrum value of one bin: Y = mask * A0 * (cos th0, sin th0)
__device__ __forceinline__ void bin_spectrum_vals(const BackParams& p, float m, float s0, float c0, float s1, float c1,
                                                  float& yr, float& yi) {
  float R = amp_from_norm(m);
  const float h = sqrtf(s0 * s0 + c0 * c0);
  float u = 1.0f, v = 0.0f;                       // atan2(0,0) = 0
  if (h > 0.0f) { u = c0 / h; v = s0 / h; }
  if (p.use_mask) {
    const float d = atan2f(s0, c0) - atan2f(s1, c1);          // phm.py:41, not wrapped
    R *= 1.0f / (1.0f + expf(-p.beta * d));
  }
  yr = R * u; yi = R * v;
}
// the five values of bin k of one frame the spectrum needs: m, s0, c0, s1, c1 (independent loads: issue them all, then compute)
__device__ __forceinline__ void bin_load(const BackParams& p, const float* fr, int k, float (&v)[5]) {
  v[0] = __ldg(fr + p.ch_m * NB + k); v[1] = __ldg(fr + p.ch_s * NB + k); v[2] = __ldg(fr + p.ch_c * NB + k);
  v[3] = p.use_mask ? __ldg(fr + p.ch_s1 * NB + k) : 0.f; v[4] = p.use_mask ? __ldg(fr + p.ch_c1 * NB + k) : 1.f;
}
__device__ __forceinline__ void bin_spectrum(const BackParams& p, const float* fr, int k, float& yr, float& yi) {
  float v[5];
  bin_load(p, fr, k, v);
  bin_spectrum_vals(p, v[0], v[1], v[2], v[3], v[4], yr, yi);
}

__device__ __forceinline__ int frames_covering(int pp, int T) {
  // number of frames t in [0,T) with 128 t <= pp < 128 t + 512
  const int hi = min(pp / HOP, T - 1);
  const int lo = pp >= NFFT ? (pp - NFFT) / HOP + 1 : 0;
  return hi - lo + 1;
}

__global__ void __launch_bounds__(NT) backend_fwd_kernel(BackParams p) {
  extern __shared__ __align__(16) float smem[];
  float* td = smem;                                  // [NFR][512] time-domain frames
  float* fre = td + NFR * NFFT;
  float* fim = fre + 4 * FPAD;
  float2* tw = (float2*)(fim + 4 * FPAD);
  const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
  const int b = blockIdx.x / p.nchunks, c = blockIdx.x % p.nchunks;
  const int Q = p.T - 1;                             // output blocks of 128 samples
  const int q0 = c * JC;
  const int nq = min(JC, Q - q0);
  const int tfirst = q0 - 1;                         // first frame touching block q0
  const int nfr = nq + 3;
  for (int k = tid; k < NFFT; k += NT) tw[k] = p.tw[k * (2048 / NFFT)];
  float* z = fre + g * (2 * FPAD);                      // interleaved (re, im) pairs, tru_fft.cuh
  const float* net = p.net + (size_t)b * p.T * p.C * NB;
  const float invn = 1.0f / NFFT;

  for (int round = 0; round * 8 < nfr; ++round) {
    const int la = round * 8 + g * 2, lb = la + 1;   // local frame slots
    const int ta = tfirst + la, tb = tfirst + lb;
    const bool va = la < nfr && ta >= 0 && ta < p.T;
    const bool vb = lb < nfr && tb >= 0 && tb < p.T;
    __syncthreads();
    // all 40 loads of the thread's 4 bins x 2 frames first (memory-level parallelism: ncu showed "long scoreboard" as the top
    // stall when every bin loaded and then computed its atan2 / exp chain), then the arithmetic; bin 256 is thread 0's extra
    auto put_bin = [&](int k, float ar, float ai, float br, float bi) {
      if (k == 0 || k == NFFT / 2) {                 // c2r ignores Im of DC / Nyquist
        TRU_FFT_RE(z, k) = ar; TRU_FFT_IM(z, k) = br;
      } else {
        TRU_FFT_RE(z, k) = ar - bi; TRU_FFT_IM(z, k) = ai + br;
        TRU_FFT_RE(z, NFFT - k) = ar + bi; TRU_FFT_IM(z, NFFT - k) = br - ai;
      }
    };
    const float* fa = net + (size_t)ta * p.C * NB;
    const float* fb = net + (size_t)tb * p.C * NB;
    float xa[4][5], xb[4][5];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (va) bin_load(p, fa, l + 64 * i, xa[i]);
      if (vb) bin_load(p, fb, l + 64 * i, xb[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
      if (va) bin_spectrum_vals(p, xa[i][0], xa[i][1], xa[i][2], xa[i][3], xa[i][4], ar, ai);
      if (vb) bin_spectrum_vals(p, xb[i][0], xb[i][1], xb[i][2], xb[i][3], xb[i][4], br, bi);
      put_bin(l + 64 * i, ar, ai, br, bi);
    }
    if (l == 0) {
      float ar = 0.f, ai = 0.f, br = 0.f, bi = 0.f;
      if (va) bin_spectrum(p, fa, NFFT / 2, ar, ai);
      if (vb) bin_spectrum(p, fb, NFFT / 2, br, bi);
      put_bin(NFFT / 2, ar, ai, br, bi);
    }
    fft_smem<NFFT, 1>(z, tw, l);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = l + 64 * j;
      if (la < nfr) td[la * NFFT + n] = va ? TRU_FFT_RE(z, n) * invn : 0.0f;
      if (lb < nfr) td[lb * NFFT + n] = vb ? TRU_FFT_IM(z, n) * invn : 0.0f;
    }
  }
  __syncthreads();

  // overlap-add in fixed order (deterministic), divide by the rectangular envelope
  float* out = p.audio + (size_t)b * Q * HOP + (size_t)q0 * HOP;
  for (int i = tid; i < nq * HOP; i += NT) {
    const int pp = (q0 + 2) * HOP + i;               // position in the padded signal
    const int q = i / HOP, n = i % HOP;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j)                      // local slots q..q+3 hold frames q0+q-1 .. q0+q+2
      acc += td[(q + j) * NFFT + (3 - j) * HOP + n];
    out[i] = acc / (float)frames_covering(pp, p.T);
  }
}

__global__ void __launch_bounds__(NT) backend_bwd_kernel(BackParams p) {
  extern __shared__ __align__(16) float smem[];
  float* fre = smem;
  float* fim = fre + 4 * FPAD;
  float2* tw = (float2*)(fim + 4 * FPAD);
  const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
  const int b = blockIdx.x / p.nchunks, c = blockIdx.x % p.nchunks;
  const int t0 = c * TCB;
  const int nfr = min(TCB, p.T - t0);
  const int NOUT = (p.T - 1) * HOP;
  for (int k = tid; k < NFFT; k += NT) tw[k] = p.tw[k * (2048 / NFFT)];
  float* z = fre + g * (2 * FPAD);                      // interleaved (re, im) pairs, tru_fft.cuh
  const float* ga = p.gaudio + (size_t)b * NOUT;
  const float ln10_20x50 = 2.302585092994046f / 20.0f * 50.0f;

  for (int round = 0; round * 8 < nfr; ++round) {
    const int la = round * 8 + g * 2, lb = la + 1;
    const int ta = t0 + la, tb = t0 + lb;
    const bool va = la < nfr, vb = lb < nfr;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = l + 64 * j;
      float xa = 0.f, xb = 0.f;
      if (va) {
        const int pp = ta * HOP + n, o = pp - NFFT / 2;
        if (o >= 0 && o < NOUT) xa = __ldg(ga + o) / (float)frames_covering(pp, p.T);
      }
      if (vb) {
        const int pp = tb * HOP + n, o = pp - NFFT / 2;
        if (o >= 0 && o < NOUT) xb = __ldg(ga + o) / (float)frames_covering(pp, p.T);
      }
      TRU_FFT_RE(z, n) = xa; TRU_FFT_IM(z, n) = xb;
    }
    fft_smem<NFFT, -1>(z, tw, l);
    // (loading all 40 values of the thread's bins ahead of the chain rule, as the forward kernel does, was measured slower here:
    // 0.15 -> 0.19 ms at 121 registers - the kernel writes 8 channels per bin and lives on occupancy)
    // gradient of bin k of frame t (h = 0: the frame in the real part of the packed transform, 1: imaginary part)
    auto do_bin = [&](int k, int h, int t, const float (&x)[5]) {
      const int kn = (NFFT - k) & (NFFT - 1);
      const float zr = TRU_FFT_RE(z, k), zi = TRU_FFT_IM(z, k);
      const float wr = TRU_FFT_RE(z, kn), wi = TRU_FFT_IM(z, kn);
      const bool edge = (k == 0 || k == NFFT / 2);
      const float ck = (edge ? 1.0f : 2.0f) / NFFT;
      // G = FFT(g)[k]; dL/dRe Y = ck Re G, dL/dIm Y = ck Im G (0 at DC / Nyquist)
      const float dRe = ck * (h == 0 ? 0.5f * (zr + wr) : 0.5f * (zi + wi));
      const float dIm = edge ? 0.0f : ck * (h == 0 ? 0.5f * (zi - wi) : -0.5f * (zr - wr));
      float* gr = p.gnet + ((size_t)b * p.T + t) * p.C * NB;
      const float m = x[0], s0 = x[1], c0 = x[2], s1 = x[3], c1 = x[4];
      const float A0 = amp_from_norm(m);
      const float h2 = s0 * s0 + c0 * c0, hh = sqrtf(h2);
      float u = 1.0f, v = 0.0f;
      if (hh > 0.0f) { u = c0 / hh; v = s0 / hh; }
      float mask = 1.0f, dmask_dth = 0.0f;
      if (p.use_mask) {
        const float d = atan2f(s0, c0) - atan2f(s1, c1);
        mask = 1.0f / (1.0f + expf(-p.beta * d));
        dmask_dth = p.beta * mask * (1.0f - mask);
      }
      const float R = mask * A0;
      const float dR = dRe * u + dIm * v;
      const float dth_dir = R * (dIm * u - dRe * v);
      const float dmask = dR * A0;
      const float dA0 = dR * mask;
      const float dth0 = dth_dir + dmask * dmask_dth;
      const float dth1 = -dmask * dmask_dth;
      const float dm = (m >= -1.0f && m <= 1.0f) ? dA0 * A0 * ln10_20x50 : 0.0f;
      // theta = atan2(s, c): d/ds = c/(s^2+c^2), d/dc = -s/(s^2+c^2)
      const float i0 = h2 > 0.0f ? 1.0f / h2 : 0.0f;
      for (int ch = 0; ch < p.C; ++ch) {
        float val = 0.0f;
        if (ch == p.ch_m) val = dm;
        else if (ch == p.ch_s) val = dth0 * c0 * i0;
        else if (ch == p.ch_c) val = -dth0 * s0 * i0;
        else if (p.use_mask && (ch == p.ch_s1 || ch == p.ch_c1)) {
          const float h1 = s1 * s1 + c1 * c1, i1 = h1 > 0.0f ? 1.0f / h1 : 0.0f;
          val = (ch == p.ch_s1) ? dth1 * c1 * i1 : -dth1 * s1 * i1;
        }
        gr[ch * NB + k] = val;
      }
    };
    for (int k = l; k <= NFFT / 2; k += 64) {
      float x[5];
      if (va) { bin_load(p, p.net + ((size_t)b * p.T + ta) * p.C * NB, k, x); do_bin(k, 0, ta, x); }
      if (vb) { bin_load(p, p.net + ((size_t)b * p.T + tb) * p.C * NB, k, x); do_bin(k, 1, tb, x); }
    }
  }
}

// Streaming step: one new network-output frame per stream.  ola (S,384) holds the overlap-add sums of
// the three hops that still wait for later frames; adding frame t completes output block t-2 (two hops of
// look-ahead, exactly the offline centre=True framing).  4 streams per CTA, ONE stream per complex transform (Hermitian
// extension of its own spectrum): streams are independent signals and must not share a transform (see frontend_step_kernel).
__global__ void __launch_bounds__(NT) backend_step_kernel(BackParams p, float* __restrict__ ola, int frame_index, int add_frame) {
  __shared__ __align__(16) float fre[8 * FPAD];
  __shared__ float2 tw[NFFT];
  const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
  for (int k = tid; k < NFFT; k += NT) tw[k] = p.tw[k * (2048 / NFFT)];
  float* z = fre + g * (2 * FPAD);                      // interleaved (re, im) pairs, tru_fft.cuh
  const int sidx = blockIdx.x * 4 + g;
  const bool valid = sidx < p.B;
  __syncthreads();
  if (add_frame) {
    for (int k = l; k <= NFFT / 2; k += 64) {
      float ar = 0.f, ai = 0.f;
      if (valid) bin_spectrum(p, p.net + (size_t)sidx * p.C * NB, k, ar, ai);
      if (k == 0 || k == NFFT / 2) {                 // c2r ignores Im of DC / Nyquist
        TRU_FFT_RE(z, k) = ar; TRU_FFT_IM(z, k) = 0.f;
      } else {
        TRU_FFT_RE(z, k) = ar; TRU_FFT_IM(z, k) = ai;
        TRU_FFT_RE(z, NFFT - k) = ar; TRU_FFT_IM(z, NFFT - k) = -ai;
      }
    }
    fft_smem<NFFT, 1>(z, tw, l);
  }
  // frames contributing to block q = frame_index - 2 are q-1 .. q+2 = frame_index-3 .. frame_index (those that exist)
  const int newest = add_frame ? frame_index : frame_index - 1;
  const int cnt = newest - max(frame_index - 3, 0) + 1;
  const float invn = 1.0f / NFFT, invc = (frame_index >= 2 && cnt > 0) ? 1.0f / (float)cnt : 0.0f;
  if (!valid) return;
  float* o = ola + (size_t)sidx * 384;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = l + 64 * j;
    const float x = add_frame ? TRU_FFT_RE(z, n) * invn : 0.0f;
    v[j] = x + (n < 384 ? o[n] : 0.0f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = l + 64 * j;
    if (n < HOP) p.audio[(size_t)sidx * HOP + n] = v[j] * invc;
    else o[n - HOP] = v[j];
  }
}

int fill(const TruBackendDesc* d, BackParams& p) {
  TRU_REQUIRE(d && d->batch > 0, TRU_ERR_ARG, "backend: need batch > 0");
  TRU_REQUIRE(d->n_channels >= 3 && d->n_channels <= 16, TRU_ERR_ARG, "backend: n_channels out of range");
  const int C = d->n_channels;
  TRU_REQUIRE(d->ch_mag >= 0 && d->ch_mag < C && d->ch_sin >= 0 && d->ch_sin < C && d->ch_cos >= 0 && d->ch_cos < C,
              TRU_ERR_ARG, "backend: bad channel index");
  if (d->use_mask)
    TRU_REQUIRE(d->ch_sin1 >= 0 && d->ch_sin1 < C && d->ch_cos1 >= 0 && d->ch_cos1 < C, TRU_ERR_ARG,
                "backend: bad set-1 channel index");
  p.B = d->batch; p.T = d->n_frames; p.C = C;
  p.ch_m = d->ch_mag; p.ch_s = d->ch_sin; p.ch_c = d->ch_cos; p.ch_s1 = d->ch_sin1; p.ch_c1 = d->ch_cos1;
  p.use_mask = d->use_mask; p.beta = (float)d->beta; p.tw = twiddle_table();
  return TRU_OK;
}

constexpr size_t FWD_SMEM = (size_t)(NFR * NFFT + 8 * FPAD) * 4 + NFFT * 8;
constexpr size_t BWD_SMEM = (size_t)(8 * FPAD) * 4 + NFFT * 8;

}  // namespace
}  // namespace tru

using namespace tru;

extern "C" int tru_backend_step(const TruBackendDesc* d, const float* net_out, float* ola_state, float* audio,
                                int frame_index, int add_frame, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  BackParams p{};
  if ((rc = fill(d, p))) return rc;
  TRU_REQUIRE(ola_state && audio && (net_out || !add_frame) && frame_index >= 0, TRU_ERR_ARG, "backend_step: null pointer / negative frame index");
  p.net = net_out; p.audio = audio;
  ProfScope prof("backend_step", 4.0 * p.B * ((double)p.C * NB + 2.0 * 384 + HOP), 0.5 * p.B * 5.0 * NFFT * 9, (cudaStream_t)stream);
  backend_step_kernel<<<(p.B + 3) / 4, NT, 0, (cudaStream_t)stream>>>(p, ola_state, frame_index, add_frame);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

extern "C" int tru_backend_fwd(const TruBackendDesc* d, const float* net_out, float* audio, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  BackParams p{};
  if ((rc = fill(d, p))) return rc;
  TRU_REQUIRE(d->n_frames >= 2, TRU_ERR_ARG, "backend: need >= 2 frames");
  TRU_REQUIRE(net_out && audio, TRU_ERR_ARG, "backend_fwd: null pointer");
  p.net = net_out; p.audio = audio;
  p.nchunks = (p.T - 1 + JC - 1) / JC;
  TRU_CUDA(cudaFuncSetAttribute(backend_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM));
  ProfScope prof("backend_fwd", 4.0 * p.B * ((double)p.C * NB * p.T + (double)HOP * (p.T - 1)), 0.5 * p.B * p.T * 5.0 * NFFT * 9,
                 (cudaStream_t)stream);
  backend_fwd_kernel<<<p.B * p.nchunks, NT, FWD_SMEM, (cudaStream_t)stream>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

extern "C" int tru_backend_bwd(const TruBackendDesc* d, const float* net_out, const float* grad_audio,
                               float* grad_net_out, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  BackParams p{};
  if ((rc = fill(d, p))) return rc;
  TRU_REQUIRE(d->n_frames >= 2, TRU_ERR_ARG, "backend: need >= 2 frames");
  TRU_REQUIRE(net_out && grad_audio && grad_net_out, TRU_ERR_ARG, "backend_bwd: null pointer");
  p.net = net_out; p.gaudio = grad_audio; p.gnet = grad_net_out;
  p.nchunks = (p.T + TCB - 1) / TCB;
  ProfScope prof("backend_bwd", 4.0 * p.B * (2.0 * p.C * NB * p.T + (double)HOP * (p.T - 1)), 0.5 * p.B * p.T * 5.0 * NFFT * 9,
                 (cudaStream_t)stream);
  backend_bwd_kernel<<<p.B * p.nchunks, NT, BWD_SMEM, (cudaStream_t)stream>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
