// Fused front end: framing + rectangular-window STFT-512 (shared-memory radix-8
// FFT, two real frames per complex transform) + log-magnitude + PCEN (per-bin
// IIR smoother as a chunked time-axis scan with decoupled look-back) +
// sin/cos phase, one pass, feature tile staged in shared memory and written
// with 16-byte stores.
//
// Replaces dataset.py:246-272 (ProcessAudio.forward: torch.stft, abs, angle,
// amp_to_db, norm, sin, cos, permute, cat) and dataset.py:56-76 (pcenfunc).
// Output layout (B, T', 4, 257): ch0 log-mag, ch1 PCEN, ch2 sin, ch3 cos.
#include "tru_common.cuh"
#include "tru_fft.cuh"

namespace tru {
namespace {

constexpr int NFFT = TRU_NFFT, HOP = TRU_HOP, NB = TRU_NBINS;
constexpr int TC = 16;                    // frames per CTA
constexpr int NT = 256;                   // threads per CTA (4 FFT groups of 64)
constexpr int FEAT = 4 * NB;              // floats per frame (1028, 16-byte multiple)
constexpr int STAGE = (TC - 1) * HOP + NFFT;
constexpr int FPAD = TRU_FFT_PAD(NFFT);

struct FrontParams {
  const float* audio; const float* state_in; float* feats; float* state_out;
  float* agg; int* flags; int* counter; const float2* tw;
  int B, N, T, nchunks;
  float eps, s, oms, alpha, delta, r, delta_r, decay_chunk;
};

__device__ __forceinline__ void bin_features(float re, float im, float& lm, float& mag,
                                             float& sn, float& cs) {
  mag = sqrtf(re * re + im * im);
  const float db = 20.0f * log10f(fmaxf(mag, 1e-7f)) - 25.0f;          // dataset.py:207-211
  lm = fminf(fmaxf(((db + 100.0f) / 100.0f) * 2.0f - 1.0f, -1.0f), 1.0f);  // :229-235
  if (mag > 0.0f) { const float inv = 1.0f / mag; sn = im * inv; cs = re * inv; }
  else { sn = 0.0f; cs = 1.0f; }                                         // angle(0) = 0
}

__device__ __forceinline__ float pcen_out(float x, float M, const FrontParams& p) {
  // dataset.py:73: (x / (M + eps)^alpha + delta)^r - delta^r.  The powers as exp2f(y log2f(x)): 2 + 2 special-function ops instead
  // of two full powf calls (the error of log2f is ~1 ulp, amplified by |alpha log2(M + eps)| <= 20: ~5e-6 relative, well inside the
  // 1e-4 feature tolerance); r = 0.5 (the reference's default) is a square root.
  const float d = x * exp2f(-p.alpha * log2f(M + p.eps)) + p.delta;
  return (p.r == 0.5f ? sqrtf(d) : exp2f(p.r * log2f(d))) - p.delta_r;
}

__global__ void __launch_bounds__(NT) frontend_kernel(FrontParams p) {
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                               // [TC][4][257]
  float* stage = tile + TC * FEAT;                  // [STAGE]
  float* fre = stage + STAGE;                       // [4][FPAD]
  float* fim = fre + 4 * FPAD;                      // [4][FPAD]
  float2* tw = (float2*)(fim + 4 * FPAD);           // [512]
  __shared__ int s_ticket;

  const int tid = threadIdx.x;
  if (tid == 0) s_ticket = atomicAdd(p.counter, 1);   // ticket order => look-back cannot deadlock
  for (int k = tid; k < NFFT; k += NT) tw[k] = p.tw[k * (2048 / NFFT)];
  __syncthreads();
  const int ticket = s_ticket;
  const int b = ticket / p.nchunks, c = ticket % p.nchunks;
  const int t0 = c * TC;
  const int nfr = min(TC, p.T - t0);
  const float* x = p.audio + (size_t)b * p.N;

  // ---- stage the audio span of this chunk (reflect padding, dataset.py:260) ----
  const int span = (nfr - 1) * HOP + NFFT;
  for (int i = tid; i < span; i += NT)
    stage[i] = __ldg(x + reflect_idx(t0 * HOP + i - NFFT / 2, p.N));

  // ---- STFT: two frames per complex FFT, 4 groups x 64 threads ----------------
  const int g = tid >> 6, l = tid & 63;
  float* z = fre + g * (2 * FPAD);                      // interleaved (re, im) pairs, tru_fft.cuh
  const int npairs = (nfr + 1) >> 1;
  for (int round = 0; round * 4 < npairs; ++round) {
    const int pair = round * 4 + g;
    const int ta = 2 * pair, tb = ta + 1;
    __syncthreads();                                 // previous round's readers are done
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = l + 64 * j;
      TRU_FFT_RE(z, n) = (ta < nfr) ? stage[ta * HOP + n] : 0.0f;
      TRU_FFT_IM(z, n) = (tb < nfr) ? stage[tb * HOP + n] : 0.0f;
    }
    fft_smem<NFFT, -1>(z, tw, l);
    if (ta < nfr) {
      for (int k = l; k <= NFFT / 2; k += 64) {
        const int kn = (NFFT - k) & (NFFT - 1);
        const float zr = TRU_FFT_RE(z, k), zi = TRU_FFT_IM(z, k);
        const float wr = TRU_FFT_RE(z, kn), wi = TRU_FFT_IM(z, kn);
        float lm, mg, sn, cs;
        bin_features(0.5f * (zr + wr), 0.5f * (zi - wi), lm, mg, sn, cs);
        float* o = tile + ta * FEAT + k;
        o[0] = lm; o[NB] = mg; o[2 * NB] = sn; o[3 * NB] = cs;
        if (tb < nfr) {
          bin_features(0.5f * (zi + wi), -0.5f * (zr - wr), lm, mg, sn, cs);
          o += FEAT;
          o[0] = lm; o[NB] = mg; o[2 * NB] = sn; o[3 * NB] = cs;
        }
      }
    }
  }
  __syncthreads();

  // ---- PCEN smoother: local scan -> publish aggregate -> look-back -> finalise --
  float* agg = p.agg + ((size_t)b * p.nchunks + c) * NB;
  for (int f = tid; f < NB; f += NT) {
    float M = 0.0f;
    for (int t = 0; t < nfr; ++t) M = p.oms * M + p.s * tile[t * FEAT + NB + f];
    __stcg(agg + f, M);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) st_release(p.flags + b * p.nchunks + c, 1);

  // Wait for every predecessor's aggregate with one flag per thread (in parallel: a chain of c dependent acquire loads per
  // thread cost ~1400 cycles per predecessor, 28 us per CTA on 10-s clips), then fold the aggregates with independent loads.
  for (int cc = tid; cc < c; cc += NT)
    while (ld_acquire(p.flags + b * p.nchunks + cc) == 0) { }
  __syncthreads();
  // The smoother is a serial chain per bin (16 FMAs per chunk); the compression (two powers per element) is not: the chain writes
  // M_t into the FFT scratch planes, which are free by now, and ALL threads then compress the chunk's nfr x 257 elements in
  // parallel (the previous version ran both inside the per-bin loop: 257 threads, one of them with two bins, each with 16
  // dependent pow pairs - ncu showed the kernel waiting at barriers behind those chains).
  float* mbuf = fre;                                   // [TC][NB] (4112 floats <= 8 FPAD)
  for (int f = tid; f < NB; f += NT) {
    float M = p.state_in ? __ldg(p.state_in + (size_t)b * NB + f) : 0.0f;
    const float* pa = p.agg + (size_t)b * p.nchunks * NB + f;
#pragma unroll 8
    for (int cc = 0; cc < c; ++cc) M = p.decay_chunk * M + __ldcg(pa + (size_t)cc * NB);
    for (int t = 0; t < nfr; ++t) {
      M = p.oms * M + p.s * tile[t * FEAT + NB + f];  // dataset.py:66-68
      mbuf[t * NB + f] = M;
    }
    if (p.state_out && c == p.nchunks - 1) p.state_out[(size_t)b * NB + f] = M;
  }
  __syncthreads();
  for (int i = tid; i < nfr * NB; i += NT) {
    const int t = i / NB, f = i - t * NB;
    float* o = tile + t * FEAT + NB + f;
    *o = pcen_out(*o, mbuf[i], p);
  }
  __syncthreads();

  // ---- one contiguous, 16-byte aligned span of nfr*1028 floats ------------------
  float4* dst = (float4*)(p.feats + ((size_t)b * p.T + t0) * FEAT);
  const float4* src = (const float4*)tile;
  for (int i = tid; i < nfr * (FEAT / 4); i += NT) __stcs(dst + i, src[i]);
}

// Streaming step (D11): one already-framed 512-sample window per stream.  ONE stream per complex transform (imaginary part
// zero): packing two streams into one FFT - the two-for-one trick the offline kernel plays with neighbouring frames of the
// SAME clip - leaks rounding noise of a loud stream (~1e-7 of its level) into a quiet neighbour; measured 3e-2 in the PCEN
// feature of a stream 40 dB below its partner.  Streams are independent signals, so they do not share a transform; the step
// kernel is ~1 % of a streaming step, the extra FFT work does not show.
__global__ void __launch_bounds__(NT) frontend_step_kernel(FrontParams p, const float* frames, int S) {
  extern __shared__ __align__(16) float smem[];
  float* fre = smem;
  float* fim = fre + 4 * FPAD;
  float2* tw = (float2*)(fim + 4 * FPAD);
  const int tid = threadIdx.x, g = tid >> 6, l = tid & 63;
  for (int k = tid; k < NFFT; k += NT) tw[k] = p.tw[k * (2048 / NFFT)];
  float* z = fre + g * (2 * FPAD);                      // interleaved (re, im) pairs, tru_fft.cuh
  const int sidx = blockIdx.x * 4 + g;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = l + 64 * j;
    TRU_FFT_RE(z, n) = (sidx < S) ? __ldg(frames + (size_t)sidx * NFFT + n) : 0.0f;
    TRU_FFT_IM(z, n) = 0.0f;
  }
  fft_smem<NFFT, -1>(z, tw, l);
  if (sidx >= S) return;
  for (int k = l; k <= NFFT / 2; k += 64) {
    float lm, mg, sn, cs;
    bin_features(TRU_FFT_RE(z, k), TRU_FFT_IM(z, k), lm, mg, sn, cs);
    float* st = p.state_out + (size_t)sidx * NB + k;
    const float M = p.oms * (*st) + p.s * mg;
    *st = M;
    float* o = p.feats + (size_t)sidx * FEAT + k;
    o[0] = lm; o[NB] = pcen_out(mg, M, p); o[2 * NB] = sn; o[3 * NB] = cs;
  }
}

int fill_params(const TruFrontendDesc* d, FrontParams& p) {
  TRU_REQUIRE(d && d->batch > 0, TRU_ERR_ARG, "frontend: bad descriptor");
  p.eps = (float)d->pcen_eps; p.s = (float)d->pcen_s; p.oms = (float)(1.0 - d->pcen_s);
  p.alpha = (float)d->pcen_alpha; p.delta = (float)d->pcen_delta; p.r = (float)d->pcen_r;
  p.delta_r = (float)pow(d->pcen_delta, d->pcen_r);
  p.decay_chunk = (float)pow(1.0 - d->pcen_s, (double)TC);
  p.tw = twiddle_table();
  return TRU_OK;
}

constexpr size_t FRONT_SMEM = (size_t)(TC * FEAT + STAGE + 8 * FPAD) * 4 + NFFT * 8;
constexpr size_t STEP_SMEM = (size_t)(8 * FPAD) * 4 + NFFT * 8;

}  // namespace
}  // namespace tru

using namespace tru;

extern "C" size_t tru_frontend_workspace_bytes(const TruFrontendDesc* d) {
  if (!d || d->batch <= 0 || d->n_samples <= 0) return 0;
  const int T = 1 + d->n_samples / HOP;
  const size_t nchunks = (T + TC - 1) / TC;
  return align_up((size_t)d->batch * nchunks * NB * 4, 256) + align_up((size_t)d->batch * nchunks * 4 + 4, 256);
}

extern "C" int tru_frontend_fwd(const TruFrontendDesc* d, const float* audio, const float* state_in,
                                float* feats, float* state_out, void* ws, size_t ws_bytes, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  FrontParams p{};
  if ((rc = fill_params(d, p))) return rc;
  TRU_REQUIRE(audio && feats && ws, TRU_ERR_ARG, "frontend_fwd: null pointer");
  TRU_REQUIRE(d->n_samples > NFFT / 2, TRU_ERR_ARG, "frontend_fwd: reflect padding needs N > 256 (got %d)", d->n_samples);
  TRU_REQUIRE(aligned16(feats), TRU_ERR_ALIGN, "frontend_fwd: feats must be 16-byte aligned");
  TRU_REQUIRE(ws_bytes >= tru_frontend_workspace_bytes(d), TRU_ERR_WORKSPACE, "frontend_fwd: workspace too small");
  p.B = d->batch; p.N = d->n_samples; p.T = 1 + p.N / HOP; p.nchunks = (p.T + TC - 1) / TC;
  p.audio = audio; p.state_in = state_in; p.feats = feats; p.state_out = state_out;
  const size_t agg_bytes = align_up((size_t)p.B * p.nchunks * NB * 4, 256);
  p.agg = (float*)ws;
  p.flags = (int*)((char*)ws + agg_bytes);
  p.counter = p.flags + (size_t)p.B * p.nchunks;
  cudaStream_t st = (cudaStream_t)stream;
  TRU_CUDA(cudaMemsetAsync(p.flags, 0, (size_t)p.B * p.nchunks * 4 + 4, st));
  TRU_SMEM_OPT_IN(frontend_kernel, FRONT_SMEM);
  ProfScope prof("frontend", 4.0 * p.B * ((double)p.N + 4.0 * NB * p.T), 0.5 * p.B * p.T * 5.0 * NFFT * 9, st);
  frontend_kernel<<<p.B * p.nchunks, NT, FRONT_SMEM, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

extern "C" int tru_frontend_step(const TruFrontendDesc* d, const float* frames, float* state,
                                 float* feats, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  FrontParams p{};
  if ((rc = fill_params(d, p))) return rc;
  TRU_REQUIRE(frames && state && feats, TRU_ERR_ARG, "frontend_step: null pointer");
  p.feats = feats; p.state_out = state;
  const int S = d->batch;
  frontend_step_kernel<<<(S + 3) / 4, NT, STEP_SMEM, (cudaStream_t)stream>>>(p, frames, S);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
