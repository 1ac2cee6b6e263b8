/* tru_b200.h - C ABI of libtru_b200.so: the B200 (sm_100a) TRU-Net hot path.
 *
 * Every entry point replaces a piece of the reference's Python hot path
 * (Okrio/tinyrecurrentunet; file:line cited per function).  The reference has
 * no FFI of its own (it is pure Python on PyTorch); the binding a maintainer
 * would add is a ctypes stub inside torch.autograd.Function - see
 * INTEGRATION.md and tinyrecurrentunet_b200/_lib.py.
 *
 * Conventions
 *  - All tensors are contiguous fp32 device memory owned by the caller; the
 *    library never allocates, frees or retains device memory (except static
 *    twiddle tables).  Workspaces are sized by the *_workspace_bytes calls.
 *  - All work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *    host synchronisation except the one-time tru_init().
 *  - Return value: 0 on success, TRU_ERR_* (<0) otherwise; CUDA errors are
 *    returned as -1000 - cudaError_t.  tru_last_error() gives a thread-local
 *    message.  Nothing throws or exits.  There is NO CPU fallback: devices
 *    other than sm_100 return TRU_ERR_ARCH.
 *  - Activation layout inside the network is channels-last [B*T][L][C]; the
 *    network boundary keeps the reference layouts (B,T,4,F) in and (B,T,8,F)
 *    out.
 */
#ifndef TRU_B200_H
#define TRU_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TRU_ABI_VERSION 1
#define TRU_OK 0
#define TRU_ERR_ARG (-1)        /* bad shape / null pointer / unsupported size */
#define TRU_ERR_WORKSPACE (-2)  /* workspace too small */
#define TRU_ERR_ARCH (-3)       /* device is not sm_100 */
#define TRU_ERR_ALIGN (-4)      /* pointer not 16-byte aligned */

#define TRU_NFFT 512
#define TRU_HOP 128
#define TRU_NBINS 257

int tru_abi_version(void);
const char* tru_last_error(void);
/* One-time per-device table initialisation (synchronous).  Called lazily by
 * every entry point; call it explicitly before CUDA-graph capture. */
int tru_init(void);

/* Optional per-kernel profiler used by bench.py for the roofline numbers: when
 * enabled every kernel launch is bracketed by CUDA events on its stream.
 * tru_profile_report synchronises the device and writes one line per kernel:
 * "name launches total_ms algorithmic_bytes flops". */
long long tru_launch_count(void);   /* kernels launched by this library so far (this process) */
int tru_profile_enable(int on);
int tru_profile_report(char* buf, size_t cap);

/* ------------------------------------------------------------------ *
 * Front end: dataset.py:246-272 (ProcessAudio.forward), :56-76 (pcenfunc)
 * audio (B,N) -> feats (B,T',4,257), T' = 1 + N/128.
 * ------------------------------------------------------------------ */
typedef struct {
  int batch;       /* B */
  int n_samples;   /* N  (> 256) */
  double pcen_eps, pcen_s, pcen_alpha, pcen_delta, pcen_r; /* dataset.py:57 defaults 1e-6 .025 .98 2 .5 */
} TruFrontendDesc;

size_t tru_frontend_workspace_bytes(const TruFrontendDesc* d);
/* pcen_state_in / pcen_state_out: optional (B,257) smoother state M (D11). */
int tru_frontend_fwd(const TruFrontendDesc* d, const float* audio,
                     const float* pcen_state_in, float* feats,
                     float* pcen_state_out, void* workspace,
                     size_t workspace_bytes, void* stream);
/* Streaming step: frames (S,512) already framed by the caller (centre = the
 * reference's reflect-padded framing), state (S,257) in/out, feats (S,4,257). */
int tru_frontend_step(const TruFrontendDesc* d, const float* frames,
                      float* pcen_state, float* feats, void* stream);

/* ------------------------------------------------------------------ *
 * Back end: dataset.py:182-203 (mod_phase), phm.py:31-45 (PhaseAwareMask),
 * dataset.py:293-296 (torch.istft, rectangular window, centre).
 * net_out (B,T',C,257) -> audio (B, 128 (T'-1)).
 * ------------------------------------------------------------------ */
typedef struct {
  int batch;      /* B */
  int n_frames;   /* T' */
  int n_channels; /* C: 8 for the network output, 3 for ProcessAudio.backward */
  int ch_mag, ch_sin, ch_cos;   /* set 0 ("mixture"): 0,2,3 */
  int ch_sin1, ch_cos1;         /* set 1 ("noise"):   6,7 ; ignored if !use_mask */
  int use_mask;                 /* 1: beta-sigmoid mask; 0: plain mod_phase+iSTFT */
  double beta;
} TruBackendDesc;

int tru_backend_fwd(const TruBackendDesc* d, const float* net_out, float* audio,
                    void* stream);
/* Streaming step (stream.py:83-109 intent): one network-output frame per stream,
 * net_out (S,C,257) with S = d->batch; ola_state (S,384) in/out = overlap-add sums still waiting for
 * later frames (zero-initialised by the caller); audio (S,128) = output block frame_index-2 of every
 * stream (zeros while frame_index < 2: two hops of look-ahead, identical to the offline centre=True
 * iSTFT).  add_frame = 0 flushes: no new frame, emits the block from the frames already seen. */
int tru_backend_step(const TruBackendDesc* d, const float* net_out, float* ola_state, float* audio,
                     int frame_index, int add_frame, void* stream);
/* grad_audio (B, 128 (T'-1)) -> grad_net_out (B,T',C,257) (fully written). */
int tru_backend_bwd(const TruBackendDesc* d, const float* net_out,
                    const float* grad_audio, float* grad_net_out, void* stream);

/* ------------------------------------------------------------------ *
 * Loss: stft_loss.py:9-166 (MultiResolutionSTFTLoss, band="full") fused with
 * util.py:239-240 (L1).  x = prediction, y = target, both (B,N).
 * sums layout (double[16]): for r<3: [4r+0] sum (|Y|-|X|)^2, [4r+1] sum |Y|^2,
 * [4r+2] sum |ln|Y| - ln|X||, [4r+3] unused;  [12] sum |x-y|.
 * out (float[3]) = {l1, sc_loss, mag_loss}.
 * ------------------------------------------------------------------ */
typedef struct {
  int batch;
  int n_samples;
  int n_res;              /* <= 3 */
  int fft_size[3];        /* each in {512,1024,2048} */
  int hop_size[3];
  int win_length[3];      /* <= fft_size */
  double sc_lambda, mag_lambda;
} TruLossDesc;

int tru_loss_fwd(const TruLossDesc* d, const float* x, const float* y,
                 const float* const* windows, double* sums, float* out,
                 void* stream);
/* grad_out (float[3], device): upstream grads of {l1, sc_loss, mag_loss}. */
int tru_loss_bwd(const TruLossDesc* d, const float* x, const float* y,
                 const float* const* windows, const double* sums,
                 const float* grad_out, float* grad_x, void* stream);

/* ------------------------------------------------------------------ *
 * TRU-Net: network.py:122-171 (TRUNet.forward, repaired per SURVEY D4/D10/D11)
 * and its backward.  x (B,T,4,257) -> out (B,T,8,257).
 *
 * params: 108 device pointers in the canonical order of
 *   tinyrecurrentunet_b200/network.py:PARAM_ORDER (the state-dict order of
 *   network.py:134-150: encoder, decoder, FGRU, TGRU), PyTorch layouts.
 * bn_*:   23 BatchNorm layers in the order encoder.1..5 (pw, dw), decoder.0..4
 *   (pw, convT), decoder.5 (pw), FGRU, TGRU.  Running stats are updated in
 *   training mode exactly like nn.BatchNorm1d (momentum, unbiased variance);
 *   bn_num_batches entries may be null.
 * The workspace holds every saved activation; the SAME workspace must be
 * passed to tru_trunet_backward (sized with with_backward = 1).
 * grads: 108 pointers, same order/shapes as params, zero-initialised by the
 *   caller; the backward ACCUMULATES into them.
 * h0 / h_last: optional TGRU state (B*16, 128) for streaming (D11); the
 *   backward assumes h0 == NULL.
 * ------------------------------------------------------------------ */
#define TRU_NET_NPARAMS 108
#define TRU_NET_NBN 23
typedef struct {
  int batch;      /* B */
  int n_frames;   /* T */
  int training;   /* 1: batch statistics + running-stat update; 0: running stats */
  double bn_eps, bn_momentum;   /* 1e-5, 0.1 */
} TruNetDesc;

size_t tru_trunet_workspace_bytes(const TruNetDesc* d, int with_backward);
int tru_trunet_forward(const TruNetDesc* d, const float* const* params,
                       float* const* bn_running_mean, float* const* bn_running_var,
                       long long* const* bn_num_batches, const float* x,
                       const float* h0, float* out, float* h_last,
                       void* workspace, size_t workspace_bytes, void* stream);
int tru_trunet_backward(const TruNetDesc* d, const float* const* params,
                        const float* x, const float* grad_out,
                        float* const* grads, void* workspace,
                        size_t workspace_bytes, void* stream);
/* Test aid: byte offset of a named saved buffer inside the workspace (-1 if unknown). */
long long tru_trunet_buffer_offset(const TruNetDesc* d, const char* name, int index);

/* ------------------------------------------------------------------ *
 * Optimizer step on flat buffers: train.py:138
 * (nn.utils.clip_grad_norm_(params, 1e9): the L2 norm of all gradients, and
 * the in-place scaling by min(1, max/(norm+1e-6)) when it exceeds max) fused
 * with train.py:140 (torch.optim.AdamW.step, decoupled weight decay, no
 * amsgrad).  params / grads / exp_avg / exp_avg_sq are flat fp32 buffers of n
 * elements in one common layout (n a multiple of 4; padding elements stay 0).
 * lr is the value util.py:81-156 (LinearWarmupCosineDecay) set for this step;
 * step is the 1-based update count.  grad_norm_out (device float, may be
 * null) receives the pre-clip norm.  Two launches, no host synchronisation.
 * ------------------------------------------------------------------ */
typedef struct {
  long long n;          /* elements per flat buffer, multiple of 4 */
  long long step;       /* t >= 1 */
  double lr, beta1, beta2, eps, weight_decay;   /* train.py:68: lr 4e-4; torch defaults .9 .999 1e-8 1e-2 */
  double max_grad_norm; /* <= 0: report the norm only */
} TruAdamWDesc;

size_t tru_flat_adamw_workspace_bytes(const TruAdamWDesc* d);
int tru_flat_adamw_step(const TruAdamWDesc* d, float* params, float* grads,
                        float* exp_avg, float* exp_avg_sq, float* grad_norm_out,
                        void* workspace, size_t workspace_bytes, void* stream);
/* The norm alone (what train.py:138 returns): grads (n floats) -> norm_out. */
int tru_flat_grad_norm(long long n, const float* grads, float* norm_out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ *
 * Batch assembly on the device: dataset.py:79-126
 * (DataAugment.__call__: F.gain, F.lowpass_biquad, F.highpass_biquad of
 * torchaudio, each biquad output clamped to [-1,1]) for B noise clips at
 * once, and dataset.py:367-379 (random crop of the clean clip,
 * noisy = clean + augmented noise).
 * coef (B, TRU_AUGMENT_NCOEF) per clip: {10^(gain_db/20), low-pass row, high-pass row}; a row is
 *   {b0,b1,b2,a1,a2}/a0 followed by six 2x2 matrices (row major) A^(TRU_AUGMENT_CHUNK * p),
 *   p = 1, 8, 16, 32, 64, 128, with A = [[-a1,-a2],[1,0]]: the zero-input response of the
 *   recurrence over p chunks, which the kernel uses to stitch the chunks it filters in
 *   parallel (tinyrecurrentunet_b200/dataset.py builds the rows, powers in float64).
 * ------------------------------------------------------------------ */
#define TRU_AUGMENT_CHUNK 63
#define TRU_AUGMENT_NCOEF 59
int tru_augment_fwd(int batch, int n_samples, const float* noise, const float* coef,
                    float* out, void* stream);
/* clean (B,n_clean), aug_noise (B,n_noise), clean_start / noise_start (B) device ints (null = 0):
 * clean_out[b,i] = clean[b, clean_start[b]+i], noisy_out[b,i] = clean_out[b,i] +
 * aug_noise[b, (noise_start[b]+i) mod n_noise], i < n_out <= n_clean. */
int tru_mix_crop(int batch, int n_clean, int n_noise, int n_out, const float* clean,
                 const float* aug_noise, const int* clean_start, const int* noise_start,
                 float* clean_out, float* noisy_out, void* stream);

/* ------------------------------------------------------------------ *
 * cos_loss.py:41-56 (CosSimLoss.forward): mean over the
 * segments [bounds[i], bounds[i+1]) and the batch rows of 1 - cosine similarity
 * (nn.CosineSimilarity: each norm clamped to eps) of x (prediction) and y
 * (target), both (B, n_samples).  The reference only runs for one row and
 * detaches the result; this one averages rows and has a backward w.r.t. x.
 * stats (B, n_seg, 3) doubles = {x.y, |x|^2, |y|^2}, written by fwd, read by bwd.
 * ------------------------------------------------------------------ */
typedef struct {
  int batch, n_samples, n_seg;   /* n_seg <= 8 */
  int bounds[9];                 /* cos_loss.py:25 default g = [508,1016,2032,4062] -> {0,508,1016,2032,4062} */
  double eps;                    /* 1e-5 */
} TruCosSimDesc;
int tru_cossim_fwd(const TruCosSimDesc* d, const float* x, const float* y, double* stats,
                   float* loss, void* stream);
/* grad_loss: device float (upstream gradient of the scalar); grad_x (B, n_samples) fully written. */
int tru_cossim_bwd(const TruCosSimDesc* d, const float* x, const float* y, const double* stats,
                   const float* grad_loss, float* grad_x, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRU_B200_H */
