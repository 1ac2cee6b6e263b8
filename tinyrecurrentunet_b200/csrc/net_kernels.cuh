// Internal launch API of the TRU-Net layer kernels (channels-last activations
// [B*T][L][C], fp32).  See DESIGN.md "network kernels" for the scheme:
//   * every conv layer stores its PRE-BatchNorm output Z; BN + ReLU of the
//     producer is applied by the consumer while loading (a = relu(p0*z + p2)),
//     so each activation is written once and read once per consumer;
//   * BN batch statistics are accumulated by the producing kernel's epilogue
//     (fp64 atomics) and turned into (p0, p2) by a tiny finalize kernel;
//   * the backward mirrors this: a layer's incoming gradient dY (w.r.t. the BN
//     output, already ReLU-masked) is turned into dZ = q0*dY + q1*Z + q2 on load.
#pragma once
#include "tru_common.cuh"

namespace tru {

// ---- implicit GEMM over channels with gathered rows --------------------------
struct Seg {
  const float* src;      // values (Z for a forward load, dY for a backward load)
  const float* src2;     // backward load: Z of the same layer (else null)
  const float* p0; const float* p1; const float* p2;   // per-channel coefficients, p0 == null: identity
  const float* W;        // weights of this segment: W[wbase + c*wsc + n*wsn]
  int Lsrc, ld, coff, C; // rows per frame, row stride (floats), first channel, channels
  int fs;                // frame stride in floats (0: Lsrc * ld); tap-shared launches may view a frame as fewer, wider rows
  int cmod;              // wide rows (several source rows side by side): per-channel coefficients repeat every cmod channels (0: no)
  int smul, sadd;        // source row li = q*smul + sadd (must satisfy 0 <= li < Lsrc, else zero row)
  int relu;
  int wbase, wsc, wsn;
};

struct IgemmParams {
  Seg seg[5]; int nseg;
  int BT, Lq, N;               // rows m = bt*Lq + q, output channels
  const float* bias;
  float* out; int Lout, ldo, ocoff, omul, oadd, planar;
  double* stats;               // forward: [2N] sum, sum of squares of the output (null: none)
  const float* extra; int ext_ld;                 // backward: gradient to add (skip connections)
  const float* zmask; const float* mp0; const float* mp2; int use_mask;   // ReLU mask of the layer receiving the gradient
  const float* bmean; const float* binv; double* bstats;                  // backward BN sums: sum g, sum g*xhat
  float src_frac;              // profiler only: fraction of each source this launch needs (0 -> 1)
  // ---- tap-shared mode (transposed convs; tensor-core kernel only) -------------------------------------------
  // ntap > 0: seg[0] is THE source.  Rows m = bt*Lq + q are VIRTUAL rows: the tile stages source row q of frame bt
  // (zero where q >= the row limit of the channel block) for virtual rows [m0 + row_base, m0 + row_base + 128 + max shift) ONCE,
  // and tap j multiplies channels [tap_c0[j], tap_c0[j] + tap_C[j]) of the rows shifted by tap_shift[j] (a row-shifted shared-memory
  // descriptor) with W[tap_wbase[j] + c*wsc + n*wsn].  Virtual rows q >= Lvalid produce no output.  The caller chooses Lq so that a
  // shifted read that leaves its frame lands on a zero row or feeds a discarded output (trunet.cu: convt_fwd / convt_bwd).
  int ntap, tap_shift[5], tap_wbase[5], tap_c0[5], tap_C[5];
  int row_base, Lvalid;        // Lvalid 0: every virtual row is an output row
  int c_hi, lmax_hi;           // source channels >= c_hi exist for rows q < lmax_hi only (c_hi 0: Lsrc for all channels)
  // ---- depthwise epilogue (inference, encoder blocks network.py:28-40; tensor-core kernel only) ---------------------------
  // dw_k > 0: the product (+bias) is NOT stored.  The epilogue applies a = relu(mp0*z + mp2) (the folded BatchNorm of the pointwise
  // conv) and the depthwise conv over the frame's rows,  out[bt][lo][n] = dw_b[n] + sum_j dw_w[n*dw_k + j] a[bt][lo*dw_s - dw_k/2 + j][n]
  // (zero padding), Lout rows per frame: the pointwise output never reaches HBM.  Needs Lq in {16,32,64,128}, 64 < N <= 128.
  const float* dw_w; const float* dw_b; int dw_k, dw_s;
};
int launch_igemm(const IgemmParams& p, cudaStream_t st);       // tensor-core path when eligible, else FFMA
bool igemm_tc_eligible(const IgemmParams& p);
int launch_igemm_tc(const IgemmParams& p, cudaStream_t st);    // 0 launched, 1 not eligible, <0 error
int launch_igemm_simt(const IgemmParams& p, cudaStream_t st);
void set_tc_enabled(bool on);
int read_mbar_debug(unsigned* out, int n);   // debug builds (-DTRU_MBAR_TIMEOUT): stuck mbarrier waits of the GEMM kernel
void set_tc_debug_flags(int f);  // ablation switches (tuning aid)   // 8 or 16 (tuning aid)
bool tc_enabled();

struct WgradJob {
  const float* a_src; const float* a_p0; const float* a_p2; int a_relu;
  int a_L, a_ld, a_coff, a_mul, a_add, C;
  const float* z_src; const float* z_src2; const float* z_p0; const float* z_p1; const float* z_p2;
  int z_L, z_ld, z_coff, z_mul, z_add, N;
  float* dW; int wbase, wsc, wsn;
  float* db;
};
struct WgradParams { WgradJob job[8]; int njobs; int BT, Lq; };
int launch_wgrad(const WgradParams& p, cudaStream_t st);        // shapes the streaming kernel below does not take: small-shape / column-sum / FFMA kernels
int launch_wgrad_simt(const WgradParams& p, cudaStream_t st);

// Streaming weight gradient of a conv layer (tcwgrad2.cu): up to two activation sources (skip concat) or up
// to five taps (transposed conv) in one pass.  dW[wbase[s] + tap*wtap + c*wsc + n*wsn] += sum_m a_s(m,c) dz(zs*m + tap - zpad, n)
struct WgStream {
  int nsrc; const float* a_src[2]; const float* a_p0[2]; const float* a_p2[2];     // a = relu(p0*src + p2) (p0 null: src)
  int a_L[2], a_ld[2], a_add[2], a_C[2], wbase[2];                                  // source row l = q + a_add
  int a_coff[2], z_coff;                                                             // first channel inside a wider row (ld > C: per-row copies)
  const float* z_src; const float* z_src2; const float* z_p0; const float* z_p1; const float* z_p2;   // dz = p0*src + p1*src2 + p2
  int z_L, z_ld, N, ntap, zs, zpad;
  float* dW; int wsc, wsn, wtap; float* db;
  float* scratch; size_t scratch_floats;    // optional: per-CTA partial results (>= SMs x 128 x 384 floats) -> ordered 2-stage reduction instead of atomics
  int n_split; float* dW2; float* db2;      // optional: dz columns n >= n_split belong to a second weight / bias tensor (FGRU directions)
  int BT, Lq;
};
int launch_wgrad_stream(const WgStream& w, cudaStream_t st);     // 0 launched, 1 not eligible

// ---- BatchNorm bookkeeping -----------------------------------------------------
struct BnFwdParams {
  const double* stats; double count; int C; int training;
  const float* gamma; const float* beta; float* running_mean; float* running_var; long long* nbt;
  float eps, momentum;
  float* p0; float* p2; float* mean; float* invstd;
};
int launch_bn_finalize(const BnFwdParams& p, cudaStream_t st);
// inference: the coefficients of all BN layers depend only on parameters and running statistics -> one launch for all of them
struct BnEvalAll { const float* gamma[TRU_NET_NBN]; const float* beta[TRU_NET_NBN]; const float* rmean[TRU_NET_NBN]; const float* rvar[TRU_NET_NBN];
                   float* p0[TRU_NET_NBN]; float* p2[TRU_NET_NBN]; float* mean[TRU_NET_NBN]; float* inv[TRU_NET_NBN]; int C[TRU_NET_NBN]; float eps; };
int launch_bn_finalize_eval_all(const BnEvalAll& p, cudaStream_t st);
struct BnBwdParams {
  const double* bstats; double count; int C;
  const float* gamma; const float* mean; const float* invstd;
  float* q0; float* q1; float* q2; float* dgamma; float* dbeta;
  const double* db_acc; float* db_out;      // optional: fp64 bias-gradient sums to add to a bias gradient
  // optional: gradient of the bias of the conv in front of this BN = sum dz = q0 sum dY + q1 sum Z + q2 count, which
  // cancels to g inv^2 s2 (mean - sum Z / count): the true value is zero, what is left is the rounding of the stored
  // fp32 mean (PyTorch's autograd holds rounding noise of the same size there); no pass over the data needed
  const double* fstats; float* conv_db;
};
int launch_bn_bwd_finalize(const BnBwdParams& p, cudaStream_t st);

// ---- encoder stem (network.py:9-21) ---------------------------------------------
int launch_enc0_fwd(const float* x, const float* w, const float* b, float* out, int BT, cudaStream_t st);
int launch_enc0_wgrad(const float* x, const float* dy, float* dw, float* db, int BT, cudaStream_t st);

// ---- small-channel transposed conv (LastTrCNN: 8 -> 8, k5 s2), convt_small.cu ------------
bool convt_small_eligible(int Cin, int Cout, int k, int s, int L, int Lout);
int launch_convt_small_fwd(const float* z, const float* p0, const float* p2, const float* W, const float* bias, float* out,
                           int BT, int L, int Lout, int planar, cudaStream_t st);
int launch_convt_small_bwd_data(const float* dy, const float* W, float* dx, const float* zmask, const float* mp0, const float* mp2,
                                const float* bmean, const float* binv, double* bstats, int BT, int L, int Lout, int dy_planar,
                                cudaStream_t st);   // dy_planar: dy is (BT, 8, Lout) as autograd hands the output gradient over
int launch_convt_small_wgrad(const float* z, const float* p0, const float* p2, const float* dy, float* dW, float* db,
                             int BT, int L, int Lout, int dy_planar, cudaStream_t st);

// ---- depthwise conv (network.py:33-40), C = 128 ----------------------------------
struct DwParams {
  const float* src; const float* src2; const float* p0; const float* p1; const float* p2;  // loader (fwd or bwd)
  const float* w; const float* bias;    // (C,1,k), (C)
  float* out;
  int BT, Lin, Lout, C, k, stride, pad;
  double* stats;
  // backward-data epilogue
  const float* zmask; const float* mp0; const float* mp2; const float* bmean; const float* binv; double* bstats;
  // backward-weight
  const float* a_src; const float* a_p0; const float* a_p2; float* dw; float* db;
};
int launch_dw_fwd(const DwParams& p, cudaStream_t st);
int launch_dw_bwd_data(const DwParams& p, cudaStream_t st);
int launch_dw_wgrad(const DwParams& p, cudaStream_t st);
int launch_dw_bwd_fused(const DwParams& p, cudaStream_t st);   // bwd_data + wgrad in one pass
int launch_dw_fwd_stream(const DwParams& p, cudaStream_t st);  // dwstream.cu: TMA-fed versions; 1 = shape not covered
int launch_dw_bwd_stream(const DwParams& p, cudaStream_t st);

int launch_colsum(const float* src, const float* src2, const float* q0, const float* q1, const float* q2, float* db,
                  long rows, int ld, int coff, int N, cudaStream_t st);
// planar (BT, C, L) <-> channels-last (BT, L, C), small C
int launch_planar_to_cl(const float* src, float* dst, int BT, int C, int L, cudaStream_t st);

// ---- GRUs (network.py:45-58; torch.nn.GRU gate order r,z,n) ----------------------
struct GruParams {
  const float* G;          // input projections [rows][ldg] (b_ih already added)
  const float* whh[2]; const float* bhh[2];
  float* H;                // [rows][ldh] hidden outputs
  float* cache;            // [rows][ndir*4*Hd]: r, z, n, hn   (null in inference)
  const float* h0; float* hlast;      // TGRU only: [nseq][Hd] (nullable)
  int nseq, steps;
  // backward
  const float* dH; float* dGi; float* dGh;
};
int launch_fgru_fwd(const GruParams& p, cudaStream_t st);   // nseq = B*T, steps = 16, Hd = 64, 2 directions
int launch_fgru_bwd(const GruParams& p, cudaStream_t st);
int launch_tgru_fwd(const GruParams& p, int B, int T, cudaStream_t st);   // sequences (b,l), l<16, Hd = 128
int launch_tgru_bwd(const GruParams& p, int B, int T, cudaStream_t st);
// T = 1 (streaming): gate arithmetic on precomputed projections; Gh null = zero initial state (uses b_hh directly)
int launch_tgru_step_gates(const float* Gi, const float* Gh, const float* bhh, const float* h0, float* H, float* hlast, long nseq,
                           cudaStream_t st);

}  // namespace tru
