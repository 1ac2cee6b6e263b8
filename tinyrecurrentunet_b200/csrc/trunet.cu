// TRU-Net forward / backward orchestration (network.py:122-171 repaired per
// SURVEY D4/D10/D11) on top of the layer kernels, exported through the C ABI.
//
// Activations are channels-last [B*T][L][C]; each conv stores its pre-BN output Z
// in the caller-provided workspace (saved for the backward), BN+ReLU is applied by
// consumers on load.  The host code below is the only "runtime": one C call walks
// the whole layer list and enqueues every kernel on the caller's stream.
#include <algorithm>
#include <stdlib.h>
#include <string>
#include "net_kernels.cuh"
#include "../../include/tru_b200_debug.h"

namespace tru {
namespace {

constexpr int NBN = TRU_NET_NBN, NPARAM = TRU_NET_NPARAMS;

// geometry
const int ENC_K[6] = {5, 3, 5, 3, 5, 3}, ENC_S[6] = {2, 1, 2, 1, 2, 2};
const int ENC_L[6] = {128, 128, 64, 64, 32, 16};      // output length of encoder block i
const int ENC_CIN[6] = {4, 64, 128, 128, 128, 128};
const int DEC_K[6] = {3, 5, 3, 5, 3, 5}, DEC_S[6] = {2, 2, 1, 2, 1, 2};
const int DEC_LP[6] = {16, 32, 64, 64, 128, 128};     // rows of the pointwise conv (= skip length)
const int DEC_LT[6] = {31, 65, 66, 129, 130, 257};    // rows after the transposed conv
const int DEC_CIN[6] = {64, 192, 192, 192, 192, 128};
const int DEC_COUT[6] = {64, 64, 64, 64, 64, 8};
const int DEC_SKIP[6] = {-1, 4, 3, 2, 1, 0};          // encoder block feeding decoder d

// canonical parameter order (Python side: network.PARAM_ORDER)
inline int P_ENC(int i, int j) { return i == 0 ? j : 2 + 8 * (i - 1) + j; }   // i>=1: pw.w pw.b bn1.w bn1.b dw.w dw.b bn2.w bn2.b
inline int P_DEC(int d, int j) { return 42 + 8 * d + j; }                    // pw.w pw.b bn1.w bn1.b ct.w ct.b [bn2.w bn2.b]
constexpr int P_FGRU = 88, P_TGRU = 100;
inline int BN_ENC(int i, int which) { return 2 * (i - 1) + which; }
inline int BN_DEC(int d, int which) { return 10 + 2 * d + which; }
constexpr int BN_FGRU = 21, BN_TGRU = 22;

struct BnSlot { double* stats; double* bstats; float *p0, *p2, *mean, *inv, *q0, *q1, *q2; };

constexpr size_t WG_SCRATCH_FLOATS = (size_t)160 * 128 * 384;     // up to 160 CTAs x (128 x 384) accumulator

struct Plan {
  size_t total = 0;
  size_t take(size_t bytes) { size_t o = total; total += align_up(bytes, 256); return o; }
  // offsets (bytes)
  size_t A0, Zp[6], Zd[6], GF, HF, CF, ZFp, GT, HT, CT, ZTp, ZDp[6], ZDt[5];
  size_t stats, bstats, small;
  // backward
  size_t dA0, dZp[6], dZd[6], dGFi, dGFh, dHF, dZFp, dGTi, dGTh, dHT, dZTp, dZDp[6], dZDt[5], dOUT, dSkip[5], wgScratch;
  void build(long BT, bool bwd) {
    auto f = [&](long per_frame) { return take((size_t)BT * per_frame * 4); };
    A0 = f(128 * 64);
    for (int i = 1; i <= 5; ++i) { Zp[i] = f((long)ENC_L[i - 1] * 128); Zd[i] = f((long)ENC_L[i] * 128); }
    GF = f(16 * 384); HF = f(16 * 128); CF = f(16 * 512); ZFp = f(16 * 64);
    GT = f(16 * 384); HT = f(16 * 128); CT = f(16 * 512); ZTp = f(16 * 64);
    for (int d = 0; d <= 5; ++d) { ZDp[d] = f((long)DEC_LP[d] * DEC_COUT[d]); if (d < 5) ZDt[d] = f((long)DEC_LT[d] * 64); }
    // bstats slot: sum g, sum g*xhat, fp64 bias-gradient scratch
    stats = take(NBN * 256 * 8); bstats = take(NBN * 384 * 8); small = take(NBN * 7 * 128 * 4);
    if (!bwd) return;
    dA0 = f(128 * 64);
    for (int i = 1; i <= 5; ++i) { dZp[i] = f((long)ENC_L[i - 1] * 128); dZd[i] = f((long)ENC_L[i] * 128); }
    dGFi = f(16 * 384); dGFh = f(16 * 384); dHF = f(16 * 128); dZFp = f(16 * 64);
    dGTi = f(16 * 384); dGTh = f(16 * 384); dHT = f(16 * 128); dZTp = f(16 * 64);
    for (int d = 0; d <= 5; ++d) { dZDp[d] = f((long)DEC_LP[d] * DEC_COUT[d]); if (d < 5) dZDt[d] = f((long)DEC_LT[d] * 64); }
    dOUT = f(257 * 8);
    dSkip[0] = f(128 * 64); dSkip[1] = f(128 * 128); dSkip[2] = f(64 * 128); dSkip[3] = f(64 * 128); dSkip[4] = f(32 * 128);
    wgScratch = take(WG_SCRATCH_FLOATS * 4);      // per-CTA partial weight gradients (2-stage reduction)
  }
};

// an activation: pre-BN tensor + the affine that turns it into the post-BN-ReLU value
struct Act { const float* z; int L, C; const float* p0; const float* p2; int bn; };
// a gradient w.r.t. a layer output: dY (BN output grad, ReLU-masked) + what turns it into dZ
struct Grad { const float* dy; const float* z; const float* q0; const float* q1; const float* q2; int L, C; };

struct Ctx {
  const TruNetDesc* d; cudaStream_t st; char* ws; Plan plan; long BT; int B, T;
  const float* const* prm; float* const* grd;
  float* const* rmean; float* const* rvar; long long* const* nbt;
  BnSlot bn[NBN];
  float* F(size_t off) const { return (float*)(ws + off); }
  void slots() {
    for (int i = 0; i < NBN; ++i) {
      bn[i].stats = (d && !d->training) ? nullptr : (double*)(ws + plan.stats) + i * 256;    // inference: no batch statistics
      bn[i].bstats = (double*)(ws + plan.bstats) + i * 384;
      float* s = (float*)(ws + plan.small) + (size_t)i * 7 * 128;
      bn[i].p0 = s; bn[i].p2 = s + 128; bn[i].mean = s + 256; bn[i].inv = s + 384;
      bn[i].q0 = s + 512; bn[i].q1 = s + 640; bn[i].q2 = s + 768;
    }
  }
  Act act(size_t off, int L, int C, int bnidx) const {
    Act a{F(off), L, C, nullptr, nullptr, bnidx};
    if (bnidx >= 0) { a.p0 = bn[bnidx].p0; a.p2 = bn[bnidx].p2; }
    return a;
  }
  Grad grad(size_t doff, size_t zoff, int L, int C, int bnidx) const {
    Grad g{F(doff), F(zoff), nullptr, nullptr, nullptr, L, C};
    if (bnidx >= 0) { g.q0 = bn[bnidx].q0; g.q1 = bn[bnidx].q1; g.q2 = bn[bnidx].q2; }
    return g;
  }
};

#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)

Seg fwd_seg(const Act& a, const float* W, int wbase, int wsc, int wsn, int smul, int sadd) {
  Seg s{};
  s.src = a.z; s.p0 = a.p0; s.p2 = a.p2; s.relu = 1; s.W = W;
  s.Lsrc = a.L; s.ld = a.C; s.coff = 0; s.C = a.C; s.smul = smul; s.sadd = sadd;
  s.wbase = wbase; s.wsc = wsc; s.wsn = wsn;
  return s;
}
Seg bwd_seg(const Grad& g, int coff, int C, int ld, const float* W, int wbase, int wsc, int wsn, int smul, int sadd) {
  Seg s{};
  s.src = g.dy; s.src2 = g.q0 ? g.z : nullptr; s.p0 = g.q0; s.p1 = g.q1; s.p2 = g.q2; s.relu = 0; s.W = W;
  s.Lsrc = g.L; s.ld = ld; s.coff = coff; s.C = C; s.smul = smul; s.sadd = sadd;
  s.wbase = wbase; s.wsc = wsc; s.wsn = wsn;
  return s;
}

// inference: all 23 BN layers in one launch at the start of the forward (nothing depends on batch statistics)
int bn_fin_eval_all(Ctx& c) {
  BnEvalAll p{};
  auto set = [&](int idx, int pg, int C) {
    p.gamma[idx] = c.prm[pg]; p.beta[idx] = c.prm[pg + 1]; p.rmean[idx] = c.rmean[idx]; p.rvar[idx] = c.rvar[idx];
    p.p0[idx] = c.bn[idx].p0; p.p2[idx] = c.bn[idx].p2; p.mean[idx] = c.bn[idx].mean; p.inv[idx] = c.bn[idx].inv; p.C[idx] = C;
  };
  for (int i = 1; i <= 5; ++i) { set(BN_ENC(i, 0), P_ENC(i, 2), 128); set(BN_ENC(i, 1), P_ENC(i, 6), 128); }
  for (int d = 0; d <= 5; ++d) { set(BN_DEC(d, 0), P_DEC(d, 2), DEC_COUT[d]); if (d < 5) set(BN_DEC(d, 1), P_DEC(d, 6), 64); }
  set(BN_FGRU, P_FGRU + 10, 64); set(BN_TGRU, P_TGRU + 6, 64);
  p.eps = (float)c.d->bn_eps;
  return launch_bn_finalize_eval_all(p, c.st);
}
int bn_fin(Ctx& c, int idx, int C, long rows, const float* gamma, const float* beta) {
  if (!c.d->training) return TRU_OK;        // done by bn_fin_eval_all
  BnFwdParams p{};
  p.stats = c.bn[idx].stats; p.count = (double)rows; p.C = C; p.training = c.d->training;
  p.gamma = gamma; p.beta = beta; p.running_mean = c.rmean[idx]; p.running_var = c.rvar[idx];
  p.nbt = c.nbt ? c.nbt[idx] : nullptr;
  p.eps = (float)c.d->bn_eps; p.momentum = (float)c.d->bn_momentum;
  p.p0 = c.bn[idx].p0; p.p2 = c.bn[idx].p2; p.mean = c.bn[idx].mean; p.invstd = c.bn[idx].inv;
  return launch_bn_finalize(p, c.st);
}
int bn_bfin(Ctx& c, int idx, int C, long rows, int pgamma, int pbias_acc = -1, int pconv_bias = -1) {
  BnBwdParams p{};
  if (pbias_acc >= 0) { p.db_acc = c.bn[idx].bstats + 256; p.db_out = c.grd[pbias_acc]; }
  if (pconv_bias >= 0) { p.fstats = c.bn[idx].stats; p.conv_db = c.grd[pconv_bias]; }
  p.bstats = c.bn[idx].bstats; p.count = (double)rows; p.C = C;
  p.gamma = c.prm[pgamma]; p.mean = c.bn[idx].mean; p.invstd = c.bn[idx].inv;
  p.q0 = c.bn[idx].q0; p.q1 = c.bn[idx].q1; p.q2 = c.bn[idx].q2;
  p.dgamma = c.grd[pgamma]; p.dbeta = c.grd[pgamma + 1];
  return launch_bn_bwd_finalize(p, c.st);
}

// ---- forward building blocks ---------------------------------------------------
// pointwise conv over [x1 (shifted by padL) ; skip] -> out [BT][Lq][N]
int pw_fwd(Ctx& c, const Act& x1, int padL, const Act* skip, const float* W, const float* bias, int N, int Lq,
           float* out, int ldo, int ocoff, double* stats) {
  IgemmParams p{};
  const int K = x1.C + (skip ? skip->C : 0);
  p.seg[0] = fwd_seg(x1, W, 0, 1, K, 1, -padL);
  p.nseg = 1;
  if (skip) { p.seg[1] = fwd_seg(*skip, W, x1.C, 1, K, 1, 0); p.nseg = 2; }
  p.BT = (int)c.BT; p.Lq = Lq; p.N = N; p.bias = bias;
  p.out = out; p.Lout = Lq; p.ldo = ldo; p.ocoff = ocoff; p.omul = 1; p.oadd = 0;
  p.stats = stats;
  return launch_igemm(p, c.st);
}

// Tap-shared launch parameters common to the transposed-conv forward and data gradient (net_kernels.cuh: IgemmParams::ntap):
// taps (delta_j, channel slice, weight offset); virtual rows per frame Lq_v chosen by the caller.
struct Tap { int delta, c0, C, wbase; };
int g_eval_fusion = 1;          // tru_debug_set_eval_fusion: 0 keeps the layer-by-layer schedule in inference (A/B aid; tests that read the pointwise outputs)
// TRU_TAP_SHARED_OFF (bit 0: forward, bit 1: data gradient) sends the transposed convs down the one-segment-per-tap launches (A/B aid)
int tap_shared_off() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TRU_TAP_SHARED_OFF"); v = e ? atoi(e) : 0; }
  return v;
}
void set_taps(IgemmParams& p, const Tap* taps, int ntap) {
  int dmin = 0;
  for (int j = 0; j < ntap; ++j) dmin = std::min(dmin, taps[j].delta);
  p.ntap = ntap; p.row_base = dmin;
  for (int j = 0; j < ntap; ++j) {
    p.tap_shift[j] = taps[j].delta - dmin; p.tap_c0[j] = taps[j].c0; p.tap_C[j] = taps[j].C; p.tap_wbase[j] = taps[j].wbase;
  }
}

// transposed conv (weight (Cin,Cout,k), stride s, pad s/2): x [BT][L][Cin] -> out [BT][Lout][Cout]
//   out[lo] = sum_j x[(lo + pad - j) / s] W_j  over the taps j with (lo + pad - j) divisible by s.
// Per output-row parity class (lo = s q + par) the taps read source rows q + delta_j, delta_j = (par + pad - j) / s.  With two or
// more taps the launch is TAP-SHARED: rows m = (bt, q) are virtual rows, Lq_v per frame; the kernel stages source row q of every
// virtual row once (zero for q >= L) and each tap reads the stage shifted by delta_j rows.  Lq_v = max(Lq + max delta, L - min delta)
// makes every shifted read that leaves its frame land on a zero row (q >= L of the neighbouring frame) or belong to a virtual
// row beyond the Lq real ones, which is not stored.
int convt_fwd(Ctx& c, const Act& x, const float* W, const float* bias, int Cout, int k, int s, int Lout, float* out,
              double* stats, int planar) {
  const int pad = s / 2;
  if (!stats && convt_small_eligible(x.C, Cout, k, s, x.L, Lout))
    return launch_convt_small_fwd(x.z, x.p0, x.p2, W, bias, out, (int)c.BT, x.L, Lout, planar, c.st);
  for (int par = 0; par < s; ++par) {
    IgemmParams p{};
    Tap taps[5];
    int nt = 0, dmax = 0, dmin = 0;
    p.nseg = 0;
    for (int j = 0; j < k; ++j) {
      if (((par + pad - j) % s + s) % s != 0) continue;
      // lo = s*q + par ; li = (lo + pad - j)/s = q + (par + pad - j)/s   (exact division)
      const int num = par + pad - j;
      const int add = num >= 0 ? num / s : -((-num) / s);
      p.seg[p.nseg++] = fwd_seg(x, W, j, Cout * k, k, 1, add);
      taps[nt++] = Tap{add, 0, x.C, j};
      dmax = std::max(dmax, add); dmin = std::min(dmin, add);
    }
    const int Lq = (Lout - par + s - 1) / s;
    p.BT = (int)c.BT; p.Lq = Lq; p.N = Cout; p.bias = bias;
    p.out = out; p.Lout = Lout; p.ldo = Cout; p.ocoff = 0; p.omul = s; p.oadd = par; p.planar = planar;
    p.stats = stats; p.src_frac = 1.0f / s;
    if (Lq <= 0 || p.nseg == 0) continue;
    if (nt >= 2 && !planar && tc_enabled() && !(tap_shared_off() & 1)) {
      IgemmParams q = p;
      q.nseg = 1; q.seg[0] = fwd_seg(x, W, 0, Cout * k, k, 1, 0);
      set_taps(q, taps, nt);
      q.Lq = std::max(Lq + dmax, x.L - dmin); q.Lvalid = Lq;
      if (igemm_tc_eligible(q)) { TRY(launch_igemm(q, c.st)); continue; }
    }
    TRY(launch_igemm(p, c.st));
  }
  return TRU_OK;
}

int forward(Ctx& c, const float* x, const float* h0, float* out, float* hlast) {
  const Plan& P = c.plan;
  const long BT = c.BT;
  if (c.d->training) TRU_CUDA(cudaMemsetAsync(c.ws + P.stats, 0, NBN * 256 * 8, c.st));
  else TRY(bn_fin_eval_all(c));
  // encoder stem
  TRY(launch_enc0_fwd(x, c.prm[0], c.prm[1], c.F(P.A0), (int)BT, c.st));
  Act cur = c.act(P.A0, 128, 64, -1);
  for (int i = 1; i <= 5; ++i) {
    const int Lin = ENC_L[i - 1], Lo = ENC_L[i];
    const int b1 = BN_ENC(i, 0), b2 = BN_ENC(i, 1);
    if (!c.d->training && g_eval_fusion && tc_enabled()) {
      // inference: no batch statistics between the two convs -> the depthwise conv runs in the pointwise GEMM's epilogue and
      // the pointwise output (the block's largest tensor) is never written (tcgemm.cu, EPI 3)
      IgemmParams p{};
      p.seg[0] = fwd_seg(cur, c.prm[P_ENC(i, 0)], 0, 1, cur.C, 1, 0);
      p.nseg = 1; p.BT = (int)BT; p.Lq = Lin; p.N = 128; p.bias = c.prm[P_ENC(i, 1)];
      p.mp0 = c.bn[b1].p0; p.mp2 = c.bn[b1].p2;
      p.dw_w = c.prm[P_ENC(i, 4)]; p.dw_b = c.prm[P_ENC(i, 5)]; p.dw_k = ENC_K[i]; p.dw_s = ENC_S[i];
      p.out = c.F(P.Zd[i]); p.Lout = Lo; p.ldo = 128; p.omul = 1;
      if (igemm_tc_eligible(p)) {
        TRY(launch_igemm(p, c.st));
        cur = c.act(P.Zd[i], Lo, 128, b2);
        continue;
      }
    }
    TRY(pw_fwd(c, cur, 0, nullptr, c.prm[P_ENC(i, 0)], c.prm[P_ENC(i, 1)], 128, Lin, c.F(P.Zp[i]), 128, 0, c.bn[b1].stats));
    TRY(bn_fin(c, b1, 128, BT * Lin, c.prm[P_ENC(i, 2)], c.prm[P_ENC(i, 3)]));
    DwParams dp{};
    dp.src = c.F(P.Zp[i]); dp.p0 = c.bn[b1].p0; dp.p2 = c.bn[b1].p2;
    dp.w = c.prm[P_ENC(i, 4)]; dp.bias = c.prm[P_ENC(i, 5)]; dp.out = c.F(P.Zd[i]);
    dp.BT = (int)BT; dp.Lin = Lin; dp.Lout = Lo; dp.C = 128; dp.k = ENC_K[i]; dp.stride = ENC_S[i]; dp.pad = ENC_K[i] / 2;
    dp.stats = c.bn[b2].stats;
    TRY(launch_dw_fwd(dp, c.st));
    TRY(bn_fin(c, b2, 128, BT * Lo, c.prm[P_ENC(i, 6)], c.prm[P_ENC(i, 7)]));
    cur = c.act(P.Zd[i], Lo, 128, b2);
  }
  // FGRU (network.py:149): input projection for both directions, recurrence, pw conv
  for (int dir = 0; dir < 2; ++dir)
    TRY(pw_fwd(c, cur, 0, nullptr, c.prm[P_FGRU + 4 * dir], c.prm[P_FGRU + 4 * dir + 2], 192, 16, c.F(P.GF), 384, 192 * dir, nullptr));
  {
    GruParams g{};
    g.G = c.F(P.GF); g.whh[0] = c.prm[P_FGRU + 1]; g.whh[1] = c.prm[P_FGRU + 5];
    g.bhh[0] = c.prm[P_FGRU + 3]; g.bhh[1] = c.prm[P_FGRU + 7];
    g.H = c.F(P.HF); g.cache = c.d->training ? c.F(P.CF) : nullptr; g.nseq = (int)BT; g.steps = 16;
    TRY(launch_fgru_fwd(g, c.st));
  }
  Act hf = c.act(P.HF, 16, 128, -1);
  TRY(pw_fwd(c, hf, 0, nullptr, c.prm[P_FGRU + 8], c.prm[P_FGRU + 9], 64, 16, c.F(P.ZFp), 64, 0, c.bn[BN_FGRU].stats));
  TRY(bn_fin(c, BN_FGRU, 64, BT * 16, c.prm[P_FGRU + 10], c.prm[P_FGRU + 11]));
  Act fo = c.act(P.ZFp, 16, 64, BN_FGRU);
  // TGRU (network.py:150, wiring D4): sequences (b, l) over t
  TRY(pw_fwd(c, fo, 0, nullptr, c.prm[P_TGRU], c.prm[P_TGRU + 2], 384, 16, c.F(P.GT), 384, 0, nullptr));
  if (c.T == 1 && !c.d->training) {
    // streaming step: hidden projection of all sequences as one GEMM (scratch: the gate-cache region, unused in inference)
    const float* Gh = nullptr;
    if (h0) {
      Act hp{h0, 16, 128, nullptr, nullptr, -1};
      TRY(pw_fwd(c, hp, 0, nullptr, c.prm[P_TGRU + 1], c.prm[P_TGRU + 3], 384, 16, c.F(P.CT), 384, 0, nullptr));
      Gh = c.F(P.CT);
    }
    TRY(launch_tgru_step_gates(c.F(P.GT), Gh, c.prm[P_TGRU + 3], h0, c.F(P.HT), hlast, (long)c.B * 16, c.st));
  } else {
    GruParams g{};
    g.G = c.F(P.GT); g.whh[0] = c.prm[P_TGRU + 1]; g.bhh[0] = c.prm[P_TGRU + 3];
    g.H = c.F(P.HT); g.cache = c.d->training ? c.F(P.CT) : nullptr; g.h0 = h0; g.hlast = hlast; g.nseq = c.B * 16; g.steps = c.T;
    TRY(launch_tgru_fwd(g, c.B, c.T, c.st));
  }
  Act ht = c.act(P.HT, 16, 128, -1);
  TRY(pw_fwd(c, ht, 0, nullptr, c.prm[P_TGRU + 4], c.prm[P_TGRU + 5], 64, 16, c.F(P.ZTp), 64, 0, c.bn[BN_TGRU].stats));
  TRY(bn_fin(c, BN_TGRU, 64, BT * 16, c.prm[P_TGRU + 6], c.prm[P_TGRU + 7]));
  cur = c.act(P.ZTp, 16, 64, BN_TGRU);
  // decoder (network.py:141-146, skip concat :95-98)
  for (int d = 0; d <= 5; ++d) {
    const int Lp = DEC_LP[d], Co = DEC_COUT[d];
    const int b1 = BN_DEC(d, 0);
    Act skip{};
    const Act* sp = nullptr;
    int padL = 0;
    if (d >= 1) {
      const int e = DEC_SKIP[d];
      skip = e == 0 ? c.act(P.A0, 128, 64, -1) : c.act(P.Zd[e], ENC_L[e], 128, BN_ENC(e, 1));
      sp = &skip;
      const int diff = Lp - cur.L;
      padL = diff >= 0 ? diff / 2 : -((-diff + 1) / 2);       // Python floor division (network.py:97)
    }
    TRY(pw_fwd(c, cur, padL, sp, c.prm[P_DEC(d, 0)], c.prm[P_DEC(d, 1)], Co, Lp, c.F(P.ZDp[d]), Co, 0, c.bn[b1].stats));
    TRY(bn_fin(c, b1, Co, BT * Lp, c.prm[P_DEC(d, 2)], c.prm[P_DEC(d, 3)]));
    Act pw = c.act(P.ZDp[d], Lp, Co, b1);
    if (d < 5) {
      const int b2 = BN_DEC(d, 1);
      TRY(convt_fwd(c, pw, c.prm[P_DEC(d, 4)], c.prm[P_DEC(d, 5)], Co, DEC_K[d], DEC_S[d], DEC_LT[d], c.F(P.ZDt[d]),
                    c.bn[b2].stats, 0));
      TRY(bn_fin(c, b2, Co, BT * DEC_LT[d], c.prm[P_DEC(d, 6)], c.prm[P_DEC(d, 7)]));
      cur = c.act(P.ZDt[d], DEC_LT[d], Co, b2);
    } else {
      TRY(convt_fwd(c, pw, c.prm[P_DEC(d, 4)], c.prm[P_DEC(d, 5)], Co, DEC_K[d], DEC_S[d], DEC_LT[d], out, nullptr, 1));
    }
  }
  return TRU_OK;
}

// ---- backward building blocks ------------------------------------------------------
void set_mask(IgemmParams& p, Ctx& c, const Act& a, bool stats) {
  p.use_mask = 1; p.zmask = a.z; p.mp0 = a.p0; p.mp2 = a.p2;
  if (stats && a.bn >= 0) { p.bstats = c.bn[a.bn].bstats; p.bmean = c.bn[a.bn].mean; p.binv = c.bn[a.bn].inv; }
}

// GRU weight / bias gradient dW (N,C) += dz^T a, db += column sums of dz, on the streaming weight-gradient kernel: a and dz
// may be channel slices of wider rows (gate / direction slices) and a may be shifted by a_add rows inside its frame (h_{t-1}).
// Returns 1 if the shape does not fit the kernel.
int gru_wgrad(Ctx& c, const Act& a, int a_L, int a_ld, int a_coff, int a_add, int C, const float* dz, int z_L, int z_ld, int z_coff,
              int N, float* dW, float* db, int BT, int Lq) {
  WgStream w{};
  w.scratch = c.F(c.plan.wgScratch); w.scratch_floats = WG_SCRATCH_FLOATS;
  w.nsrc = 1; w.a_src[0] = a.z; w.a_p0[0] = a.p0; w.a_p2[0] = a.p2; w.a_L[0] = a_L; w.a_ld[0] = a_ld; w.a_coff[0] = a_coff;
  w.a_add[0] = a_add; w.a_C[0] = C; w.wbase[0] = 0;
  w.z_src = dz; w.z_L = z_L; w.z_ld = z_ld; w.z_coff = z_coff; w.N = N; w.ntap = 1; w.zs = 1; w.zpad = 0;
  w.dW = dW; w.wsc = 1; w.wsn = C; w.wtap = 0; w.db = db; w.BT = BT; w.Lq = Lq;
  return launch_wgrad_stream(w, c.st);
}

// pointwise conv backward: weight grads (+bias), data grads to x1 (masked, BN sums) and to skip (raw)
int pw_bwd(Ctx& c, const Grad& g, const Act& x1, int padL, const Act* skip, int pw_param, float* dX1, bool mask_x1,
           const float* extra, float* dSkip) {
  const float* W = c.prm[pw_param];
  const int N = g.C, K = x1.C + (skip ? skip->C : 0);
  int streamed = 1;
  {
    WgStream w{};
    w.scratch = c.F(c.plan.wgScratch); w.scratch_floats = WG_SCRATCH_FLOATS;
    w.nsrc = skip ? 2 : 1;
    w.a_src[0] = x1.z; w.a_p0[0] = x1.p0; w.a_p2[0] = x1.p2; w.a_L[0] = x1.L; w.a_ld[0] = x1.C; w.a_add[0] = -padL; w.a_C[0] = x1.C; w.wbase[0] = 0;
    if (skip) {
      w.a_src[1] = skip->z; w.a_p0[1] = skip->p0; w.a_p2[1] = skip->p2; w.a_L[1] = skip->L; w.a_ld[1] = skip->C; w.a_add[1] = 0;
      w.a_C[1] = skip->C; w.wbase[1] = x1.C;
    }
    w.z_src = g.dy; w.z_src2 = g.q0 ? g.z : nullptr; w.z_p0 = g.q0; w.z_p1 = g.q1; w.z_p2 = g.q2;
    w.z_L = g.L; w.z_ld = N; w.N = N; w.ntap = 1; w.zs = 1; w.zpad = 0;
    w.dW = c.grd[pw_param]; w.wsc = 1; w.wsn = K; w.wtap = 0; w.db = c.grd[pw_param + 1];
    w.BT = (int)c.BT; w.Lq = g.L;
    streamed = launch_wgrad_stream(w, c.st);
    if (streamed < 0) return streamed;
  }
  if (streamed == 1) {
    WgradParams w{};
    WgradJob& j0 = w.job[0];
    j0.a_src = x1.z; j0.a_p0 = x1.p0; j0.a_p2 = x1.p2; j0.a_relu = 1; j0.a_L = x1.L; j0.a_ld = x1.C; j0.a_coff = 0;
    j0.a_mul = 1; j0.a_add = -padL; j0.C = x1.C;
    j0.z_src = g.dy; j0.z_src2 = g.q0 ? g.z : nullptr; j0.z_p0 = g.q0; j0.z_p1 = g.q1; j0.z_p2 = g.q2;
    j0.z_L = g.L; j0.z_ld = N; j0.z_coff = 0; j0.z_mul = 1; j0.z_add = 0; j0.N = N;
    j0.dW = c.grd[pw_param]; j0.wbase = 0; j0.wsc = 1; j0.wsn = K; j0.db = skip ? nullptr : c.grd[pw_param + 1];
    w.njobs = 1;
    if (skip) {
      WgradJob& j1 = w.job[1];
      j1 = j0;
      j1.a_src = skip->z; j1.a_p0 = skip->p0; j1.a_p2 = skip->p2; j1.a_L = skip->L; j1.a_ld = skip->C; j1.a_add = 0; j1.C = skip->C;
      j1.wbase = x1.C; j1.db = c.grd[pw_param + 1];
      w.njobs = 2;
    }
    if (N < 32) {     // tiny output width: weight grads on the small-shape kernel, bias grad as a column sum
      float* db = c.grd[pw_param + 1];
      for (int j = 0; j < w.njobs; ++j) w.job[j].db = nullptr;
      WgradJob& jb = w.job[w.njobs++];
      jb = w.job[0];
      jb.a_src = nullptr; jb.C = 4; jb.a_ld = 4; jb.dW = nullptr; jb.db = db;
    }
    w.BT = (int)c.BT; w.Lq = g.L;
    TRY(launch_wgrad(w, c.st));
  }
  {
    IgemmParams p{};
    p.seg[0] = bwd_seg(g, 0, N, N, W, 0, K, 1, 1, padL);
    p.nseg = 1; p.BT = (int)c.BT; p.Lq = x1.L; p.N = x1.C;
    p.out = dX1; p.Lout = x1.L; p.ldo = x1.C; p.omul = 1;
    p.extra = extra; p.ext_ld = x1.C;
    if (mask_x1) set_mask(p, c, x1, true);
    TRY(launch_igemm(p, c.st));
  }
  if (skip) {
    IgemmParams p{};
    p.seg[0] = bwd_seg(g, 0, N, N, W, x1.C, K, 1, 1, 0);
    p.nseg = 1; p.BT = (int)c.BT; p.Lq = skip->L; p.N = skip->C;
    p.out = dSkip; p.Lout = skip->L; p.ldo = skip->C; p.omul = 1;
    TRY(launch_igemm(p, c.st));
  }
  return TRU_OK;
}

// Tap-shared form of a transposed conv's data gradient: dX[li] = sum_j dZ[s li + j - pad] W_j^T.  The gradient rows of a frame are
// viewed as g.L / s "wide rows" of s consecutive rows (s * Cout channels: one contiguous run), wide row r = rows s r .. s r + s - 1;
// tap j reads wide row li + delta_j, delta_j = floor((j - pad) / s), channels [h Cout, (h + 1) Cout) with h = (j - pad) - s delta_j.
// When g.L is not a multiple of s the last wide row is incomplete: channel blocks >= c_hi exist for fewer rows (lmax_hi).
// Virtual rows per frame Lq_v = max(L + max delta, rows of the slice a negative-delta tap reads - delta): a tap that leaves its
// frame finds a zero (non-existent) wide row, or feeds one of the Lq_v - L virtual rows that are not stored.
// p = the classic launch (one segment per tap) of the same product; returns false if the tensor-core kernel does not take q.
bool convt_dgrad_shared(IgemmParams& q, const IgemmParams& p, const Grad& g, int xL, int Cout, int k, int s, const float* W) {
  const int pad = s / 2;
  if (!(k >= 2 && k <= 5 && tc_enabled() && (s == 1 || s == 2) && Cout % 32 == 0) || (tap_shared_off() & 2)) return false;
  q = p;
  q.nseg = 1;
  q.seg[0] = bwd_seg(g, 0, s * Cout, s * Cout, W, 0, k, Cout * k, 1, 0);
  const int wide_full = g.L / s, wide_any = (g.L + s - 1) / s;      // wide rows with all s rows / with at least the first
  q.seg[0].Lsrc = wide_any; q.seg[0].fs = g.L * Cout; q.seg[0].cmod = Cout;
  if (wide_full != wide_any) { q.c_hi = Cout; q.lmax_hi = wide_full; }     // (s = 2: the second half of the last wide row is missing)
  Tap taps[5];
  int dmax = 0, Lqv = xL;
  for (int j = 0; j < k; ++j) {
    const int t = j - pad;
    const int delta = t >= 0 ? t / s : -((-t + s - 1) / s);
    const int h = t - s * delta;
    taps[j] = Tap{delta, h * Cout, Cout, j};
    dmax = std::max(dmax, delta);
    if (delta < 0) Lqv = std::max(Lqv, (h == 0 ? wide_any : wide_full) - delta);
  }
  Lqv = std::max(Lqv, xL + dmax);
  set_taps(q, taps, k);
  q.Lq = Lqv; q.Lvalid = xL;
  return igemm_tc_eligible(q);
}

// transposed conv backward: g = grad of the convT output (Lout rows, Cout ch), x = its input activation
int convt_bwd(Ctx& c, const Grad& g, const Act& x, int ct_param, int k, int s, float* dX, const float* planar_dy = nullptr) {
  const float* W = c.prm[ct_param];
  const int Cout = g.C, Cin = x.C, pad = s / 2;
  if (!g.q0 && convt_small_eligible(Cin, Cout, k, s, x.L, g.L)) {
    const float* dy = planar_dy ? planar_dy : g.dy;
    TRY(launch_convt_small_wgrad(x.z, x.p0, x.p2, dy, c.grd[ct_param], c.grd[ct_param + 1], (int)c.BT, x.L, g.L, planar_dy != nullptr, c.st));
    const bool st = x.bn >= 0;
    return launch_convt_small_bwd_data(dy, W, dX, x.z, x.p0, x.p2, st ? c.bn[x.bn].mean : nullptr, st ? c.bn[x.bn].inv : nullptr,
                                       st ? c.bn[x.bn].bstats : nullptr, (int)c.BT, x.L, g.L, planar_dy != nullptr, c.st);
  }
  {
    WgStream ws{};
    ws.scratch = c.F(c.plan.wgScratch); ws.scratch_floats = WG_SCRATCH_FLOATS;
    ws.nsrc = 1;
    ws.a_src[0] = x.z; ws.a_p0[0] = x.p0; ws.a_p2[0] = x.p2; ws.a_L[0] = x.L; ws.a_ld[0] = Cin; ws.a_add[0] = 0; ws.a_C[0] = Cin; ws.wbase[0] = 0;
    ws.z_src = g.dy; ws.z_src2 = g.q0 ? g.z : nullptr; ws.z_p0 = g.q0; ws.z_p1 = g.q1; ws.z_p2 = g.q2;
    ws.z_L = g.L; ws.z_ld = Cout; ws.N = Cout; ws.ntap = k; ws.zs = s; ws.zpad = pad;
    ws.dW = c.grd[ct_param]; ws.wsc = Cout * k; ws.wsn = k; ws.wtap = 1; ws.db = nullptr;
    ws.BT = (int)c.BT; ws.Lq = x.L;
    const int streamed = launch_wgrad_stream(ws, c.st);
    if (streamed < 0) return streamed;
    if (streamed == 1) {
    WgradParams w{};
    for (int j = 0; j < k; ++j) {
      WgradJob& J = w.job[j];
      J.a_src = x.z; J.a_p0 = x.p0; J.a_p2 = x.p2; J.a_relu = 1; J.a_L = x.L; J.a_ld = Cin; J.a_mul = 1; J.a_add = 0; J.C = Cin;
      J.z_src = g.dy; J.z_src2 = g.q0 ? g.z : nullptr; J.z_p0 = g.q0; J.z_p1 = g.q1; J.z_p2 = g.q2;
      J.z_L = g.L; J.z_ld = Cout; J.z_mul = s; J.z_add = j - pad; J.N = Cout;
      J.dW = c.grd[ct_param]; J.wbase = j; J.wsc = Cout * k; J.wsn = k; J.db = nullptr;
    }
    w.njobs = k; w.BT = (int)c.BT; w.Lq = x.L;
    TRY(launch_wgrad(w, c.st));
    }
    if (!g.q0) {      // bias gradient = column sums of dz; behind a BN it comes out of bn_bwd_finalize instead (conv_db)
    WgradParams b{};
    WgradJob& J = b.job[0];
    J.a_src = nullptr; J.C = 4; J.a_ld = 4;
    J.z_src = g.dy; J.z_src2 = nullptr;
    J.z_L = g.L; J.z_ld = Cout; J.z_mul = 1; J.z_add = 0; J.N = Cout; J.dW = nullptr; J.db = c.grd[ct_param + 1];
    b.njobs = 1; b.BT = (int)c.BT; b.Lq = g.L;
    TRY(launch_wgrad(b, c.st));
    }
  }
  IgemmParams p{};
  for (int j = 0; j < k; ++j) p.seg[j] = bwd_seg(g, 0, Cout, Cout, W, j, k, Cout * k, s, j - pad);
  p.nseg = k; p.BT = (int)c.BT; p.Lq = x.L; p.N = Cin;
  p.out = dX; p.Lout = x.L; p.ldo = Cin; p.omul = 1;
  set_mask(p, c, x, true);
  IgemmParams q;
  if (convt_dgrad_shared(q, p, g, x.L, Cout, k, s, W)) return launch_igemm(q, c.st);
  return launch_igemm(p, c.st);
}

int backward(Ctx& c, const float* x, const float* gout) {
  const Plan& P = c.plan;
  const long BT = c.BT;
  TRU_CUDA(cudaMemsetAsync(c.ws + P.bstats, 0, NBN * 384 * 8, c.st));
  const bool small5 = convt_small_eligible(8, 8, DEC_K[5], DEC_S[5], DEC_LP[5], DEC_LT[5]);   // the last block reads the planar gradient directly
  if (!small5) TRY(launch_planar_to_cl(gout, c.F(P.dOUT), (int)BT, 8, 257, c.st));

  // ---- decoder, last block first ----
  for (int d = 5; d >= 0; --d) {
    const int Lp = DEC_LP[d], Co = DEC_COUT[d];
    const int b1 = BN_DEC(d, 0);
    Act pw = c.act(P.ZDp[d], Lp, Co, b1);
    Grad gt = d == 5 ? Grad{c.F(P.dOUT), nullptr, nullptr, nullptr, nullptr, 257, 8}
                     : c.grad(P.dZDt[d], P.ZDt[d], DEC_LT[d], Co, BN_DEC(d, 1));
    if (d < 5) TRY(bn_bfin(c, BN_DEC(d, 1), Co, BT * DEC_LT[d], P_DEC(d, 6), -1, P_DEC(d, 5)));
    TRY(convt_bwd(c, gt, pw, P_DEC(d, 4), DEC_K[d], DEC_S[d], c.F(P.dZDp[d]), (d == 5 && small5) ? gout : nullptr));
    TRY(bn_bfin(c, b1, Co, BT * Lp, P_DEC(d, 2)));
    Grad gp = c.grad(P.dZDp[d], P.ZDp[d], Lp, Co, b1);
    if (d >= 1) {
      const int e = DEC_SKIP[d];
      Act skip = e == 0 ? c.act(P.A0, 128, 64, -1) : c.act(P.Zd[e], ENC_L[e], 128, BN_ENC(e, 1));
      Act x1 = c.act(P.ZDt[d - 1], DEC_LT[d - 1], 64, BN_DEC(d - 1, 1));
      const int diff = Lp - x1.L;
      const int padL = diff >= 0 ? diff / 2 : -((-diff + 1) / 2);
      TRY(pw_bwd(c, gp, x1, padL, &skip, P_DEC(d, 0), c.F(P.dZDt[d - 1]), true, nullptr, c.F(P.dSkip[e])));
    } else {
      Act x1 = c.act(P.ZTp, 16, 64, BN_TGRU);
      TRY(pw_bwd(c, gp, x1, 0, nullptr, P_DEC(0, 0), c.F(P.dZTp), true, nullptr, nullptr));
    }
  }
  // ---- TGRU block ----
  TRY(bn_bfin(c, BN_TGRU, 64, BT * 16, P_TGRU + 6));
  {
    Grad gp = c.grad(P.dZTp, P.ZTp, 16, 64, BN_TGRU);
    Act ht = c.act(P.HT, 16, 128, -1);
    TRY(pw_bwd(c, gp, ht, 0, nullptr, P_TGRU + 4, c.F(P.dHT), false, nullptr, nullptr));
    GruParams g{};
    g.whh[0] = c.prm[P_TGRU + 1]; g.H = c.F(P.HT); g.cache = c.F(P.CT); g.nseq = c.B * 16; g.steps = c.T;
    g.dH = c.F(P.dHT); g.dGi = c.F(P.dGTi); g.dGh = c.F(P.dGTh);
    TRY(launch_tgru_bwd(g, c.B, c.T, c.st));
    Act fo = c.act(P.ZFp, 16, 64, BN_FGRU);
    const int TL16 = c.T * 16;            // a clip is one "frame" of T*16 rows: h_{t-1} of row r is row r - 16
    Act hprev = c.act(P.HT, 16, 128, -1);
    int r1 = gru_wgrad(c, fo, TL16, 64, 0, 0, 64, c.F(P.dGTi), TL16, 384, 0, 384, c.grd[P_TGRU], c.grd[P_TGRU + 2], c.B, TL16);
    if (r1 < 0) return r1;
    int r2 = gru_wgrad(c, hprev, TL16, 128, 0, -16, 128, c.F(P.dGTh), TL16, 384, 0, 384, c.grd[P_TGRU + 1], c.grd[P_TGRU + 3], c.B, TL16);
    if (r2 < 0) return r2;
    if (r1 == 1 || r2 == 1) {
    WgradParams w{};
    w.njobs = 0;
    if (r1 == 1) {   // W_ih, b_ih
      WgradJob& J = w.job[w.njobs++];
      J.a_src = fo.z; J.a_p0 = fo.p0; J.a_p2 = fo.p2; J.a_relu = 1; J.a_L = c.T * 16; J.a_ld = 64; J.a_mul = 1; J.C = 64;
      J.z_src = c.F(P.dGTi); J.z_L = c.T * 16; J.z_ld = 384; J.z_mul = 1; J.N = 384;
      J.dW = c.grd[P_TGRU]; J.wsc = 1; J.wsn = 64; J.db = c.grd[P_TGRU + 2];
    }
    if (r2 == 1) {
    {   // W_hh: h_{t-1} rows are 16 rows up inside a clip
      WgradJob& J = w.job[w.njobs++];
      J.a_src = c.F(P.HT); J.a_L = c.T * 16; J.a_ld = 128; J.a_mul = 1; J.a_add = -16; J.C = 128;
      J.z_src = c.F(P.dGTh); J.z_L = c.T * 16; J.z_ld = 384; J.z_mul = 1; J.N = 384;
      J.dW = c.grd[P_TGRU + 1]; J.wsc = 1; J.wsn = 128;
    }
    {   // b_hh
      WgradJob& J = w.job[w.njobs++];
      J.C = 4; J.a_ld = 4;
      J.z_src = c.F(P.dGTh); J.z_L = c.T * 16; J.z_ld = 384; J.z_mul = 1; J.N = 384; J.db = c.grd[P_TGRU + 3];
    }
    }
    w.BT = c.B; w.Lq = c.T * 16;
    TRY(launch_wgrad(w, c.st));
    }
    IgemmParams p{};
    Grad gi{c.F(P.dGTi), nullptr, nullptr, nullptr, nullptr, 16, 384};
    p.seg[0] = bwd_seg(gi, 0, 384, 384, c.prm[P_TGRU], 0, 64, 1, 1, 0);
    p.nseg = 1; p.BT = (int)BT; p.Lq = 16; p.N = 64; p.out = c.F(P.dZFp); p.Lout = 16; p.ldo = 64; p.omul = 1;
    set_mask(p, c, fo, true);
    TRY(launch_igemm(p, c.st));
  }
  // ---- FGRU block ----
  TRY(bn_bfin(c, BN_FGRU, 64, BT * 16, P_FGRU + 10));
  Act e5 = c.act(P.Zd[5], 16, 128, BN_ENC(5, 1));
  {
    Grad gp = c.grad(P.dZFp, P.ZFp, 16, 64, BN_FGRU);
    Act hf = c.act(P.HF, 16, 128, -1);
    TRY(pw_bwd(c, gp, hf, 0, nullptr, P_FGRU + 8, c.F(P.dHF), false, nullptr, nullptr));
    GruParams g{};
    g.whh[0] = c.prm[P_FGRU + 1]; g.whh[1] = c.prm[P_FGRU + 5]; g.H = c.F(P.HF); g.cache = c.F(P.CF);
    g.nseq = (int)BT; g.steps = 16; g.dH = c.F(P.dHF); g.dGi = c.F(P.dGFi); g.dGh = c.F(P.dGFh);
    TRY(launch_fgru_bwd(g, c.st));
    bool streamed = true;
    Act hfa = c.act(P.HF, 16, 128, -1);
    {   // W_ih, b_ih of both directions in one pass: dGFi rows are [gates fwd (192) | gates bwd (192)], the input is shared
      WgStream w{};
      w.scratch = c.F(c.plan.wgScratch); w.scratch_floats = WG_SCRATCH_FLOATS;
      w.nsrc = 1; w.a_src[0] = e5.z; w.a_p0[0] = e5.p0; w.a_p2[0] = e5.p2; w.a_L[0] = 16; w.a_ld[0] = 128; w.a_C[0] = 128;
      w.z_src = c.F(P.dGFi); w.z_L = 16; w.z_ld = 384; w.N = 384; w.ntap = 1; w.zs = 1;
      w.dW = c.grd[P_FGRU]; w.wsc = 1; w.wsn = 128; w.db = c.grd[P_FGRU + 2];
      w.n_split = 192; w.dW2 = c.grd[P_FGRU + 4]; w.db2 = c.grd[P_FGRU + 6];
      w.BT = (int)BT; w.Lq = 16;
      const int r0 = launch_wgrad_stream(w, c.st);
      if (r0 < 0) return r0;
      if (r0 == 1) streamed = false;
    }
    for (int dir = 0; dir < 2 && streamed; ++dir) {
      int r1 = 0;
      // W_hh, b_hh: h_prev is the neighbouring frequency position, channels 64 dir .. of the 128-wide HF rows
      int r2 = r1 ? 1 : gru_wgrad(c, hfa, 16, 128, 64 * dir, dir ? 1 : -1, 64, c.F(P.dGFh), 16, 384, 192 * dir, 192, c.grd[P_FGRU + 4 * dir + 1],
                                  c.grd[P_FGRU + 4 * dir + 3], (int)BT, 16);
      if (r2 < 0) return r2;
      if (r1 == 1 || r2 == 1) {
        if (dir != 0 || r1 == 0) return set_error(TRU_ERR_ARG, "FGRU weight gradient: streaming kernel took only part of the jobs");
        streamed = false;
      }
    }
    if (!streamed) {
    WgradParams w{};
    for (int dir = 0; dir < 2; ++dir) {
      WgradJob& Ji = w.job[dir];          // W_ih, b_ih
      Ji.a_src = e5.z; Ji.a_p0 = e5.p0; Ji.a_p2 = e5.p2; Ji.a_relu = 1; Ji.a_L = 16; Ji.a_ld = 128; Ji.a_mul = 1; Ji.C = 128;
      Ji.z_src = c.F(P.dGFi); Ji.z_L = 16; Ji.z_ld = 384; Ji.z_coff = 192 * dir; Ji.z_mul = 1; Ji.N = 192;
      Ji.dW = c.grd[P_FGRU + 4 * dir]; Ji.wsc = 1; Ji.wsn = 128; Ji.db = c.grd[P_FGRU + 4 * dir + 2];
      WgradJob& Jh = w.job[2 + dir];      // W_hh: h_prev is the neighbouring frequency position
      Jh.a_src = c.F(P.HF); Jh.a_L = 16; Jh.a_ld = 128; Jh.a_coff = 64 * dir; Jh.a_mul = 1; Jh.a_add = dir ? 1 : -1; Jh.C = 64;
      Jh.z_src = c.F(P.dGFh); Jh.z_L = 16; Jh.z_ld = 384; Jh.z_coff = 192 * dir; Jh.z_mul = 1; Jh.N = 192;
      Jh.dW = c.grd[P_FGRU + 4 * dir + 1]; Jh.wsc = 1; Jh.wsn = 64;
      WgradJob& Jb = w.job[4 + dir];      // b_hh
      Jb.C = 4; Jb.a_ld = 4;
      Jb.z_src = c.F(P.dGFh); Jb.z_L = 16; Jb.z_ld = 384; Jb.z_coff = 192 * dir; Jb.z_mul = 1; Jb.N = 192;
      Jb.db = c.grd[P_FGRU + 4 * dir + 3];
    }
    w.njobs = 6; w.BT = (int)BT; w.Lq = 16;
    TRY(launch_wgrad(w, c.st));
    }
    IgemmParams p{};
    Grad gi{c.F(P.dGFi), nullptr, nullptr, nullptr, nullptr, 16, 384};
    p.seg[0] = bwd_seg(gi, 0, 192, 384, c.prm[P_FGRU], 0, 128, 1, 1, 0);
    p.seg[1] = bwd_seg(gi, 192, 192, 384, c.prm[P_FGRU + 4], 0, 128, 1, 1, 0);
    p.nseg = 2; p.BT = (int)BT; p.Lq = 16; p.N = 128; p.out = c.F(P.dZd[5]); p.Lout = 16; p.ldo = 128; p.omul = 1;
    set_mask(p, c, e5, true);
    TRY(launch_igemm(p, c.st));
  }
  // ---- encoder blocks 5..1 ----
  for (int i = 5; i >= 1; --i) {
    const int Lin = ENC_L[i - 1], Lo = ENC_L[i];
    const int b1 = BN_ENC(i, 0), b2 = BN_ENC(i, 1);
    TRY(bn_bfin(c, b2, 128, BT * Lo, P_ENC(i, 6), -1, P_ENC(i, 5)));   // (+ the depthwise bias gradient, from the sums: it cancels exactly)
    DwParams dp{};
    dp.src = c.F(P.dZd[i]); dp.src2 = c.F(P.Zd[i]); dp.p0 = c.bn[b2].q0; dp.p1 = c.bn[b2].q1; dp.p2 = c.bn[b2].q2;
    dp.w = c.prm[P_ENC(i, 4)]; dp.out = c.F(P.dZp[i]);
    dp.BT = (int)BT; dp.Lin = Lin; dp.Lout = Lo; dp.C = 128; dp.k = ENC_K[i]; dp.stride = ENC_S[i]; dp.pad = ENC_K[i] / 2;
    dp.zmask = c.F(P.Zp[i]); dp.mp0 = c.bn[b1].p0; dp.mp2 = c.bn[b1].p2;
    dp.bmean = c.bn[b1].mean; dp.binv = c.bn[b1].inv; dp.bstats = c.bn[b1].bstats;
    dp.a_src = c.F(P.Zp[i]); dp.a_p0 = c.bn[b1].p0; dp.a_p2 = c.bn[b1].p2;
    dp.dw = c.grd[P_ENC(i, 4)]; dp.db = c.grd[P_ENC(i, 5)];
    TRY(launch_dw_bwd_fused(dp, c.st));
    TRY(bn_bfin(c, b1, 128, BT * Lin, P_ENC(i, 2)));
    Grad gp = c.grad(P.dZp[i], P.Zp[i], Lin, 128, b1);
    if (i >= 2) {
      Act xin = c.act(P.Zd[i - 1], Lin, 128, BN_ENC(i - 1, 1));
      // encoder output i-1 also feeds a decoder skip (blocks 1..4): add that gradient
      const float* extra = (i - 1 >= 1 && i - 1 <= 4) ? c.F(P.dSkip[i - 1]) : nullptr;
      TRY(pw_bwd(c, gp, xin, 0, nullptr, P_ENC(i, 0), c.F(P.dZd[i - 1]), true, extra, nullptr));
    } else {
      Act a0 = c.act(P.A0, 128, 64, -1);
      TRY(pw_bwd(c, gp, a0, 0, nullptr, P_ENC(1, 0), c.F(P.dA0), true, c.F(P.dSkip[0]), nullptr));
    }
  }
  // ---- stem ----
  return launch_enc0_wgrad(x, c.F(P.dA0), c.grd[0], c.grd[1], (int)BT, c.st);
}

int make_ctx(Ctx& c, const TruNetDesc* d, void* ws, size_t ws_bytes, bool bwd, void* stream) {
  TRU_REQUIRE(d && d->batch > 0 && d->n_frames > 0, TRU_ERR_ARG, "trunet: bad descriptor");
  TRU_REQUIRE(ws && aligned16(ws), TRU_ERR_ALIGN, "trunet: workspace must be 16-byte aligned");
  c.d = d; c.st = (cudaStream_t)stream; c.ws = (char*)ws; c.B = d->batch; c.T = d->n_frames; c.BT = (long)d->batch * d->n_frames;
  TRU_REQUIRE(c.BT * 257 < (1L << 31) / 8, TRU_ERR_ARG, "trunet: B*T too large for 32-bit row indices");
  c.plan.build(c.BT, bwd);
  TRU_REQUIRE(ws_bytes >= c.plan.total, TRU_ERR_WORKSPACE, "trunet: workspace too small (%zu < %zu)", ws_bytes, c.plan.total);
  c.slots();
  return TRU_OK;
}

}  // namespace
}  // namespace tru

using namespace tru;

extern "C" size_t tru_trunet_workspace_bytes(const TruNetDesc* d, int with_backward) {
  if (!d || d->batch <= 0 || d->n_frames <= 0) return 0;
  Plan p;
  p.build((long)d->batch * d->n_frames, with_backward != 0);
  return p.total;
}

extern "C" int tru_trunet_forward(const TruNetDesc* d, const float* const* params, float* const* bn_running_mean,
                                  float* const* bn_running_var, long long* const* bn_num_batches, const float* x,
                                  const float* h0, float* out, float* h_last, void* ws, size_t ws_bytes, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  TRU_REQUIRE(params && bn_running_mean && bn_running_var && x && out, TRU_ERR_ARG, "trunet_forward: null pointer");
  for (int i = 0; i < NPARAM; ++i) TRU_REQUIRE(params[i] && aligned16(params[i]), TRU_ERR_ALIGN, "trunet_forward: parameter %d null/unaligned", i);
  Ctx c{};
  if ((rc = make_ctx(c, d, ws, ws_bytes, false, stream))) return rc;
  c.prm = params; c.rmean = bn_running_mean; c.rvar = bn_running_var; c.nbt = bn_num_batches;
  pdl_scope(!d->training);
  rc = forward(c, x, h0, out, h_last);
  pdl_scope(false);
  return rc;
}

extern "C" int tru_trunet_backward(const TruNetDesc* d, const float* const* params, const float* x,
                                   const float* grad_out, float* const* grads, void* ws, size_t ws_bytes, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  TRU_REQUIRE(params && grads && x && grad_out, TRU_ERR_ARG, "trunet_backward: null pointer");
  TRU_REQUIRE(d && d->training, TRU_ERR_ARG, "trunet_backward: needs a training-mode forward (batch statistics)");
  for (int i = 0; i < NPARAM; ++i) TRU_REQUIRE(params[i] && grads[i], TRU_ERR_ARG, "trunet_backward: parameter/grad %d null", i);
  Ctx c{};
  if ((rc = make_ctx(c, d, ws, ws_bytes, true, stream))) return rc;
  c.prm = params; c.grd = grads;
  return backward(c, x, grad_out);
}

// Debug/test aid: byte offset of a named workspace buffer (see tests/test_gpu_network.py).
extern "C" long long tru_trunet_buffer_offset(const TruNetDesc* d, const char* name, int index) {
  if (!d || !name) return -1;
  Plan p;
  p.build((long)d->batch * d->n_frames, true);
  const std::string n(name);
  if (n == "A0") return p.A0;   if (n == "Zp") return p.Zp[index];   if (n == "Zd") return p.Zd[index];
  if (n == "GF") return p.GF;   if (n == "HF") return p.HF;   if (n == "ZFp") return p.ZFp;
  if (n == "GT") return p.GT;   if (n == "HT") return p.HT;   if (n == "ZTp") return p.ZTp;
  if (n == "ZDp") return p.ZDp[index];   if (n == "ZDt") return p.ZDt[index];
  if (n == "dA0") return p.dA0; if (n == "dZp") return p.dZp[index]; if (n == "dZd") return p.dZd[index];
  if (n == "dHF") return p.dHF; if (n == "dHT") return p.dHT; if (n == "dZFp") return p.dZFp; if (n == "dZTp") return p.dZTp;
  if (n == "dZDp") return p.dZDp[index]; if (n == "dZDt") return p.dZDt[index];
  if (n == "dGFi") return p.dGFi; if (n == "dGTi") return p.dGTi; if (n == "dSkip") return p.dSkip[index];
  if (n == "small") return p.small;
  return -1;
}

// Test aid: transposed-conv data gradient through the implicit-GEMM path.
// dy (BT, Lout, Cout) channels-last, w (Cin, Cout, k) -> dx (BT, L, Cin); with z / q0 / q1 / q2 the gradient is
// dz = q0*dy + q1*z + q2 per channel (the BatchNorm-backward affine applied on load); with zmask / mp0 / mp2 the result is
// masked by relu'(mp0*zmask + mp2) and bstats (2 Cin doubles) += sum g, invstd * sum g*(zmask - mean) (the epilogue of the
// in-network launches).
extern "C" int tru_debug_convt_bwd_data(const float* dy, const float* z, const float* q0, const float* q1, const float* q2,
                                        const float* w, float* dx, const float* zmask, const float* mp0, const float* mp2,
                                        const float* bmean, const float* binv, double* bstats, int BT, int L, int Lout,
                                        int Cin, int Cout, int k, int s, int shared, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  IgemmParams p{};
  Grad g{dy, q0 ? z : nullptr, q0, q1, q2, Lout, Cout};
  for (int j = 0; j < k; ++j) p.seg[j] = bwd_seg(g, 0, Cout, Cout, w, j, k, Cout * k, s, j - s / 2);
  p.nseg = k; p.BT = BT; p.Lq = L; p.N = Cin;
  p.out = dx; p.Lout = L; p.ldo = Cin; p.omul = 1;
  if (zmask) { p.use_mask = 1; p.zmask = zmask; p.mp0 = mp0; p.mp2 = mp2; }
  if (bstats) { p.bstats = bstats; p.bmean = bmean; p.binv = binv; }
  IgemmParams q;
  if (shared && convt_dgrad_shared(q, p, g, L, Cout, k, s, w)) return launch_igemm(q, (cudaStream_t)stream);
  if (shared) return set_error(TRU_ERR_ARG, "debug_convt_bwd_data: shape not eligible for the tap-shared path");
  return launch_igemm(p, (cudaStream_t)stream);
}

// Test aid: ConvTranspose1d forward through the implicit-GEMM path (tap-shared launches where eligible).
extern "C" int tru_debug_convt_fwd(const float* x, const float* w, const float* bias, float* out, double* stats, int BT, int L,
                                   int Lout, int Cin, int Cout, int k, int s, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  TRU_REQUIRE(Lout == (L - 1) * s - 2 * (s / 2) + k, TRU_ERR_ARG, "debug_convt_fwd: Lout does not match (L-1)s - 2 pad + k");
  Ctx c{};
  c.BT = BT; c.st = (cudaStream_t)stream;
  Act a{x, L, Cin, nullptr, nullptr, -1};
  return convt_fwd(c, a, w, bias, Cout, k, s, Lout, out, stats, 0);
}

// Test aid: pointwise conv y = act(x) W^T + b through either GEMM path.  x (M,K), w (N,K), out (M,N).
extern "C" int tru_debug_pw(const float* x, const float* p0, const float* p2, const float* w, const float* bias,
                            float* out, double* stats, int M, int K, int N, int use_tc, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  IgemmParams p{};
  Act a{x, 1, K, p0, p2, -1};
  p.seg[0] = fwd_seg(a, w, 0, 1, K, 1, 0);
  p.nseg = 1; p.BT = M; p.Lq = 1; p.N = N; p.bias = bias;
  p.out = out; p.Lout = 1; p.ldo = N; p.omul = 1; p.stats = stats;
  if (use_tc) {
    rc = launch_igemm_tc(p, (cudaStream_t)stream);
    return rc == 1 ? set_error(TRU_ERR_ARG, "debug_pw: shape not eligible for the tensor-core path") : rc;
  }
  return launch_igemm_simt(p, (cudaStream_t)stream);
}
// Test / profiling aid: the backward data-gradient variant of the GEMM kernel on plain operands.
// dx (M,N) = mask(zmask) * ((q0*dy + q1*z + q2) (M,K) @ w (K,N) [+ extra]), BN-backward sums into bstats (2N doubles).
extern "C" int tru_debug_pw_bwd(const float* dy, const float* z, const float* q0, const float* q1, const float* q2,
                                const float* w, float* dx, const float* zmask, const float* mp0, const float* mp2,
                                const float* bmean, const float* binv, double* bstats, const float* extra,
                                int M, int K, int N, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  IgemmParams p{};
  Seg& sg = p.seg[0];
  sg.src = dy; sg.src2 = z; sg.p0 = q0; sg.p1 = q1; sg.p2 = q2; sg.W = w;
  sg.Lsrc = 1; sg.ld = K; sg.coff = 0; sg.C = K; sg.smul = 1; sg.sadd = 0; sg.relu = 0; sg.wbase = 0; sg.wsc = N; sg.wsn = 1;
  p.nseg = 1; p.BT = M; p.Lq = 1; p.N = N;
  p.out = dx; p.Lout = 1; p.ldo = N; p.omul = 1;
  p.extra = extra; p.ext_ld = N;
  if (zmask) { p.use_mask = 1; p.zmask = zmask; p.mp0 = mp0; p.mp2 = mp2; }
  if (bstats) { p.bstats = bstats; p.bmean = bmean; p.binv = binv; }
  rc = launch_igemm_tc(p, (cudaStream_t)stream);
  return rc == 1 ? set_error(TRU_ERR_ARG, "debug_pw_bwd: shape not eligible for the tensor-core path") : rc;
}
extern "C" int tru_set_tensor_cores(int on) { set_tc_enabled(on != 0); return TRU_OK; }
extern "C" int tru_debug_read_mbar(unsigned* out, int n) { return read_mbar_debug(out, n); }
extern "C" int tru_debug_set_flags(int f) { set_tc_debug_flags(f); return TRU_OK; }
extern "C" int tru_debug_set_eval_fusion(int on) { g_eval_fusion = on != 0; return TRU_OK; }

// Test aid: dW (N,C) += z^T a for a (M,C), z (M,N) through the FFMA weight-gradient kernel (the parity reference of the
// streaming tensor-core kernel below).
extern "C" int tru_debug_wgrad(const float* a, const float* z, float* dw, float* db, int M, int C, int N, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  WgradParams w{};
  WgradJob& J = w.job[0];
  J.a_src = a; J.a_L = 1; J.a_ld = C; J.a_mul = 1; J.C = C;
  J.z_src = z; J.z_L = 1; J.z_ld = N; J.z_mul = 1; J.N = N;
  J.dW = dw; J.wsc = 1; J.wsn = C; J.db = db;
  w.njobs = 1; w.BT = M; w.Lq = 1;
  return launch_wgrad_simt(w, (cudaStream_t)stream);
}

// Test aid: the streaming weight-gradient kernel on plain operands.  a (M,C), dy/z (M,N) with dz = q0*dy + q1*z + q2
// (q0 null: dz = dy); rows are grouped in frames of Lq.  dW (N,C), db (N).
extern "C" int tru_debug_wgrad_stream(const float* a, const float* dy, const float* z, const float* q0, const float* q1,
                                      const float* q2, float* dw, float* db, int M, int Lq, int C, int N, void* stream) {
  int rc = ensure_init();
  if (rc) return rc;
  WgStream w{};
  w.nsrc = 1; w.a_src[0] = a; w.a_L[0] = Lq; w.a_ld[0] = C; w.a_C[0] = C;
  w.z_src = dy; w.z_src2 = z; w.z_p0 = q0; w.z_p1 = q1; w.z_p2 = q2; w.z_L = Lq; w.z_ld = N; w.N = N; w.ntap = 1; w.zs = 1;
  w.dW = dw; w.wsc = 1; w.wsn = C; w.db = db; w.BT = M / Lq; w.Lq = Lq;
  rc = launch_wgrad_stream(w, (cudaStream_t)stream);
  return rc == 1 ? set_error(TRU_ERR_ARG, "debug_wgrad_stream: not eligible") : rc;
}
