// Persistent GRU kernels (network.py:45-58, torch.nn.GRU semantics, gate order
// r,z,n; h' = (1-z)*n + z*h; n = tanh(gi_n + r*(W_hn h + b_hn))).
//
//  FGRU: bidirectional GRU(128->64) over the 16 frequency positions of every
//        frame.  A CTA owns 64 sequences of one direction; W_hh (48 KB) stays in
//        shared memory for all 16 steps; each thread keeps its 4x4 slice of h in
//        registers.
//  TGRU: causal GRU(64->128) over time for the B*16 (batch, frequency) sequences.
//        A CTA owns SC sequences for ALL T steps; thread j keeps row j of W_hh
//        (128 floats) in registers, h lives in shared memory and is broadcast.
//        The same kernel with T = 1 and h0/hlast is the streaming step (D11).
// The input projections (W_ih x + b_ih) are hoisted out as one implicit GEMM.
// Forward stores r, z, n and hn = W_hn h + b_hn so that backward-through-time only
// needs one mat-vec per step (dGh @ W_hh).
#include "net_kernels.cuh"

namespace tru {
namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }
// Gate non-linearities from the two SFU primitives (ex2.approx, rcp.approx: ~2 ulp each; absolute error ~1e-7 on outputs in
// [-1, 1], parity tolerance 1e-4).  libm's expf / tanhf and the IEEE division are ~130 instructions per hidden unit, on the serial
// path of every recurrence step.  (Round 1 tried __expf / __frcp_rn and measured a loss: __frcp_rn is a correctly-rounded software
// reciprocal.)  TRU_GRU_LIBM=1 at compile time restores libm.
#ifndef TRU_GRU_LIBM
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoidf_(float x) { return rcpa(1.0f + ex2a(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanhf_(float x) { return fmaf(-2.0f, rcpa(1.0f + ex2a(2.8853900817779268f * x)), 1.0f); }
#else
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return tanhf(x); }
#endif

// =============================== FGRU ========================================
constexpr int FH = 64, FL = 16, FSEQ = 64, FNT = 256;
constexpr int F_WT_LD = 3 * FH + 4;      // forward: WT[k][j]
constexpr int F_H_LD = FH + 4;
constexpr int F_W_LD = FH + 4;           // backward: W[j][k]
constexpr int F_G_LD = 3 * FH + 4;

__global__ void __launch_bounds__(FNT) fgru_fwd_kernel(const __grid_constant__ GruParams p) {
  pdl_trigger();                               // a PDL-launched successor (GEMM) may stage its weights while this runs
  extern __shared__ __align__(16) float smem[];
  float* WT = smem;                       // [64][196]
  float* hs = WT + FH * F_WT_LD;          // [64][68]
  const int tid = threadIdx.x, dir = blockIdx.y;
  const float* whh = p.whh[dir];
  for (int i = tid; i < 3 * FH * FH; i += FNT) {
    const int j = i / FH, k = i % FH;
    WT[k * F_WT_LD + j] = __ldg(whh + i);
  }
  for (int i = tid; i < FSEQ * F_H_LD; i += FNT) hs[i] = 0.f;
  const int ts = tid >> 4, tu = tid & 15;
  const int seq0 = blockIdx.x * FSEQ + ts * 4;
  float4 bh[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) bh[g] = ld4(p.bhh[dir] + g * FH + tu * 4);
  float hown[4][4];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int u = 0; u < 4; ++u) hown[s][u] = 0.f;
  __syncthreads();

  for (int step = 0; step < FL; ++step) {
    const int l = dir ? FL - 1 - step : step;
    float4 gi[4][3];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const bool ok = seq0 + s < p.nseq;
      const float* g = p.G + ((long)(seq0 + s) * FL + l) * (6 * FH) + dir * 3 * FH + tu * 4;
#pragma unroll
      for (int q = 0; q < 3; ++q) gi[s][q] = ok ? ld4(g + q * FH) : make_float4(0, 0, 0, 0);
    }
    // packed FFMA2 (sm_100): two hidden units per instruction, h broadcast into both halves
    float2 acc2[4][3][2];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int g = 0; g < 3; ++g) { acc2[s][g][0] = make_float2(bh[g].x, bh[g].y); acc2[s][g][1] = make_float2(bh[g].z, bh[g].w); }
#pragma unroll 4
    for (int k4 = 0; k4 < FH / 4; ++k4) {
      float4 hv[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) hv[s] = *(const float4*)(hs + (ts * 4 + s) * F_H_LD + k4 * 4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float2 hh[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float h = kk == 0 ? hv[s].x : kk == 1 ? hv[s].y : kk == 2 ? hv[s].z : hv[s].w;
          hh[s] = make_float2(h, h);
        }
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const float4 w = *(const float4*)(WT + (k4 * 4 + kk) * F_WT_LD + g * FH + tu * 4);
          const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
            acc2[s][g][0] = __ffma2_rn(hh[s], w01, acc2[s][g][0]);
            acc2[s][g][1] = __ffma2_rn(hh[s], w23, acc2[s][g][1]);
          }
        }
      }
    }
    float acc[4][3][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int g = 0; g < 3; ++g) { acc[s][g][0] = acc2[s][g][0].x; acc[s][g][1] = acc2[s][g][0].y; acc[s][g][2] = acc2[s][g][1].x; acc[s][g][3] = acc2[s][g][1].y; }
    __syncthreads();                      // everyone has read hs
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float gr[4] = {gi[s][0].x, gi[s][0].y, gi[s][0].z, gi[s][0].w};
      const float gz[4] = {gi[s][1].x, gi[s][1].y, gi[s][1].z, gi[s][1].w};
      const float gn[4] = {gi[s][2].x, gi[s][2].y, gi[s][2].z, gi[s][2].w};
      float r[4], z[4], n[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        r[u] = sigmoidf_(gr[u] + acc[s][0][u]);
        z[u] = sigmoidf_(gz[u] + acc[s][1][u]);
        n[u] = tanhf_(gn[u] + r[u] * acc[s][2][u]);
        hown[s][u] = (1.0f - z[u]) * n[u] + z[u] * hown[s][u];
      }
      const float4 hv = make_float4(hown[s][0], hown[s][1], hown[s][2], hown[s][3]);
      *(float4*)(hs + (ts * 4 + s) * F_H_LD + tu * 4) = hv;
      if (seq0 + s < p.nseq) {
        const long row = (long)(seq0 + s) * FL + l;
        *(float4*)(p.H + row * (2 * FH) + dir * FH + tu * 4) = hv;
        if (p.cache) {
          float* c = p.cache + row * (8 * FH) + dir * 4 * FH + tu * 4;
          *(float4*)(c) = make_float4(r[0], r[1], r[2], r[3]);
          *(float4*)(c + FH) = make_float4(z[0], z[1], z[2], z[3]);
          *(float4*)(c + 2 * FH) = make_float4(n[0], n[1], n[2], n[3]);
          *(float4*)(c + 3 * FH) = make_float4(acc[s][2][0], acc[s][2][1], acc[s][2][2], acc[s][2][3]);
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(FNT) fgru_bwd_kernel(const __grid_constant__ GruParams p) {
  extern __shared__ __align__(16) float smem[];
  float* W = smem;                        // [192][68]
  float* dg = W + 3 * FH * F_W_LD;        // [64][196]
  const int tid = threadIdx.x, dir = blockIdx.y;
  const float* whh = p.whh[dir];
  for (int i = tid; i < 3 * FH * FH; i += FNT) W[(i / FH) * F_W_LD + (i % FH)] = __ldg(whh + i);
  const int ts = tid >> 4, tu = tid & 15;
  const int seq0 = blockIdx.x * FSEQ + ts * 4;
  float carry[4][4];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int u = 0; u < 4; ++u) carry[s][u] = 0.f;
  __syncthreads();

  for (int step = 0; step < FL; ++step) {
    const int l = dir ? step : FL - 1 - step;           // reverse of the forward order
    const int lp = dir ? l + 1 : l - 1;                 // position of h_prev
    float dd[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      float4 dr = make_float4(0, 0, 0, 0), dz = dr, dn = dr, dhn = dr;
      if (seq0 + s < p.nseq) {
        const long row = (long)(seq0 + s) * FL + l;
        const float4 dh4 = ld4(p.dH + row * (2 * FH) + dir * FH + tu * 4);
        const float* c = p.cache + row * (8 * FH) + dir * 4 * FH + tu * 4;
        const float4 r4 = ld4(c), z4 = ld4(c + FH), n4 = ld4(c + 2 * FH), hn4 = ld4(c + 3 * FH);
        float4 hp4 = make_float4(0, 0, 0, 0);
        if (lp >= 0 && lp < FL) hp4 = ld4(p.H + ((long)(seq0 + s) * FL + lp) * (2 * FH) + dir * FH + tu * 4);
        const float dh[4] = {dh4.x + carry[s][0], dh4.y + carry[s][1], dh4.z + carry[s][2], dh4.w + carry[s][3]};
        const float r[4] = {r4.x, r4.y, r4.z, r4.w}, z[4] = {z4.x, z4.y, z4.z, z4.w};
        const float n[4] = {n4.x, n4.y, n4.z, n4.w}, hn[4] = {hn4.x, hn4.y, hn4.z, hn4.w};
        const float hp[4] = {hp4.x, hp4.y, hp4.z, hp4.w};
        float o[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float dnu = dh[u] * (1.0f - z[u]) * (1.0f - n[u] * n[u]);
          o[0][u] = dnu * hn[u] * r[u] * (1.0f - r[u]);             // d pre-activation r
          o[1][u] = dh[u] * (hp[u] - n[u]) * z[u] * (1.0f - z[u]);  // d pre-activation z
          o[2][u] = dnu;                                            // d (gi_n)
          o[3][u] = dnu * r[u];                                     // d (hn)
          dd[s][u] = dh[u] * z[u];
        }
        dr = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
        dz = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
        dn = make_float4(o[2][0], o[2][1], o[2][2], o[2][3]);
        dhn = make_float4(o[3][0], o[3][1], o[3][2], o[3][3]);
        float* gi = p.dGi + row * (6 * FH) + dir * 3 * FH + tu * 4;
        float* gh = p.dGh + row * (6 * FH) + dir * 3 * FH + tu * 4;
        *(float4*)(gi) = dr; *(float4*)(gi + FH) = dz; *(float4*)(gi + 2 * FH) = dn;
        *(float4*)(gh) = dr; *(float4*)(gh + FH) = dz; *(float4*)(gh + 2 * FH) = dhn;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) dd[s][u] = 0.f;
      }
      float* d = dg + (ts * 4 + s) * F_G_LD + tu * 4;
      *(float4*)(d) = dr; *(float4*)(d + FH) = dz; *(float4*)(d + 2 * FH) = dhn;
    }
    __syncthreads();
    float2 acc2[4][2];                    // packed FFMA2: two hidden units per instruction
#pragma unroll
    for (int s = 0; s < 4; ++s) acc2[s][0] = acc2[s][1] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int j4 = 0; j4 < 3 * FH / 4; ++j4) {
      float4 gv[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) gv[s] = *(const float4*)(dg + (ts * 4 + s) * F_G_LD + j4 * 4);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 w = *(const float4*)(W + (j4 * 4 + jj) * F_W_LD + tu * 4);
        const float2 w01 = make_float2(w.x, w.y), w23 = make_float2(w.z, w.w);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float g = jj == 0 ? gv[s].x : jj == 1 ? gv[s].y : jj == 2 ? gv[s].z : gv[s].w;
          const float2 gg = make_float2(g, g);
          acc2[s][0] = __ffma2_rn(gg, w01, acc2[s][0]);
          acc2[s][1] = __ffma2_rn(gg, w23, acc2[s][1]);
        }
      }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      carry[s][0] = dd[s][0] + acc2[s][0].x; carry[s][1] = dd[s][1] + acc2[s][0].y;
      carry[s][2] = dd[s][2] + acc2[s][1].x; carry[s][3] = dd[s][3] + acc2[s][1].y;
    }
    __syncthreads();
  }
}

// =============================== TGRU ========================================
// One CTA carries SC sequences (1, 2 or 4: tgru_seqs_per_cta) through all T steps; W_hh (384 x 128) lives in the registers of its
// 512 threads (96 weights each) for the whole launch.  The first layout of this round - 768 threads at the 80-register cap,
// half a row of W_hh per thread - left room for ONE float4 of h per thread: SASS had LDS.128 -> four FFMAs chained on one
// accumulator -> the next LDS.128 into the same registers, a serial load-latency + FMA-latency chain 64 times per step that six
// warps per scheduler could not cover (2.56 us per step at 4 sequences per CTA, linear in the sequence count:
// profiles/r02_tgru_probe_old_layout.log).  The kernels below give every thread 128 registers, feed 12 / 16 FFMAs on 3 / 4
// independent accumulators from every float4 loaded, reduce across the k- / j-slices with shuffles so that one lane OWNS each
// (sequence, unit) element - its gate arithmetic, previous state and carry stay in registers - and need one barrier per step.
constexpr int TH = 128, TL = 16;

// Per-step operands (input gates / saved gates) are staged through shared memory with cp.async two steps ahead
// (register prefetches were spilled by the compiler, which turned every step into a synchronous wait for HBM).
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- forward -----------------------------------------------------------------------------------------------------------------
// 512 threads (128 registers each), thread = (hidden unit j, k-slice kq): the THREE gate rows j, 128 + j, 256 + j
// of W_hh over columns [32 kq, 32 kq + 32) = 96 weights.  One float4 of h feeds 12 FFMAs on 3 independent accumulators
// (x SC sequences), the loads of the next sequence are in flight meanwhile.  The four k-slices of a unit are adjacent
// lanes: a two-stage shuffle reduce-scatter leaves lane kq with the three gate sums of SEQUENCE kq, which then does that
// sequence's gate arithmetic in registers - no partial sums through shared memory, ONE barrier per step (h is double
// buffered, the input gates are staged two steps ahead in a three-buffer ring).
constexpr int T3NT = 512;
constexpr int T3_HLD = TH + 12;       // h row: k-slice kq starts at 36 kq floats (16-byte aligned, the four float4s of a warp's load hit disjoint banks)
constexpr int T3_GLD = 3 * TH + 8;    // input-gate row stride: the 4 sequences of a warp's gate phase read disjoint banks
__device__ __forceinline__ int t3_hidx(int k) { return k + (k >> 5) * 4; }

template <int SC>
__global__ void __launch_bounds__(T3NT, 1) tgru_fwd3_kernel(const __grid_constant__ GruParams p, int B, int T) {
  pdl_trigger();
  constexpr int NCH = SC * 96;                         // 16-byte chunks of one step's input gates
  __shared__ __align__(16) float hs[2][SC][T3_HLD];
  __shared__ __align__(16) float gin[3][SC][T3_GLD];
  const int tid = threadIdx.x, j = tid >> 2, kq = tid & 3;
  float w[3][32];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
      const float4 v = ld4(p.whh[0] + (long)(g * TH + j) * TH + kq * 32 + k4 * 4);
      w[g][k4 * 4] = v.x; w[g][k4 * 4 + 1] = v.y; w[g][k4 * 4 + 2] = v.z; w[g][k4 * 4 + 3] = v.w;
    }
  const int nseq = B * TL;
  const int sbase = blockIdx.x * SC;
  const int sidx = sbase + kq;                         // the sequence whose gates this lane computes
  const bool gate = kq < SC, ok = gate && sidx < nseq;
  const long ibase = ok ? ((long)(sidx / TL) * T) * TL + sidx % TL : 0;      // row(t) = ibase + 16 t
  const float b_r = gate ? __ldg(p.bhh[0] + j) : 0.f, b_z = gate ? __ldg(p.bhh[0] + TH + j) : 0.f,
              b_n = gate ? __ldg(p.bhh[0] + 2 * TH + j) : 0.f;
  float hprev = (ok && p.h0) ? __ldg(p.h0 + (long)sidx * TH + j) : 0.f;
  if (gate) hs[0][kq][t3_hidx(j)] = hprev;
  // copy plan (one 16-byte chunk per thread and step): chunk tid -> sequence tid / 96, chunk tid % 96 of its 384-float gate row
  static_assert(NCH <= T3NT, "one chunk per thread");
  const int cs = tid / 96, cq = tid % 96, csi = sbase + cs;
  const bool cok = tid < NCH && csi < nseq;
  const float* csrc = cok ? p.G + (((long)(csi / TL) * T) * TL + csi % TL) * (3 * TH) + cq * 4 : nullptr;   // step 0; + 16 rows per step
  const int cdst = cs * T3_GLD + cq * 4;
  auto stage = [&](int t, int buf) {
    if (cok) cp_async16(&gin[buf][0][0] + cdst, csrc + (long)t * (TL * 3 * TH));
    cp_async_commit();
  };
  stage(0, 0);
  if (T > 1) stage(1, 1); else cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  const int b0 = kq & 1, b1 = kq >> 1;
  int gb = 0, nb = 2;                                  // ring slots of step t and of step t + 2
  for (int t = 0; t < T; ++t) {
    if (t + 2 < T) stage(t + 2, nb); else cp_async_commit();      // always one group per step (uniform wait below)
    const float* hcur = &hs[t & 1][0][0] + kq * 36;
    float acc[3][SC];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int s = 0; s < SC; ++s) acc[g][s] = 0.f;
#pragma unroll
    for (int k4 = 0; k4 < 8; ++k4) {
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        const float4 h = *(const float4*)(hcur + s * T3_HLD + k4 * 4);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          acc[g][s] = fmaf(w[g][k4 * 4], h.x, acc[g][s]); acc[g][s] = fmaf(w[g][k4 * 4 + 1], h.y, acc[g][s]);
          acc[g][s] = fmaf(w[g][k4 * 4 + 2], h.z, acc[g][s]); acc[g][s] = fmaf(w[g][k4 * 4 + 3], h.w, acc[g][s]);
        }
      }
    }
    // reduce over the 4 k-slices (lanes kq = 0..3 of a unit), scattered: lane kq ends with the sums of sequence kq.
    // The pairing is (slice kq + slice kq^1) + (slice kq^2 + slice kq^3) for every sequence and every SC.
    float tot[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      if constexpr (SC == 4) {
        const float m0 = b0 ? acc[g][1] : acc[g][0], s0 = b0 ? acc[g][0] : acc[g][1];
        const float m1 = b0 ? acc[g][3] : acc[g][2], s1 = b0 ? acc[g][2] : acc[g][3];
        const float k0 = m0 + __shfl_xor_sync(0xffffffffu, s0, 1);          // sequence b0, slices {kq, kq^1}
        const float k1 = m1 + __shfl_xor_sync(0xffffffffu, s1, 1);          // sequence 2 + b0
        const float mine = b1 ? k1 : k0, send = b1 ? k0 : k1;
        tot[g] = mine + __shfl_xor_sync(0xffffffffu, send, 2);             // sequence 2 b1 + b0 = kq
      } else if constexpr (SC == 2) {
        const float mine = b0 ? acc[g][1] : acc[g][0], send = b0 ? acc[g][0] : acc[g][1];
        const float k = mine + __shfl_xor_sync(0xffffffffu, send, 1);       // sequence b0
        tot[g] = k + __shfl_xor_sync(0xffffffffu, k, 2);
      } else {
        const float k = acc[g][0] + __shfl_xor_sync(0xffffffffu, acc[g][0], 1);
        tot[g] = k + __shfl_xor_sync(0xffffffffu, k, 2);
      }
    }
    if (gate) {
      const float* gi = gin[gb][kq];
      const float hn = tot[2] + b_n;
      const float rr = sigmoidf_(gi[j] + (tot[0] + b_r));
      const float zz = sigmoidf_(gi[TH + j] + (tot[1] + b_z));
      const float nn = tanhf_(gi[2 * TH + j] + rr * hn);
      const float hnew = (1.0f - zz) * nn + zz * hprev;
      hprev = hnew;
      hs[(t + 1) & 1][kq][t3_hidx(j)] = hnew;
      if (ok) {
        const long row = ibase + (long)t * TL;
        p.H[row * TH + j] = hnew;
        if (p.cache) {
          float* c = p.cache + row * (4 * TH) + j;
          c[0] = rr; c[TH] = zz; c[2 * TH] = nn; c[3 * TH] = hn;
        }
      }
    }
    cp_async_wait<1>();                                 // the gates of step t + 1 (staged one step ago) have landed
    __syncthreads();                                    // ... for everybody; and h of step t + 1 is complete
    gb = gb == 2 ? 0 : gb + 1;
    nb = nb == 2 ? 0 : nb + 1;
  }
  if (ok && p.hlast) p.hlast[(long)sidx * TH + j] = hprev;
}

// ---- backward (same idea as tgru_fwd3_kernel) --------------------------------------------------------------------------------
// 512 threads, thread = (4 consecutive outputs k = 4 kp .. 4 kp + 3, j-slice jq of 16): W_hh[24 jq .. 24 jq + 24)[4 kp .. + 4) =
// 96 weights; one float4 of gate gradients feeds 16 FFMAs.  The 16 j-slices of an output group are the lanes of a half warp:
// four shuffle scatter stages (xor 1, 2: the four outputs; xor 4, 8: the sequences) leave every lane with dh_prev of ONE
// (sequence, unit) = (jq >> 2, 4 kp + (jq & 3)), and that lane also does this element's gate-gradient arithmetic: its carry
// never leaves the register.  Gate gradients are double buffered, the per-step operands (dH | r | z | n | hn | h_prev) are staged
// two steps ahead in a three-slot ring: ONE barrier per step (three before).
constexpr int TB_DLD = 3 * TH + 64;   // gate-gradient row: j-slice jq starts at 28 jq floats
constexpr int TB_SLD = 6 * TH + 8;    // staged-operand row stride: the sequences of a warp's element phase read disjoint banks
__device__ __forceinline__ int tb_didx(int j) { return j + (j / 24) * 4; }

template <int N>
__device__ __forceinline__ void tb_scatter(const float (&in)[2 * N], float (&out)[N], int bit, int lane_xor) {
#pragma unroll
  for (int n = 0; n < N; ++n) {
    const float keep = bit ? in[2 * n + 1] : in[2 * n], send = bit ? in[2 * n] : in[2 * n + 1];
    out[n] = keep + __shfl_xor_sync(0xffffffffu, send, lane_xor);
  }
}

template <int SC>
__global__ void __launch_bounds__(T3NT, 1) tgru_bwd3_kernel(const __grid_constant__ GruParams p, int B, int T) {
  extern __shared__ __align__(16) float tgru_bwd_smem[];
  float (*dgs)[SC][TB_DLD] = (float (*)[SC][TB_DLD])tgru_bwd_smem;                        // [2]
  float (*sv)[SC][TB_SLD] = (float (*)[SC][TB_SLD])(tgru_bwd_smem + 2 * SC * TB_DLD);     // [3]
  const int tid = threadIdx.x, kp = tid >> 4, jq = tid & 15;
  float w[24][4];
#pragma unroll
  for (int jj = 0; jj < 24; ++jj) {
    const float4 v = ld4(p.whh[0] + (long)(24 * jq + jj) * TH + 4 * kp);
    w[jj][0] = v.x; w[jj][1] = v.y; w[jj][2] = v.z; w[jj][3] = v.w;
  }
  const int nseq = B * TL;
  const int sbase = blockIdx.x * SC;
  const int c0 = jq & 1, c1 = (jq >> 1) & 1, c2 = (jq >> 2) & 1, c3 = jq >> 3;
  // the element this lane owns
  const int own_s = SC == 4 ? (jq >> 2) : (SC == 2 ? c2 : 0);
  const int own_u = 4 * kp + (jq & 3);
  const bool owner = SC == 4 ? true : (SC == 2 ? c3 == 0 : (jq >> 2) == 0);
  const int sidx = sbase + own_s;
  const bool ok = owner && sidx < nseq;
  const long ibase = ok ? ((long)(sidx / TL) * T) * TL + sidx % TL : 0;
  // copy plan (stage() is called for t = T - 1, T - 2, ... in order, so the source offsets just run backwards; 32-bit, in float4 units):
  //   A: every thread < 128 SC: cache chunk (sequence tid >> 7, float4 tid & 127) -> r | z | n | hn
  //   B: threads < 32 SC: dH chunk (sequence tid >> 5, float4 tid & 31); threads in [32 SC, 64 SC): the same of h_prev = H of step t - 1
  static_assert(SC * 128 <= T3NT, "one cache chunk per thread");
  const int sA = tid >> 7, qA = tid & 127;
  const bool okA = tid < SC * 128 && sbase + sA < nseq;
  const bool isHp = tid >= 32 * SC;
  const int tB = isHp ? tid - 32 * SC : tid, sB = tB >> 5, qB = tB & 31;
  const bool okB = tid < 64 * SC && sbase + sB < nseq;
  int offA = 0, offB = 0;
  if (okA) { const int si = sbase + sA; offA = (((si / TL) * T + (T - 1)) * TL + si % TL) * 128 + qA; }
  if (okB) { const int si = sbase + sB; offB = (((si / TL) * T + (T - 1)) * TL + si % TL - (isHp ? TL : 0)) * 32 + qB; }
  const float* baseB = isHp ? p.H : p.dH;
  const int dA = sA * TB_SLD + TH + qA * 4, dB = sB * TB_SLD + (isHp ? 5 * TH : 0) + qB * 4;
  auto stage = [&](int t, int buf) {
    float* s0 = &sv[buf][0][0];
    if (okA) cp_async16(s0 + dA, p.cache + 4 * (long)offA);
    if (okB) {
      if (isHp && t == 0) {                              // the state before the first step: h0 or zero
        if (p.h0) cp_async16(s0 + dB, p.h0 + (long)(sbase + sB) * TH + qB * 4);
        else *(float4*)(s0 + dB) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        cp_async16(s0 + dB, baseB + 4 * (long)offB);
      }
    }
    offA -= TL * 128; offB -= TL * 32;
    cp_async_commit();
  };
  const int d0 = tb_didx(own_u), d1 = tb_didx(TH + own_u), d2 = tb_didx(2 * TH + own_u);
  float* gi = p.dGi + (ibase + (long)(T - 1) * TL) * (3 * TH) + own_u;          // the owner's output row of step t, running backwards
  const long ghd = p.dGh - p.dGi;
  float carry = 0.f;
  stage(T - 1, 0);
  if (T > 1) stage(T - 2, 1); else cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  int gb = 0, nb = 2;                                  // ring slots of step t and of step t - 2
  for (int t = T - 1; t >= 0; --t) {
    if (t >= 2) stage(t - 2, nb); else cp_async_commit();
    float dd = 0.f;
    if (owner) {
      const float* v = &sv[gb][own_s][own_u];
      const float v_dh = v[0], v_r = v[TH], v_z = v[2 * TH], v_n = v[3 * TH], v_hn = v[4 * TH], v_hp = v[5 * TH];
      const float dh = v_dh + carry;
      const float dn = dh * (1.0f - v_z) * (1.0f - v_n * v_n);
      const float dr = dn * v_hn * v_r * (1.0f - v_r);
      const float dz = dh * (v_hp - v_n) * v_z * (1.0f - v_z);
      const float dhn = dn * v_r;
      dd = dh * v_z;
      float* d = &dgs[t & 1][own_s][0];
      d[d0] = dr; d[d1] = dz; d[d2] = dhn;
      if (ok) {
        gi[0] = dr; gi[TH] = dz; gi[2 * TH] = dn;
        float* gh = gi + ghd;
        gh[0] = dr; gh[TH] = dz; gh[2 * TH] = dhn;
        gi -= TL * 3 * TH;
      }
    }
    cp_async_wait<1>();                                 // the operands of step t - 1 (staged one step ago) have landed
    __syncthreads();                                    // ... for everybody; and this step's gate gradients are complete
    const float* dcur = &dgs[t & 1][0][0] + 28 * jq;
    float acc[SC * 4];
#pragma unroll
    for (int i = 0; i < SC * 4; ++i) acc[i] = 0.f;
#pragma unroll
    for (int q6 = 0; q6 < 6; ++q6) {
#pragma unroll
      for (int s = 0; s < SC; ++s) {
        const float4 g = *(const float4*)(dcur + s * TB_DLD + q6 * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = acc[s * 4 + i];
          a = fmaf(w[q6 * 4][i], g.x, a); a = fmaf(w[q6 * 4 + 1][i], g.y, a);
          a = fmaf(w[q6 * 4 + 2][i], g.z, a); a = fmaf(w[q6 * 4 + 3][i], g.w, a);
          acc[s * 4 + i] = a;
        }
      }
    }
    // reduce over the 16 j-slices, scattered: xor 1 / 2 pick the output (jq & 3), xor 4 / 8 the sequence (jq >> 2)
    float r1[SC * 2], r2[SC];
    tb_scatter<SC * 2>(acc, r1, c0, 1);
    tb_scatter<SC>(r1, r2, c1, 2);
    float dhp;
    if constexpr (SC == 4) {
      float r3[2], r4[1];
      tb_scatter<2>(r2, r3, c2, 4);
      tb_scatter<1>(r3, r4, c3, 8);
      dhp = r4[0];
    } else if constexpr (SC == 2) {
      float r3[1];
      tb_scatter<1>(r2, r3, c2, 4);
      dhp = r3[0] + __shfl_xor_sync(0xffffffffu, r3[0], 8);
    } else {
      const float k = r2[0] + __shfl_xor_sync(0xffffffffu, r2[0], 4);
      dhp = k + __shfl_xor_sync(0xffffffffu, k, 8);
    }
    carry = dd + dhp;
    gb = gb == 2 ? 0 : gb + 1;
    nb = nb == 2 ? 0 : nb + 1;
  }
}

// Single-step TGRU (streaming inference, T = 1): the hidden projection G_h = h W_hh^T + b_hh is one dense GEMM over all
// sequences (done by the caller on the tensor-core kernel) instead of nseq / 4 CTAs each pulling the 192 KB W_hh
// through L2; this kernel is the gate arithmetic.  Thread = (sequence, 4 hidden units).
__global__ void __launch_bounds__(256) tgru_step_gates_kernel(const float* __restrict__ Gi, const float* __restrict__ Gh,
                                                              const float* __restrict__ bhh, const float* __restrict__ h0,
                                                              float* __restrict__ H, float* __restrict__ hlast, long nseq) {
  pdl_trigger();                               // a PDL-launched successor (GEMM) may stage its weights while this runs
  const long idx = (long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= nseq * (TH / 4)) return;
  const long s = idx / (TH / 4);
  const int u = (int)(idx % (TH / 4)) * 4;
  const float* gi = Gi + s * (3 * TH) + u;
  const float4 ir = ld4(gi), iz = ld4(gi + TH), in = ld4(gi + 2 * TH);
  float4 hr, hz, hn, hp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (Gh) {
    const float* gh = Gh + s * (3 * TH) + u;
    hr = ld4(gh); hz = ld4(gh + TH); hn = ld4(gh + 2 * TH);
    hp = ld4(h0 + s * TH + u);
  } else {
    hr = ld4(bhh + u); hz = ld4(bhh + TH + u); hn = ld4(bhh + 2 * TH + u);
  }
  float4 o;
#define TRU_GATE(c) { const float rr = sigmoidf_(ir.c + hr.c), zz = sigmoidf_(iz.c + hz.c), nn = tanhf_(in.c + rr * hn.c); \
                      o.c = (1.0f - zz) * nn + zz * hp.c; }
  TRU_GATE(x) TRU_GATE(y) TRU_GATE(z) TRU_GATE(w)
#undef TRU_GATE
  *(float4*)(H + s * TH + u) = o;
  if (hlast) *(float4*)(hlast + s * TH + u) = o;
}

}  // namespace

int launch_tgru_step_gates(const float* Gi, const float* Gh, const float* bhh, const float* h0, float* H, float* hlast, long nseq,
                           cudaStream_t st) {
  ProfScope prof("tgru_step_gates", 4.0 * nseq * (384.0 * (Gh ? 2 : 1) + 128.0 * (Gh ? 3 : 2)), 0.0, st);
  tgru_step_gates_kernel<<<(unsigned)((nseq * (TH / 4) + 255) / 256), 256, 0, st>>>(Gi, Gh, bhh, h0, H, hlast, nseq);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

int launch_fgru_fwd(const GruParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(FH * F_WT_LD + FSEQ * F_H_LD) * 4;
  TRU_CUDA(cudaFuncSetAttribute(fgru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p.nseq + FSEQ - 1) / FSEQ, 2);
  ProfScope prof("fgru_fwd", 4.0 * p.nseq * FL * (384 + 128 + 512), 2.0 * p.nseq * FL * 2 * FH * 3 * FH, st);
  fgru_fwd_kernel<<<grid, FNT, smem, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

int launch_fgru_bwd(const GruParams& p, cudaStream_t st) {
  const size_t smem = (size_t)(3 * FH * F_W_LD + FSEQ * F_G_LD) * 4;
  TRU_CUDA(cudaFuncSetAttribute(fgru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((p.nseq + FSEQ - 1) / FSEQ, 2);
  ProfScope prof("fgru_bwd", 4.0 * p.nseq * FL * (128 + 512 + 128 + 768), 2.0 * p.nseq * FL * 2 * FH * 3 * FH, st);
  fgru_bwd_kernel<<<grid, FNT, smem, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

// Sequences per CTA of the TGRU recurrence kernels: one CTA per SM (512 threads x 128 registers), and a step costs
// 0.70 / 1.03 / 1.88 us forward (0.68 / 1.15 / 2.28 backward) with 1 / 2 / 4 sequences (tools/probe_tgru.py) - so small batches
// (one clip = 16 sequences: configs[0], rt.py:20-27) spread over more SMs; from 2 x SMs sequences up it is 4 per CTA (B = 32:
// 128 CTAs, B = 37: one full wave).
static int tgru_seqs_per_cta(int nseq) {
  const int sms = sm_count();
  return nseq <= sms ? 1 : (nseq <= 2 * sms ? 2 : 4);
}

int launch_tgru_fwd(const GruParams& p, int B, int T, cudaStream_t st) {
  const int nseq = B * TL;
  ProfScope prof("tgru_fwd", 4.0 * nseq * T * (384 + 128 + 512), 2.0 * nseq * T * TH * 3 * TH, st);
  switch (tgru_seqs_per_cta(nseq)) {
    case 1: tgru_fwd3_kernel<1><<<nseq, T3NT, 0, st>>>(p, B, T); break;
    case 2: tgru_fwd3_kernel<2><<<(nseq + 1) / 2, T3NT, 0, st>>>(p, B, T); break;
    default: tgru_fwd3_kernel<4><<<(nseq + 3) / 4, T3NT, 0, st>>>(p, B, T); break;
  }
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

int launch_tgru_bwd(const GruParams& p, int B, int T, cudaStream_t st) {
  const int nseq = B * TL;
  ProfScope prof("tgru_bwd", 4.0 * nseq * T * (128 + 512 + 128 + 768), 2.0 * nseq * T * TH * 3 * TH, st);
  const int sc = tgru_seqs_per_cta(nseq);
  const size_t smem = (size_t)sc * (2 * TB_DLD + 3 * TB_SLD) * 4;            // gate gradients x 2 + staged operands x 3: 12,896 B per sequence
  if (sc == 1) {
    tgru_bwd3_kernel<1><<<nseq, T3NT, smem, st>>>(p, B, T);
  } else if (sc == 2) {
    tgru_bwd3_kernel<2><<<(nseq + 1) / 2, T3NT, smem, st>>>(p, B, T);
  } else {
    TRU_SMEM_OPT_IN((tgru_bwd3_kernel<4>), smem);
    tgru_bwd3_kernel<4><<<(nseq + 3) / 4, T3NT, smem, st>>>(p, B, T);
  }
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace tru
