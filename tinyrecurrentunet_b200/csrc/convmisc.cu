// Small network kernels: BatchNorm finalize (fwd/bwd), encoder stem conv
// (network.py:9-21), depthwise convs (network.py:33-40) and a layout transpose.
#include "net_kernels.cuh"

namespace tru {
namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }

// ---- BatchNorm1d (training: biased batch variance, running_var unbiased; eps 1e-5) ----
__global__ void bn_finalize_kernel(BnFwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && p.training && p.nbt) *p.nbt += 1;
  if (c >= p.C) return;
  double mean, var;
  if (p.training) {
    mean = p.stats[c] / p.count;
    var = p.stats[p.C + c] / p.count - mean * mean;
    if (var < 0) var = 0;
    const double unb = p.count > 1 ? var * p.count / (p.count - 1.0) : var;
    p.running_mean[c] = (float)((1.0 - p.momentum) * p.running_mean[c] + p.momentum * mean);
    p.running_var[c] = (float)((1.0 - p.momentum) * p.running_var[c] + p.momentum * unb);
  } else {
    mean = p.running_mean[c];
    var = p.running_var[c];
  }
  const double inv = 1.0 / sqrt(var + (double)p.eps);
  const double g = p.gamma[c];
  p.p0[c] = (float)(g * inv);
  p.p2[c] = (float)((double)p.beta[c] - mean * g * inv);
  p.mean[c] = (float)mean;
  p.invstd[c] = (float)inv;
}

__global__ void bn_finalize_eval_all_kernel(const __grid_constant__ BnEvalAll p) {
  const int l = blockIdx.y, c = threadIdx.x;
  if (c >= p.C[l]) return;
  const double mean = p.rmean[l][c], inv = 1.0 / sqrt((double)p.rvar[l][c] + (double)p.eps), g = p.gamma[l][c];
  p.p0[l][c] = (float)(g * inv);
  p.p2[l][c] = (float)((double)p.beta[l][c] - mean * g * inv);
  p.mean[l][c] = (float)mean;
  p.inv[l][c] = (float)inv;
}

// dZ = q0*dY + q1*Z + q2 with  q0 = g*inv, q1 = -g*inv^2*s2/M, q2 = g*inv*(mean*inv*s2 - s1)/M
__global__ void bn_bwd_finalize_kernel(BnBwdParams p) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  const double s1 = p.bstats[c], s2 = p.bstats[p.C + c];
  const double g = p.gamma[c], inv = p.invstd[c], mean = p.mean[c];
  p.q0[c] = (float)(g * inv);
  p.q1[c] = (float)(-g * inv * inv * s2 / p.count);
  p.q2[c] = (float)(g * inv * (mean * inv * s2 - s1) / p.count);
  p.dgamma[c] += (float)s2;
  p.dbeta[c] += (float)s1;
  if (p.db_acc) p.db_out[c] += (float)p.db_acc[c];
  if (p.conv_db) p.conv_db[c] += (float)(g * inv * inv * s2 * (mean - p.fstats[c] / p.count));
}

// ---- encoder stem: Conv1d(4->64, k5, s2, p1) + ReLU; planar (BT,4,257) in, CL (BT,128,64) out ----
constexpr int E_F = 257, E_LO = 128, E_CO = 64, E_CI = 4, E_K = 5;

// Thread = 4 output channels x 8 output positions; its 80 weights live in registers, the input frame (4 x 257 floats)
// in shared memory, the next frame is prefetched into registers while the current one is computed.
constexpr int E_XN = E_CI * (E_F + 3);               // padded frame in shared memory
constexpr int E_XI = (E_XN + 255) / 256;             // prefetch registers per thread
__global__ void __launch_bounds__(256) enc0_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ b, float* __restrict__ out, int BT) {
  pdl_trigger();                               // a PDL-launched successor (GEMM) may stage its weights while this runs
  __shared__ float xs[E_CI][E_F + 3];        // xs[ci][1 + f], zeros at both ends
  const int tid = threadIdx.x;
  const int c4 = (tid & 15) * 4, l0 = tid >> 4;
  float wr[E_CI * E_K][4];
#pragma unroll
  for (int r = 0; r < E_CI * E_K; ++r)
#pragma unroll
    for (int e = 0; e < 4; ++e) wr[r][e] = __ldg(w + (c4 + e) * (E_CI * E_K) + r);
  const float4 bias = ld4(b + c4);
  float pre[E_XI];
  auto fetch = [&](int bt) {
#pragma unroll
    for (int k = 0; k < E_XI; ++k) {
      const int i = tid + k * 256;
      const int ci = i / (E_F + 3), f = i % (E_F + 3) - 1;
      pre[k] = (i < E_XN && f >= 0 && f < E_F) ? __ldg(x + ((size_t)bt * E_CI + ci) * E_F + f) : 0.f;
    }
  };
  int bt = blockIdx.x;
  if (bt < BT) fetch(bt);
  for (; bt < BT; bt += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < E_XI; ++k) {
      const int i = tid + k * 256;
      if (i < E_XN) (&xs[0][0])[i] = pre[k];
    }
    __syncthreads();
    if (bt + (int)gridDim.x < BT) fetch(bt + gridDim.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int lo = l0 + 16 * i;
      float4 a = bias;
#pragma unroll
      for (int ci = 0; ci < E_CI; ++ci)
#pragma unroll
        for (int j = 0; j < E_K; ++j) {
          const float xv = xs[ci][2 * lo + j];                 // input index 2*lo - 1 + j
          const int r = ci * E_K + j;
          a.x = fmaf(xv, wr[r][0], a.x); a.y = fmaf(xv, wr[r][1], a.y); a.z = fmaf(xv, wr[r][2], a.z); a.w = fmaf(xv, wr[r][3], a.w);
        }
      a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
      *(float4*)(out + ((size_t)bt * E_LO + lo) * E_CO + c4) = a;
    }
  }
}

// dW[co][ci][j] = sum dy[bt][lo][co] * x[bt][ci][2lo-1+j]; db[co] = sum dy.  dy is already ReLU-masked.
// Thread = 4 output channels x one group of 8 output positions, 80 accumulators in registers: one 16-byte load of dy
// feeds 80 FMAs and every input value (a broadcast load) feeds 4; the 16 position groups meet in shared memory at the end.
__global__ void __launch_bounds__(256) enc0_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                         float* __restrict__ dw, float* __restrict__ db, int BT) {
  __shared__ float xs[E_CI][E_F + 3];
  __shared__ float red[16][E_CO];                  // per position group, one output at a time
  const int tid = threadIdx.x;
  const int c4 = (tid & 15) * 4, lg = tid >> 4;    // channels c4..c4+3, positions lg*8 .. lg*8+7
  float acc[E_CI * E_K][4];
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < E_CI * E_K; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  float pre[E_XI];
  auto fetch = [&](int bt) {
#pragma unroll
    for (int k = 0; k < E_XI; ++k) {
      const int i = tid + k * 256;
      const int ci = i / (E_F + 3), f = i % (E_F + 3) - 1;
      pre[k] = (i < E_XN && f >= 0 && f < E_F) ? __ldg(x + ((size_t)bt * E_CI + ci) * E_F + f) : 0.f;
    }
  };
  float4 g[8], gn[8];
  auto fetch_dy = [&](int bt) {
    const float* dyf = dy + (size_t)bt * E_LO * E_CO + c4;
#pragma unroll
    for (int i = 0; i < 8; ++i) gn[i] = ld4(dyf + (size_t)(lg * 8 + i) * E_CO);
  };
  int bt = blockIdx.x;
  if (bt < BT) { fetch(bt); fetch_dy(bt); }
  for (; bt < BT; bt += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int k = 0; k < E_XI; ++k) {
      const int i = tid + k * 256;
      if (i < E_XN) (&xs[0][0])[i] = pre[k];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = gn[i];
    __syncthreads();
    if (bt + (int)gridDim.x < BT) { fetch(bt + gridDim.x); fetch_dy(bt + gridDim.x); }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int lo = lg * 8 + i;
      bacc[0] += g[i].x; bacc[1] += g[i].y; bacc[2] += g[i].z; bacc[3] += g[i].w;
#pragma unroll
      for (int ci = 0; ci < E_CI; ++ci)
#pragma unroll
        for (int j = 0; j < E_K; ++j) {
          const float xv = xs[ci][2 * lo + j];
          const int r = ci * E_K + j;
          acc[r][0] = fmaf(g[i].x, xv, acc[r][0]); acc[r][1] = fmaf(g[i].y, xv, acc[r][1]);
          acc[r][2] = fmaf(g[i].z, xv, acc[r][2]); acc[r][3] = fmaf(g[i].w, xv, acc[r][3]);
        }
    }
  }
  // reduce the 16 position groups: one (ci, j) at a time through shared memory, then one atomic per output and CTA
  for (int r = 0; r <= E_CI * E_K; ++r) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = bacc[e];
#pragma unroll
      for (int q = 0; q < E_CI * E_K; ++q) if (q == r) v = acc[q][e];
      red[lg][c4 + e] = v;
    }
    __syncthreads();
    if (tid < E_CO) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) s += red[q][tid];
      if (r < E_CI * E_K) atomicAdd(dw + tid * (E_CI * E_K) + r, s);
      else atomicAdd(db + tid, s);
    }
  }
}

// ---- depthwise conv, C = 128: thread = (row lane, 4 channels) ----------------------
constexpr int DW_C = 128, DW_ROWS = 8;   // 32 channel-quads x 8 row lanes = 256 threads

__device__ __forceinline__ float4 dw_load(const DwParams& p, long off, int c4, const float4& p0, const float4& p1,
                                          const float4& p2, bool relu) {
  float4 v = ld4(p.src + off + c4);
  if (p.p0) {
    v.x = p0.x * v.x + p2.x; v.y = p0.y * v.y + p2.y; v.z = p0.z * v.z + p2.z; v.w = p0.w * v.w + p2.w;
    if (p.p1) {
      const float4 z = ld4(p.src2 + off + c4);
      v.x += p1.x * z.x; v.y += p1.y * z.y; v.z += p1.z * z.z; v.w += p1.w * z.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  }
  return v;
}

__device__ __forceinline__ void block_chan_reduce(float (&s1)[4], float (&s2)[4], double* g1, double* g2, int c4) {
  __shared__ float red[2][DW_ROWS][DW_C];
  const int rl = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[0][rl][c4 + j] = s1[j]; red[1][rl][c4 + j] = s2[j]; }
  __syncthreads();
  if (threadIdx.x < DW_C) {
    double a = 0, b = 0;
    for (int r = 0; r < DW_ROWS; ++r) { a += red[0][r][threadIdx.x]; b += red[1][r][threadIdx.x]; }
    atomicAdd(g1 + threadIdx.x, a);
    atomicAdd(g2 + threadIdx.x, b);
  }
}

// forward: out[bt,lo,c] = b[c] + sum_j w[c][j] * a(bt, lo*s - pad + j, c)
__global__ void __launch_bounds__(256) dw_fwd_kernel(const __grid_constant__ DwParams p) {
  const int c4 = (threadIdx.x & 31) * 4, rl = threadIdx.x >> 5;
  float4 p0 = make_float4(1, 1, 1, 1), p2 = make_float4(0, 0, 0, 0), p1 = p2;
  if (p.p0) { p0 = ld4(p.p0 + c4); p2 = ld4(p.p2 + c4); }
  float w[4][5];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 5; ++t) w[j][t] = t < p.k ? __ldg(p.w + (c4 + j) * p.k + t) : 0.f;
  const float4 bias = ld4(p.bias + c4);
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  const long M = (long)p.BT * p.Lout;
  for (long m = (long)blockIdx.x * DW_ROWS + rl; m < M; m += (long)gridDim.x * DW_ROWS) {
    const int bt = (int)(m / p.Lout), lo = (int)(m - (long)bt * p.Lout);
    float4 a = bias;
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int li = lo * p.stride - p.pad + t;
      if (t < p.k && li >= 0 && li < p.Lin) {
        const float4 v = dw_load(p, ((long)bt * p.Lin + li) * DW_C, c4, p0, p1, p2, true);
        a.x = fmaf(w[0][t], v.x, a.x); a.y = fmaf(w[1][t], v.y, a.y);
        a.z = fmaf(w[2][t], v.z, a.z); a.w = fmaf(w[3][t], v.w, a.w);
      }
    }
    *(float4*)(p.out + m * DW_C + c4) = a;
    s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
    s2[0] += a.x * a.x; s2[1] += a.y * a.y; s2[2] += a.z * a.z; s2[3] += a.w * a.w;
  }
  if (p.stats) block_chan_reduce(s1, s2, p.stats, p.stats + DW_C, c4);
}

// backward data: dA[bt,li,c] = sum_j w[c][j] * dz(bt, lo, c), lo*s - pad + j = li; then ReLU mask + BN sums
__global__ void __launch_bounds__(256) dw_bwd_data_kernel(const __grid_constant__ DwParams p) {
  const int c4 = (threadIdx.x & 31) * 4, rl = threadIdx.x >> 5;
  float4 p0 = make_float4(1, 1, 1, 1), p2 = make_float4(0, 0, 0, 0), p1 = p2;
  if (p.p0) { p0 = ld4(p.p0 + c4); p2 = ld4(p.p2 + c4); if (p.p1) p1 = ld4(p.p1 + c4); }
  float4 mp0 = make_float4(1, 1, 1, 1), mp2 = make_float4(0, 0, 0, 0), bmean = mp2, binv = mp2;
  if (p.mp0) { mp0 = ld4(p.mp0 + c4); mp2 = ld4(p.mp2 + c4); }
  if (p.bstats) { bmean = ld4(p.bmean + c4); binv = ld4(p.binv + c4); }
  float w[4][5];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 5; ++t) w[j][t] = t < p.k ? __ldg(p.w + (c4 + j) * p.k + t) : 0.f;
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  const long M = (long)p.BT * p.Lin;
  for (long m = (long)blockIdx.x * DW_ROWS + rl; m < M; m += (long)gridDim.x * DW_ROWS) {
    const int bt = (int)(m / p.Lin), li = (int)(m - (long)bt * p.Lin);
    float4 g = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int num = li + p.pad - t;
      if (t < p.k && num >= 0 && num % p.stride == 0) {
        const int lo = num / p.stride;
        if (lo < p.Lout) {
          const float4 v = dw_load(p, ((long)bt * p.Lout + lo) * DW_C, c4, p0, p1, p2, false);
          g.x = fmaf(w[0][t], v.x, g.x); g.y = fmaf(w[1][t], v.y, g.y);
          g.z = fmaf(w[2][t], v.z, g.z); g.w = fmaf(w[3][t], v.w, g.w);
        }
      }
    }
    const float4 z = ld4(p.zmask + m * DW_C + c4);
    g.x = (z.x * mp0.x + mp2.x > 0.f) ? g.x : 0.f; g.y = (z.y * mp0.y + mp2.y > 0.f) ? g.y : 0.f;
    g.z = (z.z * mp0.z + mp2.z > 0.f) ? g.z : 0.f; g.w = (z.w * mp0.w + mp2.w > 0.f) ? g.w : 0.f;
    *(float4*)(p.out + m * DW_C + c4) = g;
    s1[0] += g.x; s1[1] += g.y; s1[2] += g.z; s1[3] += g.w;
    s2[0] += g.x * (z.x - bmean.x) * binv.x; s2[1] += g.y * (z.y - bmean.y) * binv.y;
    s2[2] += g.z * (z.z - bmean.z) * binv.z; s2[3] += g.w * (z.w - bmean.w) * binv.w;
  }
  if (p.bstats) block_chan_reduce(s1, s2, p.bstats, p.bstats + DW_C, c4);
}

// fused backward (data + weight + bias): one pass over the input rows.  For row (bt, li) the taps
// t with lo*s - pad + t == li give  dA += w[c][t]*dz(lo)  and  dw[c][t] += dz(lo)*a(li); the centre tap
// (t == pad, li == lo*s) visits every output row exactly once, so it also carries db.  a() is the same
// tensor as the ReLU mask (Zp with BN1's affine), so every operand is read once per use site and the
// weight-gradient pass over dY / Zd / Zp of dw_wgrad_kernel disappears.  db (whose true value is 0 in
// front of a training-mode BN: pure cancellation) is summed in fp64 next to the BN sums (scratch; the network path takes the
// gradient from bn_bwd_finalize's conv_db, which derives it from the BN sums - trunet.cu, encoder blocks).
__global__ void __launch_bounds__(256, 2) dw_bwd_fused_kernel(const __grid_constant__ DwParams p) {
  __shared__ float red[DW_ROWS][DW_C];
  const int c4 = (threadIdx.x & 31) * 4, rl = threadIdx.x >> 5;
  float4 p0 = make_float4(1, 1, 1, 1), p2 = make_float4(0, 0, 0, 0), p1 = p2;
  if (p.p0) { p0 = ld4(p.p0 + c4); p2 = ld4(p.p2 + c4); if (p.p1) p1 = ld4(p.p1 + c4); }
  const float4 mp0 = ld4(p.mp0 + c4), mp2 = ld4(p.mp2 + c4), bmean = ld4(p.bmean + c4), binv = ld4(p.binv + c4);
  float w[4][5];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int t = 0; t < 5; ++t) w[j][t] = t < p.k ? __ldg(p.w + (c4 + j) * p.k + t) : 0.f;
  float acc[8][4];                 // 0-4: dw taps, 5: db, 6: sum g, 7: sum g*xhat
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  const long M = (long)p.BT * p.Lin;
  for (long m = (long)blockIdx.x * DW_ROWS + rl; m < M; m += (long)gridDim.x * DW_ROWS) {
    const int bt = (int)(m / p.Lin), li = (int)(m - (long)bt * p.Lin);
    const float4 z = ld4(p.zmask + m * DW_C + c4);
    float4 a;
    a.x = fmaxf(z.x * mp0.x + mp2.x, 0.f); a.y = fmaxf(z.y * mp0.y + mp2.y, 0.f);
    a.z = fmaxf(z.z * mp0.z + mp2.z, 0.f); a.w = fmaxf(z.w * mp0.w + mp2.w, 0.f);
    float4 g = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int num = li + p.pad - t;
      if (t < p.k && num >= 0 && num % p.stride == 0) {
        const int lo = num / p.stride;
        if (lo < p.Lout) {
          const float4 v = dw_load(p, ((long)bt * p.Lout + lo) * DW_C, c4, p0, p1, p2, false);
          g.x = fmaf(w[0][t], v.x, g.x); g.y = fmaf(w[1][t], v.y, g.y);
          g.z = fmaf(w[2][t], v.z, g.z); g.w = fmaf(w[3][t], v.w, g.w);
          acc[t][0] = fmaf(v.x, a.x, acc[t][0]); acc[t][1] = fmaf(v.y, a.y, acc[t][1]);
          acc[t][2] = fmaf(v.z, a.z, acc[t][2]); acc[t][3] = fmaf(v.w, a.w, acc[t][3]);
          if (t == p.pad) { acc[5][0] += v.x; acc[5][1] += v.y; acc[5][2] += v.z; acc[5][3] += v.w; }
        }
      }
    }
    g.x = (z.x * mp0.x + mp2.x > 0.f) ? g.x : 0.f; g.y = (z.y * mp0.y + mp2.y > 0.f) ? g.y : 0.f;
    g.z = (z.z * mp0.z + mp2.z > 0.f) ? g.z : 0.f; g.w = (z.w * mp0.w + mp2.w > 0.f) ? g.w : 0.f;
    *(float4*)(p.out + m * DW_C + c4) = g;
    acc[6][0] += g.x; acc[6][1] += g.y; acc[6][2] += g.z; acc[6][3] += g.w;
    acc[7][0] += g.x * (z.x - bmean.x) * binv.x; acc[7][1] += g.y * (z.y - bmean.y) * binv.y;
    acc[7][2] += g.z * (z.z - bmean.z) * binv.z; acc[7][3] += g.w * (z.w - bmean.w) * binv.w;
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    if (t < 5 && t >= p.k) continue;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[rl][c4 + j] = acc[t][j];
    __syncthreads();
    if (threadIdx.x < DW_C) {
      double s = 0.0;
      for (int r = 0; r < DW_ROWS; ++r) s += (double)red[r][threadIdx.x];
      if (t < 5) atomicAdd(p.dw + threadIdx.x * p.k + t, (float)s);
      else if (t == 5) atomicAdd(p.bstats + 2 * DW_C + threadIdx.x, s);     // db: fp64 scratch (see above)
      else atomicAdd(p.bstats + (t - 6) * DW_C + threadIdx.x, s);
    }
  }
}

// backward weight: dw[c][j] = sum dz(bt,lo,c) * a(bt, lo*s-pad+j, c); db[c] = sum dz
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const __grid_constant__ DwParams p) {
  __shared__ float red[DW_ROWS][DW_C];
  const int c4 = (threadIdx.x & 31) * 4, rl = threadIdx.x >> 5;
  float4 p0 = make_float4(1, 1, 1, 1), p2 = make_float4(0, 0, 0, 0), p1 = p2;
  if (p.p0) { p0 = ld4(p.p0 + c4); p2 = ld4(p.p2 + c4); if (p.p1) p1 = ld4(p.p1 + c4); }
  float4 a0 = make_float4(1, 1, 1, 1), a2 = make_float4(0, 0, 0, 0);
  if (p.a_p0) { a0 = ld4(p.a_p0 + c4); a2 = ld4(p.a_p2 + c4); }
  float acc[6][4];
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  const long M = (long)p.BT * p.Lout;
  for (long m = (long)blockIdx.x * DW_ROWS + rl; m < M; m += (long)gridDim.x * DW_ROWS) {
    const int bt = (int)(m / p.Lout), lo = (int)(m - (long)bt * p.Lout);
    const float4 dz = dw_load(p, m * DW_C, c4, p0, p1, p2, false);
    acc[5][0] += dz.x; acc[5][1] += dz.y; acc[5][2] += dz.z; acc[5][3] += dz.w;
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      const int li = lo * p.stride - p.pad + t;
      if (t < p.k && li >= 0 && li < p.Lin) {
        float4 a = ld4(p.a_src + ((long)bt * p.Lin + li) * DW_C + c4);
        a.x = fmaxf(a0.x * a.x + a2.x, 0.f); a.y = fmaxf(a0.y * a.y + a2.y, 0.f);
        a.z = fmaxf(a0.z * a.z + a2.z, 0.f); a.w = fmaxf(a0.w * a.w + a2.w, 0.f);
        acc[t][0] = fmaf(dz.x, a.x, acc[t][0]); acc[t][1] = fmaf(dz.y, a.y, acc[t][1]);
        acc[t][2] = fmaf(dz.z, a.z, acc[t][2]); acc[t][3] = fmaf(dz.w, a.w, acc[t][3]);
      }
    }
  }
  for (int t = 0; t < 6; ++t) {
    if (t < 5 && t >= p.k) continue;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[rl][c4 + j] = acc[t][j];
    __syncthreads();
    if (threadIdx.x < DW_C) {
      float s = 0.f;
      for (int r = 0; r < DW_ROWS; ++r) s += red[r][threadIdx.x];
      if (t < 5) atomicAdd(p.dw + threadIdx.x * p.k + t, s);
      else atomicAdd(p.db + threadIdx.x, s);
    }
  }
}

// db[n] += sum over rows of dz(row, n), dz = q0*dY + q1*Z + q2 (or dY): bias gradients that cannot
// ride on a weight-gradient pass (transposed convs, GRU hidden biases).
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ src, const float* __restrict__ src2,
                                                     const float* __restrict__ q0, const float* __restrict__ q1,
                                                     const float* __restrict__ q2, float* __restrict__ db, long rows,
                                                     int ld, int coff, int N) {
  __shared__ float red[256][4];
  const int cpl = N >> 2;                       // float4 columns per row
  const int rlanes = 256 / cpl;                 // rows handled per pass by this block
  const int col = threadIdx.x % cpl, rl = threadIdx.x / cpl;
  const bool active = rl < rlanes;
  float4 a0 = make_float4(1, 1, 1, 1), a1 = make_float4(0, 0, 0, 0), a2 = a1;
  if (q0 && active) { a0 = ld4(q0 + coff + col * 4); a2 = ld4(q2 + coff + col * 4); if (q1) a1 = ld4(q1 + coff + col * 4); }
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    for (long r = (long)blockIdx.x * rlanes + rl; r < rows; r += (long)gridDim.x * rlanes) {
      float4 v = ld4(src + r * ld + coff + col * 4);
      if (q0) {
        v.x = a0.x * v.x + a2.x; v.y = a0.y * v.y + a2.y; v.z = a0.z * v.z + a2.z; v.w = a0.w * v.w + a2.w;
        if (q1) {
          const float4 z = ld4(src2 + r * ld + coff + col * 4);
          v.x += a1.x * z.x; v.y += a1.y * z.y; v.z += a1.z * z.z; v.w += a1.w * z.w;
        }
      }
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) red[threadIdx.x][e] = s[e];
  __syncthreads();
  if (threadIdx.x < cpl) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < rlanes; ++r)
#pragma unroll
      for (int e = 0; e < 4; ++e) t[e] += red[r * cpl + threadIdx.x][e];
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(db + threadIdx.x * 4 + e, t[e]);
  }
}

__global__ void planar_to_cl_kernel(const float* __restrict__ src, float* __restrict__ dst, int BT, int C, int L) {
  const long total = (long)BT * C * L;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long r = i / C;
    const int l = (int)(r % L);
    const long bt = r / L;
    dst[i] = __ldg(src + (bt * C + c) * L + l);
  }
}

int dw_grid(long rows) { return (int)std::min<long>((rows + DW_ROWS - 1) / DW_ROWS, (long)sm_count() * 8); }

}  // namespace

int launch_bn_finalize(const BnFwdParams& p, cudaStream_t st) {
  ProfScope prof("bn_finalize", 0, 0, st);
  TRU_CUDA(launch_pdl(bn_finalize_kernel, dim3((p.C + 127) / 128), dim3(128), 0, st, p));
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_bn_finalize_eval_all(const BnEvalAll& p, cudaStream_t st) {
  ProfScope prof("bn_finalize_eval_all", 0, 0, st);
  bn_finalize_eval_all_kernel<<<dim3(1, TRU_NET_NBN), 128, 0, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_bn_bwd_finalize(const BnBwdParams& p, cudaStream_t st) {
  ProfScope prof("bn_bwd_finalize", 0, 0, st);
  TRU_CUDA(launch_pdl(bn_bwd_finalize_kernel, dim3((p.C + 127) / 128), dim3(128), 0, st, p));
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_enc0_fwd(const float* x, const float* w, const float* b, float* out, int BT, cudaStream_t st) {
  ProfScope prof("enc0_fwd", 4.0 * BT * (4 * 257 + 128 * 64), 2.0 * BT * 128 * 64 * 20, st);
  enc0_fwd_kernel<<<std::min(BT, sm_count() * 4), 256, 0, st>>>(x, w, b, out, BT);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_enc0_wgrad(const float* x, const float* dy, float* dw, float* db, int BT, cudaStream_t st) {
  ProfScope prof("enc0_wgrad", 4.0 * BT * (4 * 257 + 128 * 64), 2.0 * BT * 128 * 64 * 20, st);
  enc0_wgrad_kernel<<<std::min(BT, sm_count() * 2), 256, 0, st>>>(x, dy, dw, db, BT);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_dw_fwd(const DwParams& p, cudaStream_t st) {
  TRU_REQUIRE(p.C == DW_C && p.k <= 5, TRU_ERR_ARG, "depthwise kernel supports C=128, k<=5");
  { const int rc = launch_dw_fwd_stream(p, st); if (rc != 1) return rc; }
  ProfScope prof("dw_fwd", 4.0 * p.BT * ((double)p.Lin + p.Lout) * p.C, 2.0 * p.k * p.BT * p.Lout * p.C, st);
  dw_fwd_kernel<<<dw_grid((long)p.BT * p.Lout), 256, 0, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_dw_bwd_data(const DwParams& p, cudaStream_t st) {
  TRU_REQUIRE(p.C == DW_C && p.k <= 5, TRU_ERR_ARG, "depthwise kernel supports C=128, k<=5");
  ProfScope prof("dw_bwd_data", 4.0 * p.BT * (2.0 * p.Lin + 2.0 * p.Lout) * p.C, 2.0 * p.k * p.BT * p.Lout * p.C, st);
  dw_bwd_data_kernel<<<dw_grid((long)p.BT * p.Lin), 256, 0, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_dw_bwd_fused(const DwParams& p, cudaStream_t st) {
  TRU_REQUIRE(p.C == DW_C && p.k <= 5, TRU_ERR_ARG, "depthwise kernel supports C=128, k<=5");
  TRU_REQUIRE(p.zmask && p.mp0 && p.bstats && p.dw && p.db && p.a_src == p.zmask && p.a_p0 == p.mp0 && p.a_p2 == p.mp2,
              TRU_ERR_ARG, "dw_bwd_fused: the activation and the ReLU mask must be the same tensor");
  { const int rc = launch_dw_bwd_stream(p, st); if (rc != 1) return rc; }
  ProfScope prof("dw_bwd_fused", 4.0 * p.BT * (2.0 * p.Lin + 2.0 * p.Lout) * p.C, 4.0 * p.k * p.BT * p.Lout * p.C, st);
  // 2 resident CTAs per SM (122 registers): a grid-stride loop over exactly that many CTAs also keeps the
  // number of float atomics per weight-gradient element small
  dw_bwd_fused_kernel<<<std::min(dw_grid((long)p.BT * p.Lin), sm_count() * 2), 256, 0, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_dw_wgrad(const DwParams& p, cudaStream_t st) {
  TRU_REQUIRE(p.C == DW_C && p.k <= 5, TRU_ERR_ARG, "depthwise kernel supports C=128, k<=5");
  ProfScope prof("dw_wgrad", 4.0 * p.BT * ((double)p.Lin + 2.0 * p.Lout) * p.C, 2.0 * p.k * p.BT * p.Lout * p.C, st);
  dw_wgrad_kernel<<<std::min(dw_grid((long)p.BT * p.Lout), sm_count() * 2), 256, 0, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_colsum(const float* src, const float* src2, const float* q0, const float* q1, const float* q2, float* db,
                  long rows, int ld, int coff, int N, cudaStream_t st) {
  TRU_REQUIRE(N % 4 == 0 && N >= 4 && N <= 1024 && ld % 4 == 0 && coff % 4 == 0, TRU_ERR_ARG, "colsum: bad shape");
  const int rlanes = 256 / (N / 4);
  ProfScope prof("colsum", 4.0 * rows * N * (q1 ? 2 : 1), 0, st);
  colsum_kernel<<<(int)std::min<long>((rows + rlanes - 1) / rlanes, (long)sm_count() * 8), 256, 0, st>>>(
      src, src2, q0, q1, q2, db, rows, ld, coff, N);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
int launch_planar_to_cl(const float* src, float* dst, int BT, int C, int L, cudaStream_t st) {
  const long total = (long)BT * C * L;
  ProfScope prof("planar_to_cl", 8.0 * total, 0, st);
  planar_to_cl_kernel<<<(int)std::min<long>((total + 255) / 256, (long)sm_count() * 8), 256, 0, st>>>(src, dst, BT, C, L);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace tru
