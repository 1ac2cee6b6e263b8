// Small-channel transposed conv (network.py:102-120, LastTrCNN: ConvTranspose1d(8, 8, k=5, s=2, p=1), no BN/ReLU
// behind it) as plain FP32 kernels: forward, data gradient (+ReLU mask and BN-backward sums of the layer below) and
// weight / bias gradient.
//
// With 8 channels a row is 32 bytes and the whole layer is ~0.3 % of the model's MACs; on the tensor-core GEMM path
// it paid the per-tile pipeline cost of a 128-row tile for 24-40 MACs per row (0.44 + 0.28 + 0.21 + 0.60 ms per
// training step at B = 32).  Here a thread owns one row: inputs are read with two 16-byte loads per tap, the 320
// weights sit in shared memory laid out so that one 16-byte broadcast load feeds four FMAs, and the planar network
// output (B,T,8,257) is written directly.  HBM bound: forward 66 MB in + 132 MB out per step.
#include "net_kernels.cuh"

namespace tru {
namespace {

constexpr int CC = 8, KK = 5, ST = 2, PAD = 1;       // channels (in = out), taps, stride, padding (= stride / 2)
constexpr int NTH = 256;

__device__ __forceinline__ void load_act8(const float* p, const float* sp0, const float* sp2, bool affine, float (&a)[CC]) {
  const float4 u = __ldg((const float4*)p), v = __ldg((const float4*)p + 1);
  a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w; a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w;
  if (affine) {
#pragma unroll
    for (int c = 0; c < CC; ++c) a[c] = fmaxf(fmaf(sp0[c], a[c], sp2[c]), 0.f);
  }
}

// out[bt][co][lo] = b[co] + sum_{j, ci} a[bt][(lo + PAD - j) / ST][ci] * W[ci][co][j]   (lo + PAD - j divisible by ST)
// One warp works on one parity class of lo (the valid taps are then warp-uniform).
__global__ void __launch_bounds__(NTH) convt8_fwd_kernel(const float* __restrict__ z, const float* __restrict__ p0,
                                                         const float* __restrict__ p2, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int BT, int L, int Lout, int planar) {
  __shared__ __align__(16) float w[KK][CC][CC];      // [j][ci][co]
  __shared__ float sp0[CC], sp2[CC], sb[CC];
  const int tid = threadIdx.x;
  for (int i = tid; i < CC * CC * KK; i += NTH) {
    const int ci = i / (CC * KK), co = (i / KK) % CC, j = i % KK;
    w[j][ci][co] = __ldg(W + i);
  }
  if (tid < CC) { sp0[tid] = p0 ? __ldg(p0 + tid) : 1.f; sp2[tid] = p0 ? __ldg(p2 + tid) : 0.f; sb[tid] = bias ? __ldg(bias + tid) : 0.f; }
  __syncthreads();
  const int half0 = (Lout + 1) / 2, half1 = Lout / 2;           // rows of parity 0 / 1 per frame
  // warps 0-3 of a CTA take parity 0, warps 4-7 parity 1; a CTA covers 128 consecutive (bt, q) pairs of each
  const int par = tid >> 7, t = tid & 127;
  const int nper = par ? half1 : half0;
  const long idx = (long)blockIdx.x * 128 + t;
  if (idx >= (long)BT * nper) return;
  const int bt = (int)(idx / nper), q = (int)(idx - (long)bt * nper), lo = ST * q + par;
  float acc[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) acc[c] = sb[c];
  const float* zf = z + (size_t)bt * L * CC;
#pragma unroll
  for (int j = 0; j < KK; ++j) {
    if (((PAD - j) & 1) != par) continue;            // (lo + PAD - j) even  <=>  par == (PAD - j) mod 2   (warp-uniform)
    const int num = lo + PAD - j;
    const int l = num >> 1;
    if (num >= 0 && l < L) {
      float a[CC];
      load_act8(zf + (size_t)l * CC, sp0, sp2, p0 != nullptr, a);
#pragma unroll
      for (int ci = 0; ci < CC; ++ci) {
        const float4 w0 = *(const float4*)&w[j][ci][0], w1 = *(const float4*)&w[j][ci][4];
        acc[0] = fmaf(a[ci], w0.x, acc[0]); acc[1] = fmaf(a[ci], w0.y, acc[1]); acc[2] = fmaf(a[ci], w0.z, acc[2]); acc[3] = fmaf(a[ci], w0.w, acc[3]);
        acc[4] = fmaf(a[ci], w1.x, acc[4]); acc[5] = fmaf(a[ci], w1.y, acc[5]); acc[6] = fmaf(a[ci], w1.z, acc[6]); acc[7] = fmaf(a[ci], w1.w, acc[7]);
      }
    }
  }
  if (planar) {
    float* o = out + (size_t)bt * CC * Lout + lo;
#pragma unroll
    for (int c = 0; c < CC; ++c) o[(size_t)c * Lout] = acc[c];
  } else {
    float4* o = (float4*)(out + ((size_t)bt * Lout + lo) * CC);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]); o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// dX[bt][l][ci] = mask * sum_{j, co} dY[bt][ST*l - PAD + j][co] * W[ci][co][j];  mask = relu'(mp0*zmask + mp2);
// BN-backward sums of the layer that produced zmask: sum g, invstd * sum g (z - mean)   (igemm epilogue convention).
__global__ void __launch_bounds__(NTH) convt8_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                              float* __restrict__ dx, const float* __restrict__ zmask,
                                                              const float* __restrict__ mp0, const float* __restrict__ mp2,
                                                              const float* __restrict__ bmean, const float* __restrict__ binv,
                                                              double* __restrict__ bstats, int BT, int L, int Lout, int dy_planar) {
  __shared__ __align__(16) float w[KK][CC][CC];      // [j][co][ci]
  __shared__ float s0[CC], s2[CC], sm[CC];
  __shared__ float red[NTH / 32][2 * CC];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < CC * CC * KK; i += NTH) {
    const int ci = i / (CC * KK), co = (i / KK) % CC, j = i % KK;
    w[j][co][ci] = __ldg(W + i);
  }
  if (tid < CC) { s0[tid] = mp0 ? __ldg(mp0 + tid) : 1.f; s2[tid] = mp0 ? __ldg(mp2 + tid) : 0.f; sm[tid] = bstats ? __ldg(bmean + tid) : 0.f; }
  __syncthreads();
  const long idx = (long)blockIdx.x * NTH + tid;
  const bool live = idx < (long)BT * L;
  float g[CC], zz[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) { g[c] = 0.f; zz[c] = 0.f; }
  if (live) {
    const int bt = (int)(idx / L), l = (int)(idx - (long)bt * L);
    const float* df = dy + (size_t)bt * Lout * CC;
#pragma unroll
    for (int j = 0; j < KK; ++j) {
      const int lo = ST * l - PAD + j;
      if (lo >= 0 && lo < Lout) {
        float d[CC];
        if (dy_planar) {                               // the network-output gradient as autograd hands it over: (B,T,8,257)
#pragma unroll
          for (int co = 0; co < CC; ++co) d[co] = __ldg(df + (size_t)co * Lout + lo);
        } else {
          const float4 u = __ldg((const float4*)(df + (size_t)lo * CC)), v = __ldg((const float4*)(df + (size_t)lo * CC) + 1);
          d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w; d[4] = v.x; d[5] = v.y; d[6] = v.z; d[7] = v.w;
        }
#pragma unroll
        for (int co = 0; co < CC; ++co) {
          const float4 w0 = *(const float4*)&w[j][co][0], w1 = *(const float4*)&w[j][co][4];
          g[0] = fmaf(d[co], w0.x, g[0]); g[1] = fmaf(d[co], w0.y, g[1]); g[2] = fmaf(d[co], w0.z, g[2]); g[3] = fmaf(d[co], w0.w, g[3]);
          g[4] = fmaf(d[co], w1.x, g[4]); g[5] = fmaf(d[co], w1.y, g[5]); g[6] = fmaf(d[co], w1.z, g[6]); g[7] = fmaf(d[co], w1.w, g[7]);
        }
      }
    }
    if (zmask) {
      const float4 u = __ldg((const float4*)(zmask + (size_t)idx * CC)), v = __ldg((const float4*)(zmask + (size_t)idx * CC) + 1);
      zz[0] = u.x; zz[1] = u.y; zz[2] = u.z; zz[3] = u.w; zz[4] = v.x; zz[5] = v.y; zz[6] = v.z; zz[7] = v.w;
#pragma unroll
      for (int c = 0; c < CC; ++c) g[c] = fmaf(zz[c], s0[c], s2[c]) > 0.f ? g[c] : 0.f;
    }
    float4* o = (float4*)(dx + (size_t)idx * CC);
    o[0] = make_float4(g[0], g[1], g[2], g[3]); o[1] = make_float4(g[4], g[5], g[6], g[7]);
  }
  if (bstats) {                                       // whole-CTA reduction of 16 sums, then 16 fp64 atomics
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      const float a = warp_sum(g[c]), b = warp_sum(g[c] * (zz[c] - sm[c]));
      if (lane == 0) { red[warp][c] = a; red[warp][CC + c] = b; }
    }
    __syncthreads();
    if (tid < 2 * CC) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < NTH / 32; ++wv) s += red[wv][tid];
      const int c = tid & (CC - 1);
      if (tid < CC) atomicAdd(bstats + c, (double)s);
      else atomicAdd(bstats + CC + c, (double)s * (double)__ldg(binv + c));
    }
  }
}

// dW[ci][co][j] += sum_{bt,l} a[bt][l][ci] * dY[bt][ST*l - PAD + j][co];  db[co] += sum dY.
// One frame at a time in shared memory (next frame prefetched into registers); 320 threads = 4 row groups x
// (ci, 4 output channels, tap); threads 320-327 do the bias sums.
constexpr int WG_NT = 352;
__global__ void __launch_bounds__(WG_NT) convt8_wgrad_kernel(const float* __restrict__ z, const float* __restrict__ p0,
                                                             const float* __restrict__ p2, const float* __restrict__ dy,
                                                             float* __restrict__ dW, float* __restrict__ db,
                                                             int BT, int L, int Lout, int dy_planar) {
  extern __shared__ __align__(16) float sm_w[];
  float* as = sm_w;                    // [L][8] activations (BN + ReLU applied)
  float* ds = as + L * CC;             // [Lout][8]
  const int tid = threadIdx.x;
  const int nA4 = L * CC / 4, nD4 = (Lout * CC + 3) / 4, n4 = nA4 + nD4;
  constexpr int MAXI = 3;              // float4 items per thread and frame (planner: n4 <= MAXI * WG_NT)
  float4 pre[MAXI];
  float ap0[4] = {1.f, 1.f, 1.f, 1.f}, ap2[4] = {0.f, 0.f, 0.f, 0.f};
  auto fetch = [&](int bt) {
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int it = tid + i * WG_NT;
      if (it < nA4) pre[i] = __ldg((const float4*)(z + (size_t)bt * L * CC) + it);
      else if (it < n4) {
        const int e = (it - nA4) * 4;                  // planar: 4 consecutive positions of one channel; else 4 channels of one row
        if (dy_planar) {
          const float* g = dy + (size_t)bt * Lout * CC;
          pre[i].x = __ldg(g + e); pre[i].y = e + 1 < Lout * CC ? __ldg(g + e + 1) : 0.f;
          pre[i].z = e + 2 < Lout * CC ? __ldg(g + e + 2) : 0.f; pre[i].w = e + 3 < Lout * CC ? __ldg(g + e + 3) : 0.f;
        } else {
          pre[i] = __ldg((const float4*)(dy + (size_t)bt * Lout * CC) + (it - nA4));
        }
      }
    }
  };
  // role
  const int grp = tid / 80, rem = tid % 80, ci = rem / 10, cq = (rem / 5) & 1, j = rem % 5;     // tid < 320
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bsum = 0.f;
  int bt = blockIdx.x;
  if (bt < BT) fetch(bt);
  for (; bt < BT; bt += gridDim.x) {
    __syncthreads();                                   // previous frame fully consumed
#pragma unroll
    for (int i = 0; i < MAXI; ++i) {
      const int it = tid + i * WG_NT;
      if (it < nA4) {
        float4 v = pre[i];
        if (p0) {
          const int c = (it & 1) * 4;                  // 8 channels = 2 float4 per row
          v.x = fmaxf(fmaf(__ldg(p0 + c), v.x, __ldg(p2 + c)), 0.f); v.y = fmaxf(fmaf(__ldg(p0 + c + 1), v.y, __ldg(p2 + c + 1)), 0.f);
          v.z = fmaxf(fmaf(__ldg(p0 + c + 2), v.z, __ldg(p2 + c + 2)), 0.f); v.w = fmaxf(fmaf(__ldg(p0 + c + 3), v.w, __ldg(p2 + c + 3)), 0.f);
        }
        ((float4*)as)[it] = v;
      } else if (it < n4) {
        if (dy_planar) {                               // element e of the planar frame = (channel e / Lout, position e % Lout)
          const int e = (it - nA4) * 4;
          const float v4[4] = {pre[i].x, pre[i].y, pre[i].z, pre[i].w};
          constexpr int LO = 257;                      // (planner: the planar path is only taken for Lout == 257; constant divisor)
#pragma unroll
          for (int q = 0; q < 4; ++q) if (e + q < LO * CC) ds[((e + q) % LO) * CC + (e + q) / LO] = v4[q];
        } else {
          ((float4*)ds)[it - nA4] = pre[i];
        }
      }
    }
    __syncthreads();
    if (bt + (int)gridDim.x < BT) fetch(bt + gridDim.x);
    if (tid < 320) {
      const int l0 = grp * (L / 4), l1 = l0 + L / 4;   // planner: L % 4 == 0
#pragma unroll 4
      for (int l = l0; l < l1; ++l) {
        const int lo = ST * l - PAD + j;
        if (lo >= 0 && lo < Lout) {
          const float a = as[l * CC + ci];
          const float4 d = *(const float4*)&ds[lo * CC + cq * 4];
          acc[0] = fmaf(a, d.x, acc[0]); acc[1] = fmaf(a, d.y, acc[1]); acc[2] = fmaf(a, d.z, acc[2]); acc[3] = fmaf(a, d.w, acc[3]);
        }
      }
    } else if (db) {                       // warp 10: lane = (quarter of the rows, channel); 4 independent partial sums each
      const int co = (tid - 320) & 7, qt = (tid - 320) >> 3;
      float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
      int lo = qt;
      for (; lo + 12 < Lout; lo += 16) {
        b0 += ds[lo * CC + co]; b1 += ds[(lo + 4) * CC + co]; b2 += ds[(lo + 8) * CC + co]; b3 += ds[(lo + 12) * CC + co];
      }
      for (; lo < Lout; lo += 4) b0 += ds[lo * CC + co];
      bsum += (b0 + b1) + (b2 + b3);
    }
  }
  (void)ap0; (void)ap2;
  if (tid < 320) {
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(dW + (ci * CC + cq * 4 + e) * KK + j, acc[e]);
  } else if (db) {
    atomicAdd(db + ((tid - 320) & 7), bsum);
  }
}

}  // namespace

bool convt_small_eligible(int Cin, int Cout, int k, int s, int L, int Lout) {
  return Cin == CC && Cout == CC && k == KK && s == ST && L == 128 && Lout == 257 && Lout == (L - 1) * ST - 2 * PAD + KK &&
         (L * CC / 4 + Lout * CC / 4) <= 3 * WG_NT;
}

int launch_convt_small_fwd(const float* z, const float* p0, const float* p2, const float* W, const float* bias, float* out,
                           int BT, int L, int Lout, int planar, cudaStream_t st) {
  const long nper = (Lout + 1) / 2;
  const long blocks = ((long)BT * nper + 127) / 128;
  ProfScope prof("convt8_fwd", 4.0 * BT * CC * ((double)L + Lout), 2.0 * BT * L * CC * CC * KK, st);
  convt8_fwd_kernel<<<(unsigned)blocks, NTH, 0, st>>>(z, p0, p2, W, bias, out, BT, L, Lout, planar);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

int launch_convt_small_bwd_data(const float* dy, const float* W, float* dx, const float* zmask, const float* mp0, const float* mp2,
                                const float* bmean, const float* binv, double* bstats, int BT, int L, int Lout, int dy_planar,
                                cudaStream_t st) {
  const long blocks = ((long)BT * L + NTH - 1) / NTH;
  ProfScope prof("convt8_bwd_data", 4.0 * BT * CC * ((double)Lout + 2.0 * L), 2.0 * BT * L * CC * CC * KK, st);
  convt8_bwd_data_kernel<<<(unsigned)blocks, NTH, 0, st>>>(dy, W, dx, zmask, mp0, mp2, bmean, binv, bstats, BT, L, Lout, dy_planar);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

int launch_convt_small_wgrad(const float* z, const float* p0, const float* p2, const float* dy, float* dW, float* db,
                             int BT, int L, int Lout, int dy_planar, cudaStream_t st) {
  const size_t smem = (size_t)(L + Lout) * CC * 4;
  const int grid = std::min(BT, 4 * sm_count());
  ProfScope prof("convt8_wgrad", 4.0 * BT * CC * ((double)L + Lout), 2.0 * BT * L * CC * CC * KK, st);
  convt8_wgrad_kernel<<<grid, WG_NT, smem, st>>>(z, p0, p2, dy, dW, db, BT, L, Lout, dy_planar);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace tru
