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
// Register-resident: thread = (row l, 4 input channels, 4 output channels) keeps its 4 x 4 x 5 partial sums in registers
// for every frame the CTA walks (512 threads = the 128 rows of one frame x 4), reading its 16 B of activations and the five
// dY rows it touches straight from global memory (neighbouring rows share them through L1; DRAM sees every byte once).
// No shared-memory staging and no barrier per frame: the first version staged each frame in smem behind two
// __syncthreads and ran at 0.64 TB/s.  One reduction at the end: warp shuffles, shared-memory atomics, 328 global atomics.
constexpr int WG_NT = 512;
__global__ void __launch_bounds__(WG_NT, 1) convt8_wgrad_kernel(const float* __restrict__ z, const float* __restrict__ p0,
                                                                const float* __restrict__ p2, const float* __restrict__ dy,
                                                                float* __restrict__ dW, float* __restrict__ db,
                                                                int BT, int L, int Lout, int dy_planar) {
  __shared__ float red[CC * CC * KK + CC];
  const int tid = threadIdx.x;
  const int sub = tid & 3, quad = sub & 1, coh = sub >> 1, l = tid >> 2;      // planner: L == WG_NT / 4
  for (int i = tid; i < CC * CC * KK + CC; i += WG_NT) red[i] = 0.f;
  float acc[4][4][KK];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < KK; ++j) acc[e][c][j] = 0.f;
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  float s0[4] = {1.f, 1.f, 1.f, 1.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (p0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) { s0[e] = __ldg(p0 + quad * 4 + e); s2[e] = __ldg(p2 + quad * 4 + e); }
  }
  const float floor_ = p0 ? 0.f : -__int_as_float(0x7f800000);                 // ReLU only behind a BN
  const bool last = l == L - 1;
  for (int bt = blockIdx.x; bt < BT; bt += gridDim.x) {
    const float4 av = __ldg((const float4*)(z + ((size_t)bt * L + l) * CC + quad * 4));
    const float a[4] = {fmaxf(fmaf(s0[0], av.x, s2[0]), floor_), fmaxf(fmaf(s0[1], av.y, s2[1]), floor_),
                        fmaxf(fmaf(s0[2], av.z, s2[2]), floor_), fmaxf(fmaf(s0[3], av.w, s2[3]), floor_)};
    float d[KK][4];
#pragma unroll
    for (int j = 0; j < KK; ++j) {
      const int lo = ST * l - PAD + j;
      const bool ok = lo >= 0 && lo < Lout;
      if (dy_planar) {
        const float* g = dy + (size_t)bt * Lout * CC + (size_t)(coh * 4) * Lout + lo;
#pragma unroll
        for (int c = 0; c < 4; ++c) d[j][c] = ok ? __ldg(g + (size_t)c * Lout) : 0.f;
      } else {
        const float4 v = ok ? __ldg((const float4*)(dy + ((size_t)bt * Lout + lo) * CC + coh * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        d[j][0] = v.x; d[j][1] = v.y; d[j][2] = v.z; d[j][3] = v.w;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < KK; ++j) acc[e][c][j] = fmaf(a[e], d[j][c], acc[e][c][j]);
    if (quad == 0) {                       // every output row exactly once: rows 2l and 2l+1, plus row 2L for the last l
#pragma unroll
      for (int c = 0; c < 4; ++c) bs[c] += d[1][c] + d[2][c] + (last ? d[3][c] : 0.f);
    }
  }
  // ---- reduction: lanes with the same (quad, coh) -> lanes 0..3, then shared and global atomics ----
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < KK; ++j) {
        float v = acc[e][c][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[e][c][j] = v;
      }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float v = bs[c];
    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
    bs[c] = v;
  }
  __syncthreads();                         // red[] zeroed
  if ((tid & 31) < 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < KK; ++j) atomicAdd(&red[((quad * 4 + e) * CC + coh * 4 + c) * KK + j], acc[e][c][j]);
    if (quad == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) atomicAdd(&red[CC * CC * KK + coh * 4 + c], bs[c]);
    }
  }
  __syncthreads();
  if (tid < CC * CC * KK) atomicAdd(dW + tid, red[tid]);
  else if (tid < CC * CC * KK + CC && db) atomicAdd(db + (tid - CC * CC * KK), red[tid]);
}

}  // namespace

bool convt_small_eligible(int Cin, int Cout, int k, int s, int L, int Lout) {
  return Cin == CC && Cout == CC && k == KK && s == ST && L == 128 && Lout == 257 && Lout == (L - 1) * ST - 2 * PAD + KK && L * 4 == WG_NT;
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
  const int grid = std::min(BT, sm_count());
  ProfScope prof("convt8_wgrad", 4.0 * BT * CC * ((double)L + Lout), 2.0 * BT * L * CC * CC * KK, st);
  convt8_wgrad_kernel<<<grid, WG_NT, 0, st>>>(z, p0, p2, dy, dW, db, BT, L, Lout, dy_planar);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace tru
