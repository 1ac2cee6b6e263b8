// Depthwise conv over frequency (network.py:33-40), C = 128, forward and fused backward, as
// streaming kernels: these layers are pure byte movers (2*k flop per 4 bytes), so the design goal
// is bytes in flight, not arithmetic.
//
//   warp 0        producer: per unit (one frame, 32 rows) ONE cp.async.bulk per operand (the rows a
//                 unit needs are a contiguous run of its frame) into a ring of raw shared-memory
//                 stages, completion on an mbarrier (expect_tx); ~150-200 KB in flight per SM
//   warps 1-16    compute: a warp reads whole 512-byte rows from the stage (conflict-free float4),
//                 applies BN/ReLU (forward) or the BN-backward affine (backward) on the fly, and
//                 writes its output row with one coalesced 512-byte store; per-channel sums
//                 (BN statistics, BN-backward sums, weight / bias gradients) stay in registers
//                 and are reduced once per CTA.
#include <algorithm>
#include "net_kernels.cuh"
#include "tc_common.cuh"

namespace tru {
namespace {
using namespace tc;

constexpr int DC = 128;                  // channels
constexpr int RU = 32;                   // rows per unit
constexpr int CW = 16;                   // compute warps
constexpr int NT = 32 * (1 + CW);
constexpr int MAXST = 8;
constexpr size_t SMEM_MAX = 227 * 1024;

struct DwMisc { uint64_t full[MAXST], empty[MAXST]; };

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// per-channel block reduction of one float4 per thread over the compute warps, then fp64 / fp32 atomics
__device__ __forceinline__ void reduce_channels(float* red, const float (&v)[4], int cw, int c4, double* gd, float* gf, int stride) {
  // red: [CW][DC] floats
  asm volatile("bar.sync 1, %0;" ::"n"(32 * CW));
#pragma unroll
  for (int j = 0; j < 4; ++j) red[cw * DC + c4 + j] = v[j];
  asm volatile("bar.sync 1, %0;" ::"n"(32 * CW));
  const int t = cw * 32 + (c4 >> 2);
  if (t < DC) {
    double s = 0.0;
    for (int r = 0; r < CW; ++r) s += (double)red[r * DC + t];
    if (gd) atomicAdd(gd + t, s);
    if (gf) atomicAdd(gf + (long)t * stride, (float)s);
  }
}

struct DwK {
  DwParams p;
  int nstage, upf;                    // stages, units per frame
  uint32_t stage_bytes, in2_off, in3_off, red_off, misc_off, coef_off;
  unsigned units;
};

// ------------------------------------------------------------------ forward
// unit = (frame bt, RO output rows lo0..): stage holds the input rows [ra, rb) it needs
template <int K, int S>
__global__ void __launch_bounds__(NT, 1) dw_fwd_stream_kernel(const __grid_constant__ DwK Kp) {
  pdl_trigger();
  const DwParams& p = Kp.p;
  extern __shared__ __align__(128) uint8_t smem[];
  DwMisc& mi = *(DwMisc*)(smem + Kp.misc_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int PAD = K / 2;
  const int RO = min(RU, p.Lout);
  const int nst = Kp.nstage;
  if (tid == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(&mi.full[s], 1); mbar_init(&mi.empty[s], CW); }
    fence_barrier_init();
  }
  pdl_wait();                                            // barrier setup above overlaps the predecessor's tail
  __syncthreads();
  const unsigned n_my = Kp.units > blockIdx.x ? (Kp.units - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  if (warp == 0) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (unsigned i = 0; i < n_my; ++i) {
        const unsigned u = blockIdx.x + i * gridDim.x;
        const int bt = u / Kp.upf, lo0 = (u % Kp.upf) * RO;
        const int ra = max(0, lo0 * S - PAD), rb = min(p.Lin, (lo0 + RO - 1) * S - PAD + K);
        mbar_wait(&mi.empty[st], ph ^ 1);
        const uint32_t bytes = (uint32_t)(rb - ra) * DC * 4u;
        mbar_arrive_expect_tx(&mi.full[st], bytes);
        bulk_g2s(smem + (size_t)st * Kp.stage_bytes, p.src + ((size_t)bt * p.Lin + ra) * DC, bytes, &mi.full[st]);
        if (++st == nst) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int cw = warp - 1, c4 = lane * 4;
    const float4 p0 = p.p0 ? ld4(p.p0 + c4) : make_float4(1, 1, 1, 1), p2 = p.p0 ? ld4(p.p2 + c4) : make_float4(0, 0, 0, 0);
    const float fl = p.p0 ? 0.f : -__int_as_float(0x7f800000);
    float w[4][K];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int t = 0; t < K; ++t) w[j][t] = __ldg(p.w + (c4 + j) * K + t);
    const float4 bias = ld4(p.bias + c4);
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    int st = 0; uint32_t ph = 0;
    for (unsigned i = 0; i < n_my; ++i) {
      const unsigned u = blockIdx.x + i * gridDim.x;
      const int bt = u / Kp.upf, lo0 = (u % Kp.upf) * RO;
      const int ra = max(0, lo0 * S - PAD);
      mbar_wait(&mi.full[st], ph);
      const float* in = (const float*)(smem + (size_t)st * Kp.stage_bytes);
      for (int r = cw; r < RO; r += CW) {
        const int lo = lo0 + r;
        float4 a = bias;
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const int li = lo * S - PAD + t;
          if (li >= 0 && li < p.Lin) {
            float4 v = *(const float4*)(in + (li - ra) * DC + c4);
            v.x = fmaxf(fmaf(p0.x, v.x, p2.x), fl); v.y = fmaxf(fmaf(p0.y, v.y, p2.y), fl);
            v.z = fmaxf(fmaf(p0.z, v.z, p2.z), fl); v.w = fmaxf(fmaf(p0.w, v.w, p2.w), fl);
            a.x = fmaf(w[0][t], v.x, a.x); a.y = fmaf(w[1][t], v.y, a.y);
            a.z = fmaf(w[2][t], v.z, a.z); a.w = fmaf(w[3][t], v.w, a.w);
          }
        }
        *(float4*)(p.out + ((size_t)bt * p.Lout + lo) * DC + c4) = a;
        s1[0] += a.x; s1[1] += a.y; s1[2] += a.z; s1[3] += a.w;
        s2[0] = fmaf(a.x, a.x, s2[0]); s2[1] = fmaf(a.y, a.y, s2[1]); s2[2] = fmaf(a.z, a.z, s2[2]); s2[3] = fmaf(a.w, a.w, s2[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mi.empty[st]);
      if (++st == nst) { st = 0; ph ^= 1; }
    }
    if (p.stats) {
      float* red = (float*)(smem + Kp.red_off);
      reduce_channels(red, s1, cw, c4, p.stats, nullptr, 0);
      reduce_channels(red, s2, cw, c4, p.stats + DC, nullptr, 0);
    }
  }
}

// ------------------------------------------------------------------ fused backward
// unit = (frame bt, RU input rows li0..): stage holds Zp rows [li0, li0+RI) and the dY / Zd rows [la, lb] they touch.
template <int K, int S>
__global__ void __launch_bounds__(NT, 1) dw_bwd_stream_kernel(const __grid_constant__ DwK Kp) {
  pdl_trigger();
  const DwParams& p = Kp.p;
  extern __shared__ __align__(128) uint8_t smem[];
  DwMisc& mi = *(DwMisc*)(smem + Kp.misc_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int PAD = K / 2;
  const int RI = min(RU, p.Lin);
  const int nst = Kp.nstage;
  if (tid == 0) {
    for (int s = 0; s < nst; ++s) { mbar_init(&mi.full[s], 1); mbar_init(&mi.empty[s], CW); }
    fence_barrier_init();
  }
  pdl_wait();                                            // barrier setup above overlaps the predecessor's tail
  // Per-channel coefficients that are used once per row live in shared memory, not in registers: 17 warps put five on one SM
  // sub-partition, which caps the kernel at 96 registers, and with q0-q2, the taps and the K + 3 sums resident the compiler spilled
  // the RING STATE - ncu's source view had 21 % of the warp time waiting on local-memory reloads of the stage index and phase.
  float* cf = (float*)(smem + Kp.coef_off);              // [mp0 | mp2 | bmean][DC]
  for (int i = tid; i < DC; i += NT) {
    cf[i] = __ldg(p.mp0 + i); cf[DC + i] = __ldg(p.mp2 + i); cf[2 * DC + i] = __ldg(p.bmean + i);
    if (K == 5)                                          // (five taps x four channels are 20 more registers: tap-major copy, read per use)
      for (int t = 0; t < K; ++t) cf[(3 + t) * DC + i] = __ldg(p.w + i * K + t);
  }
  __syncthreads();
  const unsigned n_my = Kp.units > blockIdx.x ? (Kp.units - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  auto lo_range = [&](int li0, int& la, int& lb) {
    const int num = li0 + PAD - (K - 1);
    la = num <= 0 ? 0 : (num + S - 1) / S;
    lb = min(p.Lout - 1, (li0 + RI - 1 + PAD) / S);
  };
  if (warp == 0) {
    if (lane < 3) {
      int st = 0; uint32_t ph = 0;
      for (unsigned i = 0; i < n_my; ++i) {
        const unsigned u = blockIdx.x + i * gridDim.x;
        const int bt = u / Kp.upf, li0 = (u % Kp.upf) * RI;
        int la, lb;
        lo_range(li0, la, lb);
        const uint32_t zb = (uint32_t)RI * DC * 4u, db = (uint32_t)max(0, lb - la + 1) * DC * 4u;
        if (lane == 0) mbar_wait(&mi.empty[st], ph ^ 1);
        __syncwarp(0x7);
        uint8_t* sb = smem + (size_t)st * Kp.stage_bytes;
        if (lane == 0) {
          mbar_arrive_expect_tx(&mi.full[st], zb + 2 * db);
          bulk_g2s(sb, p.zmask + ((size_t)bt * p.Lin + li0) * DC, zb, &mi.full[st]);
        }
        __syncwarp(0x7);
        if (lane == 1 && db) bulk_g2s(sb + Kp.in2_off, p.src + ((size_t)bt * p.Lout + la) * DC, db, &mi.full[st]);
        if (lane == 2 && db) bulk_g2s(sb + Kp.in3_off, p.src2 + ((size_t)bt * p.Lout + la) * DC, db, &mi.full[st]);
        if (++st == nst) { st = 0; ph ^= 1; }
      }
    }
  } else {
    const int cw = warp - 1, c4 = lane * 4;
    const float4 q0 = ld4(p.p0 + c4), q1 = ld4(p.p1 + c4), q2 = ld4(p.p2 + c4);
    const uint32_t cfa = smem_u32(cf + c4);
    auto lds4 = [](uint32_t a) {                         // (volatile: the load stays next to its use)
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
      return v;
    };
    float w[4][K == 5 ? 1 : K];
    if (K != 5) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < (K == 5 ? 1 : K); ++t) w[j][t] = __ldg(p.w + (c4 + j) * K + t);
    }
    // (the depthwise bias gradient - exactly zero in front of a training-mode BatchNorm - is produced by bn_bwd_finalize from the
    // sums it already has, like the transposed convs': no per-row accumulation here)
    float acc[K + 2][4];             // 0..K-1: dw taps, K: sum g, K+1: sum g*(z - mean)
#pragma unroll
    for (int t = 0; t < K + 2; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
    int st = 0; uint32_t ph = 0;
    for (unsigned i = 0; i < n_my; ++i) {
      const unsigned u = blockIdx.x + i * gridDim.x;
      const int bt = u / Kp.upf, li0 = (u % Kp.upf) * RI;
      int la, lb;
      lo_range(li0, la, lb);
      mbar_wait(&mi.full[st], ph);
      const uint8_t* sb = smem + (size_t)st * Kp.stage_bytes;
      const float* zp = (const float*)sb;
      const float* dy = (const float*)(sb + Kp.in2_off);
      const float* zd = (const float*)(sb + Kp.in3_off);
      for (int r = cw; r < RI; r += CW) {
        const int li = li0 + r;
        const float4 z = *(const float4*)(zp + r * DC + c4);
        const float4 mp0 = lds4(cfa), mp2 = lds4(cfa + DC * 4);
        float4 a;
        a.x = fmaxf(fmaf(z.x, mp0.x, mp2.x), 0.f); a.y = fmaxf(fmaf(z.y, mp0.y, mp2.y), 0.f);
        a.z = fmaxf(fmaf(z.z, mp0.z, mp2.z), 0.f); a.w = fmaf(z.w, mp0.w, mp2.w); a.w = fmaxf(a.w, 0.f);
        float4 g = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const int num = li + PAD - t;
          if (num >= 0 && num % S == 0) {
            const int lo = num / S;
            if (lo < p.Lout) {
              const float4 y = *(const float4*)(dy + (lo - la) * DC + c4), zz = *(const float4*)(zd + (lo - la) * DC + c4);
              float4 v;
              v.x = fmaf(q1.x, zz.x, fmaf(q0.x, y.x, q2.x)); v.y = fmaf(q1.y, zz.y, fmaf(q0.y, y.y, q2.y));
              v.z = fmaf(q1.z, zz.z, fmaf(q0.z, y.z, q2.z)); v.w = fmaf(q1.w, zz.w, fmaf(q0.w, y.w, q2.w));
              if (K == 5) {
                const float4 wt = lds4(cfa + (3 + t) * DC * 4);
                g.x = fmaf(wt.x, v.x, g.x); g.y = fmaf(wt.y, v.y, g.y); g.z = fmaf(wt.z, v.z, g.z); g.w = fmaf(wt.w, v.w, g.w);
              } else {
                g.x = fmaf(w[0][K == 5 ? 0 : t], v.x, g.x); g.y = fmaf(w[1][K == 5 ? 0 : t], v.y, g.y);
                g.z = fmaf(w[2][K == 5 ? 0 : t], v.z, g.z); g.w = fmaf(w[3][K == 5 ? 0 : t], v.w, g.w);
              }
              acc[t][0] = fmaf(v.x, a.x, acc[t][0]); acc[t][1] = fmaf(v.y, a.y, acc[t][1]);
              acc[t][2] = fmaf(v.z, a.z, acc[t][2]); acc[t][3] = fmaf(v.w, a.w, acc[t][3]);
            }
          }
        }
        g.x = a.x > 0.f ? g.x : 0.f; g.y = a.y > 0.f ? g.y : 0.f; g.z = a.z > 0.f ? g.z : 0.f; g.w = a.w > 0.f ? g.w : 0.f;
        *(float4*)(p.out + ((size_t)bt * p.Lin + li) * DC + c4) = g;
        acc[K][0] += g.x; acc[K][1] += g.y; acc[K][2] += g.z; acc[K][3] += g.w;
        const float4 bmean = lds4(cfa + 2 * DC * 4);
        acc[K + 1][0] = fmaf(g.x, z.x - bmean.x, acc[K + 1][0]); acc[K + 1][1] = fmaf(g.y, z.y - bmean.y, acc[K + 1][1]);
        acc[K + 1][2] = fmaf(g.z, z.z - bmean.z, acc[K + 1][2]); acc[K + 1][3] = fmaf(g.w, z.w - bmean.w, acc[K + 1][3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mi.empty[st]);
      if (++st == nst) { st = 0; ph ^= 1; }
    }
    const float4 binv = ld4(p.binv + c4);
    acc[K + 1][0] *= binv.x; acc[K + 1][1] *= binv.y; acc[K + 1][2] *= binv.z; acc[K + 1][3] *= binv.w;
    float* red = (float*)(smem + Kp.red_off);
#pragma unroll
    for (int t = 0; t < K; ++t) reduce_channels(red, acc[t], cw, c4, nullptr, p.dw + t, K);
    reduce_channels(red, acc[K], cw, c4, p.bstats, nullptr, 0);
    reduce_channels(red, acc[K + 1], cw, c4, p.bstats + DC, nullptr, 0);
  }
}

template <int K, int S>
int launch_fwd_t(DwK& Kp, int grid, size_t smem, cudaStream_t st) {
  TRU_SMEM_OPT_IN((dw_fwd_stream_kernel<K, S>), SMEM_MAX);
  TRU_CUDA(launch_pdl(dw_fwd_stream_kernel<K, S>, dim3(grid), dim3(NT), smem, st, Kp));
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}
template <int K, int S>
int launch_bwd_t(DwK& Kp, int grid, size_t smem, cudaStream_t st) {
  TRU_SMEM_OPT_IN((dw_bwd_stream_kernel<K, S>), SMEM_MAX);
  TRU_CUDA(launch_pdl(dw_bwd_stream_kernel<K, S>, dim3(grid), dim3(NT), smem, st, Kp));
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

bool layout(DwK& Kp, bool bwd, size_t& smem) {
  const DwParams& p = Kp.p;
  const size_t redb = (size_t)CW * DC * 4, miscb = align_up(sizeof(DwMisc), 128);
  size_t stage;
  if (!bwd) {
    const int RO = std::min(RU, p.Lout);
    if (p.Lout % RO) return false;
    Kp.upf = p.Lout / RO;
    stage = align_up((size_t)((RO - 1) * p.stride + p.k) * DC * 4, 128);
  } else {
    const int RI = std::min(RU, p.Lin);
    if (p.Lin % RI) return false;
    Kp.upf = p.Lin / RI;
    const size_t zb = (size_t)RI * DC * 4, db = align_up((size_t)(RI / p.stride + p.k) * DC * 4, 128);
    Kp.in2_off = (uint32_t)zb; Kp.in3_off = (uint32_t)(zb + db);
    stage = zb + 2 * db;
  }
  Kp.stage_bytes = (uint32_t)stage;
  const size_t coefb = bwd ? (size_t)(3 + 5) * DC * 4 : 0;    // backward: ReLU-mask affine + BN mean of the layer in front, 5-tap weights (see the kernel)
  Kp.nstage = (int)std::min<size_t>(MAXST, (SMEM_MAX - redb - miscb - coefb - 128) / stage);
  if (Kp.nstage < 2) return false;
  Kp.red_off = (uint32_t)(Kp.nstage * stage);
  Kp.misc_off = (uint32_t)(Kp.red_off + redb);
  Kp.coef_off = (uint32_t)(Kp.misc_off + miscb);
  smem = Kp.coef_off + coefb;
  Kp.units = (unsigned)p.BT * Kp.upf;
  return true;
}

}  // namespace

// 0 launched, 1 shape not covered (caller uses the direct kernels)
int launch_dw_fwd_stream(const DwParams& p, cudaStream_t st) {
  if (p.C != DC || !((p.k == 3 && (p.stride == 1 || p.stride == 2)) || (p.k == 5 && p.stride == 2)) || p.pad != p.k / 2) return 1;
  DwK Kp{};
  Kp.p = p;
  size_t smem = 0;
  if (!layout(Kp, false, smem)) return 1;
  const int grid = (int)std::min<unsigned>(sm_count(), Kp.units);
  ProfScope prof("dw_fwd", 4.0 * p.BT * ((double)p.Lin + p.Lout) * p.C, 2.0 * p.k * p.BT * p.Lout * p.C, st);
  if (p.k == 3 && p.stride == 1) return launch_fwd_t<3, 1>(Kp, grid, smem, st);
  if (p.k == 3) return launch_fwd_t<3, 2>(Kp, grid, smem, st);
  return launch_fwd_t<5, 2>(Kp, grid, smem, st);
}

int launch_dw_bwd_stream(const DwParams& p, cudaStream_t st) {
  if (p.C != DC || !((p.k == 3 && (p.stride == 1 || p.stride == 2)) || (p.k == 5 && p.stride == 2)) || p.pad != p.k / 2) return 1;
  if (!(p.zmask && p.mp0 && p.bstats && p.dw && p.p0 && p.p1 && p.src2 && p.a_src == p.zmask && p.a_p0 == p.mp0 && p.a_p2 == p.mp2)) return 1;
  DwK Kp{};
  Kp.p = p;
  size_t smem = 0;
  if (!layout(Kp, true, smem)) return 1;
  const int grid = (int)std::min<unsigned>(sm_count(), Kp.units);
  ProfScope prof("dw_bwd_fused", 4.0 * p.BT * (2.0 * p.Lin + 2.0 * p.Lout) * p.C, 4.0 * p.k * p.BT * p.Lout * p.C, st);
  if (p.k == 3 && p.stride == 1) return launch_bwd_t<3, 1>(Kp, grid, smem, st);
  if (p.k == 3) return launch_bwd_t<3, 2>(Kp, grid, smem, st);
  return launch_bwd_t<5, 2>(Kp, grid, smem, st);
}

}  // namespace tru
