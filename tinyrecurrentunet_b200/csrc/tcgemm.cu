// Tensor-core implicit GEMM (tcgen05 / TMEM), fp32-accurate via the 3xTF32 split.
//
// Same contract as igemm.cu (gathered rows x resident weights, BN/ReLU applied on
// load, bias / BN statistics / ReLU mask / skip-gradient epilogue), for the shapes
// where every K-segment has a multiple of 32 channels and N is a multiple of 32.
//
// One persistent CTA per SM, 13 warps:
//   warp 0      MMA issuer (one elected thread) + TMEM allocation
//   warps 1-8   loaders: gather 128 rows x 32 channels from HBM, apply the affine
//               (+ReLU or BN-backward) transform, split into tf32 hi/lo and write both
//               into the 128B-swizzled K-major operand tiles of a shared-memory ring
//   warps 9-12  epilogue: TMEM -> registers -> 16-byte stores (one accumulator row =
//               one 128-byte line per thread), BN statistics by a shuffle butterfly
// The CTA's slice of the weights (hi and lo) stays resident in shared memory.
// For every 32-wide k-block the MMA thread issues 4 k-steps x 3 products
// (lo*hi + hi*lo + hi*hi) of tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) into one of
// two TMEM accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <algorithm>
#include "net_kernels.cuh"
#include "tc_common.cuh"

namespace tru {
namespace {
using namespace tc;

constexpr int BM = 128, KBLK = 32;
constexpr int LOAD_WARPS = 8, EPI_WARPS = 4;
constexpr int RPT = BM * 8 / (32 * LOAD_WARPS);           // rows per loader thread and k-block (4)
constexpr int NT = 32 * (1 + LOAD_WARPS + EPI_WARPS);    // 416
constexpr int NLOAD = 32 * LOAD_WARPS, NEPI = 32 * EPI_WARPS;
constexpr int A_TILE = BM * 128;                          // bytes of one hi (or lo) operand tile
constexpr int STAGE = 2 * A_TILE;
constexpr int MAXKB = 16;
constexpr int EPI_LD = 36;                                // floats per staged row (conflict-free 16-byte accesses)
constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_LD * 4;

struct TcLayout { int BN, nkb, nstage; uint32_t w_off, a_off, epi_off, misc_off; };

struct Misc {
  uint64_t full[8], empty[8], tfull[2], tempty[2];
  uint32_t tmem_base;
  int kb_seg[MAXKB], kb_c0[MAXKB];
  alignas(16) float bias[128];
  alignas(16) float mp0[128];
  alignas(16) float mp2[128];
  alignas(16) float bmean[128];
  alignas(16) float binv[128];
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }

// LD2: loaders read two tensors (dY and Z) and apply the BN-backward affine; EPI: the epilogue adds the
// skip gradient / applies the ReLU mask / accumulates the BN-backward sums.
template <bool LD2, bool EPI>
__global__ void __launch_bounds__(NT, 1) tc_igemm_kernel(const __grid_constant__ IgemmParams P, const TcLayout Lo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* Wsm = smem + Lo.w_off;                      // [hi|lo][nkb][BN rows][128 B], swizzled
  uint8_t* Asm = smem + Lo.a_off;                      // ring: [stage][hi|lo][128 rows][128 B]
  Misc& mi = *(Misc*)(smem + Lo.misc_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = Lo.BN, nkb = Lo.nkb, nstage = Lo.nstage;
  const int n0 = blockIdx.y * BN;
  const long M = (long)P.BT * P.Lq;
  const int ntiles = (int)((M + BM - 1) / BM);
  const int n_my = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // ---- one-time setup -------------------------------------------------------------
  {
    // segments whose channel count is not a multiple of 32 are zero padded to whole k-blocks
    for (uint32_t i = tid; i < (uint32_t)nkb * BN * 64; i += NT) ((uint32_t*)Wsm)[i] = 0u;
    __syncthreads();
    int kbb = 0;                                  // first k-block of the segment
    for (int s = 0; s < P.nseg; ++s) {
      const Seg& sg = P.seg[s];
      const int skb = (sg.C + KBLK - 1) / KBLK;
      if (tid == 0)
        for (int b = 0; b < skb; ++b) { mi.kb_seg[kbb + b] = s; mi.kb_c0[kbb + b] = b * KBLK; }
      for (int i = tid; i < sg.C * BN; i += NT) {
        int c, n;
        if (sg.wsc == 1) { c = i % sg.C; n = i / sg.C; } else { n = i % BN; c = i / BN; }
        float v = 0.f;
        if (n0 + n < P.N) v = __ldg(sg.W + sg.wbase + (long)c * sg.wsc + (long)(n0 + n) * sg.wsn);
        const int kb = kbb + (c >> 5), kk = c & 31;
        const uint32_t off = (uint32_t)kb * BN * 128 + n * 128 + ((((kk >> 2) ^ (n & 7)) << 4) | ((kk & 3) << 2));
        const uint32_t hi = f2tf32(v);
        *(uint32_t*)(Wsm + off) = hi;
        *(float*)(Wsm + (uint32_t)nkb * BN * 128 + off) = v - __uint_as_float(hi);
      }
      kbb += skb;
    }
    for (int i = tid; i < BN; i += NT) {
      const bool in = n0 + i < P.N;
      mi.bias[i] = (P.bias && in) ? __ldg(P.bias + n0 + i) : 0.f;
      mi.mp0[i] = (P.use_mask && P.mp0 && in) ? __ldg(P.mp0 + n0 + i) : 1.f;
      mi.mp2[i] = (P.use_mask && P.mp0 && in) ? __ldg(P.mp2 + n0 + i) : 0.f;
      mi.bmean[i] = (P.bstats && in) ? __ldg(P.bmean + n0 + i) : 0.f;
      mi.binv[i] = (P.bstats && in) ? __ldg(P.binv + n0 + i) : 0.f;
    }
  }
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < nstage; ++s) { mbar_init(&mi.full[s], LOAD_WARPS); mbar_init(&mi.empty[s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&mi.tfull[a], 1); mbar_init(&mi.tempty[a], NEPI); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&mi.tmem_base, 2 * BN);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mi.tmem_base;

  if (warp == 0) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(BM, BN, 0, 0);
      // descriptor = constant high word | 14-bit (address >> 4); K-steps advance the address by 32 B
      const uint64_t dhi = (uint64_t)(smem_desc_sw128(0, 16, 1024) >> 32) << 32 | (1ull << 16);
      const uint32_t a_base = smem_u32(Asm) >> 4, w_base = smem_u32(Wsm) >> 4;
      const uint32_t w_lo_off = ((uint32_t)nkb * BN * 128) >> 4;
      int st = 0;
      uint32_t ph = 0;
      for (int ti = 0; ti < n_my; ++ti) {
        const int acc = ti & 1;
        mbar_wait(&mi.tempty[acc], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&mi.full[st], ph);
          tc_fence_after();
          const uint32_t a_hi = a_base + st * (STAGE >> 4), a_lo = a_hi + (A_TILE >> 4);
          const uint32_t w_hi = w_base + (((uint32_t)kb * BN * 128) >> 4), w_lo = w_hi + w_lo_off;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t dah = dhi | (a_hi + 2 * j), dal = dhi | (a_lo + 2 * j);
            const uint64_t dbh = dhi | (w_hi + 2 * j), dbl = dhi | (w_lo + 2 * j);
            mma_tf32(d, dal, dbh, idesc, (kb | j) != 0);
            mma_tf32(d, dah, dbl, idesc, 1);
            mma_tf32(d, dah, dbh, idesc, 1);
          }
          mma_commit(&mi.empty[st]);
          if (++st == nstage) { st = 0; ph ^= 1; }
        }
        mma_commit(&mi.tfull[acc]);
      }
    }
  } else if (warp <= LOAD_WARPS) {
    // ================================== loaders ====================================
    const int lt = tid - 32, chunk = lt & 7, r0 = lt >> 3;
    const int total = n_my * nkb;
    const unsigned Mu = (unsigned)M, Lq = (unsigned)P.Lq;
    constexpr int NS = LD2 ? 2 : 3;                     // register slots; prefetch distance NS-1 k-blocks
    constexpr int RSTEP = BM / RPT;
    float4 va[NS][RPT], vb[LD2 ? NS : 1][RPT];
    unsigned vmask[NS];
    // issue-side cursor (runs two k-blocks ahead of the commit-side cursor)
    int i_kb = 0, i_ti = 0, i_seg = -1;
    unsigned rbt[RPT], rq[RPT];
    bool rok[RPT];
    int roff[RPT];
    auto decode_rows = [&]() {
      const unsigned m0 = ((unsigned)blockIdx.x + (unsigned)i_ti * gridDim.x) * BM;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const unsigned m = m0 + r0 + RSTEP * i;
        rok[i] = m < Mu;
        rbt[i] = rok[i] ? m / Lq : 0u;
        rq[i] = rok[i] ? m - rbt[i] * Lq : 0u;
      }
      i_seg = -1;
    };
    decode_rows();
    auto issue = [&](float4 (&a)[RPT], float4 (&b)[RPT], unsigned& msk) {
      const int s = mi.kb_seg[i_kb];
      const Seg& sg = P.seg[s];
      if (s != i_seg) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int li = (int)rq[i] * sg.smul + sg.sadd;
          roff[i] = (rok[i] && li >= 0 && li < sg.Lsrc) ? ((int)rbt[i] * sg.Lsrc + li) * sg.ld + sg.coff : -1;
        }
        i_seg = s;
      }
      const int c = mi.kb_c0[i_kb] + chunk * 4;
      msk = 0;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        if (roff[i] >= 0 && c < sg.C) {
          a[i] = ld4(sg.src + (unsigned)(roff[i] + c));
          if (LD2 && sg.src2) b[i] = ld4(sg.src2 + (unsigned)(roff[i] + c));
          msk |= 1u << i;
        }
      }
      if (++i_kb == nkb) { i_kb = 0; ++i_ti; decode_rows(); }
    };
    // commit-side cursor
    int c_kb = 0, c_st = 0;
    uint32_t c_ph = 0;
    auto commit = [&](float4 (&a)[RPT], float4 (&b)[RPT], unsigned msk) {
      const Seg& sg = P.seg[mi.kb_seg[c_kb]];
      const int c = sg.coff + mi.kb_c0[c_kb] + chunk * 4;
      float4 p0 = make_float4(1, 1, 1, 1), p1 = make_float4(0, 0, 0, 0), p2 = p1;
      if (sg.p0 && mi.kb_c0[c_kb] + chunk * 4 < sg.C) {
        p0 = ld4(sg.p0 + c); p2 = ld4(sg.p2 + c);
        if (LD2 && sg.p1) p1 = ld4(sg.p1 + c);
      }
      mbar_wait(&mi.empty[c_st], c_ph ^ 1);
      uint8_t* ah = Asm + c_st * STAGE;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (msk & (1u << i)) {
          v = a[i];
          if (sg.p0) {
            v.x = p0.x * v.x + p2.x; v.y = p0.y * v.y + p2.y; v.z = p0.z * v.z + p2.z; v.w = p0.w * v.w + p2.w;
            if (LD2 && sg.p1) { v.x += p1.x * b[i].x; v.y += p1.y * b[i].y; v.z += p1.z * b[i].z; v.w += p1.w * b[i].w; }
            if (sg.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          }
        }
        uint4 hi, lo;
        split_tf32(v, hi, lo);
        const int row = r0 + RSTEP * i;
        const uint32_t off = row * 128 + ((chunk ^ (row & 7)) << 4);
        *(uint4*)(ah + off) = hi;
        *(uint4*)(ah + A_TILE + off) = lo;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mi.full[c_st]);
      if (++c_kb == nkb) c_kb = 0;
      if (++c_st == nstage) { c_st = 0; c_ph ^= 1; }
    };

#pragma unroll
    for (int u = 0; u < NS - 1; ++u)
      if (u < total) issue(va[u], vb[LD2 ? u : 0], vmask[u]);
    for (int w = 0; w < total; w += NS) {
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        const int ww = w + u;
        if (ww < total) {
          constexpr int dummy = 0; (void)dummy;
          if (ww + NS - 1 < total) issue(va[(u + NS - 1) % NS], vb[LD2 ? (u + NS - 1) % NS : 0], vmask[(u + NS - 1) % NS]);
          commit(va[u], vb[LD2 ? u : 0], vmask[u]);
        }
      }
    }
  } else {
    // ================================== epilogue ===================================
    // Each warp drains its 32 accumulator rows in 32-column chunks: TMEM -> registers
    // (+bias) -> warp-private padded smem tile -> transposed read so that every global
    // access (store, ReLU-mask load, skip-gradient load) covers whole 128-byte lines.
    // BN statistics: a lane keeps the same 4 columns for all rows, sums stay in registers.
    const int lgrp = warp & 3;                          // TMEM lane group this warp may read
    float* tile = (float*)(smem + Lo.epi_off) + (warp - 1 - LOAD_WARPS) * (32 * EPI_LD);
    const int c4 = (lane & 7) * 4, rsub = lane >> 3;
    const unsigned Mu = (unsigned)M, Lq = (unsigned)P.Lq;
    float st1[4][4], st2[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) { st1[a][b] = 0.f; st2[a][b] = 0.f; }
    const bool do_stats = P.stats != nullptr, do_bstats = EPI && P.bstats != nullptr;
    for (int ti = 0; ti < n_my; ++ti) {
      const int acc = ti & 1;
      const unsigned mw = ((unsigned)blockIdx.x + (unsigned)ti * gridDim.x) * BM + lgrp * 32;   // first row of this warp
      // output row index of the 8 rows this lane touches in the transposed phase: rows rsub + 4*i
      int orow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const unsigned m = mw + rsub + 4 * i;
        orow[i] = -1;
        if (m < Mu) {
          const unsigned bt = m / Lq, q = m - bt * Lq;
          orow[i] = (int)(bt * P.Lout + q * P.omul + P.oadd);
        }
      }
      mbar_wait(&mi.tfull[acc], (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc * 32 < BN) {
          uint32_t v[32];
          tmem_ld32(tmem + ((uint32_t)(lgrp * 32) << 16) + acc * BN + cc * 32, v);
          if ((cc + 1) * 32 >= BN) { tc_fence_before(); mbar_arrive(&mi.tempty[acc]); }
          __syncwarp();                                 // previous chunk's readers are done
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            *(float4*)(tile + lane * EPI_LD + j4 * 4) =
                make_float4(__uint_as_float(v[j4 * 4]), __uint_as_float(v[j4 * 4 + 1]), __uint_as_float(v[j4 * 4 + 2]),
                            __uint_as_float(v[j4 * 4 + 3]));
          __syncwarp();
          const int col = cc * 32 + c4;                 // column inside the CTA's N tile
          const float4 bias = *(const float4*)&mi.bias[col];
          float4 mp0 = make_float4(1, 1, 1, 1), mp2 = make_float4(0, 0, 0, 0), bmean = mp2, binv = mp2;
          if (EPI && P.use_mask) { mp0 = *(const float4*)&mi.mp0[col]; mp2 = *(const float4*)&mi.mp2[col]; }
          if (do_bstats) { bmean = *(const float4*)&mi.bmean[col]; binv = *(const float4*)&mi.binv[col]; }
          // two half-chunks of 4 rows; the global loads of a half are issued before its stores
          // (a load cannot be hoisted above a store the compiler cannot prove disjoint)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float4 zv[4], xv[4];
            if (EPI) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = orow[h * 4 + i];
                zv[i] = make_float4(0.f, 0.f, 0.f, 0.f); xv[i] = zv[i];
                if (r >= 0 && n0 + col < P.N) {
                  if (P.use_mask) zv[i] = ld4(P.zmask + (unsigned)r * (unsigned)P.ldo + P.ocoff + n0 + col);
                  if (P.extra) xv[i] = ld4(P.extra + (unsigned)r * (unsigned)P.ext_ld + n0 + col);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = orow[h * 4 + i];
              if (r < 0) continue;
              float4 o = *(const float4*)(tile + (rsub + 4 * (h * 4 + i)) * EPI_LD + c4);
              o.x += bias.x; o.y += bias.y; o.z += bias.z; o.w += bias.w;
              const unsigned gofs = (unsigned)r * (unsigned)P.ldo + P.ocoff + n0 + col;
              if (EPI) {
                o.x += xv[i].x; o.y += xv[i].y; o.z += xv[i].z; o.w += xv[i].w;
                if (P.use_mask) {
                  const float4 z = zv[i];
                  o.x = (z.x * mp0.x + mp2.x > 0.f) ? o.x : 0.f; o.y = (z.y * mp0.y + mp2.y > 0.f) ? o.y : 0.f;
                  o.z = (z.z * mp0.z + mp2.z > 0.f) ? o.z : 0.f; o.w = (z.w * mp0.w + mp2.w > 0.f) ? o.w : 0.f;
                  if (do_bstats) {
                    st1[cc][0] += o.x; st1[cc][1] += o.y; st1[cc][2] += o.z; st1[cc][3] += o.w;
                    st2[cc][0] += o.x * (z.x - bmean.x) * binv.x; st2[cc][1] += o.y * (z.y - bmean.y) * binv.y;
                    st2[cc][2] += o.z * (z.z - bmean.z) * binv.z; st2[cc][3] += o.w * (z.w - bmean.w) * binv.w;
                  }
                }
              }
              if (do_stats) {
                st1[cc][0] += o.x; st1[cc][1] += o.y; st1[cc][2] += o.z; st1[cc][3] += o.w;
                st2[cc][0] += o.x * o.x; st2[cc][1] += o.y * o.y; st2[cc][2] += o.z * o.z; st2[cc][3] += o.w * o.w;
              }
              if (n0 + col < P.N) {
                if (P.planar) {                       // (BT, N, Lout) layout of the network output
                  const unsigned bt = (unsigned)r / (unsigned)P.Lout, lo = (unsigned)r - bt * P.Lout;
                  float* op = P.out + ((size_t)bt * P.N + n0 + col) * P.Lout + lo;
                  op[0] = o.x; op[P.Lout] = o.y; op[2 * (size_t)P.Lout] = o.z; op[3 * (size_t)P.Lout] = o.w;
                } else {
                  *(float4*)(P.out + gofs) = o;
                }
              }
            }
          }
        }
      }
    }
    double* gst = P.stats ? P.stats : (EPI ? P.bstats : nullptr);
    if (gst) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
        if (cc * 32 < BN)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a = st1[cc][j], b = st2[cc][j];
            a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
            b += __shfl_xor_sync(0xffffffffu, b, 8); b += __shfl_xor_sync(0xffffffffu, b, 16);
            const int n = n0 + cc * 32 + c4 + j;
            if (lane < 8 && n < P.N) {
              atomicAdd(gst + n, (double)a);
              atomicAdd(gst + P.N + n, (double)b);
            }
          }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 2 * BN);
  }
}

constexpr size_t SMEM_MAX = 227 * 1024;

bool plan(const IgemmParams& p, TcLayout& L, dim3& grid, size_t& smem_bytes) {
  if ((p.N % 32 != 0 && p.N > 32) || p.N % 4 != 0 || p.ldo % 4 != 0 || p.ocoff % 4 != 0) return false;
  if (p.planar && p.N > 32) return false;
  int nkb = 0;
  for (int s = 0; s < p.nseg; ++s) {
    if (p.seg[s].C % 4 != 0 || p.seg[s].ld % 4 != 0 || p.seg[s].coff % 4 != 0) return false;
    nkb += (p.seg[s].C + KBLK - 1) / KBLK;
  }
  if (nkb > MAXKB) return false;
  // 32-bit element offsets inside the kernel
  for (int s = 0; s < p.nseg; ++s)
    if ((double)p.BT * p.seg[s].Lsrc * p.seg[s].ld >= 2147483648.0) return false;
  if ((double)p.BT * p.Lout * p.ldo >= 4294967296.0 || (double)p.BT * p.Lq + BM >= 4294967296.0) return false;
  const size_t fixed = 1024 /* alignment slack */ + EPI_BYTES + sizeof(Misc) + 256;
  for (int BN : {128, 64, 32}) {
    if (p.N % BN != 0 && !(BN == 32 && p.N < 32)) continue;
    const size_t w = (size_t)2 * nkb * BN * 128;
    if (fixed + w + 2 * STAGE > SMEM_MAX) continue;
    int nstage = (int)std::min<size_t>(8, (SMEM_MAX - fixed - w) / STAGE);
    L.BN = BN; L.nkb = nkb; L.nstage = nstage;
    L.w_off = 0;
    L.a_off = (uint32_t)((w + 1023) / 1024 * 1024);
    L.epi_off = L.a_off + nstage * STAGE;
    L.misc_off = L.epi_off + EPI_BYTES;
    smem_bytes = 1024 + L.misc_off + sizeof(Misc);
    const int ny = (p.N + BN - 1) / BN;
    const long M = (long)p.BT * p.Lq;
    const int ntiles = (int)((M + BM - 1) / BM);
    grid = dim3(std::max(1, std::min(ntiles, sm_count() / ny)), ny);
    return smem_bytes <= SMEM_MAX;
  }
  return false;
}

}  // namespace

bool igemm_tc_eligible(const IgemmParams& p) {
  TcLayout L{};
  dim3 grid;
  size_t smem = 0;
  return plan(p, L, grid, smem);
}

// returns TRU_OK if launched, 1 if the shape is not eligible (caller falls back to the FFMA kernel)
int launch_igemm_tc(const IgemmParams& p, cudaStream_t st) {
  TcLayout L{};
  dim3 grid;
  size_t smem = 0;
  if (!plan(p, L, grid, smem)) return 1;
  const bool epi = p.use_mask || p.extra != nullptr || p.bstats != nullptr;
  bool ld2 = false;
  for (int s = 0; s < p.nseg; ++s) ld2 |= (p.seg[s].src2 != nullptr);
  static bool attr_done = false;
  if (!attr_done) {
    TRU_CUDA(cudaFuncSetAttribute(tc_igemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    TRU_CUDA(cudaFuncSetAttribute(tc_igemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    TRU_CUDA(cudaFuncSetAttribute(tc_igemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    TRU_CUDA(cudaFuncSetAttribute(tc_igemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    attr_done = true;
  }
  if (ld2 && epi) tc_igemm_kernel<true, true><<<grid, NT, smem, st>>>(p, L);
  else if (ld2) tc_igemm_kernel<true, false><<<grid, NT, smem, st>>>(p, L);
  else if (epi) tc_igemm_kernel<false, true><<<grid, NT, smem, st>>>(p, L);
  else tc_igemm_kernel<false, false><<<grid, NT, smem, st>>>(p, L);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace tru
