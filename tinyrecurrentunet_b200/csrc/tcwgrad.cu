// Tensor-core weight gradient (tcgen05 / TMEM, 3xTF32):  dW[c][n] += sum_m a(m,c) * dz(m,n).
//
// The reduction runs over ROWS, so both operands are consumed "MN-major": a k-block is
// 32 consecutive rows of the two row-major activation tensors, written as-is (channels
// contiguous) into SWIZZLE_128B_BASE32B shared-memory atoms of 4 rows x 32 channels.  One
// side ("P", 128 channels, zero padded) becomes the M dimension of the MMA, the other
// ("Q", 32..128 channels) the N dimension; the roles are chosen per job so that the
// wider tensor sits on P.  A CTA owns a contiguous range of rows and keeps ONE fp32
// accumulator (128 x Q) in TMEM for its whole range; at the end the accumulator is
// added to dW with atomics.  a() is the forward activation (BN + ReLU applied on load),
// dz() the BN-backward-transformed gradient, exactly as in wgrad_kernel (igemm.cu).
//
// 25 warps: warp 0 MMA issuer, warps 1-8 load the activation side, warps 9-24 the dz side
// (which reads two tensors when a BatchNorm sits behind the layer); the bias gradient
// (column sums of dz) is accumulated by the dz-side loaders on the fly.
#include <algorithm>
#include "net_kernels.cuh"
#include "tc_common.cuh"

namespace tru {
namespace {
using namespace tc;

constexpr int KROWS = 32;                       // rows per k-block
constexpr int NT = 32 * 20;                     // 4-warp MMA group (warp 0 issues) + 8 loader warps per side
constexpr int REG_MMA = 32, REG_LOAD = 112;     // setmaxnreg budgets: 4*32*32 + 16*32*112 = 61440 = 640 threads x 96
constexpr int PW = 128;                         // P tile width (channels)
constexpr int P_TILE = KROWS * PW * 4;          // 16 KB per hi (or lo)
constexpr int NSTAGE = 3;

struct Side {          // one operand of a job, as seen by its loaders
  const float* src; const float* src2; const float* p0; const float* p1; const float* p2;
  int L, ld, coff, mul, add, C;      // rows per frame, row stride, first channel, row map, channels
  int relu, c0;                      // c0: first channel of this CTA's tile
};
struct TcWJob {
  Side P, Q;
  float* dW; int wbase, sp, sq;      // dW[wbase + p*sp + q*sq]
  float* db; int db_on_p;            // bias gradient = column sums of the dz side
  int ptiles, qtiles, QW;            // tiles along P (128 wide) and Q (QW wide)
};
struct TcWParams { TcWJob job[8]; int njobs, BT, Lq, nsplit; };

struct WMisc { uint64_t full[NSTAGE], empty[NSTAGE], done; uint32_t tmem_base; };

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }

// Loader of one side.  W = tile width in channels (128 for P, QW for Q); NTHR threads cooperate.
// The unit of the software pipeline is UR rows (32 = a whole k-block, or 16 = half of one, which
// halves the registers a two-tensor side needs for the same number of bytes in flight); NS-1 units
// are kept in flight.  Rows are decoded incrementally (one division per thread and pass at the
// start, none per unit); the tf32 split is the truncating one (hi = x & ~0x1fff, lo = x - hi).
template <int NTHR, int UR, int MAXPASS, int NS, bool HAS2>
__device__ __forceinline__ void side_loader(const TcWParams& P, const TcWJob& J, const Side& S, int c0, int W, int lt, int lane,
                                            uint8_t* ring, int side_off, WMisc& mi, unsigned kb0, unsigned kb1,
                                            bool want_db) {
  constexpr int UPK = KROWS / UR;                // units per k-block
  const int cpr = W >> 2;                        // 16-byte chunks per row
  const int cidx = lt % cpr, rsub = lt / cpr, rstep = NTHR / cpr;
  const int npass = rstep >= UR ? 1 : UR / rstep;
  const bool t_ok = rsub < UR;
  const int mb = cidx >> 3, ch = cidx & 7, NB = W >> 5;
  const int c = c0 + cidx * 4;                   // channel inside the tensor (before coff)
  const bool c_ok = c < S.C && t_ok;
  float4 p0 = make_float4(1, 1, 1, 1), p1 = make_float4(0, 0, 0, 0), p2 = p1;
  const bool affine = S.p0 != nullptr;
  if (affine && c_ok) {
    p0 = ld4(S.p0 + S.coff + c); p2 = ld4(S.p2 + S.coff + c);
    if (HAS2 && S.p1) p1 = ld4(S.p1 + S.coff + c);
  }
  const float fl = S.relu ? 0.f : -__int_as_float(0x7f800000);
  float bs[4] = {0.f, 0.f, 0.f, 0.f};
  const unsigned Lq = (unsigned)P.Lq, Mrows = (unsigned)P.BT * Lq;
  const unsigned dq = UR % Lq, dbt = UR / Lq;                  // row advance per unit, as (frames, rows)
  const float* src = S.src + S.coff + c;
  const float* src2 = (HAS2 && S.src2) ? S.src2 + S.coff + c : nullptr;
  const int mul = S.mul, add = S.add, SL = S.L, ld = S.ld;
  // issue-side row cursor of every pass: row m = unit*UR + rsub + rstep*i = bt*Lq + q
  unsigned rm0 = kb0 * KROWS + rsub, rbt[MAXPASS], rq[MAXPASS];
#pragma unroll
  for (int i = 0; i < MAXPASS; ++i) {
    const unsigned m = rm0 + rstep * i;
    rbt[i] = m / Lq; rq[i] = m - rbt[i] * Lq;
  }
  // shared-memory offset of (row rsub, this thread's chunk); pass i adds i*rstep rows = i*(rstep/4) atoms
  const uint32_t st_off = (uint32_t)((rsub >> 2) * NB + mb) * 512 + (rsub & 3) * 128 + ((((ch >> 1) ^ (rsub & 3)) << 5) | ((ch & 1) << 4));
  const uint32_t st_step = (uint32_t)(rstep >> 2) * NB * 512, st_unit = (uint32_t)(UR >> 2) * NB * 512;
  float4 va[NS][MAXPASS], vb[HAS2 ? NS : 1][HAS2 ? MAXPASS : 1];
  unsigned msk[NS];
  auto issue = [&](float4 (&a)[MAXPASS], float4 (&b)[HAS2 ? MAXPASS : 1], unsigned& mk) {
    mk = 0;
#pragma unroll
    for (int i = 0; i < MAXPASS; ++i) {
      if (i < npass) {
        const int l = (int)rq[i] * mul + add;
        if (c_ok && rm0 + rstep * i < Mrows && (unsigned)l < (unsigned)SL) {
          const unsigned off = (rbt[i] * SL + l) * ld;
          a[i] = ld4(src + off);
          if (HAS2) b[HAS2 ? i : 0] = src2 ? ld4(src2 + off) : make_float4(0.f, 0.f, 0.f, 0.f);
          mk |= 1u << i;
        }
        rq[i] += dq; rbt[i] += dbt;
        if (rq[i] >= Lq) { rq[i] -= Lq; ++rbt[i]; }
      }
    }
    rm0 += UR;
  };
  int st = 0;
  uint32_t ph = 0;
  const unsigned u1 = (kb1 - kb0) * UPK;        // units of this CTA
#pragma unroll
  for (int u = 0; u < NS - 1; ++u)
    if ((unsigned)u < u1) issue(va[u], vb[HAS2 ? u : 0], msk[u]);
  for (unsigned ub = 0; ub < u1; ub += NS) {
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const unsigned k = ub + u;
      if (k < u1) {
        if (k + NS - 1 < u1) issue(va[(u + NS - 1) % NS], vb[HAS2 ? (u + NS - 1) % NS : 0], msk[(u + NS - 1) % NS]);
        const unsigned sub = UPK == 1 ? 0u : (k & (UPK - 1));
        if (sub == 0) mbar_wait(&mi.empty[st], ph ^ 1);
        uint8_t* base = ring + st * (4 * P_TILE) + side_off + st_off + sub * st_unit;   // stage = [P hi | P lo | Q hi | Q lo]
        if (t_ok) {
#pragma unroll
          for (int i = 0; i < MAXPASS; ++i) {
            if (i < npass) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (msk[u] & (1u << i)) {
                v = va[u][i];
                if (affine) {
                  if (HAS2) {
                    const float4 z = vb[HAS2 ? u : 0][HAS2 ? i : 0];
                    v.x = fmaf(p1.x, z.x, fmaf(p0.x, v.x, p2.x)); v.y = fmaf(p1.y, z.y, fmaf(p0.y, v.y, p2.y));
                    v.z = fmaf(p1.z, z.z, fmaf(p0.z, v.z, p2.z)); v.w = fmaf(p1.w, z.w, fmaf(p0.w, v.w, p2.w));
                  } else {
                    v.x = fmaf(p0.x, v.x, p2.x); v.y = fmaf(p0.y, v.y, p2.y); v.z = fmaf(p0.z, v.z, p2.z); v.w = fmaf(p0.w, v.w, p2.w);
                  }
                  v.x = fmaxf(v.x, fl); v.y = fmaxf(v.y, fl); v.z = fmaxf(v.z, fl); v.w = fmaxf(v.w, fl);
                }
                bs[0] += v.x; bs[1] += v.y; bs[2] += v.z; bs[3] += v.w;
              }
              uint4 hi, lo;
              hi.x = __float_as_uint(v.x) & 0xffffe000u; hi.y = __float_as_uint(v.y) & 0xffffe000u;
              hi.z = __float_as_uint(v.z) & 0xffffe000u; hi.w = __float_as_uint(v.w) & 0xffffe000u;
              lo.x = __float_as_uint(v.x - __uint_as_float(hi.x)); lo.y = __float_as_uint(v.y - __uint_as_float(hi.y));
              lo.z = __float_as_uint(v.z - __uint_as_float(hi.z)); lo.w = __float_as_uint(v.w - __uint_as_float(hi.w));
              *(uint4*)(base + i * st_step) = hi;
              *(uint4*)(base + P_TILE + i * st_step) = lo;
            }
          }
        }
        if (sub == UPK - 1) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&mi.full[st]);
          if (++st == NSTAGE) { st = 0; ph ^= 1; }
        }
      }
    }
  }
  if (want_db && c_ok) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (c + e < S.C) atomicAdd(J.db + c + e, bs[e]);
  }
}

__global__ void __launch_bounds__(NT, 1) tc_wgrad_kernel(const __grid_constant__ TcWParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const TcWJob& J = P.job[blockIdx.z];
  const int ntile = J.ptiles * J.qtiles;
  if ((int)blockIdx.y >= ntile) return;
  const int pt = blockIdx.y % J.ptiles, qt = blockIdx.y / J.ptiles;
  const int QW = J.QW;
  const int pc0 = pt * PW, qc0 = qt * QW;                     // first channel of this CTA's P / Q tile
  uint8_t* ring = smem;                                       // NSTAGE x 64 KB
  WMisc& mi = *(WMisc*)(smem + NSTAGE * 4 * P_TILE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned Mrows = (unsigned)P.BT * (unsigned)P.Lq;
  const unsigned nkb = (Mrows + KROWS - 1) / KROWS;
  const unsigned per = (nkb + P.nsplit - 1) / P.nsplit;
  const unsigned kb0 = min(nkb, blockIdx.x * per), kb1 = min(nkb, kb0 + per);

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) { mbar_init(&mi.full[s], 16); mbar_init(&mi.empty[s], 1); }
      mbar_init(&mi.done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&mi.tmem_base, 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mi.tmem_base;

  if (warp < 4) {
    reg_dec<REG_MMA>();
    if (tid == 0 && kb0 < kb1) {
      const uint32_t idesc = idesc_tf32(128, QW, 1, 1);
      // MN-major SW128_32B: LBO = 512 B between 32-channel blocks, SBO = distance between 4-row groups
      const uint64_t dP = ((smem_desc_sw128_32b(0, 512, (PW / 32) * 512) >> 16) << 16);
      const uint64_t dQ = ((smem_desc_sw128_32b(0, 512, (QW / 32) * 512) >> 16) << 16);
      const uint32_t rbase = smem_u32(ring) >> 4;
      int st = 0;
      uint32_t ph = 0;
      for (unsigned kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&mi.full[st], ph);
        tc_fence_after();
        const uint32_t p_hi = rbase + st * ((4 * P_TILE) >> 4), p_lo = p_hi + (P_TILE >> 4);
        const uint32_t q_hi = p_hi + ((2 * P_TILE) >> 4), q_lo = q_hi + (P_TILE >> 4);
#pragma unroll
        for (int g = 0; g < 4; ++g) {                         // 4 groups of 8 rows = 4 k-steps
          const uint32_t po = g * (PW / 32) * 64, qo = g * (QW / 32) * 64;     // (blocks * 1024 B) >> 4
          mma_tf32(tmem, dP | (p_lo + po), dQ | (q_hi + qo), idesc, (kb != kb0) || g != 0);
          mma_tf32(tmem, dP | (p_hi + po), dQ | (q_lo + qo), idesc, 1);
          mma_tf32(tmem, dP | (p_hi + po), dQ | (q_hi + qo), idesc, 1);
        }
        mma_commit(&mi.empty[st]);
        if (++st == NSTAGE) { st = 0; ph ^= 1; }
      }
      mma_commit(&mi.done);
    }
  } else {
    // warps 4-11 load the activation side, warps 12-19 the dz side (two tensors when a BN sits behind the layer)
    reg_inc<REG_LOAD>();
    const bool p_is_dz = J.db_on_p != 0;
    const bool big = warp >= 12;
    const bool do_p = (big == p_is_dz);
    const Side& S = do_p ? J.P : J.Q;
    const int W = do_p ? PW : QW, soff = do_p ? 0 : 2 * P_TILE;
    const bool wdb = J.db && (do_p ? (J.db_on_p && qt == 0) : (!J.db_on_p && pt == 0));
    const int c0 = do_p ? pc0 : qc0;
    if (big) side_loader<256, 16, 2, 3, true>(P, J, S, c0, W, tid - 384, lane, ring, soff, mi, kb0, kb1, wdb);
    else side_loader<256, 32, 4, 3, false>(P, J, S, c0, W, tid - 128, lane, ring, soff, mi, kb0, kb1, wdb);
    // ---- epilogue: warps 4-7 drain the accumulator and add it to dW ---------------------
    if (warp >= 4 && warp <= 7 && kb0 < kb1) {
      mbar_wait(&mi.done, 0);
      tc_fence_after();
      const int lgrp = warp & 3;
      const int p = pc0 + lgrp * 32 + lane;
      for (int cc = 0; cc * 32 < QW; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(lgrp * 32) << 16) + cc * 32, v);
        if (p < J.P.C) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = qc0 + cc * 32 + j;
            if (q < J.Q.C) atomicAdd(J.dW + J.wbase + (long)p * J.sp + (long)q * J.sq, __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

}  // namespace

// Converts the generic wgrad jobs into tensor-core jobs.  Returns 1 if some job is not eligible.
int launch_wgrad_tc(const WgradParams& p, cudaStream_t st) {
  TcWParams T{};
  T.njobs = 0; T.BT = p.BT; T.Lq = p.Lq;
  int maxt = 1;
  for (int j = 0; j < p.njobs; ++j) {
    const WgradJob& J = p.job[j];
    if (!J.a_src) return 1;                                          // bias-only jobs stay on the FFMA kernel
    if (J.C % 4 || J.N % 4 || J.a_ld % 4 || J.z_ld % 4 || J.a_coff % 4 || J.z_coff % 4) return 1;
    if ((double)p.BT * J.a_L * J.a_ld >= 4294967296.0 || (double)p.BT * J.z_L * J.z_ld >= 4294967296.0) return 1;
    if (J.db && !(J.z_mul == 1 && J.z_add == 0 && J.z_L == p.Lq)) return 1;   // db needs every dz row exactly once
    Side A{J.a_src, nullptr, J.a_p0, nullptr, J.a_p2, J.a_L, J.a_ld, J.a_coff, J.a_mul, J.a_add, J.C, J.a_relu, 0};
    Side Z{J.z_src, J.z_src2, J.z_p0, J.z_p1, J.z_p2, J.z_L, J.z_ld, J.z_coff, J.z_mul, J.z_add, J.N, 0, 0};
    TcWJob& O = T.job[T.njobs++];
    const bool a_on_p = J.C >= J.N;                                  // wider tensor on the 128-wide P side
    O.P = a_on_p ? A : Z; O.Q = a_on_p ? Z : A;
    O.dW = J.dW; O.wbase = J.wbase; O.sp = a_on_p ? J.wsc : J.wsn; O.sq = a_on_p ? J.wsn : J.wsc;
    O.db = J.db; O.db_on_p = a_on_p ? 0 : 1;
    O.QW = O.Q.C % 128 == 0 ? 128 : (O.Q.C % 64 == 0 ? 64 : 32);
    O.ptiles = (O.P.C + PW - 1) / PW; O.qtiles = (O.Q.C + O.QW - 1) / O.QW;
    maxt = std::max(maxt, O.ptiles * O.qtiles);
  }
  const long M = (long)p.BT * p.Lq;
  const long nkb = (M + KROWS - 1) / KROWS;
  T.nsplit = (int)std::max<long>(1, std::min<long>(nkb, sm_count() / (maxt * T.njobs)));
  const size_t smem = 1024 + (size_t)NSTAGE * 4 * P_TILE + sizeof(WMisc) + 64;
  static bool attr = false;
  if (!attr) {
    TRU_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid(T.nsplit, maxt, T.njobs);
  tc_wgrad_kernel<<<grid, NT, smem, st>>>(T);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace tru
