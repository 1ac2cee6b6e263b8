// Tensor-core implicit GEMM (tcgen05 / TMEM), fp32-accurate via the 3xTF32 split.
//
// Same contract as igemm.cu (gathered rows x resident weights, BN/ReLU applied on
// load, bias / BN statistics / ReLU mask / skip-gradient epilogue).
//
// The product is computed TRANSPOSED: the CTA's weight slice is the M operand of the
// MMA (M = 64 or 128 output channels, resident in shared memory as hi/lo tf32 tiles)
// and a tile of 128 gathered activation rows is the N operand, so the accumulator in
// TMEM holds  D[channel][row].  An epilogue thread therefore owns ONE output channel
// (TMEM lane) and walks over rows: bias, BN scale/shift of the ReLU mask and the BN
// statistics are per-thread scalars, a warp stores 32 consecutive channels of one row
// with one coalesced instruction, and no shared-memory transpose is needed.
//
// One persistent CTA per SM:
//   warp 0            MMA issuer (one elected thread) + TMEM allocation (warps 1-3 idle: register
//                     budgets are re-balanced per 4-warp group with setmaxnreg)
//   warps 4..3+LW     loaders: gather 128 rows x 32 channels from HBM (float4, NS
//                     k-blocks kept in flight in registers), apply the affine (+ReLU or
//                     BN-backward) transform, split into tf32 hi/lo and write both into
//                     the 128B-swizzled K-major operand tiles of a shared-memory ring
//   last 8 warps      epilogue: tcgen05.ld (32 lanes x 32 columns) -> registers -> HBM
// For every 32-wide k-block the MMA thread issues 4 k-steps x 3 products
// (lo*hi + hi*lo + hi*hi) of tcgen05.mma.kind::tf32 (M=64|128, N=128, K=8) into one of
// two TMEM accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Split: weights are rounded to tf32 (cvt.rna) once per CTA; activations use the cheap
// truncating split hi = x & ~0x1fff, lo = x - hi (exact), 2 instructions per element.
// The dropped lo*lo term is <= 2^-21 relative per product.
#include <algorithm>
#include <string.h>
#include <type_traits>
#include "net_kernels.cuh"
#include "tc_common.cuh"

namespace tru {
namespace {
using namespace tc;

constexpr int BM = 128, KBLK = 32;
constexpr int EW = 8;                                     // epilogue warps
constexpr int MAXKB = 16, MAXCOEF = 12, MAXWK = 16;
constexpr uint32_t ATILE_C = BM * 128, STAGE_C = 2 * ATILE_C;    // classic launches: 128 rows per plane
constexpr int COEF_FLOATS = 96;                           // p0[32] p2[32] p1[32]

// A tile of 128 gathered rows x 32 channels is staged ONCE per k-block (hi and lo tf32 planes, 128 B per row, SWIZZLE_128B) and
// multiplied by one or several resident weight k-blocks.  Several = the taps of a transposed conv: tap j reads the same stage
// through a descriptor whose start address is shifted by tap_shift[j] rows (the swizzle of a tcgen05 operand is a function of
// the shared-memory ADDRESS bits - measured, profiles/r02_probe_desc_shift.log - so a row-shifted start, or a plane that does
// not start on a 1024-byte boundary, reads correctly as long as the writer swizzles by address too).  The stage then holds
// 128 + max shift rows.
struct TcLayout {
  int MW;                       // MMA M = weight rows (output channels) per CTA: 64 or 128
  int nkb, nstage, ncoef, nwk;  // source k-blocks per tile, ring stages, coefficient rows, resident weight k-blocks
  int arows, shared;            // rows of a staged plane; tap-shared mode
  uint32_t atile, stage;        // bytes of one plane (arows * 128), bytes of a stage (hi + lo plane)
  unsigned lq_magic;            // ceil(2^32 / Lq) (0: Lq == 1) for the loaders' row decode
  int dbg;                      // ablation switches for bottleneck hunting (tru_debug_set_flags): 1 no MMA, 2 no global loads, 4 no smem stores, 8 no epilogue stores
  uint32_t a_off, coef_off, misc_off;                      // W tiles at offset 0
  int8_t kb_seg[MAXKB], kb_coef[MAXKB], kb_relu[MAXKB], kb_nw[MAXKB];   // kb_nw: weight k-blocks fed by this source k-block (consecutive in wk_*)
  int8_t kb_pf[MAXKB];          // rows of this k-block's source are contiguous in tile order (row m at (m + sadd) * ld): the loaders can
                                // prefetch the CTA's next tile into L2 with plain address arithmetic
  int8_t kb_lin[MAXKB];         // ... and sadd == 0, all 32 channels exist: problem row m IS source row m, no row decode / validity test
  int all_lin;                  // every k-block is linear: the per-tile row decode is skipped altogether
  int wtmem;                    // the weight slice (hi | lo, 2 x 32 nwk columns) lives in TENSOR MEMORY behind the two accumulators, not in shared memory
  int16_t kb_c0[MAXKB], kb_valid[MAXKB];
  int kb_lmax[MAXKB];           // source rows li >= kb_lmax are zero rows
  int8_t wk_seg[MAXWK], wk_shift[MAXWK];
  int16_t wk_c0[MAXWK], wk_valid[MAXWK];
  int wk_wbase[MAXWK];          // weight element of (channel c, output n) of weight k-block w: W[wk_wbase + (wk_c0 + c) * wsc + n * wsn]
  int8_t coef_seg[MAXCOEF];
  int16_t coef_c0[MAXCOEF];
};

struct Misc {
  uint64_t full[8], empty[16], tfull[2], tempty[2];    // empty[(turn & 1) * 8 + stage]: see the loader comment
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }


#ifndef TRU_CACHE_ROWS
#define TRU_CACHE_ROWS 1
#endif
#ifndef TRU_L2_PREFETCH
#define TRU_L2_PREFETCH 1
#endif
#ifndef TRU_REG_MMA12
#define TRU_REG_MMA12 40
#define TRU_REG_LOAD12 88
#define TRU_REG_EPI12 88
#endif
#ifndef TRU_W_IN_TMEM
#define TRU_W_IN_TMEM 1
#endif
#ifndef TRU_LINEAR_ROWS
#define TRU_LINEAR_ROWS 1
#endif
#ifndef TRU_EPI2_PIPE
#define TRU_EPI2_PIPE 1
#endif
#ifndef TRU_EPI2_REGS
#define TRU_EPI2_REGS 1
#endif
// loader warps LW (template parameter: 12, or 8 for the single-segment masked backward launches, measured ~5 % faster there);
// loader groups of 4 warps: group g owns k-blocks g, g + LW/4, ... (global k-block counter)

// LD2: loaders read two tensors (dY and Z) and apply the BN-backward affine.  EPI (0 plain, 1 mask, 2 mask +
// added tensor): the epilogue adds the skip gradient / applies the ReLU mask / accumulates the BN-backward sums.
//
// The kernel is bound by instruction issue, not by HBM or the tensor pipe (ablation: with MMAs, loads and
// stores all disabled the old skeleton still took the full HBM time), so both producer and consumer roles are
// written for instructions per element:
//   * 12 loader warps as 3 groups (16 warps at 72 registers spilled as soon as anything was added; 12 at 88 registers
//     hold the per-tile row decode and are slightly faster); a loader GROUP (4 warps) owns a whole k-block: 8 rows x one 16-byte channel chunk per thread, loads
//     issued back to back, then transformed and stored.  The per-k-block bookkeeping (descriptor fetch, row
//     decode, barrier handshake) is paid once per 8 float4 instead of once per 2, and memory-level
//     parallelism comes from the groups working on different k-blocks, not from register slots.
//   * the epilogue addresses rows with one 32-bit element offset (lane j computes row j's, broadcast by
//     shuffle) and IMAD.WIDE + STG; full tiles skip the per-row validity test.
// SH: tap-shared launch (compile-time, so that the classic launches keep their constant stage geometry, their per-thread
// constant swizzle and one weight k-block per staged k-block: the generalised loader cost them ~10 % when it was a run-time mode)
template <int LW, bool LD2, int EPI, bool SH>
__global__ void __launch_bounds__(32 * (4 + LW + EW), 1) tc_igemm_kernel(const __grid_constant__ IgemmParams P,
                                                         const __grid_constant__ TcLayout Lo) {
  // register budgets per role (launch value 72 for 28 warps): 4*24 + 16*64 + 8*112 = 2016 = 28*72
  //                                                      or 4*24 + 16*72 + 8*96 when the epilogue has no added tensor
  constexpr int NG = LW / 4, NT = 32 * (4 + LW + EW);
  // register budgets per role; launch value = 65536 / threads rounded down to a multiple of 8:
  //   LW = 12: 24 warps x 80 = 1920 >= 4*24 + 12*88 + 8*96      (EPI 2: 4*24 + 12*72 + 8*112 = 1856)
  //   LW =  8: 20 warps x 96 = 1920 >= 4*24 +  8*104 + 8*120
  constexpr int REG_LAUNCH = LW == 8 ? 96 : 80;
  // (the issuing thread is the critical role of the 12-loader-warp launches - ncu's source view has it busy 3/4 of the time - and at
  // 24 registers its loop spilled in the tap-shared variants (local-memory reloads in front of the MMAs) and any change of the
  // prologue pushed the forward variants over the edge (15-40 % slower, twice).  The 12-warp variants give the MMA warps 40 registers
  // and take them from the epilogue (96 -> 88): 4 x 40 + 12 x 88 + 8 x 88 = 1920 = 24 x 80)
  constexpr int REG_MMA = LW == 8 ? 24 : ((EPI == 2 && TRU_EPI2_REGS) ? 40 : TRU_REG_MMA12);
  constexpr int REG_LOAD = LW == 8 ? 104 : ((EPI == 2 && TRU_EPI2_REGS) ? 72 : TRU_REG_LOAD12);
  constexpr int REG_EPI = LW == 8 ? 120 : ((EPI == 2 && TRU_EPI2_REGS) ? 112 : TRU_REG_EPI12);
  extern __shared__ __align__(1024) uint8_t smem[];    // (1024-byte aligned: the weight tiles and the classic stages are whole swizzle atoms)
  uint8_t* Wsm = smem;                                 // [hi|lo][nwk][MW rows][128 B], swizzled
  uint8_t* Asm = smem + Lo.a_off;                      // ring: [stage][hi|lo][arows][128 B]
  float* coef = (float*)(smem + Lo.coef_off);          // [ncoef][p0 | p2 | p1][32]
  Misc& mi = *(Misc*)(smem + Lo.misc_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int MW = Lo.MW, nkb = Lo.nkb, nstage = Lo.nstage;
  const int n0 = blockIdx.y * MW;
  const long M = (long)P.BT * P.Lq;
  const int ntiles = (int)((M + BM - 1) / BM);
  const int n_my = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  if (smem_u32(smem) & 1023u) { if (tid == 0) printf("tc_igemm_kernel: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }

  // ---- one-time setup: resident weight slice (hi/lo), affine coefficient table, barriers, TMEM ------
  pdl_trigger();       // the weights are parameters: nothing in the step writes them, so staging them may overlap the predecessor's tail
  if (!Lo.wtmem) {
    const int msh = MW == 64 ? 11 : 12;                 // log2(32 channels x MW rows)
    const uint32_t nel = (uint32_t)Lo.nwk << msh;       // weight elements (one hi and one lo word each)
    for (uint32_t i = tid; i < nel / 2; i += NT) ((uint4*)Wsm)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const uint32_t lo_off = nel * 4u;
    constexpr int WU = 8;          // weight loads in flight per thread (the prologue is on the critical path of every launch)
    for (uint32_t i0 = tid; i0 < nel; i0 += NT * WU) {
      float v[WU];
      uint32_t oo[WU];
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        const uint32_t i = i0 + u * NT;
        v[u] = 0.f; oo[u] = 0xffffffffu;
        if (i < nel) {
          const int w = (int)(i >> msh), rem = (int)(i & ((1u << msh) - 1u));
          const Seg& sg = P.seg[Lo.wk_seg[w]];
          int c, n;
          if (sg.wsc == 1) { c = rem & 31; n = rem >> 5; } else { n = rem & (MW - 1); c = rem >> (msh - 5); }
          if (c < Lo.wk_valid[w] && n0 + n < P.N) {
            v[u] = __ldg(sg.W + Lo.wk_wbase[w] + (long)(Lo.wk_c0[w] + c) * sg.wsc + (long)(n0 + n) * sg.wsn);
            oo[u] = (uint32_t)w * MW * 128 + n * 128 + ((((c >> 2) ^ (n & 7)) << 4) | ((c & 3) << 2));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < WU; ++u) {
        if (oo[u] != 0xffffffffu) {
          const uint32_t hi = f2tf32(v[u]);
          *(uint32_t*)(Wsm + oo[u]) = hi;
          *(float*)(Wsm + lo_off + oo[u]) = v[u] - __uint_as_float(hi);
        }
      }
    }
  }
  auto load_coef = [&]() {
    pdl_wait();        // everything below reads what earlier kernels of the step produced (BN coefficients first)
    for (int i = tid; i < Lo.ncoef * 32; i += NT) {
      const int e = i >> 5, j = i & 31;
      const Seg& sg = P.seg[Lo.coef_seg[e]];
      const int c = Lo.coef_c0[e] + j;
      const bool ok = c < sg.C;
      const int cc = sg.coff + (sg.cmod ? c % sg.cmod : c);
      coef[e * COEF_FLOATS + j] = ok ? __ldg(sg.p0 + cc) : 1.f;
      coef[e * COEF_FLOATS + 32 + j] = ok ? __ldg(sg.p2 + cc) : 0.f;
      coef[e * COEF_FLOATS + 64 + j] = (ok && sg.p1) ? __ldg(sg.p1 + cc) : 0.f;
    }
  };
  if (!Lo.wtmem) load_coef();
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < nstage; ++s) { mbar_init(&mi.full[s], LW / NG); mbar_init(&mi.empty[s], 1); mbar_init(&mi.empty[8 + s], 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(&mi.tfull[a], 1); mbar_init(&mi.tempty[a], 32 * EW); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&mi.tmem_base, Lo.wtmem ? 512 : 2 * BM);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mi.tmem_base;
  if (Lo.wtmem) {
    // Weights as the MMA's A operand IN TENSOR MEMORY (M = 128 launches with <= 128 reduction channels): row n of the slice in
    // lane n, its channels in consecutive columns - hi plane at column 256, lo plane behind it.  The MMAs then read only the
    // gathered rows from shared memory (per 128 x 128 x 128 tile the operand reads drop from 384 KB to 192 KB of a 512 KB total:
    // the pipe was ~55 % of the shared-memory bandwidth), and the 128 KB the slice occupied become ring stages.
    // A warp writes the 32 lanes of its quadrant (warp & 3); the 6 warps of a quadrant share the 16-column groups.
    // (M = 64: row n of the slice lives in lane 32 (n / 16) + n % 16, the accumulator's own lane mapping; lanes 16-31 of a quadrant hold zeros)
    const int q = warp & 3, n = MW == 64 ? q * 16 + (lane & 15) : q * 32 + lane, ngrp = Lo.nwk * 4;      // groups: [hi | lo][nwk][2 x 16 channels]
    const bool nok = n0 + n < P.N && (MW == 128 || lane < 16);
    for (int gi = warp >> 2; gi < ngrp; gi += NT / 128) {
      const int plane = gi >= 2 * Lo.nwk, gg = gi - plane * 2 * Lo.nwk, w = gg >> 1, c0 = (gg & 1) * 16;
      const Seg& sg = P.seg[Lo.wk_seg[w]];
      const float* wp = sg.W + Lo.wk_wbase[w] + (long)(Lo.wk_c0[w] + c0) * sg.wsc + (long)(n0 + n) * sg.wsn;
      uint32_t v[16];
      float x[16];
      if (sg.wsc == 1 && nok && c0 + 16 <= Lo.wk_valid[w] && (((size_t)wp) & 15) == 0) {      // the thread's 16 channels are contiguous: 4 x 16 bytes
#pragma unroll
        for (int c = 0; c < 4; ++c) { const float4 t = __ldg((const float4*)wp + c); x[4 * c] = t.x; x[4 * c + 1] = t.y; x[4 * c + 2] = t.z; x[4 * c + 3] = t.w; }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = (nok && c0 + c < Lo.wk_valid[w]) ? __ldg(wp + (long)c * sg.wsc) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const uint32_t hi = f2tf32(x[c]);
        v[c] = plane ? __float_as_uint(x[c] - __uint_as_float(hi)) : hi;
      }
      tmem_st16(tmem + ((uint32_t)(q * 32) << 16) + 2 * BM + (uint32_t)plane * Lo.nwk * 32 + (uint32_t)w * 32 + c0, v);
    }
    tmem_st_wait();
    load_coef();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (warp < 4) {
    // ================================ MMA issuer ==================================
    reg_dec<REG_MMA>();
    if (tid == 0) {
      const uint32_t idesc = idesc_tf32(MW, BM, 0, 0);
      // descriptor = constant high word | 14-bit (address >> 4); K-steps advance the address by 32 B
      const uint64_t dhi = (uint64_t)(smem_desc_sw128(0, 16, 1024) >> 32) << 32 | (1ull << 16);
      const uint32_t a_base = smem_u32(Asm) >> 4, w_base = smem_u32(Wsm) >> 4;
      const uint32_t w_lo_off = ((uint32_t)Lo.nwk * MW * 128) >> 4, w_kb = ((uint32_t)MW * 128) >> 4;
      const uint32_t stage16 = (SH ? Lo.stage : STAGE_C) >> 4, atile16 = (SH ? Lo.atile : ATILE_C) >> 4;
      int st = 0;
      uint32_t ph = 0;
      for (int ti = 0; ti < n_my; ++ti) {
        const int acc = ti & 1;
        mbar_wait(&mi.tempty[acc], ((ti >> 1) & 1) ^ 1, 100 + ti);
        tc_fence_after();
        const uint32_t d = tmem + acc * BM;
        int w = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&mi.full[st], ph & 1, 200 + st * 10 + kb);
          tc_fence_after();
          const uint32_t x_st = a_base + st * stage16;
          for (int u = SH ? Lo.kb_nw[kb] : 1; u > 0; --u, ++w) {
            // tap w reads the stage from row wk_shift[w] on: start address + shift * 128 B (8 units of 16 B)
            const uint32_t x_hi = x_st + (SH ? (uint32_t)Lo.wk_shift[w] * 8u : 0u), x_lo = x_hi + atile16;
            const uint32_t w_hi = w_base + (uint32_t)w * w_kb, w_lo = w_hi + w_lo_off;
            if (Lo.wtmem) {
              const uint32_t t_hi = tmem + 2 * BM + (uint32_t)w * 32, t_lo = t_hi + (uint32_t)Lo.nwk * 32;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t dxh = dhi | (x_hi + 2 * j), dxl = dhi | (x_lo + 2 * j);
                if (Lo.dbg & 1) continue;
                mma_tf32_ts(d, t_lo + 8 * j, dxh, idesc, (w | j) != 0);
                mma_tf32_ts(d, t_hi + 8 * j, dxl, idesc, 1);
                mma_tf32_ts(d, t_hi + 8 * j, dxh, idesc, 1);
              }
              continue;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint64_t dxh = dhi | (x_hi + 2 * j), dxl = dhi | (x_lo + 2 * j);
              const uint64_t dwh = dhi | (w_hi + 2 * j), dwl = dhi | (w_lo + 2 * j);
              if (Lo.dbg & 1) continue;
              mma_tf32(d, dwl, dxh, idesc, (w | j) != 0);
              mma_tf32(d, dwh, dxl, idesc, 1);
              mma_tf32(d, dwh, dxh, idesc, 1);
            }
          }
          mma_commit(&mi.empty[(ph & 1) * 8 + st]);
          if (++st == nstage) { st = 0; ++ph; }
        }
        mma_commit(&mi.tfull[acc]);
      }
    }
  } else if (warp < 4 + LW) {
    // ================================== loaders ====================================
    if (REG_LOAD > REG_LAUNCH) reg_inc<REG_LOAD>();
    if (REG_LOAD < REG_LAUNCH) reg_dec<REG_LOAD>();
    constexpr int R = LD2 ? 4 : 8;                      // rows per thread and pass (LD2: two passes of 4 rows, two tensors)
    const int lt = tid - 128, g = lt >> 7, gt = lt & 127, chunk = gt & 7, rbase = gt >> 3;   // rows rbase + 16 i
    const unsigned Lq = (unsigned)P.Lq, magic = Lo.lq_magic, Mu = (unsigned)M;
    constexpr bool shared = SH;
    const int xrows = SH ? Lo.arows - BM : 0;           // tap-shared launches stage a few extra rows (the largest tap shift)
    const uint32_t a_addr = smem_u32(Asm), stage_b = SH ? Lo.stage : STAGE_C, atile_b = SH ? Lo.atile : ATILE_C;
    const uint32_t sw_c = (uint32_t)(chunk ^ (rbase & 7)) << 4;      // classic stages are whole 1024-byte atoms: constant swizzle
    int ti = 0, kb = g;
    while (kb >= nkb) { kb -= nkb; ++ti; }
    // Ring bookkeeping.  A group advances NG k-blocks at a time, which can be more than one turn of the ring,
    // so with one parity barrier per stage it could be two completions behind and mis-read the parity.  The
    // "slot free" barriers therefore alternate between two mbarriers per stage (even / odd turns): each one
    // completes every other turn and a waiter is never more than one completion behind.
    int st = g;
    uint32_t ph = 0;                 // turn of the ring this group's k-block belongs to
    while (st >= nstage) { st -= nstage; ++ph; }
    int cur = -1;
    unsigned bt0 = 0, qb = 0, qmax = 0;
    int v0 = 0;                      // tap-shared: virtual row of the thread's first stage row
    constexpr bool CACHE_ROWS = TRU_CACHE_ROWS != 0;
    unsigned rbt[CACHE_ROWS ? 8 : 1], rqq[CACHE_ROWS ? 8 : 1];
    const float ninf = -__int_as_float(0x7f800000);
    constexpr unsigned NOROW = 0x3fffffffu;              // row decode of a virtual row outside [0, M): fails every "li < rows" test
    // (frame, position) of virtual / tile row vv; tap-shared launches stage rows in front of / behind the problem as zero rows,
    // the classic path clamps rows beyond M (last tile only) to row M-1: they load valid memory and the epilogue drops them
    auto decode = [&](int ii, unsigned& bt, unsigned& q) {
      if (shared) {
        const int vv = v0 + 16 * ii;
        const bool in = (unsigned)vv < Mu;
        const unsigned uv = in ? (unsigned)vv : 0u;
        const unsigned bq = magic ? __umulhi(uv, magic) : uv;       // uv / Lq (exact: uv * Lq < 2^32)
        bt = bq; q = in ? uv - bq * Lq : NOROW;
      } else {
        const unsigned qq = min(qb + 16u * ii, qmax);
        const unsigned bq = magic ? __umulhi(qq, magic) : qq;        // qq / Lq (exact: qq * Lq < 2^32)
        bt = bt0 + bq; q = qq - bq * Lq;
      }
    };
    while (ti < n_my) {
      if (ti != cur) {               // new tile: decode its first row once; the thread's rows follow by a small exact division
        const unsigned m0 = ((unsigned)blockIdx.x + (unsigned)ti * gridDim.x) * BM;
        bt0 = m0 / Lq; qb = m0 - bt0 * Lq + rbase; qmax = Mu - 1u - bt0 * Lq;
        v0 = (int)m0 + P.row_base + rbase;
        cur = ti;
        if (CACHE_ROWS && (SH || !Lo.all_lin)) {            // the (frame, position) of the thread's 8 rows is the same for every k-block of the tile
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) decode(ii, rbt[ii], rqq[ii]);
        }
      }
      const Seg& sg = P.seg[Lo.kb_seg[kb]];
      const float* src = sg.src;
      const float* src2 = LD2 ? sg.src2 : nullptr;
      const unsigned ld = (unsigned)sg.ld, fs = sg.fs ? (unsigned)sg.fs : (unsigned)sg.Lsrc * ld;
      const int smul = sg.smul, sadd = sg.sadd;
      const unsigned cb = (unsigned)(sg.coff + Lo.kb_c0[kb] + chunk * 4);
      const bool cok = chunk * 4 < Lo.kb_valid[kb] && !(Lo.dbg & 2);
      const unsigned Lok = cok ? (unsigned)Lo.kb_lmax[kb] : 0u;        // channel chunk beyond the segment: every row is a zero row
      const int e = Lo.kb_coef[kb];
      const float fl = Lo.kb_relu[kb] ? 0.f : ninf;
      float4 p0 = make_float4(1.f, 1.f, 1.f, 1.f), p1 = make_float4(0.f, 0.f, 0.f, 0.f), p2 = p1;
      if (e >= 0) {
        const float* ce = coef + e * COEF_FLOATS + chunk * 4;
        p0 = *(const float4*)ce; p2 = *(const float4*)(ce + 32);
        if (LD2) p1 = *(const float4*)(ce + 64);
      }
      // Swizzle by ADDRESS (bits 4-6 ^= bits 7-9): a plane of a tap-shared stage need not start on a 1024-byte boundary.
      // Rows rbase + 16 i of a plane share their phase, so the two chunk offsets are per-k-block constants.
      const uint32_t s_addr = a_addr + (uint32_t)st * stage_b;
      const uint32_t sw_hi = SH ? (uint32_t)(chunk ^ (((s_addr >> 7) + rbase) & 7)) << 4 : sw_c;
      const uint32_t sw_lo = SH ? (uint32_t)(chunk ^ ((((s_addr + atile_b) >> 7) + rbase) & 7)) << 4 : sw_c;
      uint8_t* ah = Asm + (size_t)st * stage_b + (uint32_t)rbase * 128;
      // transform (affine + ReLU, or BN-backward affine), truncating tf32 split, store row `row` of the stage
      auto put = [&](float4 v, const float4& bz, bool keep, int row) {
        if (e >= 0 && !(Lo.dbg & 32)) {
          if (LD2) {
            v.x = fmaf(p1.x, bz.x, fmaf(p0.x, v.x, p2.x)); v.y = fmaf(p1.y, bz.y, fmaf(p0.y, v.y, p2.y));
            v.z = fmaf(p1.z, bz.z, fmaf(p0.z, v.z, p2.z)); v.w = fmaf(p1.w, bz.w, fmaf(p0.w, v.w, p2.w));
          } else {
            v.x = fmaf(p0.x, v.x, p2.x); v.y = fmaf(p0.y, v.y, p2.y); v.z = fmaf(p0.z, v.z, p2.z); v.w = fmaf(p0.w, v.w, p2.w);
            v.x = fmaxf(v.x, fl); v.y = fmaxf(v.y, fl); v.z = fmaxf(v.z, fl); v.w = fmaxf(v.w, fl);
          }
        }
        if (!keep) v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint4 hi, lo;
        hi.x = __float_as_uint(v.x) & 0xffffe000u; hi.y = __float_as_uint(v.y) & 0xffffe000u;
        hi.z = __float_as_uint(v.z) & 0xffffe000u; hi.w = __float_as_uint(v.w) & 0xffffe000u;
        lo.x = __float_as_uint(v.x - __uint_as_float(hi.x)); lo.y = __float_as_uint(v.y - __uint_as_float(hi.y));
        lo.z = __float_as_uint(v.z - __uint_as_float(hi.z)); lo.w = __float_as_uint(v.w - __uint_as_float(hi.w));
        if (Lo.dbg & 4) return;
        *(uint4*)(ah + row * (16 * 128) + sw_hi) = hi;
        *(uint4*)(ah + atile_b + row * (16 * 128) + sw_lo) = lo;
      };
      if (!SH && Lo.kb_lin[kb]) {
        // LINEAR k-block (pointwise convs and their data gradients: the bulk of the step): problem row m is source row m, every
        // row and channel exists - no row decode, no validity mask (rows beyond M, last tile only, re-read row M-1; the epilogue
        // drops them).  The profile had ~480 instructions per warp and k-block in this loop for 150 of loads, math and stores.
        const unsigned mr = ((unsigned)blockIdx.x + (unsigned)ti * gridDim.x) * BM + rbase;
#pragma unroll
        for (int h = 0; h < 8 / R; ++h) {
          float4 a[R], b[LD2 ? R : 1];
#pragma unroll
          for (int i = 0; i < R; ++i) {
            const unsigned off = min(mr + 16u * (h * R + i), Mu - 1u) * ld + cb;
            a[i] = ldg4_off(src, off);
            if (LD2) b[i] = src2 ? ldg4_off(src2, off) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (h == 0 && ph > 0) mbar_wait(&mi.empty[((ph - 1) & 1) * 8 + st], ((ph - 1) >> 1) & 1, 300 + st * 10 + kb + 1000 * ti);
#pragma unroll
          for (int i = 0; i < R; ++i) put(a[i], b[LD2 ? i : 0], true, h * R + i);
        }
      } else {
#pragma unroll
      for (int h = 0; h < 8 / R; ++h) {
        float4 a[R], b[LD2 ? R : 1];
        unsigned msk = 0;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int ii = h * R + i;
          unsigned bt, q;
          if (CACHE_ROWS) { bt = rbt[ii]; q = rqq[ii]; } else decode(ii, bt, q);
          const unsigned li = (unsigned)((int)q * smul + sadd);
          const bool ok = li < Lok;
          unsigned off = bt * fs + li * ld + cb;
          off = ok ? off : 0u;               // padding rows read element 0 (always mapped) and are zeroed below
          if (ok) msk |= 1u << i;
          a[i] = ldg4_off(src, off);
          if (LD2) b[i] = src2 ? ldg4_off(src2, off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (h == 0 && ph > 0) mbar_wait(&mi.empty[((ph - 1) & 1) * 8 + st], ((ph - 1) >> 1) & 1, 300 + st * 10 + kb + 1000 * ti);
        const bool all_ok = __all_sync(0xffffffffu, msk == (1u << R) - 1u);
        if (all_ok) {
#pragma unroll
          for (int i = 0; i < R; ++i) put(a[i], b[LD2 ? i : 0], true, h * R + i);
        } else {
#pragma unroll
          for (int i = 0; i < R; ++i) put(a[i], b[LD2 ? i : 0], (msk >> i) & 1u, h * R + i);
        }
      }
      }
      // L2 prefetch of the same k-block of the CTA's NEXT tile (one 128-byte line per row: the chunk-0 threads issue it).  The
      // loaders keep 8 (LD2: 2 x 4) 16-byte loads per thread in flight - ~32-48 KB per SM, borderline for 22 B/clk x ~1,000 cycles
      // of loaded DRAM latency - and more would cost registers; a prefetch costs none and turns the next tile's loads into L2 hits.
      // (measured: the same prefetch for gathered rows - skip alignment, tap-shared virtual rows; with a row decode per line or as
      // a linear sweep from the tile's first row - gained nothing, and for the epilogue's mask / added-tensor rows of 128-channel
      // layers it cost more than it hid: 0.95 -> 1.06 ms on the masked 128 -> 128 data gradient; the epilogue prefetch pays for
      // <= 64 output channels only, see below)
      if (!SH && TRU_L2_PREFETCH && Lo.kb_pf[kb] && ti + 1 < n_my) {
        // (one line per LANE: the warp's 32 rows are (lane >> 3) + 4 (warp of the group) + 16 ii; lane j takes ii = j & 7)
        const unsigned m1 = ((unsigned)blockIdx.x + (unsigned)(ti + 1) * gridDim.x) * BM + (unsigned)(rbase + 16 * chunk);
        const unsigned mm = min(m1, Mu - 1u);
        const unsigned off = (unsigned)((int)mm + sadd) * ld + (unsigned)(sg.coff + Lo.kb_c0[kb]);
        prefetch_l2_off(src, off);
        if (LD2 && src2) prefetch_l2_off(src2, off);
      }
      if (shared && rbase < xrows) {         // stage rows 128 + rbase (rbase < largest tap shift): one more row for the first threads
        unsigned bt, q;
        decode(8, bt, q);
        const unsigned li = (unsigned)((int)q * smul + sadd);
        const bool ok = li < Lok;
        const unsigned off = ok ? bt * fs + li * ld + cb : 0u;
        const float4 a9 = ldg4_off(src, off);
        const float4 b9 = (LD2 && src2) ? ldg4_off(src2, off) : make_float4(0.f, 0.f, 0.f, 0.f);
        put(a9, b9, ok, 8);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mi.full[st]);
      kb += NG;
      while (kb >= nkb) { kb -= nkb; ++ti; }
      st += NG;
      while (st >= nstage) { st -= nstage; ++ph; }
    }
  } else if constexpr (EPI == 3) {
    // ====================== epilogue with the depthwise conv (inference) ======================
    // Encoder block network.py:28-40 in eval mode: pointwise conv -> BN -> ReLU -> depthwise conv.  The accumulator is
    // D[channel][row] and a thread owns one channel, so the rows of a frame are CONSECUTIVE REGISTERS of the thread and the
    // 3 / 5-tap depthwise filter along the frame is register-local: a = relu(mp0 * (acc + bias) + mp2) (the folded BatchNorm,
    // the same two roundings as the consumers' load-time affine), out[lo] = dw_b + sum_j w[j] a[lo*S - K/2 + j], zero padding.
    // Frames (Lq rows, Lq | 128, 16 | Lq) never straddle a tile; a warp walks its 64 tile rows in chunks of 16 carrying the
    // last 4 activations (e[0..3]); a chunk emits the outputs whose window ENDS in it, the frame's last chunk also the output
    // whose window ends in the padding.  The warp of the second half tile starts mid-frame when Lq = 128: it reads its 4-row
    // halo from TMEM.  The pointwise output never reaches HBM (halves the encoder traffic of an inference pass).
    reg_inc<REG_EPI>();
    const int ew = warp - 4 - LW, lg = warp & 3, half = ew >> 2;
    const int n = n0 + lg * 32 + lane;
    const bool nok = n < P.N;
    const float bias = (P.bias && nok) ? __ldg(P.bias + n) : 0.f;
    const float a0 = (P.mp0 && nok) ? __ldg(P.mp0 + n) : 1.f, a2 = (P.mp2 && nok) ? __ldg(P.mp2 + n) : 0.f;
    const float dwb = (P.dw_b && nok) ? __ldg(P.dw_b + n) : 0.f;
    float* outb = P.out + (size_t)n + P.ocoff;
    const unsigned Lq = (unsigned)P.Lq, Lout = (unsigned)P.Lout, BT = (unsigned)P.BT, ldo = (unsigned)P.ldo;
    auto act = [&](uint32_t v) { return fmaxf(fmaf(__uint_as_float(v) + bias, a0, a2), 0.f); };
    auto run = [&](auto KK, auto SS) {
      constexpr int K = decltype(KK)::value, S = decltype(SS)::value, PD = K / 2;
      constexpr int D = -((PD) / S);                    // first output of a chunk at q: lo = q/S + D  (ceil(-PD / S))
      constexpr int I0 = D * S - PD + 4;                // its window starts at e[I0]
      constexpr int NO = 16 / S;
      static_assert(I0 >= 0 && I0 + (NO - 1) * S + K - 1 <= 19, "window outside the 4 + 16 register buffer");
      float w[K];
#pragma unroll
      for (int j = 0; j < K; ++j) w[j] = nok ? __ldg(P.dw_w + (size_t)n * K + j) : 0.f;
      for (int ti = 0; ti < n_my; ++ti) {
        const int acc = ti & 1;
        const unsigned m0 = ((unsigned)blockIdx.x + (unsigned)ti * gridDim.x) * BM + half * 64;
        unsigned bt = m0 / Lq, q = m0 - bt * Lq;
        mbar_wait(&mi.tfull[acc], (ti >> 1) & 1, 400 + ti);
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + acc * BM + half * 64;
        float e[20];
        uint32_t va[16], vb[16];
        tmem_ld16_issue(taddr, va);
        if (q != 0) {
          uint32_t h[4];
          tmem_ld4(taddr - 4, h);
#pragma unroll
          for (int j = 0; j < 4; ++j) e[j] = act(h[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) e[j] = 0.f;
        }
        auto chunk = [&](uint32_t (&v)[16]) {
#pragma unroll
          for (int j = 0; j < 16; ++j) e[4 + j] = act(v[j]);
          const bool ok = nok && bt < BT;               // (rows beyond M: the last tile of the problem only)
          const int lo0 = (int)(q / S) + D;
          const unsigned rb = bt * Lout;
#pragma unroll
          for (int i = 0; i < NO; ++i) {
            float o = dwb;
#pragma unroll
            for (int j = 0; j < K; ++j) o = fmaf(w[j], e[I0 + i * S + j], o);
            const int lo = lo0 + i;
            if (ok && lo >= 0) outb[(size_t)((rb + (unsigned)lo) * ldo)] = o;
          }
          q += 16;
          if (q == Lq) {                                // frame complete: the output whose window ends in the padding, then a new frame
            if (D < 0) {
              float o = dwb;
#pragma unroll
              for (int j = 0; j < K; ++j)
                if (I0 + j < 4) o = fmaf(w[j], e[16 + I0 + j], o);
              if (ok) outb[(size_t)((bt * Lout + Lout - 1u) * ldo)] = o;
            }
            q = 0; ++bt;
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = e[16 + j];
          }
        };
        tmem_ld_wait(va);
        tmem_ld16_issue(taddr + 16, vb);
        chunk(va);
        tmem_ld_wait(vb);
        tmem_ld16_issue(taddr + 32, va);
        chunk(vb);
        tmem_ld_wait(va);
        tmem_ld16_issue(taddr + 48, vb);
        chunk(va);
        tmem_ld_wait(vb);
        tc_fence_before();
        mbar_arrive(&mi.tempty[acc]);
        chunk(vb);
      }
    };
    if (P.dw_k == 5) run(std::integral_constant<int, 5>{}, std::integral_constant<int, 2>{});
    else if (P.dw_s == 1) run(std::integral_constant<int, 3>{}, std::integral_constant<int, 1>{});
    else run(std::integral_constant<int, 3>{}, std::integral_constant<int, 2>{});
  } else {
    // ================================== epilogue ===================================
    // Thread = one output channel (TMEM lane); per tile a warp drains 2 chunks of 32 tile rows
    // (TMEM columns), each as two sub-chunks of 16 rows.  M = 64: the accumulator occupies lanes
    // 0-15 of every 32-lane quadrant (rows 16q..16q+15); a 16x32bx2 load hands lanes 16-31 the
    // second 16 columns of a chunk, so a chunk is one sub-step and no lane idles.  The ReLU-mask values (Z of the layer that
    // receives the gradient) do not depend on the accumulator, so they are fetched one sub-chunk
    // ahead - across tile boundaries too - and their latency hides behind the previous sub-chunk.
    reg_inc<REG_EPI>();
    const int ew = warp - 4 - LW, lg = warp & 3, half = ew >> 2;
    const bool m64 = MW == 64;
    const int nl = m64 ? lg * 16 + (lane & 15) : lg * 32 + lane;
    const int n = n0 + nl;
    const bool nok = n < P.N && !(Lo.dbg & 8);
    const int src64 = lane & 16;                        // M = 64: lanes 16-31 hold the second 16 rows of a chunk
    const unsigned Mu = (unsigned)M, Lq = (unsigned)P.Lq;      // (P.Lvalid: the launcher sets it to Lq for classic launches)
    const float bias = (P.bias && nok) ? __ldg(P.bias + n) : 0.f;
    float mp0 = 1.f, mp2 = 0.f;
    const bool use_mask = EPI && P.use_mask, has_extra = EPI == 2 && P.extra != nullptr;
    if (use_mask && P.mp0 && nok) { mp0 = __ldg(P.mp0 + n); mp2 = __ldg(P.mp2 + n); }
    const bool do_stats = P.stats != nullptr, do_bstats = EPI && P.bstats != nullptr;
    const float bmean = (do_bstats && nok) ? __ldg(P.bmean + n) : 0.f;
    float s1 = 0.f, s2 = 0.f;      // stats: sum o, sum o*o;  bstats: sum g, sum g*(z - mean) (scaled at the end)
    // per-thread base pointers; per-row element offsets are computed by lane j for row j of a chunk
    const size_t obase = P.planar ? (size_t)n * P.Lout : (size_t)n;
    const float* zbase = use_mask ? P.zmask + obase : nullptr;
    const float* xbase = has_extra ? P.extra + n - P.ocoff : nullptr;     // ext_ld == ldo (checked by the planner): same row offsets as the output
    float* outb = P.out + obase;
    asm("" : "+l"(outb));            // opaque: keep the finished pointer in registers (no re-association with P.out)
    asm("" : "+l"(zbase));
    asm("" : "+l"(xbase));

    // lane j: element offsets of tile row cc*32 + j of tile ti (0xffffffff: row does not exist)
    auto row_offsets = [&](int ti, int cc, unsigned& ooff, unsigned& eoff) {
      ooff = 0xffffffffu; eoff = 0u;
      if (ti < n_my) {
        const unsigned m = ((unsigned)blockIdx.x + (unsigned)ti * gridDim.x) * BM + cc * 32 + lane;
        if (m < Mu) {
          const unsigned bt = m / Lq, q = m - bt * Lq;
          const unsigned lo = q * P.omul + P.oadd, r = bt * P.Lout + lo;
          ooff = P.planar ? bt * (unsigned)P.N * (unsigned)P.Lout + lo : r * (unsigned)P.ldo + P.ocoff;
          eoff = r * (unsigned)P.ext_ld;
          if (SH && q >= (unsigned)P.Lvalid) ooff = 0xffffffffu;   // (tap-shared launches: virtual rows that only exist to be read by shifted taps)
        }
      }
    };
    auto zfetch = [&](float (&z)[16], float (&x)[EPI == 2 ? 16 : 1], unsigned ooff, int src0) {
      if (EPI && (use_mask || has_extra)) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const unsigned off = __shfl_sync(0xffffffffu, ooff, src0 + r);
          const bool ok = nok && off != 0xffffffffu;
          z[r] = (ok && use_mask) ? ldg_off(zbase, off) : 0.f;
          if (EPI == 2) x[EPI == 2 ? r : 0] = (ok && has_extra) ? ldg_off(xbase, off) : 0.f;
        }
      }
    };
    // the added tensor is folded into the accumulator registers as soon as they arrive, which frees its buffer
    // for the next prefetch (one x buffer instead of two: the epilogue stays inside its register budget)
    auto addx = [&](uint32_t (&v)[16], const float (&x)[EPI == 2 ? 16 : 1]) {
      if (EPI == 2) {
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = __float_as_uint(__uint_as_float(v[r]) + x[EPI == 2 ? r : 0]);
      }
    };
    // CHK: the tile may contain rows beyond M (only the last tile of the problem).  ST: 0 no sums, 1 forward BN
    // statistics, 2 BN-backward sums.  Both are compile-time so the per-element code is shuffle, address, add,
    // store (+2-3 for the sums) with the loop-invariant channel predicate on the store.
    auto process = [&](auto CHK, auto STM, const uint32_t (&v)[16], int src0, const float (&z)[16], unsigned ooff) {
      constexpr bool chk = decltype(CHK)::value;
      constexpr int stm = decltype(STM)::value;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const unsigned off = __shfl_sync(0xffffffffu, ooff, src0 + r);
        const bool ok = nok && (!chk || off != 0xffffffffu);
        float o = __uint_as_float(v[r]) + bias;
        if (EPI) {
          if (use_mask) o = (fmaf(z[r], mp0, mp2) > 0.f) ? o : 0.f;
        }
        if (chk) {
          if (ok) {
            outb[off] = o;
            if (stm == 1) { s1 += o; s2 = fmaf(o, o, s2); }
            if (stm == 2) { s1 += o; s2 = fmaf(o, z[r] - bmean, s2); }
          }
        } else {       // full tile: only the store carries the (loop-invariant) channel predicate; idle lanes' sums are never used
          if (nok) outb[off] = o;
          if (stm == 1) { s1 += o; s2 = fmaf(o, o, s2); }
          if (stm == 2) { s1 += o; s2 = fmaf(o, z[r] - bmean, s2); }
        }
      }
    };
    const int stmode = do_stats ? 1 : ((EPI && do_bstats) ? 2 : 0);
    float za[16], zb[16], xx[EPI == 2 ? 16 : 1];
    unsigned oa, ea, ob, eb;
    row_offsets(0, half * 2, oa, ea);
    zfetch(za, xx, oa, m64 ? src64 : 0);
    for (int ti = 0; ti < n_my; ++ti) {
      const int acc = ti & 1;
      const bool part = (((unsigned)blockIdx.x + (unsigned)ti * gridDim.x) + 1u) * BM > Mu || (SH && P.Lvalid < P.Lq);    // tile has missing rows
      row_offsets(ti, half * 2 + 1, ob, eb);
      if (TRU_L2_PREFETCH && EPI && m64 && (use_mask || has_extra) && ti + 1 < n_my && !P.planar) {
        // L2 prefetch of the NEXT tile's mask / added-tensor rows of this warp (64 rows x one 128-byte line of 32 channels; lane j
        // takes rows j and j + 32 of the warp's half tile)
        const int cbase = lg * 16;                                     // first channel of the warp's 16 (M = 64 accumulators only: measured)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const unsigned m = ((unsigned)blockIdx.x + (unsigned)(ti + 1) * gridDim.x) * BM + half * 64 + rr * 32 + lane;
          if (m < Mu) {
            const unsigned bt = m / Lq, q = m - bt * Lq;
            const unsigned off = (bt * P.Lout + q * P.omul + P.oadd) * (unsigned)P.ldo + P.ocoff + n0 + cbase;
            if (use_mask) prefetch_l2_off(P.zmask, off);
            if (has_extra) prefetch_l2_off(P.extra - P.ocoff, off);
          }
        }
      }
      mbar_wait(&mi.tfull[acc], (ti >> 1) & 1, 400 + ti);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + acc * BM + half * 64;
      uint32_t v[16];
      if (Lo.dbg & 16) { tc_fence_before(); mbar_arrive(&mi.tempty[acc]); continue; }
#define TRU_PROCESS(V, S0, Z, O) do { \
        if (part) { if (stmode == 1) process(std::true_type{}, std::integral_constant<int, 1>{}, V, S0, Z, O); \
                    else if (stmode == 2) process(std::true_type{}, std::integral_constant<int, EPI ? 2 : 0>{}, V, S0, Z, O); \
                    else process(std::true_type{}, std::integral_constant<int, 0>{}, V, S0, Z, O); } \
        else if (stmode == 1) process(std::false_type{}, std::integral_constant<int, 1>{}, V, S0, Z, O); \
        else if (stmode == 2) process(std::false_type{}, std::integral_constant<int, EPI ? 2 : 0>{}, V, S0, Z, O); \
        else process(std::false_type{}, std::integral_constant<int, 0>{}, V, S0, Z, O); } while (0)
      if (EPI != 2 || TRU_EPI2_PIPE) {
        // TMEM loads are software-pipelined: the load of sub-chunk s+1 is in flight while s is processed (a TMEM load
        // that competes with the MMAs of the next tile takes ~1.5k cycles - the profile showed the epilogue, and
        // behind it the whole ring, waiting on it four times per tile)
        uint32_t w[16];
        if (m64) {
          tmem_ld16x2_issue(taddr, v); tmem_ld_wait(v); addx(v, xx);
          tmem_ld16x2_issue(taddr + 32, w);
          zfetch(zb, xx, ob, src64);
          TRU_PROCESS(v, src64, za, oa);
          tmem_ld_wait(w); addx(w, xx);
          tc_fence_before();
          mbar_arrive(&mi.tempty[acc]);
          row_offsets(ti + 1, half * 2, oa, ea);
          zfetch(za, xx, oa, src64);
          TRU_PROCESS(w, src64, zb, ob);
        } else {
          tmem_ld16_issue(taddr, v); tmem_ld_wait(v); addx(v, xx);
          tmem_ld16_issue(taddr + 16, w);
          zfetch(zb, xx, oa, 16);
          TRU_PROCESS(v, 0, za, oa);
          tmem_ld_wait(w); addx(w, xx);
          tmem_ld16_issue(taddr + 32, v);
          zfetch(za, xx, ob, 0);
          TRU_PROCESS(w, 16, zb, oa);
          tmem_ld_wait(v); addx(v, xx);
          tmem_ld16_issue(taddr + 48, w);
          zfetch(zb, xx, ob, 16);
          TRU_PROCESS(v, 0, za, ob);
          tmem_ld_wait(w); addx(w, xx);
          tc_fence_before();
          mbar_arrive(&mi.tempty[acc]);
          const unsigned ob2 = ob;
          row_offsets(ti + 1, half * 2, oa, ea);
          zfetch(za, xx, oa, 0);
          TRU_PROCESS(w, 16, zb, ob2);
        }
      } else if (m64) {                // 2 sub-steps of 32 columns, all 32 lanes busy
        tmem_ld16x2(taddr, v); addx(v, xx);
        zfetch(zb, xx, ob, src64);
        TRU_PROCESS(v, src64, za, oa);
        tmem_ld16x2(taddr + 32, v); addx(v, xx);
        tc_fence_before();
        mbar_arrive(&mi.tempty[acc]);
        row_offsets(ti + 1, half * 2, oa, ea);
        zfetch(za, xx, oa, src64);
        TRU_PROCESS(v, src64, zb, ob);
      } else {                         // 4 sub-steps of 16 columns
        tmem_ld16(taddr, v); addx(v, xx);
        zfetch(zb, xx, oa, 16);
        TRU_PROCESS(v, 0, za, oa);
        tmem_ld16(taddr + 16, v); addx(v, xx);
        zfetch(za, xx, ob, 0);
        TRU_PROCESS(v, 16, zb, oa);
        tmem_ld16(taddr + 32, v); addx(v, xx);
        zfetch(zb, xx, ob, 16);
        TRU_PROCESS(v, 0, za, ob);
        tmem_ld16(taddr + 48, v); addx(v, xx);
        tc_fence_before();
        mbar_arrive(&mi.tempty[acc]);
        const unsigned ob2 = ob;
        row_offsets(ti + 1, half * 2, oa, ea);
        zfetch(za, xx, oa, 0);
        TRU_PROCESS(v, 16, zb, ob2);
      }
#undef TRU_PROCESS
    }
    double* gst = P.stats ? P.stats : (EPI ? P.bstats : nullptr);
    if (gst && nok) {
      if (EPI && do_bstats) {       // sum g*xhat = invstd * sum g*(z - mean)
        atomicAdd(gst + n, (double)s1);
        atomicAdd(gst + P.N + n, (double)s2 * (double)__ldg(P.binv + n));
      } else {
        atomicAdd(gst + n, (double)s1);
        atomicAdd(gst + P.N + n, (double)s2);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, Lo.wtmem ? 512 : 2 * BM);
  }
}

constexpr size_t SMEM_MAX = 227 * 1024;
constexpr int STAGE = 2 * BM * 128;       // bytes of a classic stage (hi + lo plane of 128 rows)
int g_dbg_flags = 0;        // ablation switches (tru_debug_set_flags)

int total_kblocks(const IgemmParams& p) {
  int nkb = 0;
  for (int s = 0; s < p.nseg; ++s) nkb += (p.seg[s].C + KBLK - 1) / KBLK;
  return nkb;
}

int max_kblocks(int MW);

bool shape_ok(const IgemmParams& p) {
  if (p.nseg < 1 || p.nseg > 5 || p.N < 1 || p.Lq < 1 || p.Lq > 32768) return false;
  for (int s = 0; s < p.nseg; ++s) {
    if (p.seg[s].C % 4 != 0 || p.seg[s].ld % 4 != 0 || p.seg[s].coff % 4 != 0 || p.seg[s].fs % 4 != 0) return false;
    // 32-bit element offsets inside the kernel
    const double fs = p.seg[s].fs ? p.seg[s].fs : (double)p.seg[s].Lsrc * p.seg[s].ld;
    if ((double)p.BT * fs >= 2147483648.0) return false;
  }
  if ((double)p.BT * p.Lout * std::max(p.ldo, 1) >= 4294967296.0 || (double)p.BT * p.Lq + BM >= 2147483648.0) return false;
  if (p.planar && (double)p.BT * p.N * p.Lout >= 4294967296.0) return false;
  if (p.extra && (p.ext_ld != p.ldo || p.planar)) return false;     // the added tensor shares the output's row offsets
  if (p.ntap) {                // tap-shared mode: one source, whole 32-channel blocks, shifts inside the 8 spare rows, one N block
    if (p.ntap > 5 || p.nseg != 1 || p.N > 128 || p.planar || p.seg[0].smul != 1 || p.seg[0].sadd != 0 || p.seg[0].C % KBLK != 0) return false;
    if (p.Lvalid < 0 || p.Lvalid > p.Lq || p.row_base > 0 || p.row_base < -7) return false;
    for (int j = 0; j < p.ntap; ++j)
      if (p.tap_shift[j] < 0 || p.tap_shift[j] > 7 || p.tap_c0[j] % KBLK != 0 || p.tap_C[j] % KBLK != 0 || p.tap_C[j] < KBLK ||
          p.tap_c0[j] < 0 || p.tap_c0[j] + p.tap_C[j] > p.seg[0].C) return false;
  } else if (p.Lvalid != 0 && p.Lvalid != p.Lq) return false;
  if (p.dw_k) {                // depthwise epilogue: whole frames per half tile, one 128-lane accumulator, nothing else in the epilogue
    if (p.ntap || p.planar || p.extra || p.use_mask || p.stats || p.bstats || !p.dw_w || p.N <= 64 || p.N > 128) return false;
    if (!((p.dw_k == 5 && p.dw_s == 2) || (p.dw_k == 3 && (p.dw_s == 1 || p.dw_s == 2)))) return false;
    if (p.Lq < 16 || p.Lq % 16 != 0 || 128 % p.Lq != 0) return false;
    if (p.Lout != (p.Lq + 2 * (p.dw_k / 2) - p.dw_k) / p.dw_s + 1) return false;
    if (total_kblocks(p) > max_kblocks(128)) return false;       // (never split into k passes)
  }
  return true;
}

int weight_rows(const IgemmParams& p) { return p.N <= 64 ? 64 : 128; }

// the most k-blocks one classic launch can keep resident next to a 2-stage ring
int max_kblocks(int MW) {
  const size_t fixed = 2 * STAGE + sizeof(Misc) + 64 + (size_t)MAXCOEF * COEF_FLOATS * 4;
  return (int)std::min<size_t>(MAXKB, (SMEM_MAX - fixed) / ((size_t)MW * 256));
}

// TRU_W_TMEM_OFF=1 keeps the weight slice in shared memory (A/B aid)
bool wtmem_off() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("TRU_W_TMEM_OFF"); v = (e && atoi(e)) ? 1 : 0; }
  return v != 0;
}

bool plan(const IgemmParams& p, TcLayout& L, dim3& grid, size_t& smem_bytes) {
  if (!shape_ok(p)) return false;
  memset(&L, 0, sizeof(L));
  L.MW = weight_rows(p);
  L.shared = p.ntap > 0;
  int nkb = 0, ncoef = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const Seg& sg = p.seg[s];
    for (int c0 = 0; c0 < sg.C; c0 += KBLK) {
      if (nkb >= MAXKB) return false;
      L.kb_seg[nkb] = (int8_t)s; L.kb_c0[nkb] = (int16_t)c0; L.kb_valid[nkb] = (int16_t)std::min(KBLK, sg.C - c0);
      L.kb_relu[nkb] = (int8_t)(sg.p0 && sg.relu);
      L.kb_lmax[nkb] = (L.shared && p.c_hi > 0 && c0 >= p.c_hi) ? p.lmax_hi : sg.Lsrc;
      L.kb_pf[nkb] = (int8_t)(!L.shared && sg.smul == 1 && sg.Lsrc == p.Lq && (sg.fs == 0 || sg.fs == sg.Lsrc * sg.ld));
      L.kb_lin[nkb] = (int8_t)(L.kb_pf[nkb] && sg.sadd == 0 && L.kb_valid[nkb] == KBLK && TRU_LINEAR_ROWS);
      L.kb_coef[nkb] = -1;
      if (sg.p0) {
        int e = -1;
        for (int t = 0; t < ncoef && e < 0; ++t) {
          const Seg& o = p.seg[L.coef_seg[t]];
          if (o.p0 == sg.p0 && o.p1 == sg.p1 && o.p2 == sg.p2 && o.coff + L.coef_c0[t] == sg.coff + c0 &&
              std::min(KBLK, o.C - L.coef_c0[t]) == L.kb_valid[nkb]) e = t;
        }
        if (e < 0) {
          if (ncoef >= MAXCOEF) return false;
          e = ncoef++;
          L.coef_seg[e] = (int8_t)s; L.coef_c0[e] = (int16_t)c0;
        }
        L.kb_coef[nkb] = (int8_t)e;
      }
      ++nkb;
    }
  }
  // resident weight k-blocks, grouped by the source k-block that feeds them (the MMA thread walks them in this order)
  int nwk = 0, maxshift = 0;
  for (int kb = 0; kb < nkb; ++kb) {
    if (!L.shared) {
      const Seg& sg = p.seg[L.kb_seg[kb]];
      L.wk_seg[nwk] = L.kb_seg[kb]; L.wk_shift[nwk] = 0; L.wk_c0[nwk] = L.kb_c0[kb]; L.wk_valid[nwk] = L.kb_valid[kb];
      L.wk_wbase[nwk] = sg.wbase;
      L.kb_nw[kb] = 1; ++nwk;
      continue;
    }
    const int c0 = L.kb_c0[kb];
    for (int j = 0; j < p.ntap; ++j) {
      if (c0 < p.tap_c0[j] || c0 >= p.tap_c0[j] + p.tap_C[j]) continue;
      if (nwk >= MAXWK) return false;
      L.wk_seg[nwk] = 0; L.wk_shift[nwk] = (int8_t)p.tap_shift[j]; L.wk_c0[nwk] = (int16_t)(c0 - p.tap_c0[j]); L.wk_valid[nwk] = KBLK;
      L.wk_wbase[nwk] = p.tap_wbase[j];
      maxshift = std::max(maxshift, p.tap_shift[j]);
      ++L.kb_nw[kb]; ++nwk;
    }
    if (L.kb_nw[kb] == 0) return false;       // a staged channel block no tap reads
  }
  L.nkb = nkb; L.ncoef = ncoef; L.nwk = nwk; L.dbg = g_dbg_flags;
  L.all_lin = 1;
  for (int kb = 0; kb < nkb; ++kb) L.all_lin &= L.kb_lin[kb];
  L.arows = BM + maxshift; L.atile = (uint32_t)L.arows * 128; L.stage = 2 * L.atile;
  L.lq_magic = p.Lq == 1 ? 0u : (unsigned)((0x100000000ull + (unsigned)p.Lq - 1) / (unsigned)p.Lq);
  L.wtmem = (TRU_W_IN_TMEM && !wtmem_off() && nwk * 64 <= 512 - 2 * BM) ? 1 : 0;
  const size_t w = L.wtmem ? 0 : (size_t)2 * nwk * L.MW * 128;
  const size_t coefb = (size_t)ncoef * COEF_FLOATS * 4;
  const size_t fixed = coefb + sizeof(Misc) + 64;
  if (fixed + w + 2 * (size_t)L.stage > SMEM_MAX) return false;
  L.nstage = (int)std::min<size_t>(6, (SMEM_MAX - fixed - w) / L.stage);
  L.a_off = (uint32_t)w;                                   // multiple of 1024
  L.coef_off = L.a_off + L.nstage * L.stage;               // multiple of 128
  L.misc_off = (uint32_t)align_up(L.coef_off + coefb, 16);
  smem_bytes = L.misc_off + sizeof(Misc);
  const int ny = (p.N + L.MW - 1) / L.MW;
  const long M = (long)p.BT * p.Lq;
  const int ntiles = (int)((M + BM - 1) / BM);
  grid = dim3(std::max(1, std::min(ntiles, sm_count() / ny)), ny);
  return smem_bytes <= SMEM_MAX;
}

template <int LW, bool LD2, int EPI, bool SH>
int launch_inst(const IgemmParams& p, const TcLayout& L, dim3 grid, size_t smem, cudaStream_t st) {
  TRU_SMEM_OPT_IN((tc_igemm_kernel<LW, LD2, EPI, SH>), SMEM_MAX);
  TRU_CUDA(launch_pdl(tc_igemm_kernel<LW, LD2, EPI, SH>, grid, dim3(32 * (4 + LW + EW)), smem, st, p, L));
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

int launch_one(const IgemmParams& p, const TcLayout& L, dim3 grid, size_t smem, cudaStream_t st) {
  const int epi = p.extra != nullptr ? 2 : ((p.use_mask || p.bstats != nullptr) ? 1 : 0);
  bool ld2 = false;
  for (int s = 0; s < p.nseg; ++s) ld2 |= (p.seg[s].src2 != nullptr);
  if (p.dw_k) return launch_inst<12, false, 3, false>(p, L, grid, smem, st);      // inference: pointwise conv with the depthwise conv in the epilogue
  if (L.shared) {        // transposed convs: forward (affine + ReLU on load, plain / statistics epilogue), data gradient (masked)
    if (!ld2 && epi == 0) return launch_inst<12, false, 0, true>(p, L, grid, smem, st);
    if (ld2 && epi == 0) return launch_inst<12, true, 0, true>(p, L, grid, smem, st);
    if (ld2 && epi == 1) return launch_inst<8, true, 1, true>(p, L, grid, smem, st);      // (12 loader warps measured slower: 0.57 vs 0.51 ms on dec4 - the masked epilogue is the slow role)
    if (!ld2 && epi == 1) return launch_inst<12, false, 1, true>(p, L, grid, smem, st);     // (no BN behind the conv: tests only)
    return set_error(TRU_ERR_ARG, "igemm_tc: no tap-shared kernel variant for this loader / epilogue combination");
  }
  if (ld2 && epi >= 1 && p.nseg == 1)      // masked backward of a pointwise conv: the epilogue is the slow role, 8 loader warps leave it more registers
    return epi == 1 ? launch_inst<8, true, 1, false>(p, L, grid, smem, st) : launch_inst<8, true, 2, false>(p, L, grid, smem, st);
  if (ld2) return epi == 0 ? launch_inst<12, true, 0, false>(p, L, grid, smem, st) : epi == 1 ? launch_inst<12, true, 1, false>(p, L, grid, smem, st) : launch_inst<12, true, 2, false>(p, L, grid, smem, st);
  return epi == 0 ? launch_inst<12, false, 0, false>(p, L, grid, smem, st) : epi == 1 ? launch_inst<12, false, 1, false>(p, L, grid, smem, st) : launch_inst<12, false, 2, false>(p, L, grid, smem, st);
}

int launch_planned(const IgemmParams& p0, cudaStream_t st) {
  TcLayout L;
  dim3 grid;
  size_t smem = 0;
  if (!plan(p0, L, grid, smem)) return 1;
  IgemmParams p = p0;
  if (p.Lvalid == 0) p.Lvalid = p.Lq;        // classic launches: every row is an output row
  return launch_one(p, L, grid, smem, st);
}

// k-blocks [k0, k1) of p as a launch of their own
IgemmParams slice_kblocks(const IgemmParams& p, int k0, int k1) {
  IgemmParams q = p;
  q.nseg = 0;
  int kb = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const Seg& sg = p.seg[s];
    const int skb = (sg.C + KBLK - 1) / KBLK;
    const int a = std::max(k0, kb), b = std::min(k1, kb + skb);
    if (a < b) {
      Seg t = sg;
      const int c0 = (a - kb) * KBLK, c1 = std::min(sg.C, (b - kb) * KBLK);
      t.coff = sg.coff + c0; t.C = c1 - c0; t.wbase = sg.wbase + c0 * sg.wsc;
      q.seg[q.nseg++] = t;
    }
    kb += skb;
  }
  return q;
}

}  // namespace

void set_tc_debug_flags(int f) { g_dbg_flags = f; }
int read_mbar_debug(unsigned* out, int n) {
#ifdef TRU_MBAR_TIMEOUT
  cudaMemcpyFromSymbol(out, tc::g_mbar_dbg, sizeof(unsigned) * (size_t)std::min(n, 1028));
  return 1;
#else
  (void)out; (void)n;
  return 0;
#endif
}

bool igemm_tc_eligible(const IgemmParams& p) {
  if (!shape_ok(p)) return false;
  if (p.ntap) { TcLayout L; dim3 g; size_t s = 0; return plan(p, L, g, s); }      // tap-shared launches are never split
  const int nkb = total_kblocks(p), cap = max_kblocks(weight_rows(p));
  if (nkb <= cap) { TcLayout L; dim3 g; size_t s = 0; return plan(p, L, g, s); }
  return !p.planar && cap >= 1 && (nkb + cap - 1) / cap <= 4;
}

// returns TRU_OK if launched, 1 if the shape is not eligible (caller falls back to the FFMA kernel).
// When the weight slice does not fit in shared memory the reduction is split into passes over
// k-block ranges: pass 0 writes the partial product (+bias, +skip gradient), later passes add to it
// in place, and the last one applies the mask / accumulates the statistics.
int launch_igemm_tc(const IgemmParams& p, cudaStream_t st) {
  if (!igemm_tc_eligible(p)) return 1;
  const int nkb = total_kblocks(p), cap = max_kblocks(weight_rows(p));
  if (p.ntap || nkb <= cap) return launch_planned(p, st);
  const int npass = (nkb + cap - 1) / cap, per = (nkb + npass - 1) / npass;
  for (int i = 0; i < npass; ++i) {
    IgemmParams q = slice_kblocks(p, i * per, std::min(nkb, (i + 1) * per));
    const bool first = i == 0, last = i == npass - 1;
    if (!first) { q.bias = nullptr; q.extra = p.out + p.ocoff; q.ext_ld = p.ldo; }
    if (!last) { q.stats = nullptr; q.use_mask = 0; q.zmask = nullptr; q.bstats = nullptr; }
    const int rc = launch_planned(q, st);
    if (rc) return rc < 0 ? rc : set_error(TRU_ERR_ARG, "igemm_tc: k-split pass %d not plannable", i);
  }
  return TRU_OK;
}

}  // namespace tru
