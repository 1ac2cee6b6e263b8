// Streaming tensor-core weight gradient (tcgen05 / TMEM, 3xTF32) for the conv layers:
//
//   dW[s][c][j][n] += sum_m  a_s(m, c) * dz(zs*m + j - zpad, n)
//
// for up to two activation sources s (the two halves of a skip concat) and up to five taps
// j (a transposed conv), all in ONE pass over the rows: every operand row is read from HBM
// exactly once per 16-row unit (plus the k-s halo rows of a transposed conv).
//
// A weight gradient is a pure streaming reduction (M ~ 10^6 rows, a 64..320 x 64..192 result),
// so the kernel is organised around keeping ~100 KB of loads in flight per SM:
//   warp 1       producer: per 16-row unit, one cp.async.bulk (TMA, no tensor map) per operand
//                (the rows of a unit that exist are one contiguous run in memory) into a ring of
//                RAW fp32 stages, completion on an mbarrier (expect_tx); rows that fall outside
//                their frame are skipped and flagged
//   warps 4-19   transform: raw stage -> BN/ReLU (activations) or BN-backward affine (dz, applied
//                ONCE per row, then scattered to every tap slot it feeds) -> truncating tf32
//                split -> MN-major SWIZZLE_128B_BASE32B operand tiles (2 stages)
//   warp 0       MMA issuer: per unit 2 k-steps x 3 products of tcgen05.mma.kind::tf32 into ONE
//                fp32 accumulator (P channels x Q columns) that lives in TMEM for the CTA's
//                whole row range; warps 4-7 add it to dW with atomics at the end.
// P (the MMA M side, 64 or 128 channels) is the activation tile when there are several taps and
// the dz tile when there are two sources; Q (up to 384 columns) is the other one.
#include <algorithm>
#include <map>
#include <string>
#include <string.h>
#include "net_kernels.cuh"
#include "tc_common.cuh"

namespace tru {
namespace {
using namespace tc;

constexpr int UR = 16;                          // rows per unit (2 k-steps of 8)
constexpr int TRW = 16;                         // transform warps
constexpr int NTR = 32 * TRW;
constexpr int NT = 32 * (4 + TRW);
constexpr int MAXRAW = 8;

struct WgK {
  int nsrc; const float* a_src[2]; const float* a_p0[2]; const float* a_p2[2];
  int a_L[2], a_ld[2], a_add[2], a_C[2], a_c0[2], a_coff[2], Ca;
  const float* z_src; const float* z_src2; const float* z_p0; const float* z_p1; const float* z_p2;
  int z_L, z_ld, z_coff, N, ntap, zs, zpad, NZ;
  int p_is_z, PWt, QWt, Mmma, Pvalid, Qvalid, tmem_cols;
  uint32_t op_stage, p_tile, q_tile, raw_off, raw_stage, raw_a[2], raw_dy, raw_z, coef_off, misc_off;
  int nraw, strided;            // strided: bit o set if operand o (0/1 A sources, 2 dY, 3 Z) is a channel slice of wider rows
  float* dW; int wbase[2], wsc, wsn, wtap; float* db;
  int n_split; float* dW2; float* db2;
  float* scratch;                // per-CTA partial accumulators [CTA][Mmma][QWt] (null: atomics straight into dW)
  int BT, Lq; unsigned units_total, units_per_cta;
  int ur;                        // rows per unit: 16, or 32 for narrow layers (few bytes per row: the per-unit hand-offs, not the bytes, set the pace)
  int dbg;                       // ablation switches (env TRU_WG_DBG, tuning aid; results are garbage when set): 1 no MMAs, 2 no transform math / stores, 4 no proxy fence
};

struct StageFlags { uint32_t amask[2]; uint32_t zmask[2]; };
struct WMisc2 {
  uint64_t raw_full[MAXRAW], raw_empty[MAXRAW], op_full[2], op_empty[2], done;
  StageFlags flags[MAXRAW];
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// byte offset of (k-row r, channel c) in an MN-major SWIZZLE_128B_BASE32B tile of NB 32-channel blocks
__device__ __forceinline__ uint32_t tile_off(int r, int c, int NB) {
  return (uint32_t)((r >> 2) * NB + (c >> 5)) * 512u + (uint32_t)(r & 3) * 128u +
         ((((uint32_t)(c >> 3) & 3u) ^ (uint32_t)(r & 3)) << 5) + (uint32_t)(c & 7) * 4u;
}
__device__ __forceinline__ void split_store(uint8_t* hi_ptr, uint32_t lo_delta, const float4& v) {
  uint4 hi, lo;
  hi.x = __float_as_uint(v.x) & 0xffffe000u; hi.y = __float_as_uint(v.y) & 0xffffe000u;
  hi.z = __float_as_uint(v.z) & 0xffffe000u; hi.w = __float_as_uint(v.w) & 0xffffe000u;
  lo.x = __float_as_uint(v.x - __uint_as_float(hi.x)); lo.y = __float_as_uint(v.y - __uint_as_float(hi.y));
  lo.z = __float_as_uint(v.z - __uint_as_float(hi.z)); lo.w = __float_as_uint(v.w - __uint_as_float(hi.w));
  *(uint4*)hi_ptr = hi;
  *(uint4*)(hi_ptr + lo_delta) = lo;
}

// destination of accumulator element (p, q)
__device__ __forceinline__ float* wgrad_dst(const WgK& K, int p, int q) {
  const int acol = K.p_is_z ? q : p, zcol = K.p_is_z ? p : q;
  const int s = (K.nsrc == 2 && acol >= K.a_c0[1]) ? 1 : 0, c = acol - K.a_c0[s];
  const int tap = zcol / K.N, n = zcol - tap * K.N;
  if (K.n_split > 0 && n >= K.n_split) return K.dW2 + K.wbase[s] + (long)c * K.wsc + (long)(n - K.n_split) * K.wsn;
  return K.dW + K.wbase[s] + (long)tap * K.wtap + (long)c * K.wsc + (long)n * K.wsn;
}

// second stage of the weight-gradient reduction: dW[p][q] += sum over CTAs of scratch[cta][p][q] (fixed order: deterministic).
// Block = one accumulator row p x 32 columns; 8 warps each sum every 8th partial (short load chains), then meet in shared memory.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgK K, int nctas) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][32];
  const int nqb = K.QWt >> 5;
  const int p = blockIdx.x / nqb, q = (blockIdx.x - p * nqb) * 32 + (threadIdx.x & 31), cg = threadIdx.x >> 5;
  const float* src = K.scratch + (size_t)p * K.QWt + q;
  const size_t stride = (size_t)K.Mmma * K.QWt;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = cg;
  for (; c + 24 < nctas; c += 32) {
    s0 += src[(size_t)c * stride]; s1 += src[(size_t)(c + 8) * stride]; s2 += src[(size_t)(c + 16) * stride]; s3 += src[(size_t)(c + 24) * stride];
  }
  for (; c < nctas; c += 8) s0 += src[(size_t)c * stride];
  red[cg][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (cg == 0 && p < K.Pvalid && q < K.Qvalid) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += red[g][threadIdx.x];
    *wgrad_dst(K, p, q) += s;
  }
}

// IA / IZ: per-thread item slots (activation / dz float4s per unit), NTAP: taps (compile-time so that unused
// slots and taps cost nothing; with one slot each the BN coefficients live in registers)
template <int IA, int IZ, int NTAP, int URT>
__global__ void __launch_bounds__(NT, 1) tc_wgrad_stream_kernel(const __grid_constant__ WgK K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ops = smem;                                   // 2 x [P hi | P lo | Q hi | Q lo]
  uint8_t* raws = smem + K.raw_off;                      // nraw x [A rows | dY rows | Z rows]
  float* coef = (float*)(smem + K.coef_off);             // A: p0[Ca] p2[Ca] floor[Ca]; Z: q0[N] q1[N] q2[N]
  WMisc2& mi = *(WMisc2*)(smem + K.misc_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned u0 = min(K.units_total, blockIdx.x * K.units_per_cta), u1 = min(K.units_total, u0 + K.units_per_cta);
  const unsigned nun = u1 - u0;
  const int nraw = K.nraw;

  // ---- setup: coefficient tables, barriers, TMEM -----------------------------------------
  pdl_trigger();
  pdl_wait();                                            // the coefficient tables come from bn_finalize / bn_bwd_finalize
  for (int i = tid; i < K.Ca; i += NT) {
    const int s = (K.nsrc == 2 && i >= K.a_c0[1]) ? 1 : 0, c = i - K.a_c0[s];
    const bool aff = K.a_p0[s] != nullptr;
    coef[i] = aff ? __ldg(K.a_p0[s] + c) : 1.f;
    coef[K.Ca + i] = aff ? __ldg(K.a_p2[s] + c) : 0.f;
    coef[2 * K.Ca + i] = aff ? 0.f : -__int_as_float(0x7f800000);     // ReLU only behind a BN (trunet.cu: fwd_seg)
  }
  for (uint32_t i = tid; i < 2 * K.op_stage / 16; i += NT) ((uint4*)ops)[i] = make_uint4(0u, 0u, 0u, 0u);   // pad columns stay zero
  float* zc = coef + 3 * K.Ca;
  for (int i = tid; i < K.N; i += NT) {
    zc[i] = K.z_p0 ? __ldg(K.z_p0 + i) : 1.f;
    zc[K.N + i] = (K.z_p0 && K.z_p1) ? __ldg(K.z_p1 + i) : 0.f;
    zc[2 * K.N + i] = K.z_p0 ? __ldg(K.z_p2 + i) : 0.f;
  }
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < nraw; ++s) { mbar_init(&mi.raw_full[s], 1); mbar_init(&mi.raw_empty[s], TRW); }
      for (int s = 0; s < 2; ++s) { mbar_init(&mi.op_full[s], TRW); mbar_init(&mi.op_empty[s], 1); }
      mbar_init(&mi.done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&mi.tmem_base, K.tmem_cols);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mi.tmem_base;
  const int NBp = K.PWt >> 5, NBq = K.QWt >> 5;

  if (warp < 4) {
  reg_dec<48>();                               // 4*32*48 + 16*32*104 = 59392 <= 640 threads x 96 registers
  if (warp == 0) {
    // ================================ MMA issuer ==================================
    if (lane == 0 && nun > 0) {
      const uint64_t dP = ((smem_desc_sw128_32b(0, 512, NBp * 512) >> 16) << 16);
      const uint64_t dQ = ((smem_desc_sw128_32b(0, 512, NBq * 512) >> 16) << 16);
      const int nsplit = (K.QWt + 255) / 256, per = (NBq + nsplit - 1) / nsplit;     // Q blocks per MMA
      const uint32_t obase = smem_u32(ops) >> 4;
      for (unsigned u = 0; u < nun; ++u) {
        const int os = u & 1;
        mbar_wait(&mi.op_full[os], (u >> 1) & 1);
        tc_fence_after();
        const uint32_t p_hi = obase + os * (K.op_stage >> 4), p_lo = p_hi + (K.p_tile >> 4);
        const uint32_t q_hi = p_lo + (K.p_tile >> 4), q_lo = q_hi + (K.q_tile >> 4);
        for (int sp = 0; sp < nsplit; ++sp) {
          const int b0 = sp * per, nb = min(per, NBq - b0);
          const uint32_t idesc = idesc_tf32(K.Mmma, nb * 32, 1, 1);
          const uint32_t d = tmem + b0 * 32;
#pragma unroll
          for (int g = 0; g < URT / 8; ++g) {                  // k-steps of 8 rows = 2 atoms
            const uint32_t po = g * NBp * 64, qo = g * NBq * 64 + b0 * 32;     // (bytes >> 4)
            if (K.dbg & 1) continue;
            mma_tf32(d, dP | (p_lo + po), dQ | (q_hi + qo), idesc, (u | g) != 0);
            mma_tf32(d, dP | (p_hi + po), dQ | (q_lo + qo), idesc, 1);
            mma_tf32(d, dP | (p_hi + po), dQ | (q_hi + qo), idesc, 1);
          }
        }
        mma_commit(&mi.op_empty[os]);
      }
      mma_commit(&mi.done);
    }
  } else {
    // ================================ producers (TMA bulk copies) ====================
    // Three warps take turns (unit u belongs to warp 1 + u % 3): one warp's serial per-unit chain (plan, shuffles, flags,
    // expect_tx, copies) was as long as the unit's HBM time, so the ring could not be kept full.
    // (a producer that advances NPW units at a time must not run two ring turns ahead of the consumers, or it mis-reads
    // the parity of raw_empty: NPW <= nraw)
    const int NPW = min(3, nraw);
    const int pw = warp - 1;
    // Rows of a unit that exist are one contiguous run in their frame.  Dense operands (ld == channels) need ONE bulk
    // copy per unit (lane 0/1 = activation sources, lane 2 = dY, lane 3 = Z); an operand that is a channel slice of wider
    // rows (GRU gate / hidden-state slices) is copied row by row, the rows spread over the 32 lanes.
    const unsigned Lq = (unsigned)K.Lq;
    struct Run { int lo, hi, ld; const float* src; uint32_t dst, rowb; };
    // operand o of the unit starting at (bt, q0): 0/1 activation sources, 2 dY, 3 Z
    auto plan = [&](int o, unsigned bt, unsigned q0) {
      Run r{0, 0, 0, nullptr, 0u, 0u};
      if (o < K.nsrc) {
        const int l0 = (int)q0 + K.a_add[o];
        r.lo = max(0, -l0); r.hi = min(URT, K.a_L[o] - l0);
        r.rowb = (uint32_t)K.a_C[o] * 4u; r.ld = K.a_ld[o];
        r.src = K.a_src[o] + ((size_t)bt * K.a_L[o] + (l0 + r.lo)) * r.ld + K.a_coff[o];
        r.dst = K.raw_a[o] + (uint32_t)r.lo * r.rowb;
      } else if (o == 2 || (o == 3 && K.z_src2)) {
        const int r0 = K.zs * (int)q0 - K.zpad;
        r.lo = max(0, -r0); r.hi = min(K.NZ, K.z_L - r0);
        r.rowb = (uint32_t)K.N * 4u; r.ld = K.z_ld;
        r.src = (o == 2 ? K.z_src : K.z_src2) + ((size_t)bt * K.z_L + (r0 + r.lo)) * r.ld + K.z_coff;
        r.dst = (o == 2 ? K.raw_dy : K.raw_z) + (uint32_t)r.lo * r.rowb;
      }
      return r;
    };
    int rs = pw;
    uint32_t ph = 0;
    while (rs >= nraw) { rs -= nraw; ph ^= 1; }
    for (unsigned u = pw; pw < NPW && u < nun; u += NPW) {
      const unsigned m0 = (u0 + u) * URT, bt = m0 / Lq, q0 = m0 - bt * Lq;
      const Run r = plan(lane, bt, q0);              // lane o < 4 plans operand o (lanes >= 4: empty run)
      const bool has = lane < 4 && r.hi > r.lo;
      const uint32_t mybytes = has ? (uint32_t)(r.hi - r.lo) * r.rowb : 0u;
      uint32_t bytes = mybytes;
      bytes += __shfl_xor_sync(0xffffffffu, bytes, 1);
      bytes += __shfl_xor_sync(0xffffffffu, bytes, 2);
      const unsigned long long run = has ? (((r.hi >= 64 ? ~0ull : (1ull << r.hi) - 1ull)) & ~((1ull << r.lo) - 1ull)) : 0ull;
      const unsigned long long r0m = __shfl_sync(0xffffffffu, run, 0), r1m = __shfl_sync(0xffffffffu, run, 1), r2m = __shfl_sync(0xffffffffu, run, 2);
      mbar_wait(&mi.raw_empty[rs], ph ^ 1);
      uint8_t* st = raws + (size_t)rs * K.raw_stage;
      if (lane == 0) {              // flags and the expected byte count are posted before the first copy can land
        mi.flags[rs].amask[0] = (uint32_t)r0m; mi.flags[rs].amask[1] = (uint32_t)r1m;
        mi.flags[rs].zmask[0] = (uint32_t)r2m; mi.flags[rs].zmask[1] = (uint32_t)(r2m >> 32);
        mbar_arrive_expect_tx(&mi.raw_full[rs], bytes);
      }
      __syncwarp();
      if (has && !((K.strided >> lane) & 1)) bulk_g2s(st + r.dst, r.src, mybytes, &mi.raw_full[rs]);     // dense rows: one copy
      if (K.strided) {              // channel slices of wider rows: one copy per row, rows spread over the lanes
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          if (!((K.strided >> o) & 1)) continue;
          const int n = __shfl_sync(0xffffffffu, has ? r.hi - r.lo : 0, o);
          const unsigned long long sp = __shfl_sync(0xffffffffu, (unsigned long long)r.src, o);
          const uint32_t dst = __shfl_sync(0xffffffffu, r.dst, o), rowb = __shfl_sync(0xffffffffu, r.rowb, o);
          const int ld = __shfl_sync(0xffffffffu, r.ld, o);
          for (int i = lane; i < n; i += 32)
            bulk_g2s(st + dst + (uint32_t)i * rowb, (const float*)sp + (size_t)i * ld, rowb, &mi.raw_full[rs]);
        }
      }
      rs += NPW;
      while (rs >= nraw) { rs -= nraw; ph ^= 1; }
    }
  }
  } else {
    // ================================ transform =======================================
    reg_inc<104>();
    const int tt = tid - 128;
    const int NBa = (K.Ca + 31) >> 5, NBz = (K.N + 31) >> 5;     // item grids are padded to 32 channels; pad slots stay idle
    // Per-thread item slots, fully decoded once: raw-stage byte offset, operand-stage byte offset(s), channel and
    // validity bit.  Lane = (row & 3) * 8 + 16-byte chunk, so a quarter warp reads 128 contiguous bytes of a raw
    // row and writes 128 bytes of one swizzle atom (conflict-free both ways).
    const uint32_t a_tile = K.p_is_z ? 2 * K.p_tile : 0u, a_lo = K.p_is_z ? K.q_tile : K.p_tile;
    const uint32_t z_tile = K.p_is_z ? 0u : 2 * K.p_tile, z_lo = K.p_is_z ? K.p_tile : K.q_tile;
    const int NBat = K.p_is_z ? NBq : NBp, NBzt = K.p_is_z ? NBp : NBq;
    constexpr uint32_t NONE = 0xffffffffu;
    uint32_t a_raw[IA], a_dst[IA], a_cb[IA];          // a_cb: channel | row << 16 | source << 24
    uint32_t z_raw[IZ], z_cb[IZ], z_dst[IZ][NTAP];    // z_cb: channel | t << 16
    const int nA = URT * NBa * 8, nZ = ((K.NZ + 3) & ~3) * NBz * 8;
#pragma unroll
    for (int k = 0; k < IA; ++k) {
      const int it = tt + k * NTR;
      a_raw[k] = NONE; a_dst[k] = 0; a_cb[k] = 0;
      if (it < nA) {
        const int blk = (it >> 5) % NBa, rg = (it >> 5) / NBa, r = rg * 4 + ((it >> 3) & 3), c = blk * 32 + (it & 7) * 4;
        const int s = (K.nsrc == 2 && c >= K.a_c0[1]) ? 1 : 0;
        if (c < K.Ca) {
          a_raw[k] = K.raw_a[s] + (uint32_t)(r * K.a_C[s] + c - K.a_c0[s]) * 4u;
          a_dst[k] = a_tile + tile_off(r, c, NBat);
          a_cb[k] = (uint32_t)c | ((uint32_t)r << 16) | ((uint32_t)s << 24);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < IZ; ++k) {
      const int it = tt + k * NTR;
      z_raw[k] = NONE; z_cb[k] = 0;
#pragma unroll
      for (int j = 0; j < NTAP; ++j) z_dst[k][j] = NONE;
      if (it < nZ) {
        const int blk = (it >> 5) % NBz, rg = (it >> 5) / NBz, t = rg * 4 + ((it >> 3) & 3), c = blk * 32 + (it & 7) * 4;
        if (t < K.NZ && c < K.N) {
          z_raw[k] = (uint32_t)(t * K.N + c) * 4u;
          z_cb[k] = (uint32_t)c | ((uint32_t)t << 16);
          // row t of the unit's dz window feeds tap j at unit row (t - j) / zs
#pragma unroll
          for (int j = 0; j < NTAP; ++j) {
            const int d = t - j;
            if (d >= 0 && d % K.zs == 0 && d / K.zs < URT) z_dst[k][j] = z_tile + tile_off(d / K.zs, j * K.N + c, NBzt);
          }
        }
      }
    }
    const bool has_z2 = K.z_src2 != nullptr;
    constexpr bool CREG = IA == 1 && IZ == 1;
    float4 rp0, rp2, rfl, rq0, rq1, rq2;
    if (CREG) {
      const int ca = a_cb[0] & 0xffff, cz = z_cb[0] & 0xffff;
      rp0 = *(const float4*)(coef + ca); rp2 = *(const float4*)(coef + K.Ca + ca); rfl = *(const float4*)(coef + 2 * K.Ca + ca);
      rq0 = *(const float4*)(zc + cz); rq1 = *(const float4*)(zc + K.N + cz); rq2 = *(const float4*)(zc + 2 * K.N + cz);
    }
    float bs[IZ][4];
#pragma unroll
    for (int k = 0; k < IZ; ++k) bs[k][0] = bs[k][1] = bs[k][2] = bs[k][3] = 0.f;
    int rs = 0;
    uint32_t ph = 0;
    for (unsigned u = 0; u < nun; ++u) {
      const int os = u & 1;
      mbar_wait(&mi.raw_full[rs], ph);
      const uint32_t am0 = mi.flags[rs].amask[0], am1 = mi.flags[rs].amask[1];
      const uint32_t zm0 = mi.flags[rs].zmask[0], zm1 = mi.flags[rs].zmask[1];
      const uint8_t* st = raws + (size_t)rs * K.raw_stage;
      // raw loads first (independent of the operand stage), then wait for the stage to be free
      float4 xa[IA], xy[IZ], xz[IZ];
#pragma unroll
      for (int k = 0; k < IA; ++k)
        if (a_raw[k] != NONE) xa[k] = *(const float4*)(st + a_raw[k]);
#pragma unroll
      for (int k = 0; k < IZ; ++k)
        if (z_raw[k] != NONE) {
          xy[k] = *(const float4*)(st + K.raw_dy + z_raw[k]);
          xz[k] = has_z2 ? *(const float4*)(st + K.raw_z + z_raw[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      mbar_wait(&mi.op_empty[os], ((u >> 1) & 1) ^ 1);
      uint8_t* op = ops + (size_t)os * K.op_stage;
      if (!(K.dbg & 2)) {
#pragma unroll
      for (int k = 0; k < IA; ++k) {
        if (a_raw[k] != NONE) {
          const int c = a_cb[k] & 0xffff, r = (a_cb[k] >> 16) & 0xff;
          const bool ok = (((a_cb[k] >> 24) ? am1 : am0) >> r) & 1u;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            const float4 p0 = CREG ? rp0 : *(const float4*)(coef + c), p2 = CREG ? rp2 : *(const float4*)(coef + K.Ca + c);
            const float4 fl = CREG ? rfl : *(const float4*)(coef + 2 * K.Ca + c);
            v.x = fmaxf(fmaf(p0.x, xa[k].x, p2.x), fl.x); v.y = fmaxf(fmaf(p0.y, xa[k].y, p2.y), fl.y);
            v.z = fmaxf(fmaf(p0.z, xa[k].z, p2.z), fl.z); v.w = fmaxf(fmaf(p0.w, xa[k].w, p2.w), fl.w);
          }
          split_store(op + a_dst[k], a_lo, v);
        }
      }
#pragma unroll
      for (int k = 0; k < IZ; ++k) {
        if (z_raw[k] != NONE) {
          const int c = z_cb[k] & 0xffff, t = z_cb[k] >> 16;
          const bool ok = ((t < 32 ? zm0 >> t : zm1 >> (t - 32)) & 1u) != 0;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            const float4 q0 = CREG ? rq0 : *(const float4*)(zc + c), q1 = CREG ? rq1 : *(const float4*)(zc + K.N + c);
            const float4 q2 = CREG ? rq2 : *(const float4*)(zc + 2 * K.N + c);
            v.x = fmaf(q1.x, xz[k].x, fmaf(q0.x, xy[k].x, q2.x)); v.y = fmaf(q1.y, xz[k].y, fmaf(q0.y, xy[k].y, q2.y));
            v.z = fmaf(q1.z, xz[k].z, fmaf(q0.z, xy[k].z, q2.z)); v.w = fmaf(q1.w, xz[k].w, fmaf(q0.w, xy[k].w, q2.w));
            bs[k][0] += v.x; bs[k][1] += v.y; bs[k][2] += v.z; bs[k][3] += v.w;
          }
          uint4 hi, lo;
          hi.x = __float_as_uint(v.x) & 0xffffe000u; hi.y = __float_as_uint(v.y) & 0xffffe000u;
          hi.z = __float_as_uint(v.z) & 0xffffe000u; hi.w = __float_as_uint(v.w) & 0xffffe000u;
          lo.x = __float_as_uint(v.x - __uint_as_float(hi.x)); lo.y = __float_as_uint(v.y - __uint_as_float(hi.y));
          lo.z = __float_as_uint(v.z - __uint_as_float(hi.z)); lo.w = __float_as_uint(v.w - __uint_as_float(hi.w));
#pragma unroll
          for (int j = 0; j < NTAP; ++j)
            if (NTAP == 1 || z_dst[k][j] != NONE) { *(uint4*)(op + z_dst[k][j]) = hi; *(uint4*)(op + z_dst[k][j] + z_lo) = lo; }
        }
      }
      }
      if (!(K.dbg & 4)) fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&mi.op_full[os]); mbar_arrive(&mi.raw_empty[rs]); }
      if (++rs == nraw) { rs = 0; ph ^= 1; }
    }
    if (K.db) {                                          // (eligibility: one tap, so every dz row is seen exactly once)
#pragma unroll
      for (int k = 0; k < IZ; ++k) {
        if (z_raw[k] != NONE) {
          const int c = z_cb[k] & 0xffff;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float* dst = (K.n_split > 0 && c + e >= K.n_split) ? K.db2 + (c + e - K.n_split) : K.db + c + e;
            atomicAdd(dst, bs[k][e]);
          }
        }
      }
    }
    // ---- epilogue: warps 4-7 add the accumulator to dW -----------------------------------
    if (warp <= 7 && nun > 0) {
      mbar_wait(&mi.done, 0);
      tc_fence_after();
      const int lg = warp & 3;
      const int p = K.Mmma == 128 ? lg * 32 + lane : lg * 16 + lane;
      const bool pok = (K.Mmma == 128 || lane < 16) && p < K.Pvalid;
      for (int cb = 0; cb < NBq; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + cb * 32, v);
        if (K.scratch) {
          // 2-stage reduction: the L2 atomic units were the tail of every launch (P x Q atomics per CTA, all CTAs on the same
          // addresses: ~0.65 ms per training step); partials go to scratch with plain 16-byte stores, wgrad_reduce_kernel adds them up
          if (K.Mmma == 128 || lane < 16) {
            float4* dst = (float4*)(K.scratch + ((size_t)blockIdx.x * K.Mmma + p) * K.QWt + cb * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        } else if (pok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int q = cb * 32 + j;
            if (q < K.Qvalid) atomicAdd(wgrad_dst(K, p, q), __uint_as_float(v[j]));
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, K.tmem_cols);
  }
}

constexpr size_t SMEM_MAX = 227 * 1024;

template <int IA, int IZ, int NTAP, int URT = UR>
int launch_variant(const WgK& K, int grid, size_t smem, cudaStream_t st) {
  TRU_SMEM_OPT_IN((tc_wgrad_stream_kernel<IA, IZ, NTAP, URT>), SMEM_MAX);
  TRU_CUDA(launch_pdl(tc_wgrad_stream_kernel<IA, IZ, NTAP, URT>, dim3(grid), dim3(NT), smem, st, K));
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace

// returns TRU_OK if launched, 1 if the job does not fit this kernel (caller uses the per-job kernels)
int launch_wgrad_stream(const WgStream& w, cudaStream_t st) {
  WgK K{};
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("TRU_WG_DBG"); dbg = e ? atoi(e) : 0; } K.dbg = dbg; }
  if (w.nsrc < 1 || w.nsrc > 2 || w.ntap < 1 || w.ntap > 5 || w.Lq % UR != 0 || w.N % 4 != 0 || w.N > (w.ntap == 1 ? 384 : 128)) return 1;
  // 32-row units wherever two operand stages of 32 rows leave room for >= 3 raw stages and a kernel variant exists: a 16-row unit
  // of e.g. the 128 -> 8 layer is 9 KB, 0.2 us of HBM time, while the transform -> fence -> MMA -> commit hand-off of a unit is
  // ~0.5 us on two operand stages (ablation: 0.30 ms skeleton, 0.51 ms full)
  static int ur32_on = -1;
  if (ur32_on < 0) { const char* e = getenv("TRU_WG_UR32_OFF"); ur32_on = (e && atoi(e)) ? 0 : 1; }
  int ca_all = 0;
  for (int s = 0; s < w.nsrc; ++s) ca_all += w.a_C[s];
  int ur = UR;
  if (ur32_on && w.Lq % 32 == 0 && 31 * w.zs + w.ntap <= 64) {
    const int nz32 = 31 * w.zs + w.ntap;
    const int ia32 = (32 * ((ca_all + 31) / 32) * 8 + NTR - 1) / NTR, iz32 = (((nz32 + 3) & ~3) * ((w.N + 31) / 32) * 8 + NTR - 1) / NTR;
    const int key32 = ia32 * 100 + iz32 * 10 + (w.ntap == 1 ? 1 : (w.ntap <= 3 ? 3 : 5));
    const int have[] = {111, 211, 123, 121, 221, 131};    // (the 32-row instantiations below)
    bool okk = false;
    for (int k : have) okk |= (k == key32);
    // shared memory of the 32-row layout (same formulas as below)
    const bool pz = (w.ntap == 1 && w.N <= 128 && (ca_all > 128 || w.N > ca_all));
    const int Pc = pz ? w.N : ca_all, Qc = pz ? ca_all : w.ntap * w.N;
    const size_t pw = Pc <= 64 ? 64 : 128, qw = (size_t)(Qc + 31) / 32 * 32;
    const size_t op32 = align_up(2 * 32 * pw * 4 + 2 * 32 * qw * 4, 1024);
    const size_t zb = align_up((size_t)nz32 * w.N * 4, 128);
    const size_t raw32 = align_up(align_up((size_t)32 * ca_all * 4, 128) + zb + ((w.z_p0 && w.z_src2) ? (size_t)nz32 * w.N * 4 : 0), 1024);
    const size_t fixed32 = 1024 + 2 * op32 + align_up((size_t)(3 * ca_all + 3 * w.N) * 4, 128) + align_up(sizeof(WMisc2), 128);
    if (okk && Pc <= 128 && Qc <= 384 && fixed32 + 3 * raw32 <= SMEM_MAX) ur = 32;
  }
  K.ur = ur;
  if (w.nsrc == 2 && w.ntap > 1) return 1;
  int Ca = 0;
  for (int s = 0; s < w.nsrc; ++s) {
    if (w.a_C[s] % 4 != 0 || (w.nsrc == 2 && w.a_C[s] % 32 != 0) || w.a_ld[s] % 4 != 0 || w.a_coff[s] % 4 != 0 || !aligned16(w.a_src[s])) return 1;
    K.a_src[s] = w.a_src[s]; K.a_p0[s] = w.a_p0[s]; K.a_p2[s] = w.a_p2[s];
    K.a_L[s] = w.a_L[s]; K.a_ld[s] = w.a_ld[s]; K.a_add[s] = w.a_add[s]; K.a_C[s] = w.a_C[s]; K.a_c0[s] = Ca; K.a_coff[s] = w.a_coff[s];
    K.wbase[s] = w.wbase[s];
    Ca += w.a_C[s];
  }
  if (w.z_ld % 4 != 0 || w.z_coff % 4 != 0 || !aligned16(w.z_src) || (w.z_src2 && !aligned16(w.z_src2))) return 1;
  K.nsrc = w.nsrc; K.Ca = Ca;
  K.z_src = w.z_src; K.z_src2 = w.z_p0 ? w.z_src2 : nullptr; K.z_p0 = w.z_p0; K.z_p1 = w.z_p1; K.z_p2 = w.z_p2;
  K.z_L = w.z_L; K.z_ld = w.z_ld; K.z_coff = w.z_coff; K.N = w.N;
  K.strided = (w.z_ld != w.N ? (4 | (K.z_src2 ? 8 : 0)) : 0);
  for (int s = 0; s < w.nsrc; ++s) if (w.a_ld[s] != w.a_C[s]) K.strided |= 1 << s; K.ntap = w.ntap; K.zs = w.zs; K.zpad = w.zpad;
  K.NZ = (ur - 1) * w.zs + w.ntap;
  if (K.NZ > 64) return 1;
  const int Zcols = w.ntap * w.N;
  K.p_is_z = (w.ntap == 1 && w.N <= 128 && (Ca > 128 || w.N > Ca)) ? 1 : 0;
  const int Pc = K.p_is_z ? w.N : Ca, Qc = K.p_is_z ? Ca : Zcols;
  if (Pc > 128 || Qc > 384) return 1;
  K.Mmma = Pc <= 64 ? 64 : 128; K.PWt = K.Mmma; K.QWt = (Qc + 31) / 32 * 32; K.Pvalid = Pc; K.Qvalid = Qc;
  K.tmem_cols = K.QWt <= 32 ? 32 : K.QWt <= 64 ? 64 : K.QWt <= 128 ? 128 : K.QWt <= 256 ? 256 : 512;
  const int nA = ur * ((Ca + 31) / 32) * 8, nZ = ((K.NZ + 3) & ~3) * ((w.N + 31) / 32) * 8;
  const int ia = (nA + NTR - 1) / NTR, iz = (nZ + NTR - 1) / NTR, nt = w.ntap == 1 ? 1 : (w.ntap <= 3 ? 3 : 5);
  if (ia > 2 || iz > 3) return 1;
  if (w.db && w.ntap != 1) return 1;
  K.db = w.db;
  K.p_tile = (uint32_t)ur * K.PWt * 4; K.q_tile = (uint32_t)ur * K.QWt * 4;
  K.op_stage = (uint32_t)align_up(2 * K.p_tile + 2 * K.q_tile, 1024);
  if (w.z_ld < w.N + w.z_coff) return 1;
  for (int s = 0; s < w.nsrc; ++s) if (w.a_ld[s] < w.a_C[s] + w.a_coff[s]) return 1;
  K.raw_a[0] = 0; K.raw_a[1] = (uint32_t)ur * K.a_C[0] * 4;
  K.raw_dy = (uint32_t)align_up((size_t)ur * Ca * 4, 128);
  K.raw_z = K.raw_dy + (uint32_t)align_up((size_t)K.NZ * w.N * 4, 128);
  K.raw_stage = (uint32_t)align_up(K.raw_z + (K.z_src2 ? (size_t)K.NZ * w.N * 4 : 0), 1024);
  const size_t coefb = align_up((size_t)(3 * Ca + 3 * w.N) * 4, 128), miscb = align_up(sizeof(WMisc2), 128);
  const size_t fixed = 1024 + 2 * (size_t)K.op_stage + coefb + miscb;
  if (fixed + 2 * (size_t)K.raw_stage > SMEM_MAX) return 1;
  K.nraw = (int)std::min<size_t>(MAXRAW, (SMEM_MAX - fixed) / K.raw_stage);
  K.raw_off = 2 * K.op_stage;
  K.coef_off = K.raw_off + (uint32_t)K.nraw * K.raw_stage;
  K.misc_off = K.coef_off + (uint32_t)coefb;
  const size_t smem = 1024 + K.misc_off + miscb;
  K.dW = w.dW; K.wsc = w.wsc; K.wsn = w.wsn; K.wtap = w.wtap;
  K.n_split = (w.dW2 && w.ntap == 1) ? w.n_split : 0; K.dW2 = w.dW2; K.db2 = w.db2;
  if (w.dW2 && (w.ntap != 1 || w.n_split % 4 != 0 || (w.db && !w.db2))) return 1;
  K.BT = w.BT; K.Lq = w.Lq;
  const long M = (long)w.BT * w.Lq;
  if ((double)w.BT * w.z_L * w.z_ld >= 1.8e19) return 1;
  K.units_total = (unsigned)(M / ur);
  const int grid = (int)std::min<long>(sm_count(), K.units_total);
  K.units_per_cta = (K.units_total + grid - 1) / grid;
  const char* nm = "wgrad_stream";
  if (prof_enabled()) {
    static std::map<std::string, const char*> names;
    char buf[128];
    snprintf(buf, sizeof(buf), "wgrad_stream:M=%ld,Ca=%d,N=%d,taps=%d,s=%d%s", M, Ca, w.N, w.ntap, w.zs, K.z_src2 ? ",bnload" : "");
    auto it = names.find(buf);
    if (it == names.end()) it = names.emplace(buf, strdup(buf)).first;
    nm = it->second;
  }
  ProfScope prof(nm, 4.0 * M * (Ca + (double)w.N * w.zs * (K.z_src2 ? 2 : 1)), 2.0 * M * Ca * (double)w.N * w.ntap, st);
  K.scratch = (w.scratch && w.scratch_floats >= (size_t)grid * K.Mmma * K.QWt) ? w.scratch : nullptr;
  const int key = ia * 100 + iz * 10 + nt + (ur == 32 ? 1000 : 0);
  int rc;
  switch (key) {
    case 1111: rc = launch_variant<1, 1, 1, 32>(K, grid, smem, st); break;
    case 1211: rc = launch_variant<2, 1, 1, 32>(K, grid, smem, st); break;
    case 1123: rc = launch_variant<1, 2, 3, 32>(K, grid, smem, st); break;
    case 1121: rc = launch_variant<1, 2, 1, 32>(K, grid, smem, st); break;
    case 1221: rc = launch_variant<2, 2, 1, 32>(K, grid, smem, st); break;
    case 1131: rc = launch_variant<1, 3, 1, 32>(K, grid, smem, st); break;
    case 111: rc = launch_variant<1, 1, 1>(K, grid, smem, st); break;
    case 113: rc = launch_variant<1, 1, 3>(K, grid, smem, st); break;
    case 115: rc = launch_variant<1, 1, 5>(K, grid, smem, st); break;
    case 123: rc = launch_variant<1, 2, 3>(K, grid, smem, st); break;
    case 125: rc = launch_variant<1, 2, 5>(K, grid, smem, st); break;
    case 211: rc = launch_variant<2, 1, 1>(K, grid, smem, st); break;
    case 121: rc = launch_variant<1, 2, 1>(K, grid, smem, st); break;
    case 221: rc = launch_variant<2, 2, 1>(K, grid, smem, st); break;
    case 131: rc = launch_variant<1, 3, 1>(K, grid, smem, st); break;
    default: return set_error(TRU_ERR_ARG, "wgrad_stream: no kernel variant for slots (%d,%d) taps %d", ia, iz, w.ntap);
  }
  if (rc) return rc;
  if (K.scratch) {
    // CTAs past the last unit (units_per_cta is rounded up) have no partial result
    const int nact = (int)((K.units_total + K.units_per_cta - 1) / K.units_per_cta);
    count_launch();
    TRU_CUDA(launch_pdl(wgrad_reduce_kernel, dim3(K.Mmma * (K.QWt >> 5)), dim3(256), 0, st, K, nact));
    TRU_LAUNCH_CHECK();
  }
  return TRU_OK;
}

}  // namespace tru
