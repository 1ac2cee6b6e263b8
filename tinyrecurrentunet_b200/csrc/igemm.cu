// Implicit GEMM over channels with gathered rows (fp32 FFMA version).
//
// One kernel covers every dense contraction of the network (network.py:28,50,
// 64,67,83,86,106,109 and nn.GRU's input projection, forward and data-gradient):
// pointwise convs, the skip concat (two K-segments with the pad/crop shift of
// network.py:96-98 folded into the row map), transposed convs (one K-segment per
// tap, parity classes for stride 2) and their adjoints.  The producer's BN+ReLU
// is applied while loading A; the epilogue adds bias, accumulates BN statistics,
// or (backward) adds the skip gradient, applies the ReLU mask and accumulates the
// two BN-backward sums.  Weights for the CTA's 64 output columns stay resident in
// shared memory while the CTA walks over row tiles.
#include <map>
#include <string>
#include <string.h>
#include "net_kernels.cuh"

namespace tru {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;
constexpr int AS_LD = BM + 4, WS_LD = BN + 4;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }

__device__ __forceinline__ float4 load_transform(const Seg& s, long off, int c) {
  float4 v = ld4(s.src + off + c);
  if (s.p0) {
    const float4 a = ld4(s.p0 + s.coff + c), b = ld4(s.p2 + s.coff + c);
    v.x = a.x * v.x + b.x; v.y = a.y * v.y + b.y; v.z = a.z * v.z + b.z; v.w = a.w * v.w + b.w;
    if (s.p1) {
      const float4 z = ld4(s.src2 + off + c), q = ld4(s.p1 + s.coff + c);
      v.x += q.x * z.x; v.y += q.y * z.y; v.z += q.z * z.z; v.w += q.w * z.w;
    }
    if (s.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
  }
  return v;
}

__global__ void __launch_bounds__(NT, 2) igemm_kernel(const __grid_constant__ IgemmParams P, int ktot_pad) {
  extern __shared__ __align__(16) float smem[];
  float* Ws = smem;                                 // [ktot_pad][WS_LD]
  float* As = Ws + (size_t)ktot_pad * WS_LD;        // [BK][AS_LD]
  __shared__ double s_stat[2][BN];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * BN;
  const long M = (long)P.BT * P.Lq;
  const int ntiles = (int)((M + BM - 1) / BM);

  // ---- resident weight slice: Ws[k][n] = W_seg[wbase + c*wsc + (n0+n)*wsn] ----
  {
    int kb = 0;
    for (int s = 0; s < P.nseg; ++s) {
      const Seg& sg = P.seg[s];
      for (int i = tid; i < sg.C * BN; i += NT) {
        int c, n;
        if (sg.wsc == 1) { c = i % sg.C; n = i / sg.C; } else { n = i % BN; c = i / BN; }
        float v = 0.f;
        if (n0 + n < P.N) v = __ldg(sg.W + sg.wbase + (long)c * sg.wsc + (long)(n0 + n) * sg.wsn);
        Ws[(kb + c) * WS_LD + n] = v;
      }
      kb += sg.C;
    }
    for (int i = tid; i < (ktot_pad - kb) * WS_LD; i += NT) Ws[kb * WS_LD + i] = 0.f;
  }
  if (tid < BN) { s_stat[0][tid] = 0.0; s_stat[1][tid] = 0.0; }
  __syncthreads();

  const int kq = tid & 3, r0 = tid >> 2;            // A loader: rows r0, r0+64; k quad kq
  const int ty = tid >> 4, tx = tid & 15;           // compute: rows ty*8.., cols tx*4..
  float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int ncol = n0 + tx * 4;
  const bool col_ok = ncol < P.N;                   // N is a multiple of 4
  if (P.bias && col_ok) bias4 = ld4(P.bias + ncol);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long m0 = (long)tile * BM;
    int lbt[2], lq[2]; bool lok[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const long m = m0 + r0 + 64 * i;
      lok[i] = m < M;
      lbt[i] = lok[i] ? (int)(m / P.Lq) : 0;
      lq[i] = lok[i] ? (int)(m - (long)lbt[i] * P.Lq) : 0;
    }
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    int kb = 0;
    for (int s = 0; s < P.nseg; ++s) {
      const Seg& sg = P.seg[s];
      long off[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int li = lq[i] * sg.smul + sg.sadd;
        off[i] = (lok[i] && li >= 0 && li < sg.Lsrc) ? ((long)lbt[i] * sg.Lsrc + li) * sg.ld + sg.coff : -1;
      }
      float4 pre[2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        pre[i] = (off[i] >= 0 && kq * 4 < sg.C) ? load_transform(sg, off[i], kq * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c0 = 0; c0 < sg.C; c0 += BK) {
        __syncthreads();                             // previous step's readers are done
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float* a = As + (kq * 4) * AS_LD + r0 + 64 * i;
          a[0] = pre[i].x; a[AS_LD] = pre[i].y; a[2 * AS_LD] = pre[i].z; a[3 * AS_LD] = pre[i].w;
        }
        __syncthreads();
        const int cn = c0 + BK + kq * 4;             // prefetch next k-step of this segment
#pragma unroll
        for (int i = 0; i < 2; ++i)
          pre[i] = (c0 + BK < sg.C && off[i] >= 0 && cn < sg.C) ? load_transform(sg, off[i], cn)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
        const float* wrow = Ws + (size_t)(kb + c0) * WS_LD + tx * 4;
        const float* arow = As + ty * 8;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          const float4 a0 = *(const float4*)(arow + k * AS_LD);
          const float4 a1 = *(const float4*)(arow + k * AS_LD + 4);
          const float4 b = *(const float4*)(wrow + k * WS_LD);
          const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(av[i], b.x, acc[i][0]); acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], b.z, acc[i][2]); acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
          }
        }
      }
      kb += sg.C;
    }

    // ---- epilogue -----------------------------------------------------------
    if (col_ok) {
      float4 mp0 = make_float4(1.f, 1.f, 1.f, 1.f), mp2 = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 bmean = mp2, binv = mp2;
      if (P.use_mask && P.mp0) { mp0 = ld4(P.mp0 + ncol); mp2 = ld4(P.mp2 + ncol); }
      if (P.bstats) { bmean = ld4(P.bmean + ncol); binv = ld4(P.binv + ncol); }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long m = m0 + ty * 8 + i;
        if (m >= M) break;
        const int bt = (int)(m / P.Lq), q = (int)(m - (long)bt * P.Lq);
        const int lo = q * P.omul + P.oadd;
        const long row = (long)bt * P.Lout + lo;
        float4 v = make_float4(acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w);
        if (P.extra) {
          const float4 e = ld4(P.extra + row * P.ext_ld + ncol);
          v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        }
        if (P.use_mask) {
          const float4 z = ld4(P.zmask + row * P.ldo + P.ocoff + ncol);
          v.x = (z.x * mp0.x + mp2.x > 0.f) ? v.x : 0.f; v.y = (z.y * mp0.y + mp2.y > 0.f) ? v.y : 0.f;
          v.z = (z.z * mp0.z + mp2.z > 0.f) ? v.z : 0.f; v.w = (z.w * mp0.w + mp2.w > 0.f) ? v.w : 0.f;
          if (P.bstats) {
            st1[0] += v.x; st1[1] += v.y; st1[2] += v.z; st1[3] += v.w;
            st2[0] += v.x * (z.x - bmean.x) * binv.x; st2[1] += v.y * (z.y - bmean.y) * binv.y;
            st2[2] += v.z * (z.z - bmean.z) * binv.z; st2[3] += v.w * (z.w - bmean.w) * binv.w;
          }
        }
        if (P.stats) {
          st1[0] += v.x; st1[1] += v.y; st1[2] += v.z; st1[3] += v.w;
          st2[0] += v.x * v.x; st2[1] += v.y * v.y; st2[2] += v.z * v.z; st2[3] += v.w * v.w;
        }
        if (P.planar) {
          float* o = P.out + ((long)bt * P.N + ncol) * P.Lout + lo;
          o[0] = v.x; o[P.Lout] = v.y; o[2 * (long)P.Lout] = v.z; o[3 * (long)P.Lout] = v.w;
        } else {
          *(float4*)(P.out + row * P.ldo + P.ocoff + ncol) = v;
        }
      }
    }
  }

  double* gst = P.stats ? P.stats : P.bstats;
  if (gst) {
    if (col_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&s_stat[0][tx * 4 + j], (double)st1[j]);
        atomicAdd(&s_stat[1][tx * 4 + j], (double)st2[j]);
      }
    }
    __syncthreads();
    if (tid < BN && n0 + tid < P.N) {
      atomicAdd(gst + n0 + tid, s_stat[0][tid]);
      atomicAdd(gst + P.N + n0 + tid, s_stat[1][tid]);
    }
  }
}

// ---- weight gradient: dW[c][n] += sum_m a(m,c) * dz(m,n) ---------------------------
constexpr int WT = 64, WR = 16;     // tile 64 (c) x 64 (n), 16 rows per step

__global__ void __launch_bounds__(NT) wgrad_kernel(const __grid_constant__ WgradParams P, int rows_per_cta) {
  __shared__ __align__(16) float As[WR][WT + 4];
  __shared__ __align__(16) float Zs[WR][WT + 4];
  const WgradJob& J = P.job[blockIdx.z];
  const int ctiles = (J.C + WT - 1) / WT, ntl = (J.N + WT - 1) / WT;
  if ((int)blockIdx.y >= ctiles * ntl) return;
  const int c0 = (blockIdx.y % ctiles) * WT, n0 = (blockIdx.y / ctiles) * WT;
  const int tid = threadIdx.x;
  const int lr = tid >> 4, l4 = (tid & 15) * 4;      // loader: row lr, 4 channels at l4
  const int ty = tid >> 4, tx = tid & 15;            // compute: c = c0 + ty*4.., n = n0 + tx*4..
  const long M = (long)P.BT * P.Lq;
  const long mbeg = (long)blockIdx.x * rows_per_cta;
  const long mend = min(M, mbeg + rows_per_cta);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool a_ok = c0 + l4 < J.C, z_ok = n0 + l4 < J.N;
  float4 ap0 = make_float4(1, 1, 1, 1), ap2 = make_float4(0, 0, 0, 0);
  float4 zp0 = ap0, zp1 = ap2, zp2 = ap2;
  if (J.a_p0 && a_ok) { ap0 = ld4(J.a_p0 + J.a_coff + c0 + l4); ap2 = ld4(J.a_p2 + J.a_coff + c0 + l4); }
  if (J.z_p0 && z_ok) {
    zp0 = ld4(J.z_p0 + J.z_coff + n0 + l4); zp2 = ld4(J.z_p2 + J.z_coff + n0 + l4);
    if (J.z_p1) zp1 = ld4(J.z_p1 + J.z_coff + n0 + l4);
  }

  for (long mb = mbeg; mb < mend; mb += WR) {
    const long m = mb + lr;
    float4 a = make_float4(0, 0, 0, 0), z = a;
    if (m < mend) {
      const int bt = (int)(m / P.Lq), q = (int)(m - (long)bt * P.Lq);
      const int la = q * J.a_mul + J.a_add, lz = q * J.z_mul + J.z_add;
      const bool va = !J.a_src || (la >= 0 && la < J.a_L);       // a_src == null: bias-only job
      if (va && lz >= 0 && lz < J.z_L) {
        if (a_ok && J.a_src) {
          a = ld4(J.a_src + ((long)bt * J.a_L + la) * J.a_ld + J.a_coff + c0 + l4);
          if (J.a_p0) {
            a.x = ap0.x * a.x + ap2.x; a.y = ap0.y * a.y + ap2.y; a.z = ap0.z * a.z + ap2.z; a.w = ap0.w * a.w + ap2.w;
            if (J.a_relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
          }
        }
        if (z_ok) {
          const long zo = ((long)bt * J.z_L + lz) * J.z_ld + J.z_coff + n0 + l4;
          z = ld4(J.z_src + zo);
          if (J.z_p0) {
            z.x = zp0.x * z.x + zp2.x; z.y = zp0.y * z.y + zp2.y; z.z = zp0.z * z.z + zp2.z; z.w = zp0.w * z.w + zp2.w;
            if (J.z_p1) {
              const float4 zz = ld4(J.z_src2 + zo);
              z.x += zp1.x * zz.x; z.y += zp1.y * zz.y; z.z += zp1.z * zz.z; z.w += zp1.w * zz.w;
            }
          }
        }
      }
    }
    __syncthreads();
    *(float4*)&As[lr][l4] = a;
    *(float4*)&Zs[lr][l4] = z;
    __syncthreads();
    bsum[0] += z.x; bsum[1] += z.y; bsum[2] += z.z; bsum[3] += z.w;
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      const float4 av = *(const float4*)&As[r][ty * 4];
      const float4 zv = *(const float4*)&Zs[r][tx * 4];
      const float aa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(aa[i], zv.x, acc[i][0]); acc[i][1] = fmaf(aa[i], zv.y, acc[i][1]);
        acc[i][2] = fmaf(aa[i], zv.z, acc[i][2]); acc[i][3] = fmaf(aa[i], zv.w, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    if (c >= J.C || !J.a_src) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < J.N) atomicAdd(J.dW + J.wbase + (long)c * J.wsc + (long)n * J.wsn, acc[i][j]);
    }
  }
  // bias gradient: column sums of dz, taken by the c-tile 0 CTAs.  NOTE: a row whose
  // A-side source is out of range is skipped above, so jobs that carry db must have
  // an identity A map or no A operand at all (a_src == null, C = 4: bias-only job).
  if (J.db && c0 == 0) {                            // block-uniform condition
    __syncthreads();
    float* red = &As[0][0];                          // reuse: [16 rows][64]
    red[lr * (WT + 4) + l4 + 0] = bsum[0]; red[lr * (WT + 4) + l4 + 1] = bsum[1];
    red[lr * (WT + 4) + l4 + 2] = bsum[2]; red[lr * (WT + 4) + l4 + 3] = bsum[3];
    __syncthreads();
    if (tid < WT && n0 + tid < J.N) {
      float s = 0.f;
      for (int r = 0; r < WR; ++r) s += red[r * (WT + 4) + tid];
      atomicAdd(J.db + n0 + tid, s);
    }
  }
}

// ---- weight gradient for tiny channel counts (C*N <= 1024, e.g. the 8-channel output block) ----
// One thread per output element (up to 4 per thread); 32-row chunks staged in shared memory.
constexpr int SW_ROWS = 32;
__global__ void __launch_bounds__(256) wgrad_small_kernel(const __grid_constant__ WgradParams P) {
  __shared__ float As[SW_ROWS][64 + 1];
  __shared__ float Zs[SW_ROWS][64 + 1];
  const WgradJob& J = P.job[blockIdx.y];
  const int tid = threadIdx.x;
  const int nout = J.C * J.N;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int oc[4], on[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int o = tid + 256 * k; oc[k] = o % J.C; on[k] = o / J.C; }
  const long M = (long)P.BT * P.Lq;
  for (long mb = (long)blockIdx.x * SW_ROWS; mb < M; mb += (long)gridDim.x * SW_ROWS) {
    __syncthreads();
    for (int i = tid; i < SW_ROWS * J.C; i += 256) {
      const int r = i / J.C, c = i % J.C;
      const long m = mb + r;
      float v = 0.f;
      if (m < M) {
        const int bt = (int)(m / P.Lq), q = (int)(m - (long)bt * P.Lq);
        const int la = q * J.a_mul + J.a_add;
        if (la >= 0 && la < J.a_L) {
          v = __ldg(J.a_src + ((long)bt * J.a_L + la) * J.a_ld + J.a_coff + c);
          if (J.a_p0) {
            v = __ldg(J.a_p0 + J.a_coff + c) * v + __ldg(J.a_p2 + J.a_coff + c);
            if (J.a_relu) v = fmaxf(v, 0.f);
          }
        }
      }
      As[r][c] = v;
    }
    for (int i = tid; i < SW_ROWS * J.N; i += 256) {
      const int r = i / J.N, n = i % J.N;
      const long m = mb + r;
      float v = 0.f;
      if (m < M) {
        const int bt = (int)(m / P.Lq), q = (int)(m - (long)bt * P.Lq);
        const int lz = q * J.z_mul + J.z_add;
        if (lz >= 0 && lz < J.z_L) {
          const long zo = ((long)bt * J.z_L + lz) * J.z_ld + J.z_coff + n;
          v = __ldg(J.z_src + zo);
          if (J.z_p0) {
            v = __ldg(J.z_p0 + J.z_coff + n) * v + __ldg(J.z_p2 + J.z_coff + n);
            if (J.z_p1) v += __ldg(J.z_p1 + J.z_coff + n) * __ldg(J.z_src2 + zo);
          }
        }
      }
      Zs[r][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (tid + 256 * k < nout) {
        float a = acc[k];
#pragma unroll 8
        for (int r = 0; r < SW_ROWS; ++r) a = fmaf(As[r][oc[k]], Zs[r][on[k]], a);
        acc[k] = a;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (tid + 256 * k < nout) atomicAdd(J.dW + J.wbase + (long)oc[k] * J.wsc + (long)on[k] * J.wsn, acc[k]);
}

}  // namespace

static bool small_wgrad_job(const WgradJob& J) {
  return J.a_src && !J.db && J.C <= 64 && J.N <= 64 && J.C * J.N <= 1024 && (J.C < 32 || J.N < 32);
}
static int launch_wgrad_small(const WgradParams& p, cudaStream_t st) {
  const long M = (long)p.BT * p.Lq;
  double bytes = 0, flops = 0;
  if (prof_enabled())
    for (int j = 0; j < p.njobs; ++j) { bytes += 4.0 * M * (p.job[j].C + p.job[j].N * (p.job[j].z_src2 ? 2 : 1)); flops += 2.0 * M * p.job[j].C * p.job[j].N; }
  ProfScope prof("wgrad_small", bytes, flops, st);
  dim3 grid((unsigned)std::min<long>((M + SW_ROWS - 1) / SW_ROWS, (long)sm_count() * 8 / p.njobs + 1), p.njobs);
  wgrad_small_kernel<<<grid, 256, 0, st>>>(p);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

static bool g_tc_on = true;
void set_tc_enabled(bool on) { g_tc_on = on; }
bool tc_enabled() { return g_tc_on; }

static void igemm_cost(const IgemmParams& p, double& bytes, double& flops) {
  const long M = (long)p.BT * p.Lq;
  int ktot = 0;
  bytes = 0;
  const float* seen[5]; int ns = 0;
  for (int s = 0; s < p.nseg; ++s) {
    ktot += p.seg[s].C;
    bool dup = false;
    for (int t = 0; t < ns; ++t) dup |= (seen[t] == p.seg[s].src + p.seg[s].coff);
    if (!dup) {
      seen[ns++] = p.seg[s].src + p.seg[s].coff;
      bytes += 4.0 * p.BT * p.seg[s].Lsrc * p.seg[s].C * (p.seg[s].src2 ? 2 : 1) * (p.src_frac > 0 ? p.src_frac : 1.0f);
    }
    bytes += 4.0 * p.seg[s].C * p.N;
  }
  const double Mout = p.ntap && p.Lvalid > 0 ? (double)p.BT * p.Lvalid : (double)M;      // tap-shared: virtual rows produce no output
  if (p.ntap) { ktot = 0; for (int j = 0; j < p.ntap; ++j) ktot += p.tap_C[j]; bytes += 4.0 * ktot * p.N; }
  if (p.dw_k) bytes += 4.0 * p.BT * p.Lout * p.N;        // depthwise epilogue: only the depthwise output is written
  else bytes += 4.0 * Mout * p.N * (1 + (p.use_mask ? 1 : 0) + (p.extra ? 1 : 0));
  flops = 2.0 * Mout * ktot * p.N;
}

// interned "<kind>:M=..,K=..,N=..,seg=..[,bwd]" strings for the profiler (bench.py groups by prefix)
static const char* shape_name(const char* kind, const IgemmParams& p) {
  static std::map<std::string, const char*> names;
  int ktot = 0;
  for (int s = 0; s < p.nseg; ++s) ktot += p.seg[s].C;
  char buf[128];
  if (p.ntap) { ktot = 0; for (int j = 0; j < p.ntap; ++j) ktot += p.tap_C[j]; }
  snprintf(buf, sizeof(buf), "%s:M=%ld,K=%d,N=%d,%s=%d%s%s%s", kind, (long)p.BT * p.Lq, ktot, p.N, p.ntap ? "taps" : "seg",
           p.ntap ? p.ntap : p.nseg, p.seg[0].src2 ? ",bnload" : "", p.use_mask ? ",mask" : "", p.dw_k ? ",dw" : "");
  auto it = names.find(buf);
  if (it == names.end()) it = names.emplace(buf, strdup(buf)).first;
  return it->second;
}

int launch_igemm(const IgemmParams& p, cudaStream_t st) {
  if (g_tc_on && igemm_tc_eligible(p)) {
    double bytes = 0, flops = 0;
    if (prof_enabled()) igemm_cost(p, bytes, flops);
    ProfScope prof(prof_enabled() ? shape_name("igemm_tc", p) : "igemm_tc", bytes, flops, st);
    return launch_igemm_tc(p, st);
  }
  return launch_igemm_simt(p, st);
}

int launch_igemm_simt(const IgemmParams& p, cudaStream_t st) {
  TRU_REQUIRE(p.ntap == 0, TRU_ERR_ARG, "igemm: tap-shared launches exist on the tensor-core kernel only (the caller checks igemm_tc_eligible)");
  TRU_REQUIRE(p.dw_k == 0, TRU_ERR_ARG, "igemm: the depthwise epilogue exists on the tensor-core kernel only (the caller checks igemm_tc_eligible)");
  TRU_REQUIRE(p.nseg >= 1 && p.nseg <= 5 && p.N % 4 == 0 && p.BT > 0 && p.Lq > 0, TRU_ERR_ARG, "igemm: bad params");
  int ktot = 0;
  for (int s = 0; s < p.nseg; ++s) {
    TRU_REQUIRE(p.seg[s].C % 4 == 0 && p.seg[s].ld % 4 == 0 && p.seg[s].coff % 4 == 0, TRU_ERR_ARG,
                "igemm: channel counts must be multiples of 4");
    ktot += p.seg[s].C;
  }
  const int ktot_pad = ktot + BK;
  const size_t smem = ((size_t)ktot_pad * WS_LD + (size_t)BK * AS_LD) * sizeof(float);
  TRU_REQUIRE(smem <= 200 * 1024, TRU_ERR_ARG, "igemm: K too large (%d)", ktot);
  TRU_SMEM_OPT_IN(igemm_kernel, 200 * 1024);
  const long M = (long)p.BT * p.Lq;
  const int ntiles = (int)((M + BM - 1) / BM);
  const int per_sm = smem > 100 * 1024 ? 1 : 2;
  dim3 grid(std::min(ntiles, sm_count() * per_sm), (p.N + BN - 1) / BN);
  double bytes = 0, flops = 0;
  if (prof_enabled()) igemm_cost(p, bytes, flops);
  ProfScope prof("igemm", bytes, flops, st);
  igemm_kernel<<<grid, NT, smem, st>>>(p, ktot_pad);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

static void wgrad_cost(const WgradParams& p, double& bytes, double& flops) {
  const long M = (long)p.BT * p.Lq;
  bytes = 0; flops = 0;
  for (int j = 0; j < p.njobs; ++j) {
    const WgradJob& J = p.job[j];
    bytes += 4.0 * M * ((J.a_src ? J.C : 0) + J.N * (J.z_src2 ? 2 : 1)) + 4.0 * (J.a_src ? J.C : 0) * J.N;
    flops += J.a_src ? 2.0 * M * J.C * J.N : 0.0;
  }
}

// Weight gradients of the shapes the streaming tensor-core kernel (tcwgrad2.cu) does not take: tiny products on the
// small-shape kernel, bias-only jobs as column sums, everything else on the FFMA kernel.
int launch_wgrad(const WgradParams& p, cudaStream_t st) {
  WgradParams small{}, left{};
  small.BT = left.BT = p.BT; small.Lq = left.Lq = p.Lq;
  for (int j = 0; j < p.njobs; ++j) {
    const WgradJob& J = p.job[j];
    if (tc_enabled() && small_wgrad_job(J)) { small.job[small.njobs++] = J; continue; }
    if (!J.a_src && J.db && J.z_mul == 1 && J.z_add == 0 && J.z_L == p.Lq) {
      const int rc = launch_colsum(J.z_src, J.z_src2, J.z_p0, J.z_p1, J.z_p2, J.db, (long)p.BT * p.Lq, J.z_ld, J.z_coff, J.N, st);
      if (rc) return rc;
      continue;
    }
    left.job[left.njobs++] = J;
  }
  if (small.njobs) {
    const int rc = launch_wgrad_small(small, st);
    if (rc) return rc;
  }
  if (left.njobs) return launch_wgrad_simt(left, st);
  return TRU_OK;
}

int launch_wgrad_simt(const WgradParams& p, cudaStream_t st) {
  TRU_REQUIRE(p.njobs >= 1 && p.njobs <= 8 && p.BT > 0 && p.Lq > 0, TRU_ERR_ARG, "wgrad: bad params");
  int maxt = 1;
  for (int j = 0; j < p.njobs; ++j) {
    const WgradJob& J = p.job[j];
    TRU_REQUIRE(J.C % 4 == 0 && J.N % 4 == 0 && J.a_ld % 4 == 0 && J.z_ld % 4 == 0 && J.a_coff % 4 == 0 &&
                J.z_coff % 4 == 0, TRU_ERR_ARG, "wgrad: channel counts must be multiples of 4");
    maxt = std::max(maxt, ((J.C + WT - 1) / WT) * ((J.N + WT - 1) / WT));
  }
  const long M = (long)p.BT * p.Lq;
  // enough CTAs for ~2 waves, at least 256 rows each
  long target = std::max<long>(1, (2L * sm_count() * 4) / ((long)maxt * p.njobs));
  long rows = std::max<long>(256, (M + target - 1) / target);
  rows = (rows + WR - 1) / WR * WR;
  dim3 grid((unsigned)((M + rows - 1) / rows), maxt, p.njobs);
  double bytes = 0, flops = 0;
  if (prof_enabled()) wgrad_cost(p, bytes, flops);
  ProfScope prof("wgrad", bytes, flops, st);
  wgrad_kernel<<<grid, NT, 0, st>>>(p, (int)rows);
  TRU_LAUNCH_CHECK();
  return TRU_OK;
}

}  // namespace tru
