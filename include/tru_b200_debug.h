/* tru_b200_debug.h - test / tuning entry points of libtru_b200.so.
 *
 * NOT part of the drop-in boundary (include/tru_b200.h).  These expose single
 * kernels of the TRU-Net orchestration (csrc/trunet.cu) on plain row-major
 * operands so that tests/test_gpu_kernels.py can check each GEMM-shaped kernel
 * against an fp64 matmul, plus a few process-wide switches used when hunting
 * bottlenecks.  Same conventions as tru_b200.h (device pointers, `stream` is a
 * cudaStream_t, 0 = ok, message via tru_last_error()).
 */
#ifndef TRU_B200_DEBUG_H
#define TRU_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif

/* Pointwise conv (network.py:28-32) y = act(x) W^T + b: x (M,K), w (N,K), out (M,N); act(x) = relu(p0*x + p2) per input
 * channel when p0 != NULL; stats (2N doubles, may be NULL) += column sums / sums of squares of y.
 * use_tc = 1: tcgen05 kernel (csrc/tcgemm.cu), 0: FFMA kernel (csrc/igemm.cu). */
int tru_debug_pw(const float* x, const float* p0, const float* p2, const float* w, const float* bias, float* out,
                 double* stats, int M, int K, int N, int use_tc, void* stream);
/* Data gradient of a pointwise conv with the BatchNorm-backward affine on load and the ReLU mask / BN-backward sums /
 * skip-gradient add in the epilogue: dx (M,N) = mask(zmask) * ((q0*dy + q1*z + q2) (M,K) @ w (K,N) [+ extra]). */
int tru_debug_pw_bwd(const float* dy, const float* z, const float* q0, const float* q1, const float* q2, const float* w,
                     float* dx, const float* zmask, const float* mp0, const float* mp2, const float* bmean,
                     const float* binv, double* bstats, const float* extra, int M, int K, int N, void* stream);
/* Data gradient of ConvTranspose1d (network.py:67-73): dy (BT,Lout,Cout) channels-last, w (Cin,Cout,k) -> dx (BT,L,Cin); with
 * z / q0 / q1 / q2 (q0 != NULL) the gradient is dz = q0*dy + q1*z + q2 per channel, applied on load (BatchNorm backward).
 * shared = 0: one gathered K-segment per tap; 1: the tap-shared launch (one staged tile, row-shifted descriptors per tap). */
int tru_debug_convt_bwd_data(const float* dy, const float* z, const float* q0, const float* q1, const float* q2, const float* w,
                             float* dx, const float* zmask, const float* mp0, const float* mp2, const float* bmean,
                             const float* binv, double* bstats, int BT, int L, int Lout, int Cin, int Cout, int k, int s,
                             int shared, void* stream);   /* zmask / bstats: the ReLU mask and BN-backward sums of tru_debug_pw_bwd */
/* ConvTranspose1d forward: x (BT,L,Cin) channels-last, w (Cin,Cout,k), bias (Cout) -> out (BT,Lout,Cout), stride s, pad s/2;
 * stats (2 Cout doubles, may be NULL) += column sums / sums of squares.  Takes the tap-shared launches where eligible. */
int tru_debug_convt_fwd(const float* x, const float* w, const float* bias, float* out, double* stats, int BT, int L, int Lout,
                        int Cin, int Cout, int k, int s, void* stream);
/* Weight gradient dW (N,C) += z^T a, db (N) += column sums of z, a (M,C), z (M,N): the FFMA kernel (parity reference). */
int tru_debug_wgrad(const float* a, const float* z, float* dw, float* db, int M, int C, int N, void* stream);
/* The streaming tcgen05 weight-gradient kernel (csrc/tcwgrad2.cu): dz = q0*dy + q1*z + q2 (q0 NULL: dz = dy), rows in
 * frames of Lq. */
int tru_debug_wgrad_stream(const float* a, const float* dy, const float* z, const float* q0, const float* q1,
                           const float* q2, float* dw, float* db, int M, int Lq, int C, int N, void* stream);

/* Process-wide switches (tuning aids; defaults: tensor cores on, flags 0). */
int tru_set_tensor_cores(int on);          /* 0: every GEMM-shaped launch takes the FFMA kernels */
int tru_debug_set_flags(int flags);        /* ablation bits of tc_igemm_kernel (results are garbage when set) */
int tru_debug_set_eval_fusion(int on);     /* 0: inference keeps the layer-by-layer schedule (pointwise outputs written; default 1: depthwise conv in the GEMM epilogue) */
int tru_debug_read_mbar(unsigned* out, int n);   /* -DTRU_MBAR_TIMEOUT builds: log of stuck mbarrier waits */

#ifdef __cplusplus
}
#endif
#endif /* TRU_B200_DEBUG_H */
