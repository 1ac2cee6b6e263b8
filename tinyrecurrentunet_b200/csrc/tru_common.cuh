// Common host/device helpers for libtru_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/tru_b200.h"

namespace tru {

// ---- error reporting (thread-local, never throws) ---------------------------
char* last_error_buf();
int set_error(int code, const char* fmt, ...);
int ensure_init();                       // tables + arch check, lazily
const float2* twiddle_table();           // device pointer: exp(-2*pi*i*k/2048), k<2048
int sm_count();

#define TRU_CUDA(call)                                                         \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess)                                                    \
      return tru::set_error(-1000 - (int)e__, "%s:%d %s: %s", __FILE__,        \
                            __LINE__, #call, cudaGetErrorString(e__));         \
  } while (0)

#define TRU_LAUNCH_CHECK()                                                     \
  do {                                                                         \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess)                                                    \
      return tru::set_error(-1000 - (int)e__, "%s:%d launch: %s", __FILE__,    \
                            __LINE__, cudaGetErrorString(e__));                \
  } while (0)

#define TRU_REQUIRE(cond, code, ...)                                           \
  do {                                                                         \
    if (!(cond)) return tru::set_error(code, __VA_ARGS__);                     \
  } while (0)

// ---- optional per-kernel event profiler (bench.py roofline; off by default) -----
// prof_begin/prof_end bracket one kernel launch with CUDA events on its stream.
bool prof_enabled();
void prof_begin(const char* name, double bytes, double flops, cudaStream_t st);
void prof_end(cudaStream_t st);
void count_launch();
struct ProfScope {
  cudaStream_t st; bool on;
  ProfScope(const char* name, double bytes, double flops, cudaStream_t s) : st(s), on(prof_enabled()) {
    count_launch();
    if (on) prof_begin(name, bytes, flops, st);
  }
  ~ProfScope() { if (on) prof_end(st); }
};

// ---- programmatic dependent launch (PDL) ---------------------------------------
// The long-running kernels call pdl_trigger() first thing (the next kernel of the stream may then be scheduled onto SMs as
// they free up) and pdl_wait() before they touch anything an earlier kernel of the step produced; what runs before the
// wait (staging the weight slice into shared memory, barrier setup) overlaps the tail of the predecessor.  Only kernels
// that contain a pdl_wait() may be launched through launch_pdl().  Without the launch attribute the device instructions are
// no-ops.  pdl_scope(true) is set by the inference entry points; TRU_PDL in the environment overrides (common.cu).
bool pdl_enabled();
void pdl_scope(bool inference);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Opt a kernel in to more than 48 KB of dynamic shared memory.  Function attributes are per device (context), so the
// "already done" flag is kept per device index, like the library's init state.
#define TRU_SMEM_OPT_IN(kernel, bytes)                                                                        \
  do {                                                                                                        \
    static bool done__[64] = {};                                                                              \
    int dev__ = 0;                                                                                            \
    cudaGetDevice(&dev__);                                                                                    \
    if (dev__ < 0 || dev__ >= 64 || !done__[dev__]) {                                                         \
      TRU_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      if (dev__ >= 0 && dev__ < 64) done__[dev__] = true;                                                     \
    }                                                                                                         \
  } while (0)

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15u) == 0; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

#if defined(__CUDACC__)
// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// reflect index (torch pad_mode="reflect"): valid for -n < k < 2n-1
__device__ __forceinline__ int reflect_idx(int k, int n) {
  k = k < 0 ? -k : k;
  return k >= n ? 2 * (n - 1) - k : k;
}
#endif

}  // namespace tru
