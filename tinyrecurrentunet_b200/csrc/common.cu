// Library-wide state: error string, per-device init (twiddle table, arch check).
#include <mutex>
#include <string.h>
#include "tru_common.cuh"

namespace tru {

static thread_local char g_err[512] = "";
char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

__device__ float2 g_twiddle[2048];

__global__ void init_twiddle_kernel() {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < 2048) {
    double s, c;
    sincospi(-2.0 * (double)k / 2048.0, &s, &c);
    g_twiddle[k] = make_float2((float)c, (float)s);
  }
}

static std::mutex g_mu;
static int g_init_state[64];   // 0 = not yet, 1 = ok, <0 = error code
static int g_sms[64];
static const float2* g_tw_ptr[64];

int ensure_init() {
  int dev = 0;
  TRU_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return set_error(TRU_ERR_ARG, "device index %d unsupported", dev);
  if (g_init_state[dev] == 1) return TRU_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_init_state[dev] == 1) return TRU_OK;
  cudaDeviceProp prop;
  TRU_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return set_error(TRU_ERR_ARCH, "libtru_b200 is built for sm_100a only; device %d is sm_%d%d (no fallback)",
                     dev, prop.major, prop.minor);
  g_sms[dev] = prop.multiProcessorCount;
  init_twiddle_kernel<<<8, 256>>>();
  TRU_LAUNCH_CHECK();
  TRU_CUDA(cudaDeviceSynchronize());
  void* p = nullptr;
  TRU_CUDA(cudaGetSymbolAddress(&p, g_twiddle));
  g_tw_ptr[dev] = (const float2*)p;
  g_init_state[dev] = 1;
  return TRU_OK;
}

const float2* twiddle_table() {
  int dev = 0;
  cudaGetDevice(&dev);
  return g_tw_ptr[dev];
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  return g_sms[dev] > 0 ? g_sms[dev] : 148;
}

}  // namespace tru

extern "C" int tru_abi_version(void) { return TRU_ABI_VERSION; }
extern "C" const char* tru_last_error(void) { return tru::last_error_buf(); }
extern "C" int tru_init(void) { return tru::ensure_init(); }
