// Library-wide state: error string, per-device init (twiddle table, arch check).
#include <stdlib.h>
#include <mutex>
#include <vector>
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

// ---------------------------------------------------------------------------------
struct ProfRec { const char* name; double bytes, flops; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_pool;
static bool g_prof_on = false;

static long long g_launches = 0;
// TRU_PDL: 0 = never, 1 (default) = inference launches only, 2 = training launches too.  Measured on B200: the 4096-stream
// step gains 6.6 % (2.125 -> 1.984 ms: its 21 GEMM launches are short and their weight staging was exposed); the training step
// does not gain (30.3 ms either way) and its end-to-end figure loses 0.8 %, so training launches stay serialised.
static int pdl_env() {
  static const int v = [] { const char* e = getenv("TRU_PDL"); return e ? atoi(e) : 1; }();
  return v;
}
static thread_local bool g_pdl_inference = false;
void pdl_scope(bool inference) { g_pdl_inference = inference; }
bool pdl_enabled() { return pdl_env() >= 2 || (pdl_env() == 1 && g_pdl_inference); }

void count_launch() { ++g_launches; }
bool prof_enabled() { return g_prof_on; }

static cudaEvent_t prof_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(const char* name, double bytes, double flops, cudaStream_t st) {
  ProfRec r{name, bytes, flops, prof_event(), prof_event()};
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) { cudaEventRecord(g_prof.back().e1, st); }


}  // namespace tru

extern "C" int tru_profile_enable(int on) {
  tru::g_prof_on = on != 0;
  return TRU_OK;
}

// Synchronises the device, aggregates the recorded launches per kernel name into
// "name count total_ms total_bytes total_flops\n" lines, clears the records.
extern "C" int tru_profile_report(char* buf, size_t cap) {
  using namespace tru;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return set_error(-1000 - (int)e, "profile_report: %s", cudaGetErrorString(e));
  struct Agg { const char* name; long n; double ms, bytes, flops; };
  std::vector<Agg> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    g_pool.push_back(r.e0); g_pool.push_back(r.e1);
    Agg* a = nullptr;
    for (auto& x : agg) if (x.name == r.name || strcmp(x.name, r.name) == 0) { a = &x; break; }
    if (!a) { agg.push_back(Agg{r.name, 0, 0, 0, 0}); a = &agg.back(); }
    a->n++; a->ms += ms; a->bytes += r.bytes; a->flops += r.flops;
  }
  g_prof.clear();
  size_t off = 0;
  if (buf && cap) buf[0] = 0;
  for (auto& a : agg) {
    int w = snprintf(buf + off, off < cap ? cap - off : 0, "%s %ld %.6f %.0f %.0f\n", a.name, a.n, a.ms, a.bytes, a.flops);
    if (w < 0 || off + (size_t)w >= cap) break;
    off += (size_t)w;
  }
  return TRU_OK;
}

extern "C" long long tru_launch_count(void) { return tru::g_launches; }
extern "C" int tru_abi_version(void) { return TRU_ABI_VERSION; }
extern "C" const char* tru_last_error(void) { return tru::last_error_buf(); }
extern "C" int tru_init(void) { return tru::ensure_init(); }
