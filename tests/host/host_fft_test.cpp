// Host-side check of the Stockham butterflies in tru_fft.cuh against a naive
// double-precision DFT.  Built and run by tests/test_host_fft.py (CPU only).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../tinyrecurrentunet_b200/csrc/tru_fft.cuh"
using namespace tru;

template <int N, int DIR>
static double run() {
  std::vector<float> zbuf(2 * TRU_FFT_PAD(N) + 2);
  float* z = zbuf.data() + (((size_t)zbuf.data() & 7) ? 1 : 0);     // 8-byte aligned (the device buffers are)
  std::vector<float2> tw(N);
  std::vector<double> xr(N), xi(N);
  for (int k = 0; k < N; ++k) {
    tw[k].x = (float)cos(-2.0 * M_PI * k / N);
    tw[k].y = (float)sin(-2.0 * M_PI * k / N);
    xr[k] = rand() / (double)RAND_MAX - 0.5; xi[k] = rand() / (double)RAND_MAX - 0.5;
    TRU_FFT_RE(z, k) = (float)xr[k]; TRU_FFT_IM(z, k) = (float)xi[k];
    xr[k] = TRU_FFT_RE(z, k); xi[k] = TRU_FFT_IM(z, k);
  }
  int p = 1;
  for (int s = 0; s < FftPlan<N>::n8; ++s) {
    std::vector<float> ur(N), ui(N);
    for (int t = 0; t < N / 8; ++t) {
      float a[8], b[8];
      fft_butterfly_load<N, 8, DIR>(z, tw.data(), p, t, a, b);
      dft8<DIR>(a, b);
      for (int r = 0; r < 8; ++r) { ur[t * 8 + r] = a[r]; ui[t * 8 + r] = b[r]; }
    }
    for (int t = 0; t < N / 8; ++t) {
      float a[8], b[8];
      for (int r = 0; r < 8; ++r) { a[r] = ur[t * 8 + r]; b[r] = ui[t * 8 + r]; }
      fft_butterfly_store<N, 8>(z, p, t, a, b);
    }
    p *= 8;
  }
  for (int s = 0; s < FftPlan<N>::n4; ++s) {
    std::vector<float> ur(N), ui(N);
    for (int t = 0; t < N / 4; ++t) {
      float a[4], b[4];
      fft_butterfly_load<N, 4, DIR>(z, tw.data(), p, t, a, b);
      dft4<DIR>(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3]);
      for (int r = 0; r < 4; ++r) { ur[t * 4 + r] = a[r]; ui[t * 4 + r] = b[r]; }
    }
    for (int t = 0; t < N / 4; ++t) {
      float a[4], b[4];
      for (int r = 0; r < 4; ++r) { a[r] = ur[t * 4 + r]; b[r] = ui[t * 4 + r]; }
      fft_butterfly_store<N, 4>(z, p, t, a, b);
    }
    p *= 4;
  }
  double maxerr = 0, maxv = 0;
  for (int k = 0; k < N; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      double ang = DIR * 2.0 * M_PI * (double)((long long)k * n % N) / N;
      sr += xr[n] * cos(ang) - xi[n] * sin(ang);
      si += xr[n] * sin(ang) + xi[n] * cos(ang);
    }
    maxerr = fmax(maxerr, fmax(fabs(sr - TRU_FFT_RE(z, k)), fabs(si - TRU_FFT_IM(z, k))));
    maxv = fmax(maxv, fmax(fabs(sr), fabs(si)));
  }
  printf("N=%d DIR=%d maxerr=%.3e maxval=%.3e rel=%.3e\n", N, DIR, maxerr, maxv, maxerr / maxv);
  return maxerr / maxv;
}

int main() {
  double w = 0;
  w = fmax(w, run<512, -1>());  w = fmax(w, run<512, 1>());
  w = fmax(w, run<1024, -1>()); w = fmax(w, run<1024, 1>());
  w = fmax(w, run<2048, -1>()); w = fmax(w, run<2048, 1>());
  if (w > 2e-6) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
