// Host emulation of the 32-points-per-thread Stockham plan in fft32_core.cuh: every "thread"
// is run in a loop, phase by phase, so index arithmetic, shared-memory layouts (including the
// 16-byte alignment the 128-bit stores need) and the twiddle logic are checked without a GPU.
// Prints one line per size: log2N  relative-rms-error-vs-double  max-smem-slot.  Driven by
// tests/test_host_emulation.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include "../../linrad_b200/csrc/fft32_core.cuh"
using namespace lb;

static void fft_double(std::vector<std::complex<double>>& a)
{
  const int N = (int)a.size();
  int lg = 0;
  while ((1 << lg) < N) lg++;
  std::vector<std::complex<double>> b(N);
  for (int i = 0; i < N; i++) {
    int r = 0;
    for (int k = 0; k < lg; k++) if (i & (1 << k)) r |= 1 << (lg - 1 - k);
    b[r] = a[i];
  }
  for (int len = 2; len <= N; len <<= 1)
    for (int s = 0; s < N; s += len)
      for (int k = 0; k < len / 2; k++) {
        const std::complex<double> w = std::polar(1.0, -2.0 * M_PI * k / len);
        const std::complex<double> u = b[s + k], v = b[s + k + len / 2] * w;
        b[s + k] = u + v;
        b[s + k + len / 2] = u - v;
      }
  a = b;
}

template <int LOG2N>
static void run_case()
{
  using P = Plan32<LOG2N>;
  constexpr int N = P::N, T = P::T;
  std::vector<float2> in(N), sm(P::XCH + 64, make_float2(NAN, NAN));
  srand(77 + LOG2N);
  for (int i = 0; i < N; i++) in[i] = make_float2((float)(rand() % 4001 - 2000), (float)(rand() % 4001 - 2000));
  std::vector<float2> v((size_t)T * 32);
  auto V = [&](int t) -> float2(&)[32] { return *reinterpret_cast<float2(*)[32]>(&v[(size_t)t * 32]); };
  for (int t = 0; t < T; t++)
    for (int e = 0; e < 32; e++) V(t)[e] = in[t + T * e];
  for (int t = 0; t < T; t++) pass0<P::R0>(V(t));
  // the 128-bit stores need 16-byte aligned slots: every base used must be even
  for (int t = 0; t < T; t++)
    if (pad2<P::SH1>(t * P::R0) & 1) { printf("%d misaligned\n", LOG2N); return; }
  {
    // 128-bit stores are served 8 lanes at a time: the 8 lanes of a quarter warp must hit 8
    // different 16-byte bank groups (slot index / 2 mod 8)
    int worst = 1;
    for (int t0 = 0; t0 < T; t0 += 8) {
      int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int l = 0; l < 8 && t0 + l < T; l++) cnt[(pad2<P::SH1>((t0 + l) * P::R0) / 2) & 7]++;
      for (int g = 0; g < 8; g++) if (cnt[g] > worst) worst = cnt[g];
    }
    if (worst > 1) { printf("%d exch1 store %d-way bank conflict\n", LOG2N, worst); return; }
  }
  for (int t = 0; t < T; t++) exch1_store<LOG2N>(V(t), sm.data(), t);
  for (int t = 0; t < T; t++) exch1_load<LOG2N>(V(t), sm.data(), t);
  if (P::NPASS == 3) {
    // pass 1: Ns = R0, exact table w^r, w = exp(-2 pi i k/(32 R0))
    for (int t = 0; t < T; t++) {
      const int k = t & (P::R0 - 1);
      float2 w[32];
      for (int r = 0; r < 32; r++) {
        const double a = -2.0 * M_PI * (double)(k * r) / (32.0 * P::R0);
        w[r] = make_float2((float)cos(a), (float)sin(a));
      }
      radix32_table(V(t), w);
    }
    for (int i = 0; i < P::XCH + 64; i++) sm[i] = make_float2(NAN, NAN);
    for (int t = 0; t < T; t++) exch2_store<LOG2N>(V(t), sm.data(), t);
    for (int t = 0; t < T; t++) exch2_load<LOG2N>(V(t), sm.data(), t);
  }
  // last pass: Ns = T, k = t, w = exp(-2 pi i t / N); exact binary powers
  for (int t = 0; t < T; t++) {
    float2 wb[5];
    for (int j = 0; j < 5; j++) {
      const double a = -2.0 * M_PI * (double)(t << j) / (double)N;
      wb[j] = make_float2((float)cos(a), (float)sin(a));
    }
    radix32_gen(V(t), wb);
  }
  std::vector<std::complex<double>> ref(N);
  for (int i = 0; i < N; i++) ref[i] = std::complex<double>(in[i].x, in[i].y);
  fft_double(ref);
  double num = 0, den = 0;
  for (int t = 0; t < T; t++)
    for (int e = 0; e < 32; e++) {
      const std::complex<double> g(V(t)[e].x, V(t)[e].y);
      num += std::norm(g - ref[t + T * e]);
      den += std::norm(ref[t + T * e]);
    }
  printf("%d %.3e %d\n", LOG2N, sqrt(num / den), P::XCH);
}

int main()
{
  run_case<10>();
  run_case<11>();
  run_case<12>();
  run_case<13>();
  run_case<14>();
  return 0;
}
