// Host emulation of the register/shared-memory Stockham plan in fft_core.cuh: every
// "thread" is run in a loop, phase by phase, so the index arithmetic and twiddle logic
// can be checked on a machine without a GPU (driven by tests/test_host_emulation.py).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../linrad_b200/csrc/fft_core.cuh"
using namespace lb;

template <int LOG2N, int LOG2E>
double run_case()
{
  using P = Plan<LOG2N, LOG2E>;
  constexpr int N = P::N, E = P::E, T = P::T;
  std::vector<float2> in(N), Wn(N), sm(N + N / 32 + 64);
  std::vector<double> xr(N), xi(N);
  srand(1234 + LOG2N * 7 + LOG2E);
  for (int i = 0; i < N; i++) {
    in[i] = make_float2((float)(rand() % 2001 - 1000), (float)(rand() % 2001 - 1000));
    Wn[i] = make_float2((float)cos(-2.0 * M_PI * i / N), (float)sin(-2.0 * M_PI * i / N));
  }
  std::vector<float2> v((size_t)T * E);
  auto V = [&](int t) -> float2(&)[E] { return *reinterpret_cast<float2(*)[E]>(&v[(size_t)t * E]); };
  for (int t = 0; t < T; t++)
    for (int e = 0; e < E; e++) V(t)[e] = in[t + T * e];
  for (int t = 0; t < T; t++) pass_butterflies<E, P::R0, T>(V(t), nullptr, false);
  for (int p = 1; p < P::NPASS; p++) {
    const int NsPrev = P::ns(p - 1);
    for (int t = 0; t < T; t++) {
      if (p == 1) exchange_store<E, P::R0, T, 5>(V(t), sm.data(), t, NsPrev);
      else exchange_store<E, E, T, 5>(V(t), sm.data(), t, NsPrev);
    }
    for (int t = 0; t < T; t++) exchange_load<E, T, 5>(V(t), sm.data(), t);
    for (int t = 0; t < T; t++) {
      float2 w[8];
      for (int b = 0; b < LOG2E; b++) w[b] = Wn[tw_index<E, E, T>(t, 0, P::ns(p), b)];
      pass_butterflies<E, E, T>(V(t), w, true);
    }
  }
  // reference: O(N^2) in double for small N, else recursive double FFT
  std::vector<double> Xr(N), Xi(N);
  {
    // iterative radix-2 in double
    int lg = LOG2N;
    std::vector<double> ar(N), ai(N);
    for (int i = 0; i < N; i++) {
      int r = 0;
      for (int b = 0; b < lg; b++) if (i & (1 << b)) r |= 1 << (lg - 1 - b);
      ar[r] = in[i].x; ai[r] = in[i].y;
    }
    for (int len = 2; len <= N; len <<= 1)
      for (int s = 0; s < N; s += len)
        for (int k = 0; k < len / 2; k++) {
          double c = cos(-2.0 * M_PI * k / len), sn = sin(-2.0 * M_PI * k / len);
          double ur = ar[s + k], ui = ai[s + k];
          double vr = ar[s + k + len / 2] * c - ai[s + k + len / 2] * sn;
          double vi = ar[s + k + len / 2] * sn + ai[s + k + len / 2] * c;
          ar[s + k] = ur + vr; ai[s + k] = ui + vi;
          ar[s + k + len / 2] = ur - vr; ai[s + k + len / 2] = ui - vi;
        }
    Xr = ar; Xi = ai;
  }
  double num = 0, den = 0;
  for (int t = 0; t < T; t++)
    for (int e = 0; e < E; e++) {
      int k = t + T * e;
      double dr = V(t)[e].x - Xr[k], di = V(t)[e].y - Xi[k];
      num += dr * dr + di * di;
      den += Xr[k] * Xr[k] + Xi[k] * Xi[k];
    }
  double rel = sqrt(num / den);
  printf("N=2^%d E=2^%d passes=%d R0=%d relrms=%.3e\n", LOG2N, LOG2E, P::NPASS, P::R0, rel);
  return rel;
}

int main()
{
  double worst = 0;
#define RUN(a, b) { double r = run_case<a, b>(); if (r > worst) worst = r; }
  RUN(3, 3) RUN(4, 3) RUN(5, 3) RUN(6, 3) RUN(7, 3) RUN(8, 3) RUN(9, 3)
  RUN(4, 4) RUN(5, 4) RUN(6, 4) RUN(7, 4) RUN(8, 4) RUN(9, 4) RUN(10, 4) RUN(11, 4) RUN(12, 4) RUN(13, 4) RUN(14, 4)
  RUN(5, 5) RUN(8, 5) RUN(10, 5) RUN(13, 5) RUN(14, 5) RUN(15, 5)
  printf("WORST %.3e\n", worst);
  return worst < 3e-7 ? 0 : 1;
}
