// Host emulation of the persistent four-step kernel's arithmetic and queue (fft1_pipe.cuh): every
// "lane" is run in a loop, phase by phase.  Checks (1) the two-pass 32-points-per-lane transform
// with the warp-private (role A) and the row-interleaved (role B) exchange, (2) the four-step
// index mapping through the transposed intermediate Y[n2][k1] against a float64 FFT, (3) that the
// queue order visits every item exactly once and never places a consumer ahead of its producer.
// Prints one line per size: log2N rel-rms-error; then "queue ok".  Driven by tests/test_host_emulation.py.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <complex>
#include "../../linrad_b200/csrc/fft1_pipe.cuh"
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

static float2 tw(long m, long n)
{
  const double a = -2.0 * M_PI * (double)m / (double)n;
  return make_float2((float)cos(a), (float)sin(a));
}

template <int LN1, int LN2>
static int run_case()
{
  using C = PipeCfg<LN1, LN2, 0>;
  constexpr int N1 = C::N1, N2 = C::N2, N = C::N, T1 = C::T1, T2 = C::T2, TB = C::TB, CW = C::CW, TA = C::TA;
  std::vector<float2> x(N), Y(N), X(N);
  srand(1000 + LN1 * 16 + LN2);
  for (int i = 0; i < N; i++) x[i] = make_float2((float)(rand() % 4001 - 2000), (float)(rand() % 4001 - 2000));
  // ---- role A, tile by tile, warp by warp
  std::vector<float2> work(C::WORK_BYTES / 8);
  for (int tile = 0; tile < C::TILES_A; tile++) {
    for (int warp = 0; warp < 8; warp++) {
      std::vector<float2> regs(32 * 32);
      auto V = [&](int lane) -> float2(&)[32] { return *reinterpret_cast<float2(*)[32]>(&regs[(size_t)lane * 32]); };
      for (int lane = 0; lane < 32; lane++) {
        const int t = lane & (T1 - 1), cw = lane >> C::LT1;
        const int n2 = tile * TA + warp * CW + cw;
        for (int e = 0; e < 32; e++) V(lane)[e] = x[(size_t)(t + T1 * e) * N2 + n2];
        pass0<T1>(V(lane));
      }
      std::vector<float2> ureg(32 * 32);
      auto U = [&](int lane) -> float2(&)[32] { return *reinterpret_cast<float2(*)[32]>(&ureg[(size_t)lane * 32]); };
      for (int qq = 0; qq < C::Q1; qq++) {
        for (auto& f : work) f = make_float2(NAN, NAN);
        for (int lane = 0; lane < 32; lane++) {
          const int t = lane & (T1 - 1), cw = lane >> C::LT1;
          float2* area = work.data() + (size_t)warp * (C::AREA_A / 8) + cw * (T1 * C::XP);
          if (((uintptr_t)(area + t * C::XP) - (uintptr_t)work.data()) & 15) { printf("misaligned exchange row\n"); return 1; }
          colx_store<T1>(V(lane), area, t, qq);
        }
        for (int lane = 0; lane < 32; lane++) {
          const int t = lane & (T1 - 1), cw = lane >> C::LT1;
          float2* area = work.data() + (size_t)warp * (C::AREA_A / 8) + cw * (T1 * C::XP);
          colx_load<T1>(U(lane), area, t, qq);
        }
      }
      for (int lane = 0; lane < 32; lane++) {
        const int t = lane & (T1 - 1), cw = lane >> C::LT1;
        const int n2 = tile * TA + warp * CW + cw;
        float2 wb[5];
        for (int j = 0; j < 5; j++) wb[j] = tw((long)t << j, N1);
        radix32_gen(U(lane), wb);
        float2 sb[5];
        for (int j = 0; j < 5; j++) sb[j] = tw((long)(n2 * T1) << j, N);
        apply_power_twiddles<32>(U(lane), tw((long)n2 * t, N), sb);
        for (int e = 0; e < 32; e++) Y[(size_t)n2 * N1 + t + T1 * e] = U(lane)[e];
      }
    }
  }
  // ---- role B
  for (int tile = 0; tile < C::TILES_B; tile++) {
    std::vector<float2> regs(256 * 32), ureg(256 * 32);
    auto V = [&](int tid) -> float2(&)[32] { return *reinterpret_cast<float2(*)[32]>(&regs[(size_t)tid * 32]); };
    auto U = [&](int tid) -> float2(&)[32] { return *reinterpret_cast<float2(*)[32]>(&ureg[(size_t)tid * 32]); };
    // the tile as the TMA load leaves it: in[n2][r]
    std::vector<float2> in((size_t)N2 * TB);
    for (int n2 = 0; n2 < N2; n2++)
      for (int r = 0; r < TB; r++) in[(size_t)n2 * TB + r] = Y[(size_t)n2 * N1 + tile * TB + r];
    for (int tid = 0; tid < 256; tid++) {
      const int r = tid & (TB - 1), t = tid / TB;
      for (int e = 0; e < 32; e++) V(tid)[e] = in[(size_t)(t + T2 * e) * TB + r];
      pass0<T2>(V(tid));
    }
    {
      // all rounds side by side in the (consumed) input buffer, one barrier between stores and loads
      std::vector<float2> buf(C::IN_BYTES / 8, make_float2(NAN, NAN));
      if ((size_t)C::Q2 * C::ROUND_B > buf.size() * 8) { printf("input buffer too small for the exchange\n"); return 1; }
      for (int tid = 0; tid < 256; tid++)
        for (int qq = 0; qq < C::Q2; qq++) rowx_store<T2, TB>(V(tid), buf.data() + qq * (T2 * T2 * TB), tid / TB, tid & (TB - 1), qq);
      for (int tid = 0; tid < 256; tid++)
        for (int qq = 0; qq < C::Q2; qq++) rowx_load<T2, TB>(U(tid), buf.data() + qq * (T2 * T2 * TB), tid / TB, tid & (TB - 1), qq);
    }
    for (int tid = 0; tid < 256; tid++) {
      const int r = tid & (TB - 1), t = tid / TB;
      float2 wb[5];
      for (int j = 0; j < 5; j++) wb[j] = tw((long)t << j, N2);
      radix32_gen(U(tid), wb);
      for (int e = 0; e < 32; e++) X[(size_t)(tile * TB + r) + (size_t)N1 * (t + T2 * e)] = U(tid)[e];
    }
  }
  std::vector<std::complex<double>> ref(N);
  for (int i = 0; i < N; i++) ref[i] = std::complex<double>(x[i].x, x[i].y);
  fft_double(ref);
  double num = 0, den = 0;
  for (int i = 0; i < N; i++) {
    num += std::norm(std::complex<double>(X[i].x, X[i].y) - ref[i]);
    den += std::norm(ref[i]);
  }
  printf("%d %.3e\n", LN1 + LN2, sqrt(num / den));
  return 0;
}

static int check_queue(int nb, int lag, int slots, int IA, int IB)
{
  const int total = nb * (IA + IB);
  std::vector<int> posA((size_t)nb * IA, -1), posB((size_t)nb * IB, -1);
  for (int i = 0; i < total; i++) {
    const PipeItem it = pipe_decode(i, nb, lag, IA, IB);
    if (it.role < 0 || it.b < 0 || it.b >= nb) return 1;
    if (it.role == 0) { if (it.j >= IA || posA[(size_t)it.b * IA + it.j] >= 0) return 2; posA[(size_t)it.b * IA + it.j] = i; }
    else { if (it.j >= IB || posB[(size_t)it.b * IB + it.j] >= 0) return 3; posB[(size_t)it.b * IB + it.j] = i; }
  }
  if (pipe_decode(total, nb, lag, IA, IB).role != -1) return 4;
  for (int b = 0; b < nb; b++) {
    int lastA = -1, firstB = total, lastB = -1;
    for (int j = 0; j < IA; j++) { if (posA[(size_t)b * IA + j] < 0) return 5; if (posA[(size_t)b * IA + j] > lastA) lastA = posA[(size_t)b * IA + j]; }
    for (int j = 0; j < IB; j++) {
      const int v = posB[(size_t)b * IB + j];
      if (v < 0) return 6;
      if (v < firstB) firstB = v;
      if (v > lastB) lastB = v;
    }
    if (lastA >= firstB) return 7;                                  // a row item ahead of one of its columns
    if (b + slots < nb)                                             // the slot's next writer comes after its last reader
      for (int j = 0; j < IA; j++) if (posA[(size_t)(b + slots) * IA + j] <= lastB) return 8;
  }
  return 0;
}

int main()
{
  int rc = 0;
  rc |= run_case<8, 7>();
  rc |= run_case<8, 8>();
  rc |= run_case<9, 8>();
  rc |= run_case<9, 9>();
  rc |= run_case<10, 9>();
  rc |= run_case<10, 10>();
  const int cases[][5] = {{1, 8, 16, 32, 32}, {3, 8, 4, 32, 32}, {60, 8, 16, 32, 32}, {60, 8, 9, 32, 32}, {7, 38, 7, 4, 4},
                          {100, 38, 76, 8, 8}, {5, 1, 2, 64, 128}, {20, 3, 6, 128, 128}, {2, 5, 2, 16, 16}};
  for (auto& c : cases) {
    const int r = check_queue(c[0], c[1], c[2], c[3], c[4]);
    if (r) { printf("queue check %d failed for nb=%d lag=%d slots=%d\n", r, c[0], c[1], c[2]); rc = 1; }
  }
  if (!rc) printf("queue ok\n");
  return rc;
}
