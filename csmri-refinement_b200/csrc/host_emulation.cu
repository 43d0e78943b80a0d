// Host-side emulation of the LineFFT thread choreography (no GPU needed).
// Runs every "thread" of a tile phase by phase (a phase boundary is where the
// kernel has a __syncthreads) and compares against a naive O(N^2) DFT in
// double precision.  Built and run by tests/test_host_emulation.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "dc_core.cuh"

namespace csmri { cf h_twiddle[kTwN]; }
using namespace csmri;

static void init_tw() {
  for (int m = 0; m < kTwN; ++m) {
    double a = -2.0 * M_PI * m / kTwN;
    h_twiddle[m] = mk((float)cos(a), (float)sin(a));
  }
}

template <int N, int E, int CW = 2>
static double check(bool inv_first) {
  typedef LineFFT<N, E, CW> L;
  constexpr int T = L::T;
  std::vector<cf> sm(L::kSmemBytes / sizeof(cf));
  std::vector<float4> tw4(L::kTwBytes / sizeof(float4));
  cf* tw = reinterpret_cast<cf*>(tw4.data());
  for (int t = 0; t < T; ++t) L::fill_twiddles(tw, t, T);
  std::vector<std::vector<cf>> regs(T * CW, std::vector<cf>(E));
  std::vector<double> xr(N * CW), xi(N * CW);
  srand(N * 131 + E);
  for (auto& a : xr) a = rand() / (double)RAND_MAX - 0.5;
  for (auto& a : xi) a = rand() / (double)RAND_MAX - 0.5;
  // load n-layout
  for (int j = 0; j < T; ++j) for (int l = 0; l < CW; ++l) for (int i = 0; i < E; ++i) {
    int n = L::n_index(j, i);
    regs[j * CW + l][i] = mk((float)xr[n * CW + l], (float)xi[n * CW + l]);
  }
  for (int j = 0; j < T; ++j) for (int l = 0; l < CW; ++l) {
    if (inv_first) L::template a_front<true>(regs[j * CW + l].data(), sm.data(), tw, j, l);
    else L::template a_front<false>(regs[j * CW + l].data(), sm.data(), tw, j, l);
  }
  for (int j = 0; j < T; ++j) for (int l = 0; l < CW; ++l) {
    if (inv_first) L::template a_back<true>(regs[j * CW + l].data(), sm.data(), j, l);
    else L::template a_back<false>(regs[j * CW + l].data(), sm.data(), j, l);
  }
  // compare with the naive DFT
  double num = 0, den = 0;
  double sgn = inv_first ? 1.0 : -1.0;
  for (int l = 0; l < CW; ++l) for (int t = 0; t < T; ++t) for (int r = 0; r < E; ++r) {
    int k = L::k_index(t, r);
    double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      double a = sgn * 2.0 * M_PI * ((long long)n * k % N) / N;
      double c = cos(a), s = sin(a);
      sr += xr[n * CW + l] * c - xi[n * CW + l] * s;
      si += xr[n * CW + l] * s + xi[n * CW + l] * c;
    }
    cf g = regs[t * CW + l][r];
    num += (g.x - sr) * (g.x - sr) + (g.y - si) * (g.y - si);
    den += sr * sr + si * si;
  }
  double e1 = sqrt(num / den);
  // round trip through halfB with the opposite direction
  for (int j = 0; j < T; ++j) for (int l = 0; l < CW; ++l) {
    if (inv_first) L::template b_front<false>(regs[j * CW + l].data(), sm.data(), j, l);
    else L::template b_front<true>(regs[j * CW + l].data(), sm.data(), j, l);
  }
  for (int j = 0; j < T; ++j) for (int l = 0; l < CW; ++l) {
    if (inv_first) L::template b_back<false>(regs[j * CW + l].data(), sm.data(), tw, j, l);
    else L::template b_back<true>(regs[j * CW + l].data(), sm.data(), tw, j, l);
  }
  num = den = 0;
  for (int j = 0; j < T; ++j) for (int l = 0; l < CW; ++l) for (int i = 0; i < E; ++i) {
    int n = L::n_index(j, i);
    cf g = regs[j * CW + l][i];
    double dr = g.x / N - xr[n * CW + l], di = g.y / N - xi[n * CW + l];
    num += dr * dr + di * di;
    den += xr[n * CW + l] * xr[n * CW + l] + xi[n * CW + l] * xi[n * CW + l];
  }
  double e2 = sqrt(num / den);
  printf("N=%d E=%d inv_first=%d dft_rel_l2=%.3e roundtrip_rel_l2=%.3e\n", N, E, (int)inv_first, e1, e2);
  return e1 > e2 ? e1 : e2;
}

// two lines per thread (cf2 / LineFFTV): forward DFT vs naive, then round trip
template <int N, int E>
static double check_v2() {
  constexpr int LW = 2;
  typedef LineFFTV<N, E, LW> L;
  typedef LineFFT<N, E, LW> B;
  constexpr int T = L::T;
  std::vector<float4> sm(N * LW);
  std::vector<float4> tw4(B::kTwBytes / sizeof(float4));
  cf* tw = reinterpret_cast<cf*>(tw4.data());
  for (int t = 0; t < T; ++t) B::fill_twiddles(tw, t, T);
  std::vector<std::vector<cf2>> regs(T * LW, std::vector<cf2>(E));
  const int NL = 2 * LW;   // lines
  std::vector<double> xr(N * NL), xi(N * NL);
  srand(N * 7 + E);
  for (auto& a : xr) a = rand() / (double)RAND_MAX - 0.5;
  for (auto& a : xi) a = rand() / (double)RAND_MAX - 0.5;
  for (int j = 0; j < T; ++j) for (int l = 0; l < LW; ++l) for (int i = 0; i < E; ++i) {
    int n = B::n_index(j, i);
    regs[j * LW + l][i] = mk2(mk((float)xr[n * NL + 2 * l], (float)xr[n * NL + 2 * l + 1]),
                              mk((float)xi[n * NL + 2 * l], (float)xi[n * NL + 2 * l + 1]));
  }
  for (int j = 0; j < T; ++j) for (int l = 0; l < LW; ++l)
    L::template a_front<false>(regs[j * LW + l].data(), sm.data(), tw, j, l);
  for (int j = 0; j < T; ++j) for (int l = 0; l < LW; ++l)
    L::template a_back<false>(regs[j * LW + l].data(), sm.data(), j, l);
  double num = 0, den = 0;
  for (int l = 0; l < LW; ++l) for (int c = 0; c < 2; ++c) for (int t = 0; t < T; ++t)
    for (int r = 0; r < E; ++r) {
      int k = B::k_index(t, r), line = 2 * l + c;
      double sr = 0, si = 0;
      for (int n = 0; n < N; ++n) {
        double a = -2.0 * M_PI * ((long long)n * k % N) / N;
        sr += xr[n * NL + line] * cos(a) - xi[n * NL + line] * sin(a);
        si += xr[n * NL + line] * sin(a) + xi[n * NL + line] * cos(a);
      }
      cf2 g = regs[t * LW + l][r];
      double gr = c ? g.re.y : g.re.x, gi = c ? g.im.y : g.im.x;
      num += (gr - sr) * (gr - sr) + (gi - si) * (gi - si);
      den += sr * sr + si * si;
    }
  double e1 = sqrt(num / den);
  for (int j = 0; j < T; ++j) for (int l = 0; l < LW; ++l)
    L::template b_front<true>(regs[j * LW + l].data(), sm.data(), j, l);
  for (int j = 0; j < T; ++j) for (int l = 0; l < LW; ++l)
    L::template b_back<true>(regs[j * LW + l].data(), sm.data(), tw, j, l);
  num = den = 0;
  for (int j = 0; j < T; ++j) for (int l = 0; l < LW; ++l) for (int i = 0; i < E; ++i)
    for (int c = 0; c < 2; ++c) {
      int n = B::n_index(j, i), line = 2 * l + c;
      cf2 g = regs[j * LW + l][i];
      double dr = (c ? g.re.y : g.re.x) / N - xr[n * NL + line];
      double di = (c ? g.im.y : g.im.x) / N - xi[n * NL + line];
      num += dr * dr + di * di;
      den += xr[n * NL + line] * xr[n * NL + line] + xi[n * NL + line] * xi[n * NL + line];
    }
  double e2 = sqrt(num / den);
  printf("V2 N=%d E=%d dft_rel_l2=%.3e roundtrip_rel_l2=%.3e\n", N, E, e1, e2);
  return e1 > e2 ? e1 : e2;
}

int main() {
  init_tw();
  double worst = 0;
  auto upd = [&](double e) { if (e > worst) worst = e; };
  for (int inv = 0; inv < 2; ++inv) {
    upd(check<32, 8>(inv));
    upd(check<64, 8>(inv));
    upd(check<64, 16>(inv));
    upd(check<128, 16>(inv));
    upd(check<256, 16>(inv));
    upd(check<256, 32>(inv));
    upd(check<512, 32>(inv));
    upd(check<1024, 32>(inv));
    upd(check<320, 40>(inv));      // radix-5 path
    upd(check<320, 40, 8>(inv));
    upd(check<256, 16, 8>(inv));   // padded exchange layout (CW = 8)
    upd(check<128, 16, 8>(inv));
  }
  upd(check_v2<256, 16>());
  upd(check_v2<128, 16>());
  upd(check_v2<512, 32>());
  upd(check_v2<64, 8>());
  upd(check_v2<320, 40>());
  printf("worst=%.3e\n", worst);
  return worst < 2e-6 ? 0 : 1;
}
