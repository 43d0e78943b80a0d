// Two-pass length-N DFT shared by every kernel of the DC path.
//
// A "line" of N = E*T points is transformed by T cooperating threads, each
// holding E points in registers; CW independent lines sit side by side in the
// lanes of a warp (CW = 32: a warp is one register row of 32 lines, so every
// shared-memory access below is a contiguous, conflict-free 256-byte row).
//
//   n-layout : thread j holds x[j + T*i]               in v[i],  i < E
//   k-layout : thread t holds X[(q*T + t) + E*k2]       in u[q*T + k2]
//
//   halfA : n-layout -> k-layout   (E-point FFT, twiddle W_N^{j*k1}, exchange
//                                   through shared memory, T-point FFTs)
//   halfB : k-layout -> n-layout   (the same steps backwards)
//
// DC = halfA<forward> -> blend in k-layout -> halfB<inverse>.  halfB writes its
// exchange data to exactly the shared-memory slots the same thread read in
// halfA and reads the slots it wrote, so one tile costs two __syncthreads and
// the next tile needs none in between.
//
// Follows the arithmetic of data/reconstruction/deep_med_lib/my_pytorch/
// myfft.py:78-128 (ortho FFT2/iFFT2 as un-normalised transform + scaling).
#pragma once
#include "fft_regs.cuh"

namespace csmri {

constexpr int kTwN = 1024;  // twiddle table: W_1024^m = exp(-2 pi i m/1024)

#ifdef __CUDACC__
__constant__ cf c_twiddle[kTwN];
#endif
#ifndef __CUDA_ARCH__
extern cf h_twiddle[kTwN];  // host mirror (emulation + upload source)
#endif

CSMRI_HD cf tw_lookup(int idx) {
#ifdef __CUDA_ARCH__
  return c_twiddle[idx];
#else
  return h_twiddle[idx];
#endif
}

template <int N, int E, int CW>
struct LineFFT {
  static constexpr int T = N / E;  // threads per line
  static constexpr int Q = E / T;  // T-point sub-FFTs per thread in pass 2
  static_assert(E * T == N, "N must equal E*T");
  static_assert(Q * T == E && Q >= 1, "T must divide E");
  static_assert(kTwN % N == 0, "N must divide the twiddle table size");
  static constexpr int kSmemBytes = N * CW * (int)sizeof(cf);

  // ---- halfA ---------------------------------------------------------------
  template <bool INV>
  static CSMRI_HD void a_front(cf* v, cf* sm, int j, int lane) {
    RegFFT<E, INV>::run(v);
    const int step = j * (kTwN / N);
#pragma unroll
    for (int k1 = 1; k1 < E; ++k1) {
      cf w = tw_lookup(step * k1);
      v[k1] = INV ? cmul_conj(v[k1], w) : cmul(v[k1], w);
    }
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) sm[(k1 * T + j) * CW + lane] = v[k1];
  }
  template <bool INV>
  static CSMRI_HD void a_back(cf* u, const cf* sm, int t, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j2 = 0; j2 < T; ++j2)
        u[q * T + j2] = sm[((q * T + t) * T + j2) * CW + lane];
#pragma unroll
    for (int q = 0; q < Q; ++q) RegFFT<T, INV>::run(u + q * T);
  }
  // index held in u[r] by thread t (k-layout)
  static CSMRI_HD int k_index(int t, int r) {
    return ((r / T) * T + t) + E * (r % T);
  }
  // index held in v[i] by thread j (n-layout)
  static CSMRI_HD int n_index(int j, int i) { return j + T * i; }

  // ---- halfB ---------------------------------------------------------------
  template <bool INV>
  static CSMRI_HD void b_front(cf* u, cf* sm, int t, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q) RegFFT<T, INV>::run(u + q * T);
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j2 = 0; j2 < T; ++j2)
        sm[((q * T + t) * T + j2) * CW + lane] = u[q * T + j2];
  }
  template <bool INV>
  static CSMRI_HD void b_back(cf* v, const cf* sm, int j, int lane) {
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) v[k1] = sm[(k1 * T + j) * CW + lane];
    const int step = j * (kTwN / N);
#pragma unroll
    for (int k1 = 1; k1 < E; ++k1) {
      cf w = tw_lookup(step * k1);
      v[k1] = INV ? cmul_conj(v[k1], w) : cmul(v[k1], w);
    }
    RegFFT<E, INV>::run(v);
  }
};

}  // namespace csmri
