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

// Per-size inter-pass twiddle tables W_N^{j*k1} in [T][E] order, for N = 32 ..
// 1024 back to back (N entries each, offset N - 32).  They live in global
// memory (L2-resident, 16 KiB) and each CTA copies its table into shared memory
// with coalesced 8-byte loads.  (The first version gathered them from constant
// memory with per-thread indices: divergent, cold LDC cost ~7 us per CTA -
// found with the per-CTA timeline probe, tools/gpu_trace.py.)
constexpr int kTwLinesTotal = 2048 - 32 + 320;   // ... followed by the table for N = 320
CSMRI_HD constexpr int tw_lines_offset(int n) { return n == 320 ? 2048 - 32 : n - 32; }
#ifdef __CUDACC__
__constant__ cf c_twiddle[kTwN];
__device__ cf g_tw_lines[kTwLinesTotal];
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
  // exchange slot of (k1, j): (k1*TP + j)*CW + lane.  With CW = 8 a warp holds
  // four j (or t) values; one padding slot per k1 row keeps the k-layout
  // accesses (stride TP slots between consecutive t) on disjoint banks.
  static constexpr int TP = T + ((CW == 8 && T % 2 == 0) ? 1 : 0);
  static constexpr int kSmemBytes = E * TP * CW * (int)sizeof(cf);

  // Inter-pass twiddles W_N^{j*k1}.  Register-indexed constant loads (LDC
  // c[3][R]) turned out to be the largest single stall source in the first
  // ncu source pages (low-throughput, high-latency path), so each CTA keeps
  // the N values it needs as a [T][E] table in shared memory and every thread
  // reads its row with E/2 broadcast LDS.128.
  // Row j of the shared-memory table starts at j * kTwPitch.  With LW = 8 (or 16)
  // lanes per line a warp holds 4 (2) different j, and rows exactly E entries
  // (a multiple of 128 bytes for E = 16 / 32) apart put their broadcast
  // LDS.128 on the same banks: a 4-way conflict on every twiddle load (ncu r1:
  // 16 % of the shared-load wavefronts of the adjoint were conflicts).  Two
  // entries (16 bytes) of padding per row spread them over disjoint banks.
  static constexpr int kTwPitch = E + 2;
  static constexpr int kTwBytes = T * kTwPitch * (int)sizeof(cf);
  static CSMRI_HD void fill_twiddles(cf* tw_s, int tid, int nthreads) {
#ifdef __CUDA_ARCH__
    const cf* src = g_tw_lines + tw_lines_offset(N);
    for (int idx = tid; idx < N; idx += nthreads) {
      const int jj = idx / E, k1 = idx - jj * E;
      tw_s[jj * kTwPitch + k1] = __ldg(src + idx);
    }
#else
    for (int idx = tid; idx < N; idx += nthreads) {
      const int jj = idx / E, k1 = idx - jj * E;
      const double a = -2.0 * 3.14159265358979323846 * (double)((jj * k1) % N) / (double)N;
      tw_s[jj * kTwPitch + k1] = mk((float)__builtin_cos(a), (float)__builtin_sin(a));
    }
#endif
  }
  // multiply the k-layout registers of thread t by its E table entries
  static CSMRI_HD void apply_dtab(cf* u, const float* drow_t) {
    const float4* d4 = reinterpret_cast<const float4*>(drow_t);
#pragma unroll
    for (int p = 0; p < E / 4; ++p) {
      const float4 q = d4[p];
      u[4 * p] = cscale(u[4 * p], q.x);
      u[4 * p + 1] = cscale(u[4 * p + 1], q.y);
      u[4 * p + 2] = cscale(u[4 * p + 2], q.z);
      u[4 * p + 3] = cscale(u[4 * p + 3], q.w);
    }
  }
  template <bool INV>
  static CSMRI_HD void apply_twiddles(cf* v, const cf* tw_s, int j) {
    const float4* row = reinterpret_cast<const float4*>(tw_s + j * kTwPitch);
#pragma unroll
    for (int p = 0; p < E / 2; ++p) {
      const float4 q = row[p];
      if (p > 0) v[2 * p] = INV ? cmul_conj(v[2 * p], mk(q.x, q.y)) : cmul(v[2 * p], mk(q.x, q.y));
      v[2 * p + 1] = INV ? cmul_conj(v[2 * p + 1], mk(q.z, q.w)) : cmul(v[2 * p + 1], mk(q.z, q.w));
    }
  }

  // ---- halfA ---------------------------------------------------------------
  template <bool INV>
  static CSMRI_HD void a_front(cf* v, cf* sm, const cf* tw_s, int j, int lane) {
    RegFFT<E, INV>::run(v);
    apply_twiddles<INV>(v, tw_s, j);
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) sm[(k1 * TP + j) * CW + lane] = v[k1];
  }
  template <bool INV>
  static CSMRI_HD void a_back(cf* u, const cf* sm, int t, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j2 = 0; j2 < T; ++j2)
        u[q * T + j2] = sm[((q * T + t) * TP + j2) * CW + lane];
#pragma unroll
    for (int q = 0; q < Q; ++q) RegFFT<T, INV>::run(u + q * T);
  }
  // index held in u[r] by thread t (k-layout)
  static CSMRI_HD int k_index(int t, int r) {
    return ((r / T) * T + t) + E * (r % T);
  }
  // position of D[k] in the per-slice table csmri_dc_prepare writes: thread t
  // finds the E factors of its k-layout registers contiguously at t*E + r
  static CSMRI_HD int dtab_slot(int k) {
    const int k1 = k % E, k2 = k / E;
    return (k1 % T) * E + (k1 / T) * T + k2;
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
        sm[((q * T + t) * TP + j2) * CW + lane] = u[q * T + j2];
  }
  template <bool INV>
  static CSMRI_HD void b_back(cf* v, const cf* sm, const cf* tw_s, int j, int lane) {
    b_back_load(v, sm, j, lane);
    b_back_compute<INV>(v, tw_s, j);
  }
  // the two steps of b_back, for callers that recycle the exchange buffer in between
  static CSMRI_HD void b_back_load(cf* v, const cf* sm, int j, int lane) {
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) v[k1] = sm[(k1 * TP + j) * CW + lane];
  }
  template <bool INV>
  static CSMRI_HD void b_back_compute(cf* v, const cf* tw_s, int j) {
    apply_twiddles<INV>(v, tw_s, j);
    RegFFT<E, INV>::run(v);
  }
};

// ---------------------------------------------------------------------------
// LineFFTV: the same two-pass transform with TWO lines per thread (cf2, SoA).
// LW lanes own 2*LW adjacent lines; exchange slots are float4
// {re.x, re.y, im.x, im.y} moved with STS.128 / LDS.128 (a quarter-warp covers a
// contiguous 128 bytes in both layouts, so no padding is needed).
// ---------------------------------------------------------------------------
template <int N, int E, int LW>
struct LineFFTV {
  typedef LineFFT<N, E, LW> Base;           // index helpers + twiddle table
  static constexpr int T = N / E;
  static constexpr int Q = E / T;
  static constexpr int kSmemBytes = N * LW * (int)sizeof(float4);
  static constexpr int kTwBytes = Base::kTwBytes;

  static CSMRI_HD float4 pack(cf2 a) {
    float4 r; r.x = a.re.x; r.y = a.re.y; r.z = a.im.x; r.w = a.im.y; return r;
  }
  static CSMRI_HD cf2 unpack(float4 a) { return mk2(mk(a.x, a.y), mk(a.z, a.w)); }

  template <bool INV>
  static CSMRI_HD void apply_twiddles(cf2* v, const cf* tw_s, int j) {
    const float4* row = reinterpret_cast<const float4*>(tw_s + j * Base::kTwPitch);
#pragma unroll
    for (int p = 0; p < E / 2; ++p) {
      const float4 q = row[p];
      if (p > 0) v[2 * p] = INV ? cmul_conj(v[2 * p], mk(q.x, q.y)) : cmul(v[2 * p], mk(q.x, q.y));
      v[2 * p + 1] = INV ? cmul_conj(v[2 * p + 1], mk(q.z, q.w)) : cmul(v[2 * p + 1], mk(q.z, q.w));
    }
  }
  static CSMRI_HD void apply_dtab(cf2* u, const float* drow_t) {
    const float4* d4 = reinterpret_cast<const float4*>(drow_t);
#pragma unroll
    for (int p = 0; p < E / 4; ++p) {
      const float4 q = d4[p];
      u[4 * p] = cscale(u[4 * p], q.x);
      u[4 * p + 1] = cscale(u[4 * p + 1], q.y);
      u[4 * p + 2] = cscale(u[4 * p + 2], q.z);
      u[4 * p + 3] = cscale(u[4 * p + 3], q.w);
    }
  }
  template <bool INV>
  static CSMRI_HD void a_front(cf2* v, float4* sm, const cf* tw_s, int j, int lane) {
    RegFFT<E, INV>::run(v);
    apply_twiddles<INV>(v, tw_s, j);
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) sm[(k1 * T + j) * LW + lane] = pack(v[k1]);
  }
  template <bool INV>
  static CSMRI_HD void a_back(cf2* u, const float4* sm, int t, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j2 = 0; j2 < T; ++j2)
        u[q * T + j2] = unpack(sm[((q * T + t) * T + j2) * LW + lane]);
#pragma unroll
    for (int q = 0; q < Q; ++q) RegFFT<T, INV>::run(u + q * T);
  }
  template <bool INV>
  static CSMRI_HD void b_front(cf2* u, float4* sm, int t, int lane) {
#pragma unroll
    for (int q = 0; q < Q; ++q) RegFFT<T, INV>::run(u + q * T);
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j2 = 0; j2 < T; ++j2)
        sm[((q * T + t) * T + j2) * LW + lane] = pack(u[q * T + j2]);
  }
  template <bool INV>
  static CSMRI_HD void b_back(cf2* v, const float4* sm, const cf* tw_s, int j, int lane) {
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) v[k1] = unpack(sm[(k1 * T + j) * LW + lane]);
    apply_twiddles<INV>(v, tw_s, j);
    RegFFT<E, INV>::run(v);
  }
};

}  // namespace csmri
