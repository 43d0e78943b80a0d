// In-register complex FFT building blocks for the DC kernels (sm_100a).
//
// A complex number is a float2 {re, im}.  On the device every complex add,
// subtract, +-i rotation and multiply is issued as packed FADD2 / FMUL2 /
// FFMA2 (PTX add/mul/fma.rn.f32x2, new on sm_100): one instruction works on
// the (re, im) pair, and the per-lane swap / negate the butterflies need are
// free operand modifiers in SASS (.LO_HI, .NP).  The same templates compile as
// plain scalar C++ on the host, which is how tests/test_host_emulation.py
// checks the index logic without a GPU.
//
// Sizes 2, 4, 8, 16, 32 are fully unrolled radix-4/2 Cooley-Tukey with
// compile-time twiddles, natural order in and out.  Sign convention is
// numpy's (forward e^{-2 pi i nk/N}), the one the reference pins at
// data/reconstruction/deep_med_lib/my_pytorch/myfft.py:225,241-242.
#pragma once
#include <cuda_runtime.h>

#define CSMRI_HD __host__ __device__ __forceinline__

namespace csmri {

typedef float2 cf;

#if defined(__CUDA_ARCH__) && !defined(CSMRI_SCALAR_MATH)
#define CSMRI_PACKED 1
#else
#define CSMRI_PACKED 0
#endif

CSMRI_HD cf mk(float re, float im) { cf r; r.x = re; r.y = im; return r; }

#if CSMRI_PACKED
__device__ __forceinline__ cf f2add(cf a, cf b) {
  cf r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};"
      " add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ cf f2sub(cf a, cf b) {
  cf r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};"
      " sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ cf f2mul(cf a, cf b) {
  cf r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};"
      " mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ cf f2fma(cf a, cf b, cf c) {
  cf r;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};"
      " mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
#else
CSMRI_HD cf f2add(cf a, cf b) { return mk(a.x + b.x, a.y + b.y); }
CSMRI_HD cf f2sub(cf a, cf b) { return mk(a.x - b.x, a.y - b.y); }
CSMRI_HD cf f2mul(cf a, cf b) { return mk(a.x * b.x, a.y * b.y); }
CSMRI_HD cf f2fma(cf a, cf b, cf c) {
#ifdef __CUDA_ARCH__
  return mk(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#else
  return mk(__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y));
#endif
}
#endif

CSMRI_HD cf cadd(cf a, cf b) { return f2add(a, b); }
CSMRI_HD cf csub(cf a, cf b) { return f2sub(a, b); }
// a - i*b  and  a + i*b
CSMRI_HD cf add_mi(cf a, cf b) { return f2fma(mk(b.y, -b.x), mk(1.f, 1.f), a); }
CSMRI_HD cf add_pi(cf a, cf b) { return f2fma(mk(-b.y, b.x), mk(1.f, 1.f), a); }
// a * w  (two packed instructions: FMUL2 + FFMA2)
CSMRI_HD cf cmul(cf a, cf w) {
  cf t = f2mul(mk(a.y, a.y), mk(-w.y, w.x));
  return f2fma(mk(a.x, a.x), w, t);
}
// a * conj(w)
CSMRI_HD cf cmul_conj(cf a, cf w) {
  cf t = f2mul(mk(a.y, a.y), mk(w.y, w.x));
  return f2fma(mk(a.x, a.x), mk(w.x, -w.y), t);
}
CSMRI_HD cf cscale(cf a, float s) { return f2mul(a, mk(s, s)); }

// ---------------------------------------------------------------------------
// cf2: the same complex arithmetic on TWO independent lines at once, stored
// structure-of-arrays {re.x, re.y} / {im.x, im.y}.  Every operation is still one
// packed instruction per register pair, multiplications by a (shared) twiddle
// broadcast the scalar, and +-i rotations are pure register renaming.  A thread
// that owns two adjacent image columns moves them with 64-bit global / 128-bit
// shared accesses, halving the memory-instruction count per element.
// ---------------------------------------------------------------------------
struct cf2 {
  float2 re, im;
};
CSMRI_HD cf2 mk2(float2 re, float2 im) { cf2 r; r.re = re; r.im = im; return r; }
CSMRI_HD float2 f2neg(float2 a) { return mk(-a.x, -a.y); }
CSMRI_HD cf2 cadd(cf2 a, cf2 b) { return mk2(f2add(a.re, b.re), f2add(a.im, b.im)); }
CSMRI_HD cf2 csub(cf2 a, cf2 b) { return mk2(f2sub(a.re, b.re), f2sub(a.im, b.im)); }
CSMRI_HD cf2 add_mi(cf2 a, cf2 b) { return mk2(f2add(a.re, b.im), f2sub(a.im, b.re)); }
CSMRI_HD cf2 add_pi(cf2 a, cf2 b) { return mk2(f2sub(a.re, b.im), f2add(a.im, b.re)); }
CSMRI_HD cf2 cmul(cf2 a, cf w) {
  const float2 c = mk(w.x, w.x), s = mk(w.y, w.y);
  return mk2(f2fma(a.re, c, f2neg(f2mul(a.im, s))), f2fma(a.im, c, f2mul(a.re, s)));
}
CSMRI_HD cf2 cmul_conj(cf2 a, cf w) {
  const float2 c = mk(w.x, w.x), s = mk(w.y, w.y);
  return mk2(f2fma(a.re, c, f2mul(a.im, s)), f2fma(a.im, c, f2neg(f2mul(a.re, s))));
}
CSMRI_HD cf2 cscale(cf2 a, float s) { return mk2(f2mul(a.re, mk(s, s)), f2mul(a.im, mk(s, s))); }
CSMRI_HD cf2 rot_mi(cf2 a) { return mk2(a.im, f2neg(a.re)); }   // a * (-i)
CSMRI_HD cf2 rot_pi(cf2 a) { return mk2(f2neg(a.im), a.re); }   // a * (+i)
CSMRI_HD cf2 cneg(cf2 a) { return mk2(f2neg(a.re), f2neg(a.im)); }
CSMRI_HD cf cmadd(cf acc, cf a, float s) { return f2fma(a, mk(s, s), acc); }
CSMRI_HD cf2 cmadd(cf2 acc, cf2 a, float s) {
  return mk2(f2fma(a.re, mk(s, s), acc.re), f2fma(a.im, mk(s, s), acc.im));
}
CSMRI_HD cf rot_mi(cf a) { return mk(a.y, -a.x); }
CSMRI_HD cf rot_pi(cf a) { return mk(-a.y, a.x); }
CSMRI_HD cf cneg(cf a) { return mk(-a.x, -a.y); }

// ---------------------------------------------------------------------------
// compile-time twiddles: cos/sin(2 pi k / n) evaluated in double by Taylor
// series after octant reduction (error < 2e-15, i.e. correctly rounded floats)
// ---------------------------------------------------------------------------
__host__ __device__ constexpr double ct_pi() { return 3.14159265358979323846264338327950288; }
__host__ __device__ constexpr double ct_sin(double x) {
  double x2 = x * x, t = x, s = x;
  for (int i = 1; i < 14; ++i) { t *= -x2 / ((2 * i) * (2 * i + 1)); s += t; }
  return s;
}
__host__ __device__ constexpr double ct_cos(double x) {
  double x2 = x * x, t = 1, s = 1;
  for (int i = 1; i < 14; ++i) { t *= -x2 / ((2 * i - 1) * (2 * i)); s += t; }
  return s;
}
__host__ __device__ constexpr double ct_cos2pi(int k, int n) {
  k %= n;
  if (k < 0) k += n;
  if ((8 * k) % n == 0) {                 // exact multiples of pi/4
    const double h = 0.70710678118654752440084436210484903928;
    switch ((8 * k) / n) {
      case 0: return 1; case 1: return h; case 2: return 0; case 3: return -h;
      case 4: return -1; case 5: return -h; case 6: return 0; default: return h;
    }
  }
  int oct = (int)((8LL * k) / n);
  double r = 2 * ct_pi() * ((double)k / n) - oct * (ct_pi() / 4);
  switch (oct) {
    case 0: return ct_cos(r);
    case 1: return ct_sin(ct_pi() / 4 - r);
    case 2: return -ct_sin(r);
    case 3: return -ct_cos(ct_pi() / 4 - r);
    case 4: return -ct_cos(r);
    case 5: return -ct_sin(ct_pi() / 4 - r);
    case 6: return ct_sin(r);
    default: return ct_cos(ct_pi() / 4 - r);
  }
}
// sin(2 pi k/n) = cos(2 pi k/n - pi/2) = cos(2 pi (4k - n)/(4n))
__host__ __device__ constexpr double ct_sin2pi(int k, int n) { return ct_cos2pi(4 * k - n, 4 * n); }

// multiply by W_N^K (forward, e^{-2 pi i K/N}) or its conjugate (INV)
template <int K, int N, bool INV, typename C>
CSMRI_HD C twiddle(C a) {
  constexpr int KK = ((K % N) + N) % N;
  if constexpr (KK == 0) {
    return a;
  } else if constexpr (4 * KK == N) {            // -i (fwd) / +i (inv)
    return INV ? rot_pi(a) : rot_mi(a);
  } else if constexpr (2 * KK == N) {
    return cneg(a);
  } else if constexpr (4 * KK == 3 * N) {        // +i (fwd) / -i (inv)
    return INV ? rot_mi(a) : rot_pi(a);
  } else {
    constexpr float c = (float)ct_cos2pi(KK, N);
    constexpr float s = (float)ct_sin2pi(KK, N);
    return cmul(a, mk(c, INV ? s : -s));
  }
}

// ---------------------------------------------------------------------------
// butterflies, in place, natural order
// ---------------------------------------------------------------------------
template <bool INV, typename C>
CSMRI_HD void fft2(C& a0, C& a1) {
  C t = cadd(a0, a1);
  a1 = csub(a0, a1);
  a0 = t;
}

template <bool INV, typename C>
CSMRI_HD void fft4(C& a0, C& a1, C& a2, C& a3) {
  C t0 = cadd(a0, a2), t1 = csub(a0, a2);
  C t2 = cadd(a1, a3), t3 = csub(a1, a3);
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  if (INV) { a1 = add_pi(t1, t3); a3 = add_mi(t1, t3); }
  else     { a1 = add_mi(t1, t3); a3 = add_pi(t1, t3); }
}

// radix-5 (320 = 2^6 * 5): 4 real constants, natural order in place
template <bool INV, typename C>
CSMRI_HD void fft5(C& a0, C& a1, C& a2, C& a3, C& a4) {
  constexpr float c1 = (float)ct_cos2pi(1, 5), c2 = (float)ct_cos2pi(2, 5);
  constexpr float s1 = (float)ct_sin2pi(1, 5), s2 = (float)ct_sin2pi(2, 5);
  const C t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  const C m1 = cmadd(cmadd(a0, t1, c1), t2, c2);
  const C m2 = cmadd(cmadd(a0, t1, c2), t2, c1);
  const C n1 = cmadd(cscale(t3, s1), t4, s2);
  const C n2 = cmadd(cscale(t3, s2), t4, -s1);
  a0 = cadd(a0, cadd(t1, t2));
  if (INV) { a1 = add_pi(m1, n1); a4 = add_mi(m1, n1); a2 = add_pi(m2, n2); a3 = add_mi(m2, n2); }
  else     { a1 = add_mi(m1, n1); a4 = add_pi(m1, n1); a2 = add_mi(m2, n2); a3 = add_pi(m2, n2); }
}

template <int N, bool INV> struct RegFFT;

template <bool INV> struct RegFFT<1, INV> {
  template <typename C> static CSMRI_HD void run(C*) {}
};
template <bool INV> struct RegFFT<2, INV> {
  template <typename C> static CSMRI_HD void run(C* v) { fft2<INV>(v[0], v[1]); }
};
template <bool INV> struct RegFFT<5, INV> {
  template <typename C> static CSMRI_HD void run(C* v) { fft5<INV>(v[0], v[1], v[2], v[3], v[4]); }
};
template <bool INV> struct RegFFT<4, INV> {
  template <typename C> static CSMRI_HD void run(C* v) { fft4<INV>(v[0], v[1], v[2], v[3]); }
};

// Generic two-factor step N = A*B on a register array (all loops unroll):
//   B sub-FFTs of size A over v[n2 + B*n1], twiddle W_N^{n2*k1},
//   A sub-FFTs of size B over n2, then the compile-time transpose that puts
//   X[k1 + A*k2] at v[k1 + A*k2].
template <int N, int A, int B, bool INV>
struct RegFFT2F {
  template <int N2, typename C>
  static CSMRI_HD void stage1(C* v) {
    if constexpr (N2 < B) {
      C t[A];
#pragma unroll
      for (int n1 = 0; n1 < A; ++n1) t[n1] = v[N2 + B * n1];
      RegFFT<A, INV>::run(t);
      tw_row<N2, 0>(t);
#pragma unroll
      for (int k1 = 0; k1 < A; ++k1) v[N2 + B * k1] = t[k1];
      stage1<N2 + 1>(v);
    }
  }
  template <int N2, int K1, typename C>
  static CSMRI_HD void tw_row(C* t) {
    if constexpr (K1 < A) {
      t[K1] = twiddle<N2 * K1, N, INV>(t[K1]);
      tw_row<N2, K1 + 1>(t);
    }
  }
  template <typename C>
  static CSMRI_HD void run(C* v) {
    stage1<0>(v);
    C w[N];
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) {
      C t[B];
#pragma unroll
      for (int n2 = 0; n2 < B; ++n2) t[n2] = v[n2 + B * k1];
      RegFFT<B, INV>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < B; ++k2) w[k1 + A * k2] = t[k2];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = w[i];
  }
};

template <bool INV> struct RegFFT<8, INV> {
  template <typename C> static CSMRI_HD void run(C* v) { RegFFT2F<8, 4, 2, INV>::run(v); }
};
template <bool INV> struct RegFFT<16, INV> {
  template <typename C> static CSMRI_HD void run(C* v) { RegFFT2F<16, 4, 4, INV>::run(v); }
};
template <bool INV> struct RegFFT<40, INV> {
  template <typename C> static CSMRI_HD void run(C* v) { RegFFT2F<40, 8, 5, INV>::run(v); }
};
template <bool INV> struct RegFFT<32, INV> {
  template <typename C> static CSMRI_HD void run(C* v) { RegFFT2F<32, 4, 8, INV>::run(v); }
};

}  // namespace csmri
