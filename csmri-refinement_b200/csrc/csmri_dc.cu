// libcsmri_dc.so - hand-written sm_100a kernels for the k-space
// data-consistency (DC) path of mseitzer/csmri-refinement, behind the C ABI of
// include/csmri_dc.h.  No torch types, no CPU fallback.
//
// Reference arithmetic being replaced (paths relative to the reference root):
//   data/reconstruction/deep_med_lib/my_pytorch/myfft.py:78-128   Fft2d/Ifft2d
//   data/reconstruction/deep_med_lib/my_pytorch/myfft.py:131-142  blend
//   data/reconstruction/deep_med_lib/my_pytorch/myfft.py:145-163  perform
//   models/recnet.py:147-148                                      residual add
//   data/reconstruction/deep_med_lib/utils/compressed_sensing.py:460-512 undersample
//
// The math every Cartesian kernel relies on: Cartesian masks are constant
// along W (compressed_sensing.py:115-116), so the W-axis transforms of FFT2 /
// iFFT2 cancel around the blend and DC becomes, per image column,
//     out = iFFT_H(D * FFT_H(x) + addend),
// with D (per row) and addend = iFFT_W(c*k0) prepared once per batch.
//
// Kernels
//   dc_strip_pipev_kernel  (dc_pipev.cuh) the hot one: persistent, TMA-fed,
//                          two image columns per thread, dynamic tile
//                          scheduler, programmatic dependent launch.
//   dc_strip_pipe_kernel   (dc_pipe.cuh) one-column twin, used for 320 / 512.
//   dc_strip_row_kernel    direct global->register variant of the same math:
//                          sizes without a pipelined instantiation (32, 1024)
//                          and pointers that are not 16-byte aligned.
//   dc_strip_dense_kernel  the column pass for an arbitrary dense mask, on a
//                          row-transformed (hybrid) tensor.
//   fft_strip_kernel       single column DFT (fft2 / undersample).
//   fft_rows_kernel        single row DFT; the threads of a row sit in adjacent
//                          lanes, so HBM <-> register traffic is coalesced without
//                          a staging transpose.
//   mask_rows_kernel       proves row-constancy and builds the D table.
//   magnitude_clamp_kernel, psnr_sum_kernel   reporting-side pointwise ops.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "../../include/csmri_dc.h"
#include "dc_core.cuh"
#include "dc_pipe.cuh"
#include "dc_pipev.cuh"
#include "refine_ops.cuh"
#include "loader_tail.cuh"
#include "conv_wgrad.cuh"
#include "conv_epilogue.cuh"
#include "conv_thin.cuh"
#include "conv_tc.cuh"
#include "conv_wgrad_tc.cuh"

namespace csmri {

cf h_twiddle[kTwN];

// ---------------------------------------------------------------------------
// global memory access helpers: everything is streamed exactly once
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---------------------------------------------------------------------------
// hot kernel: Cartesian DC forward / adjoint on one H x CW column strip
// ---------------------------------------------------------------------------
template <int H, int E, int CW, int MINB, int WT>
__global__ void __launch_bounds__(CW*(H / E), MINB)
    dc_strip_row_kernel(const float* __restrict__ x, const float* __restrict__ residual,
                        const float* __restrict__ dtab, const float* __restrict__ addend,
                        float* __restrict__ out, int W_rt, int nstrips_rt) {
  typedef LineFFT<H, E, CW> L;
  constexpr int T = L::T;
  // WT != 0: the row pitch is a compile-time constant, so every row offset
  // below folds into the immediate field of LDG/STG (no address arithmetic)
  const int W = WT ? WT : W_rt;
  const int nstrips = WT ? WT / CW : nstrips_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* sm = reinterpret_cast<cf*>(smem_raw);
  cf* tw_s = reinterpret_cast<cf*>(smem_raw + L::kSmemBytes);
  L::fill_twiddles(tw_s, threadIdx.x, CW * T);

  const int lane = threadIdx.x % CW;
  const int j = threadIdx.x / CW;
  const int b = blockIdx.x / nstrips;
  const int strip = blockIdx.x - b * nstrips;
  const size_t plane = (size_t)H * W;
  const size_t base = (size_t)b * 2 * plane + (size_t)strip * CW + lane;

  cf v[E];
  {
    const float* pr = x + base;
    const float* pi = pr + plane;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const size_t o = (size_t)(j + T * i) * W;
      v[i] = mk(ld_stream(pr + o), ld_stream(pi + o));
    }
  }
  if (residual != nullptr) {
    const float* pr = residual + base;
    const float* pi = pr + plane;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const size_t o = (size_t)(j + T * i) * W;
      v[i] = cadd(v[i], mk(ld_stream(pr + o), ld_stream(pi + o)));
    }
  }
  __syncthreads();  // twiddle table ready (the loads above are already in flight)
  L::template a_front<false>(v, sm, tw_s, j, lane);
  __syncthreads();
  L::template a_back<false>(v, sm, j, lane);

  L::apply_dtab(v, dtab + (size_t)b * H + j * E);
  if (addend != nullptr) {   // hybrid-space k0 term, rows in k-layout order
    const float* pr = addend + base;
    const float* pi = pr + plane;
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const size_t o = (size_t)j * W + (size_t)L::k_index(0, r) * W;
      v[r] = cadd(v[r], mk(ld_stream(pr + o), ld_stream(pi + o)));
    }
  }

  L::template b_front<true>(v, sm, j, lane);
  __syncthreads();
  L::template b_back<true>(v, sm, tw_s, j, lane);
  {
    float* pr = out + base;
    float* pi = pr + plane;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const size_t o = (size_t)(j + T * i) * W;
      st_stream(pr + o, v[i].x);
      st_stream(pi + o, v[i].y);
    }
  }
}

// ---------------------------------------------------------------------------
// column pass for arbitrary masks, on the row-transformed tensor (in place ok)
//   ADJ = false: Y = blend(s*K, k0, mask)  (myfft.py:139 / :141 as written)
//   ADJ = true : Y = D * s*K, D = (1-m) or (1-m)+m/(1+v)
// ---------------------------------------------------------------------------
template <int H, int E, int CW, bool NOISY, bool ADJ, int WT, int MINB>
__global__ void __launch_bounds__(CW*(H / E), MINB)
    dc_strip_dense_kernel(const float* __restrict__ hyb, const float* __restrict__ k0,
                          const float* __restrict__ mask, float* __restrict__ out, int W_rt,
                          int nstrips_rt, float s, float nv) {
  typedef LineFFT<H, E, CW> L;
  constexpr int T = L::T;
  const int W = WT ? WT : W_rt;                 // compile-time pitch for square slices
  const int nstrips = WT ? WT / CW : nstrips_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* sm = reinterpret_cast<cf*>(smem_raw);
  cf* tw_s = reinterpret_cast<cf*>(smem_raw + L::kSmemBytes);
  L::fill_twiddles(tw_s, threadIdx.x, CW * T);

  const int lane = threadIdx.x % CW;
  const int j = threadIdx.x / CW;
  const int b = blockIdx.x / nstrips;
  const int strip = blockIdx.x - b * nstrips;
  const size_t plane = (size_t)H * W;
  const size_t base = (size_t)b * 2 * plane + (size_t)strip * CW + lane;

  cf v[E];
  {
    const float* pr = hyb + base;
    const float* pi = pr + plane;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const size_t o = (size_t)(j + T * i) * W;
      v[i] = mk(ld_stream(pr + o), ld_stream(pi + o));
    }
  }
  __syncthreads();  // twiddle table ready (the loads above are already in flight)
  L::template a_front<false>(v, sm, tw_s, j, lane);
  __syncthreads();
  L::template a_back<false>(v, sm, j, lane);

  {
    const float inv1pv = 1.0f / (1.0f + nv);
    const float* mr = mask + base;
    const float* mi = mr + plane;
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const size_t o = (size_t)L::k_index(j, r) * W;
      const float m0 = ld_stream(mr + o), m1 = ld_stream(mi + o);
      const cf k = cscale(v[r], s);
      if (ADJ) {
        const float d0 = NOISY ? (1.0f - m0) + m0 * inv1pv : (1.0f - m0);
        const float d1 = NOISY ? (1.0f - m1) + m1 * inv1pv : (1.0f - m1);
        v[r] = mk(d0 * k.x, d1 * k.y);
      } else {
        const float* kr = k0 + base;
        const float* ki = kr + plane;
        const float a = ld_stream(kr + o), c = ld_stream(ki + o);
        if (NOISY) {
          v[r] = mk((1.0f - m0) * k.x + m0 * (k.x + nv * a) * inv1pv,
                    (1.0f - m1) * k.y + m1 * (k.y + nv * c) * inv1pv);
        } else {
          v[r] = mk((1.0f - m0) * k.x + a, (1.0f - m1) * k.y + c);
        }
      }
    }
  }

  L::template b_front<true>(v, sm, j, lane);
  __syncthreads();
  L::template b_back<true>(v, sm, tw_s, j, lane);
  {
    float* pr = out + base;
    float* pi = pr + plane;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const size_t o = (size_t)(j + T * i) * W;
      st_stream(pr + o, v[i].x * s);
      st_stream(pi + o, v[i].y * s);
    }
  }
}

// ---------------------------------------------------------------------------
// single column DFT:  out[k, w] = scale * rowmul[k] * sum_h in[h, w] W_H^{+-hk}
//   rows  : optional (B,H) uint8 multiplier on OUTPUT rows (undersample mask)
//   out2  : optional second destination for the same values
// ---------------------------------------------------------------------------
template <int H, int E, int CW, bool INV, int WT, int MINB>
__global__ void __launch_bounds__(CW*(H / E), MINB)
    fft_strip_kernel(const float* __restrict__ in, float* __restrict__ out, int W_rt,
                     int nstrips_rt, float scale, const unsigned char* __restrict__ rows) {
  typedef LineFFT<H, E, CW> L;
  constexpr int T = L::T;
  const int W = WT ? WT : W_rt;
  const int nstrips = WT ? WT / CW : nstrips_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* sm = reinterpret_cast<cf*>(smem_raw);
  cf* tw_s = reinterpret_cast<cf*>(smem_raw + L::kSmemBytes);
  L::fill_twiddles(tw_s, threadIdx.x, CW * T);

  const int lane = threadIdx.x % CW;
  const int j = threadIdx.x / CW;
  const int b = blockIdx.x / nstrips;
  const int strip = blockIdx.x - b * nstrips;
  const size_t plane = (size_t)H * W;
  const size_t base = (size_t)b * 2 * plane + (size_t)strip * CW + lane;

  cf v[E];
  {
    const float* pr = in + base;
    const float* pi = pr + plane;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const size_t o = (size_t)(j + T * i) * W;
      v[i] = mk(ld_stream(pr + o), ld_stream(pi + o));
    }
  }
  __syncthreads();  // twiddle table ready
  L::template a_front<INV>(v, sm, tw_s, j, lane);
  __syncthreads();
  L::template a_back<INV>(v, sm, j, lane);
  {
    float* pr = out + base;
    float* pi = pr + plane;
#pragma unroll
    for (int r = 0; r < E; ++r) {
      const int k = L::k_index(j, r);
      // mask * value as a product, so that masked entries are signed zeros
      // exactly like compressed_sensing.py:510 (`mask * (x_f + nz)`)
      const float m = (rows != nullptr) ? (rows[(size_t)b * H + k] ? 1.0f : 0.0f) : 1.0f;
      const size_t o = (size_t)k * W;
      st_stream(pr + o, (v[r].x * scale) * m);
      st_stream(pi + o, (v[r].y * scale) * m);
    }
  }
}

// ---------------------------------------------------------------------------
// Loader step for Cartesian (row-constant) sampling, column part.  With
// m[h, w] = r[h] the zero-filled image is F^-1(m . F x) = F_H^-1(r . F_H x): the
// W-axis transforms cancel, so one pass over the real image column gives
//   target = (x, 0)                                   dnn_io.py:47-61
//   hyb    = r[k] * FFT_H(x)[k] / H                   masked hybrid-space rows; this is
//                                                     also the DC plan's `addend`
//   inp    = iFFT_H(hyb)                              compressed_sensing.py:511
// and k-space itself follows from the sampled rows of hyb alone (row kernel,
// PRE = 4).  Replaces row FFT -> column FFT + mask -> row iFFT -> column iFFT.
// ---------------------------------------------------------------------------
template <int H, int E, int CW, int WT, int MINB>
__global__ void __launch_bounds__(CW*(H / E), MINB)
    undersample_strip_kernel(const float* __restrict__ img, const unsigned char* __restrict__ rows,
                             float* __restrict__ target, float* __restrict__ hyb,
                             float* __restrict__ inp, int W_rt, int nstrips_rt) {
  typedef LineFFT<H, E, CW> L;
  constexpr int T = L::T;
  const int W = WT ? WT : W_rt;
  const int nstrips = WT ? WT / CW : nstrips_rt;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* sm = reinterpret_cast<cf*>(smem_raw);
  cf* tw_s = reinterpret_cast<cf*>(smem_raw + L::kSmemBytes);
  L::fill_twiddles(tw_s, threadIdx.x, CW * T);

  const int lane = threadIdx.x % CW;
  const int j = threadIdx.x / CW;
  const int b = blockIdx.x / nstrips;
  const int strip = blockIdx.x - b * nstrips;
  const size_t plane = (size_t)H * W;
  const size_t base_img = (size_t)b * plane + (size_t)strip * CW + lane;
  const size_t base = (size_t)b * 2 * plane + (size_t)strip * CW + lane;

  cf v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) v[i] = mk(ld_stream(img + base_img + (size_t)(j + T * i) * W), 0.0f);
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const size_t o = (size_t)(j + T * i) * W;
    st_stream(target + base + o, v[i].x);
    st_stream(target + base + plane + o, 0.0f);
  }
  __syncthreads();  // twiddle table ready
  L::template a_front<false>(v, sm, tw_s, j, lane);
  __syncthreads();
  L::template a_back<false>(v, sm, j, lane);
  constexpr float inv_h = 1.0f / (float)H;
#pragma unroll
  for (int r = 0; r < E; ++r) {
    const int k = L::k_index(j, r);
    const float m = rows[(size_t)b * H + k] ? 1.0f : 0.0f;
    v[r] = mk((v[r].x * inv_h) * m, (v[r].y * inv_h) * m);
    const size_t o = (size_t)k * W;
    st_stream(hyb + base + o, v[r].x);
    st_stream(hyb + base + plane + o, v[r].y);
  }
  L::template b_front<true>(v, sm, j, lane);
  __syncthreads();
  L::template b_back<true>(v, sm, tw_s, j, lane);
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const size_t o = (size_t)(j + T * i) * W;
    st_stream(inp + base + o, v[i].x);
    st_stream(inp + base + plane + o, v[i].y);
  }
}

// ---------------------------------------------------------------------------
// single row DFT along W.  A CTA owns RT = CW rows of one plane pair; the tile
// is staged through shared memory so that global accesses stay coalesced while
// the lanes of a warp own different rows.
//   PRE: 0 none, 1 add `aux` (residual), 2 multiply by cmul*aux (aux = mask),
//        3 real input (imaginary plane absent / zero),
//        4 sampled rows only: rows with rowsel[b, h] == 0 are neither read nor
//          transformed, their output is zero; the dense 2-channel mask
//          (dnn_io.py:56-59) is written alongside into mask_out
//        5 sampled rows only, COMPACT input: `in` is (B,2,L,W) and holds just the
//          L sampled lines of each slice in ascending row order (L travels in
//          cmulv); row h reads line #(sampled rows before h); no mask output
// ---------------------------------------------------------------------------
template <int W, int E, int R, bool INV, int PRE>
__global__ void __launch_bounds__(R*(W / E))
    fft_rows_kernel(const float* __restrict__ in, const float* __restrict__ aux,
                    float* __restrict__ out, int H, float scale, float cmulv,
                    const unsigned char* __restrict__ rowsel, float* __restrict__ mask_out) {
  // The T = W/E threads of one row sit in adjacent lanes (thread j holds
  // x[j + T*i]), so for every i the lanes read / write consecutive floats:
  // global traffic goes straight between HBM and registers in T*4-byte
  // segments, no staging transpose.  R rows per CTA.  The exchange buffer is
  // private to a row; its k1-rows are padded by one slot so that both the
  // n-layout (consecutive j) and the k-layout (consecutive t, stride T+1
  // slots) accesses are bank-conflict free.
  constexpr int T = W / E;
  constexpr int Q = E / T;
  constexpr int TP = T + 1;
  constexpr int NT = R * T;
  static_assert(Q * T == E, "T must divide E");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cf* sm_all = reinterpret_cast<cf*>(smem_raw);          // [R][E*TP]
  cf* tw_t = sm_all + R * E * TP;                        // [E][T]: W_W^{j*k1} at k1*T + j
  {
    const cf* src = g_tw_lines + tw_lines_offset(W);     // global table is [T][E]
    for (int idx = threadIdx.x; idx < W; idx += NT) {
      const int jj = idx / E, k1 = idx - jj * E;
      tw_t[k1 * T + jj] = __ldg(src + idx);
    }
  }
  const int j = threadIdx.x % T;
  const int r = threadIdx.x / T;
  cf* sm = sm_all + r * (E * TP);

  const int tiles_per_slice = H / R;
  const int b = blockIdx.x / tiles_per_slice;
  const int row = (blockIdx.x - b * tiles_per_slice) * R + r;
  const size_t plane = (size_t)H * W;
  size_t off_in = (PRE == 3 ? (size_t)b * plane : (size_t)b * 2 * plane) + (size_t)row * W + j;
  size_t plane_in = plane;
  const size_t off = (size_t)b * 2 * plane + (size_t)row * W + j;

  // PRE == 4 / 5: uniform over the T threads (adjacent lanes) that share a row
  const bool on = (PRE != 4 && PRE != 5) || rowsel[(size_t)b * H + row] != 0;
  if (PRE == 5) {
    // position of this row among the sampled lines of its slice: the T threads
    // of the row count disjoint parts of rowsel[b, 0..row) and fold in the lanes
    const int L = (int)cmulv;
    int cnt = 0;
    for (int h = j; h < row; h += T) cnt += rowsel[(size_t)b * H + h] != 0;
#pragma unroll
    for (int sft = T / 2; sft > 0; sft >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, sft);
    if (cnt > L - 1) cnt = L - 1;   // malformed table (reported through lines_ok): stay in bounds
    plane_in = (size_t)L * W;
    off_in = (size_t)b * 2 * plane_in + (size_t)cnt * W + j;
  }
  cf v[E];
#pragma unroll
  for (int i = 0; i < E; ++i) {
    float re = on ? ld_stream(in + off_in + T * i) : 0.0f;
    float im = (PRE == 3 || !on) ? 0.0f : ld_stream(in + off_in + plane_in + T * i);
    if (PRE == 1) {
      re += ld_stream(aux + off + T * i);
      im += ld_stream(aux + off + plane + T * i);
    } else if (PRE == 2) {
      re *= cmulv * ld_stream(aux + off + T * i);
      im *= cmulv * ld_stream(aux + off + plane + T * i);
    }
    v[i] = mk(re, im);
  }
  if (on) RegFFT<E, INV>::run(v);
  __syncthreads();   // twiddle table ready
  if (on) {
#pragma unroll
    for (int k1 = 1; k1 < E; ++k1) {
      const cf w = tw_t[k1 * T + j];
      v[k1] = INV ? cmul_conj(v[k1], w) : cmul(v[k1], w);
    }
#pragma unroll
    for (int k1 = 0; k1 < E; ++k1) sm[k1 * TP + j] = v[k1];
  }
  __syncthreads();
  if (on) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j2 = 0; j2 < T; ++j2) v[q * T + j2] = sm[(q * T + j) * TP + j2];
#pragma unroll
    for (int q = 0; q < Q; ++q) RegFFT<T, INV>::run(v + q * T);
  }
  // thread t = j holds X[(q*T + t) + E*k2] in v[q*T + k2]: consecutive t -> consecutive k
#pragma unroll
  for (int rr = 0; rr < E; ++rr) {
    const int k0i = (rr / T) * T + E * (rr % T);          // + j
    st_stream(out + off + k0i, v[rr].x * scale);
    st_stream(out + off + plane + k0i, v[rr].y * scale);
    if (PRE == 4) {
      const float m = on ? 1.0f : 0.0f;
      st_stream(mask_out + off + k0i, m);
      st_stream(mask_out + off + plane + k0i, m);
    }
  }
}

// ---------------------------------------------------------------------------
// mask analysis: one warp per (b, h) row
// ---------------------------------------------------------------------------
__global__ void set_flag_kernel(int* flag, int value) { *flag = value; }

// dtab layout: D[b, k]/H stored at b*H + dtab_slot(k) for the (E, T = H/E)
// decomposition every strip kernel of this H uses (LineFFT::dtab_slot)
__global__ void mask_rows_kernel(const float* __restrict__ mask, int B, int H, int W, float nv,
                                 int noisy, int E, float* __restrict__ dtab,
                                 int* __restrict__ flag) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32;
  const int lane = threadIdx.x % 32;
  if (warp >= B * H) return;
  const int b = warp / H, h = warp - b * H;
  const size_t plane = (size_t)H * W;
  const float* r0 = mask + (size_t)b * 2 * plane + (size_t)h * W;
  const float* r1 = r0 + plane;
  const float m = r0[0];
  bool ok = true;
  if ((((uintptr_t)r0 | (uintptr_t)r1) & 15u) == 0) {     // 128-bit loads (W % 4 == 0 always)
    const float4* q0 = reinterpret_cast<const float4*>(r0);
    const float4* q1 = reinterpret_cast<const float4*>(r1);
    for (int w = lane; w < W / 4; w += 32) {
      const float4 a = __ldcs(q0 + w), c = __ldcs(q1 + w);
      ok = ok && a.x == m && a.y == m && a.z == m && a.w == m && c.x == m && c.y == m &&
           c.z == m && c.w == m;
    }
  } else {
    for (int w = lane; w < W; w += 32) ok = ok && (r0[w] == m) && (r1[w] == m);
  }
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    if (!ok) atomicExch(flag, 0);
    const float d = noisy ? (1.0f - m) + m / (1.0f + nv) : (1.0f - m);
    const int T = H / E;
    const int k1 = h % E, k2 = h / E;
    dtab[(size_t)b * H + (k1 % T) * E + (k1 / T) * T + k2] = d / (float)H;
  }
}

// D table (noiseless: D = 1 - m) straight from the sampled-line table, in the
// strip kernels' slot order - what mask_rows_kernel derives from a dense mask
__global__ void dtab_from_rows_kernel(const unsigned char* __restrict__ rows, int B, int H, int E,
                                      float* __restrict__ dtab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, h = i - b * H, T = H / E;
  const int k1 = h % E, k2 = h / E;
  const float m = rows[i] ? 1.0f : 0.0f;
  dtab[(size_t)b * H + (k1 % T) * E + (k1 / T) * T + k2] = (1.0f - m) / (float)H;
}

// The same for a compact plan request (csmri_dc_prepare_lines): one CTA per
// slice, D = 1-m or (1-m) + m/(1+v) as in mask_rows_kernel, and the number of
// sampled rows of every slice is checked against L (lines_ok <- 0 on mismatch).
__global__ void dtab_from_rows_checked_kernel(const unsigned char* __restrict__ rows, int H, int E,
                                              int L, float nv, int noisy,
                                              float* __restrict__ dtab, int* __restrict__ lines_ok) {
  const int b = blockIdx.x, T = H / E;
  int cnt = 0;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float m = rows[(size_t)b * H + h] ? 1.0f : 0.0f;
    cnt += m != 0.0f;
    const float d = noisy ? (1.0f - m) + m / (1.0f + nv) : (1.0f - m);
    const int k1 = h % E, k2 = h / E;
    dtab[(size_t)b * H + (k1 % T) * E + (k1 / T) * T + k2] = d / (float)H;
  }
  __shared__ int total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  if (cnt) atomicAdd(&total, cnt);
  __syncthreads();
  if (threadIdx.x == 0 && total != L) atomicExch(lines_ok, 0);
}

// ---------------------------------------------------------------------------
// reporting-side pointwise ops on the DC output
//   magnitude_clamp : out = clamp(sqrt(re^2 + im^2), lo, hi)   utils/tensor_transforms.py:62-75
//                     + data/reconstruction/rec_transforms.py:79-85 (output_transform)
//   psnr_sum        : sum over all pixels of (|pred|_c - |target|_c)^2, the MSE numerator of
//                     metrics/image_metrics.py:7-19, in one pass over both tensors
// Squares, sum and square root are rounded separately (no FMA contraction) so
// the magnitudes are bit-identical to torch's (a**2 + b**2) ** 0.5.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float magnitude_clamped(float re, float im, float lo, float hi) {
  const float m = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
  return fminf(fmaxf(m, lo), hi);
}

__global__ void magnitude_clamp_kernel(const float* __restrict__ x, float* __restrict__ out,
                                       size_t plane, size_t total, float lo, float hi) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, r = i - b * plane;
    out[i] = magnitude_clamped(x[b * 2 * plane + r], x[b * 2 * plane + plane + r], lo, hi);
  }
}

__global__ void psnr_sum_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                double* __restrict__ sum_sq, size_t plane, size_t total, float lo,
                                float hi) {
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, r = i - b * plane, o = b * 2 * plane + r;
    const float d = magnitude_clamped(pred[o], pred[o + plane], lo, hi) -
                    magnitude_clamped(target[o], target[o + plane], lo, hi);
    acc += (double)d * (double)d;
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  __shared__ double part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (threadIdx.x == 0) atomicAdd(sum_sq, acc);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(CSMRI_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CSMRI_CUDA(call)                                   \
  do {                                                     \
    cudaError_t e_ = (call);                               \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);    \
  } while (0)

// elements per thread (E) of the two-pass column FFT for each supported line
// length; the dtab layout depends on it, so every strip-kernel variant of one H
// shares it
static int strip_radix(int H) { return H == 320 ? 40 : (H <= 64 ? 8 : (H <= 256 ? 16 : 32)); }

static std::mutex g_host_mutex;   // guards the lazily filled host-side tables
static bool g_tw_uploaded[64] = {false};

static int ensure_init() {
  int dev = 0;
  CSMRI_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(CSMRI_E_CUDA, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (g_tw_uploaded[dev]) return CSMRI_OK;
  static bool host_ready = false;
  if (!host_ready) {
    for (int m = 0; m < kTwN; ++m) {
      const double a = -2.0 * M_PI * (double)m / (double)kTwN;
      h_twiddle[m] = mk((float)cos(a), (float)sin(a));
    }
    // exact values on the axes and diagonals
    h_twiddle[0] = mk(1.f, 0.f);
    h_twiddle[kTwN / 4] = mk(0.f, -1.f);
    h_twiddle[kTwN / 2] = mk(-1.f, 0.f);
    h_twiddle[3 * kTwN / 4] = mk(0.f, 1.f);
    host_ready = true;
  }
  CSMRI_CUDA(cudaMemcpyToSymbol(c_twiddle, h_twiddle, sizeof(cf) * kTwN));
  {
    // [T][E] inter-pass tables for every supported line length (dc_core.cuh)
    static cf lines[kTwLinesTotal];
    const int sizes[7] = {32, 64, 128, 256, 512, 1024, 320};
    for (int si = 0; si < 7; ++si) {
      const int n = sizes[si], e = strip_radix(n);
      for (int idx = 0; idx < n; ++idx) {
        const int jj = idx / e, k1 = idx % e;
        const int m = (jj * k1) % n;
        cf w;
        if (4 * m == n) w = mk(0.f, -1.f);
        else if (2 * m == n) w = mk(-1.f, 0.f);
        else if (4 * m == 3 * n) w = mk(0.f, 1.f);
        else if (m == 0) w = mk(1.f, 0.f);
        else {
          const double a = -2.0 * M_PI * (double)m / (double)n;
          w = mk((float)cos(a), (float)sin(a));
        }
        lines[tw_lines_offset(n) + idx] = w;
      }
    }
    CSMRI_CUDA(cudaMemcpyToSymbol(g_tw_lines, lines, sizeof(lines)));
  }
  g_tw_uploaded[dev] = true;
  return CSMRI_OK;
}

static bool pow2_in(int n, int lo, int hi) { return n >= lo && n <= hi && (n & (n - 1)) == 0; }

static int check_shape(int B, int H, int W) {
  if (B <= 0) return fail(CSMRI_E_SHAPE, "batch must be positive, got %d", B);
  if (!(pow2_in(H, 32, 1024) || H == 320) || !(pow2_in(W, 32, 1024) || W == 320))
    return fail(CSMRI_E_SHAPE,
                "unsupported slice size %dx%d: H and W must be powers of two in [32, 1024] or 320",
                H, W);
  if ((long long)B * H * W * 2 >= (1LL << 40)) return fail(CSMRI_E_SHAPE, "problem too large");
  return CSMRI_OK;
}
static int check_ptr(const void* p, const char* name) {
  if (p == nullptr) return fail(CSMRI_E_NULLPTR, "%s is NULL", name);
  if (((uintptr_t)p & 3u) != 0) return fail(CSMRI_E_ALIGN, "%s is not 4-byte aligned", name);
  return CSMRI_OK;
}
#define CSMRI_TRY(expr)            \
  do {                             \
    int rc_ = (expr);              \
    if (rc_ != CSMRI_OK) return rc_; \
  } while (0)

static int check_ptr16(const void* p, const char* name) {
  CSMRI_TRY(check_ptr(p, name));
  if (((uintptr_t)p & 15u) != 0) return fail(CSMRI_E_ALIGN, "%s is not 16-byte aligned", name);
  return CSMRI_OK;
}

// cudaFuncSetAttribute is issued once per (kernel, device), never on the hot
// path: per-launch calls showed up as a ~9 us bubble between back-to-back
// launches in the per-CTA timeline (tools/gpu_trace.py).
struct SmemOptIn {
  const void* fn;
  int dev;
  int bytes;
};
static SmemOptIn g_smem_optin[512];
static int g_smem_optin_n = 0;

template <typename K>
static int set_smem(K kernel, int bytes) {
  if (bytes <= 48 * 1024) return CSMRI_OK;
  int dev = 0;
  CSMRI_CUDA(cudaGetDevice(&dev));
  const void* fn = (const void*)kernel;
  std::lock_guard<std::mutex> lock(g_host_mutex);
  for (int i = 0; i < g_smem_optin_n; ++i)
    if (g_smem_optin[i].fn == fn && g_smem_optin[i].dev == dev && g_smem_optin[i].bytes >= bytes)
      return CSMRI_OK;
  CSMRI_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (g_smem_optin_n < 512) g_smem_optin[g_smem_optin_n++] = SmemOptIn{fn, dev, bytes};
  return CSMRI_OK;
}

// Dynamic tile scheduler state: (tiles handed out, retired CTAs) pairs, zero =
// armed; the last CTA of a launch re-arms its pair, so a pair may be reused as
// soon as that launch has finished.  Two launches that can be in flight at the
// same time must not share a pair.  Launches on ONE stream are ordered (with
// programmatic dependent launch only neighbours overlap), so each stream - and
// each CUDA-graph capture, whose baked-in pair is replayed later on any stream -
// owns a range of kSchedPerRange pairs and walks it round-robin.  Ranges are
// handed out first come first served and recycled oldest-first once more than
// kSchedRanges distinct streams / captures have been seen.
constexpr int kSchedRanges = 64, kSchedPerRange = 64;
__device__ unsigned g_sched[2 * kSchedRanges * kSchedPerRange];
struct SchedRange {
  unsigned long long key;
  unsigned next;
  bool used;
};
static SchedRange g_sched_ranges[kSchedRanges];
static unsigned g_sched_victim = 0;
static int g_use_pdl = 1;             // programmatic dependent launch for the strip kernels
static int g_wgrad_cot = 8;      // output channels per thread of conv3x3_wgrad_kernel (8; 4 = A/B baseline)
static int g_strip_variant = 0;  // tuning knobs, see csmri_set_variant / csmri_set_tuning
static int g_tc_debug = 0;       // conv3x3_tc_kernel probe bits (tuning key 6)
static int g_wgrad_tc = 1;       // 32 -> 32 weight gradient on the tensor cores (tuning key 7; 0 = SIMT kernel)
static int g_thin_tma = 1;       // 32 -> 2 thin convolution staged by TMA (tuning key 8; 0 = cp.async staging)
static int g_conv_pdl = 0;       // programmatic dependent launch for the tensor-core convolution kernels
                                 // (tuning key 10; 0 = plain launches, 1 = wait before the first global read,
                                 // 2 = wait after the weight staging).  Measured: no gain - D5C5 step 13.56 /
                                 // 13.67-13.93 / 13.71-13.75 ms for 0 / 1 / 2 (profiles/r2_conv_pdl.json): a
                                 // 200 KB-smem CTA cannot share an SM with its predecessor, so there is no
                                 // prologue to overlap.  Off by default.
static int g_general_chunk_mib = 0;   // general-mask path: MiB of hybrid scratch per chunk (tuning key 9; 0 = whole batch per pass)
static long long* g_trace = nullptr;  // tuning probe: per-CTA timeline buffers (2 x 1024 x 40)
static int g_trace_launch = 0;

// ---- TMA-fed persistent strip kernel -----------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  // cuTensorMapEncodeTiled is a driver-API call: it needs the device's primary context current
  // on the CALLING thread (CUDA_ERROR_INVALID_CONTEXT otherwise).  A thread that has only ever
  // been handed work by another one - torch's autograd worker running a backward pass - may
  // not have it bound yet; cudaSetDevice binds it and is legal during stream capture.
  static thread_local int bound_dev = -1;
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != bound_dev && cudaSetDevice(dev) == cudaSuccess)
    bound_dev = dev;
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// (B,2,H,W) fp32 viewed as {W, H, 2B} (H <= 256) or {W, 256, H/256, 2B}; one
// box = the CW-column strip of both planes of one slice.
static int make_tile_map(CUtensorMap* m, const float* ptr, int B, int H, int W, int CW) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(CSMRI_E_CUDA, "cuTensorMapEncodeTiled is unavailable");
  CUresult r;
  if (H <= 256) {
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)2 * B};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)CW, (cuuint32_t)H, 2};
    cuuint32_t es[3] = {1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    // box dimensions are limited to 256: split the rows into equal chunks
    const int nch = (H + 255) / 256, rows = H / nch;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)rows, (cuuint64_t)nch, (cuuint64_t)2 * B};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * rows * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[4] = {(cuuint32_t)CW, (cuuint32_t)rows, (cuuint32_t)nch, 2};
    cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)ptr, dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return fail(CSMRI_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CSMRI_OK;
}

// (N,32,H,W) fp32 viewed as {W, H, 32 N}; one box = 32 channels x 32 pixels of one row,
// 128-byte swizzled (the K-major UMMA operand layout), out-of-range rows / columns zero-filled
// Launch with the programmatic-stream-serialization attribute: the kernel may start while the
// previous kernel in the stream drains and blocks in griddepcontrol.wait before its first
// dependent access (conv_tc.cuh, conv_wgrad_tc.cuh).
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_conv_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static int make_conv_map(CUtensorMap* m, const float* ptr, int N, int C, int H, int W) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) return fail(CSMRI_E_CUDA, "cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * C};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {32, 1, 32};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CSMRI_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return CSMRI_OK;
}

static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev < 0 || dev >= 64) ? 0 : dev;
}

static int sm_count() {
  static std::atomic<int> n[64];
  const int dev = current_device();
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// resident CTAs per SM of a persistent kernel, cached per (instantiation, device)
template <typename K>
static int resident_blocks(K kernel, int threads, int smem_bytes, std::atomic<int>* cache) {
  const int dev = current_device();
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, smem_bytes) !=
        cudaSuccess)
      v = -1;
    if (v < 1) v = -1;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// scheduler pair for a launch on stream s (see g_sched)
static int sched_slot(cudaStream_t s, unsigned** out) {
  static unsigned* base[64] = {nullptr};
  const int dev = current_device();
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  CSMRI_CUDA(cudaStreamGetCaptureInfo(s, &st, &id));
  const unsigned long long key = st == cudaStreamCaptureStatusActive
                                     ? (0x8000000000000000ULL | id)
                                     : ((unsigned long long)(uintptr_t)s & 0x7fffffffffffffffULL);
  std::lock_guard<std::mutex> lock(g_host_mutex);
  if (base[dev] == nullptr) CSMRI_CUDA(cudaGetSymbolAddress((void**)&base[dev], g_sched));
  int r = -1;
  for (int i = 0; i < kSchedRanges; ++i)
    if (g_sched_ranges[i].used && g_sched_ranges[i].key == key) { r = i; break; }
  if (r < 0) {
    for (int i = 0; i < kSchedRanges && r < 0; ++i)
      if (!g_sched_ranges[i].used) r = i;
    if (r < 0) r = (int)(g_sched_victim++ % kSchedRanges);
    g_sched_ranges[r].key = key;
    g_sched_ranges[r].next = 0;
    g_sched_ranges[r].used = true;
  }
  const unsigned slot = g_sched_ranges[r].next++ % kSchedPerRange;
  *out = base[dev] + 2 * ((size_t)r * kSchedPerRange + slot);
  return CSMRI_OK;
}

template <int H, int E, int CW, int MINB, int WT, bool ADD, bool INPL>
static int launch_strip_pipe_wt(const float* x, const float* residual, const float* dtab,
                                const float* addend, float* out, int B, int W, cudaStream_t s) {
  typedef LineFFT<H, E, CW> L;
  typedef PipeSmem<H, E, CW, ADD, INPL> S;
  auto kern = dc_strip_pipe_kernel<H, E, CW, MINB, WT, ADD, INPL>;
  CSMRI_TRY(set_smem(kern, S::kBytes));
  static std::atomic<int> occ[64];
  const int blocks_per_sm = resident_blocks(kern, CW * L::T, S::kBytes, occ);
  if (blocks_per_sm < 1) return fail(CSMRI_E_CUDA, "pipelined strip kernel does not fit an SM");
  alignas(64) CUtensorMap tm_x, tm_a;
  CSMRI_TRY(make_tile_map(&tm_x, x, B, H, W, CW));
  if (ADD) CSMRI_TRY(make_tile_map(&tm_a, addend, B, H, W, CW));
  else tm_a = tm_x;
  const int nstrips = W / CW;
  const int ntiles = B * nstrips;
  int grid = sm_count() * blocks_per_sm;
  if (grid > ntiles) grid = ntiles;
  unsigned* sched = nullptr;
  CSMRI_TRY(sched_slot(s, &sched));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CW * L::T);
  cfg.dynamicSmemBytes = S::kBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSMRI_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_x, tm_a, residual, dtab, out, W, nstrips, ntiles,
                                sched));
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

// MINB_F / MINB_A: resident CTAs per SM asked of the compiler for the forward
// (x + addend tiles in smem) and adjoint (x tile only) instantiations;
// INPL_F / INPL_A: x tile shares the exchange buffer (see dc_pipe.cuh)
template <int H, int E, int CW, int MINB_F, int MINB_A = MINB_F, bool INPL_F = false,
          bool INPL_A = false>
static int launch_strip_pipe_cfg(const float* x, const float* residual, const float* dtab,
                                 const float* addend, float* out, int B, int W, cudaStream_t s) {
  if (addend != nullptr) {
    if (W == H)
      return launch_strip_pipe_wt<H, E, CW, MINB_F, H, true, INPL_F>(x, residual, dtab, addend, out,
                                                                     B, W, s);
    return launch_strip_pipe_wt<H, E, CW, MINB_F, 0, true, INPL_F>(x, residual, dtab, addend, out, B,
                                                                   W, s);
  }
  if (W == H)
    return launch_strip_pipe_wt<H, E, CW, MINB_A, H, false, INPL_A>(x, residual, dtab, addend, out,
                                                                    B, W, s);
  return launch_strip_pipe_wt<H, E, CW, MINB_A, 0, false, INPL_A>(x, residual, dtab, addend, out, B,
                                                                  W, s);
}

template <int H, int E, int CW, int MINB, int WT, bool ADD>
static int launch_strip_pipev_wt(const float* x, const float* residual, const float* dtab,
                                 const float* addend, float* out, int B, int W, cudaStream_t s) {
  typedef PipeVSmem<H, E, CW, ADD> S;
  constexpr int threads = (CW / 2) * (H / E);
  auto kern = dc_strip_pipev_kernel<H, E, CW, MINB, WT, ADD>;
  CSMRI_TRY(set_smem(kern, S::kBytes));
  static std::atomic<int> occ[64];
  const int blocks_per_sm = resident_blocks(kern, threads, S::kBytes, occ);
  if (blocks_per_sm < 1) return fail(CSMRI_E_CUDA, "two-column strip kernel does not fit an SM");
  alignas(64) CUtensorMap tm_x, tm_a;
  CSMRI_TRY(make_tile_map(&tm_x, x, B, H, W, CW));
  if (ADD) CSMRI_TRY(make_tile_map(&tm_a, addend, B, H, W, CW));
  else tm_a = tm_x;
  const int nstrips = W / CW;
  const int ntiles = B * nstrips;
  int grid = sm_count() * blocks_per_sm;
  if (grid > ntiles) grid = ntiles;
  unsigned* sched = nullptr;
  CSMRI_TRY(sched_slot(s, &sched));
  long long* trace =
      g_trace ? g_trace + (size_t)((g_trace_launch++) & 1) * 1024 * 40 : nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = S::kBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSMRI_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_x, tm_a, residual, dtab, out, W, nstrips, ntiles,
                                trace, sched));
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

template <int H, int E, int CW, int MINB_F, int MINB_A>
static int launch_strip_pipev_cfg(const float* x, const float* residual, const float* dtab,
                                  const float* addend, float* out, int B, int W, cudaStream_t s) {
  if (addend != nullptr) {
    if (W == H)
      return launch_strip_pipev_wt<H, E, CW, MINB_F, H, true>(x, residual, dtab, addend, out, B, W, s);
    return launch_strip_pipev_wt<H, E, CW, MINB_F, 0, true>(x, residual, dtab, addend, out, B, W, s);
  }
  if (W == H)
    return launch_strip_pipev_wt<H, E, CW, MINB_A, H, false>(x, residual, dtab, addend, out, B, W, s);
  return launch_strip_pipev_wt<H, E, CW, MINB_A, 0, false>(x, residual, dtab, addend, out, B, W, s);
}

static bool tma_ok(const float* x, const float* addend, const float* dtab) {
  return (((uintptr_t)x | (uintptr_t)addend | (uintptr_t)dtab) & 15u) == 0;
}

// ---- strip (column) launches ------------------------------------------------
template <int H, int E, int CW, int MINB>
static int launch_strip_row_cfg(const float* x, const float* residual, const float* dtab,
                                const float* addend, float* out, int B, int W, cudaStream_t s) {
  typedef LineFFT<H, E, CW> L;
  const int nstrips = W / CW;
  if (W == H) {  // square slices (every shipped config): compile-time row pitch
    auto kern = dc_strip_row_kernel<H, E, CW, MINB, H>;
    CSMRI_TRY(set_smem(kern, L::kSmemBytes + L::kTwBytes));
    kern<<<B * nstrips, CW * L::T, L::kSmemBytes + L::kTwBytes, s>>>(x, residual, dtab, addend, out, W, nstrips);
  } else {
    auto kern = dc_strip_row_kernel<H, E, CW, MINB, 0>;
    CSMRI_TRY(set_smem(kern, L::kSmemBytes + L::kTwBytes));
    kern<<<B * nstrips, CW * L::T, L::kSmemBytes + L::kTwBytes, s>>>(x, residual, dtab, addend, out, W, nstrips);
  }
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}


// Kernel choice per column length H (measured, profiles/README.md):
//   variant 0 (default)  best TMA-pipelined kernel for the size
//   variant 1            one-column pipelined kernel (128 / 256; A/B comparison)
//   variant 2            direct kernel (also taken whenever a pointer is not
//                        16-byte aligned, which TMA requires)
static int launch_strip_row(const float* x, const float* residual, const float* dtab,
                            const float* addend, float* out, int B, int H, int W,
                            cudaStream_t s) {
  const bool tma = g_strip_variant != 2 && tma_ok(x, addend, dtab);
  switch (H) {
    case 32: return launch_strip_row_cfg<32, 8, 32, 1>(x, residual, dtab, addend, out, B, W, s);
    case 64:
      if (tma) return launch_strip_pipev_cfg<64, 8, 32, 4, 6>(x, residual, dtab, addend, out, B, W, s);
      return launch_strip_row_cfg<64, 8, 32, 1>(x, residual, dtab, addend, out, B, W, s);
    case 128:
      if (tma && g_strip_variant == 1)
        return launch_strip_pipe_cfg<128, 16, 32, 2>(x, residual, dtab, addend, out, B, W, s);
      if (tma) return launch_strip_pipev_cfg<128, 16, 32, 2, 3>(x, residual, dtab, addend, out, B, W, s);
      return launch_strip_row_cfg<128, 16, 32, 1>(x, residual, dtab, addend, out, B, W, s);
    case 256:
      if (tma && g_strip_variant == 1)
        return launch_strip_pipe_cfg<256, 16, 16, 2, 3>(x, residual, dtab, addend, out, B, W, s);
      // two columns per thread, 16-column tiles, 128-thread persistent CTAs,
      // two (forward) / three (adjoint) per SM
      if (tma) return launch_strip_pipev_cfg<256, 16, 16, 2, 3>(x, residual, dtab, addend, out, B, W, s);
      return launch_strip_row_cfg<256, 16, 16, 4>(x, residual, dtab, addend, out, B, W, s);
    // 320: a 40 KiB strip per buffer leaves room for one prefetching CTA of 4
    // warps; with the x tile landing in the exchange buffer (INPL) two (forward)
    // / three (adjoint) CTAs fit: forward 81 -> 68 us at B=164.  Variant 1 = the
    // prefetching layout, kept for A/B runs.  At 512 / 1024 both layouts time
    // the same (issue-bound, not occupancy-bound: profiles/README.md), so they
    // keep the prefetching one.
    case 320:
      if (tma && g_strip_variant == 1)
        return launch_strip_pipe_cfg<320, 40, 16, 1, 2>(x, residual, dtab, addend, out, B, W, s);
      if (tma)
        return launch_strip_pipe_cfg<320, 40, 16, 2, 3, true, true>(x, residual, dtab, addend, out, B, W, s);
      return launch_strip_row_cfg<320, 40, 32, 1>(x, residual, dtab, addend, out, B, W, s);
    case 512:
      // (two columns per thread at 32 points each - 253-255 registers, 128-thread CTAs,
      // no spills - was measured in round 2: 0.72 vs 0.79 of peak; not kept)
      if (tma) return launch_strip_pipe_cfg<512, 32, 16, 1>(x, residual, dtab, addend, out, B, W, s);
      return launch_strip_row_cfg<512, 32, 16, 1>(x, residual, dtab, addend, out, B, W, s);
    case 1024:
      if (tma) return launch_strip_pipe_cfg<1024, 32, 8, 1>(x, residual, dtab, addend, out, B, W, s);
      return launch_strip_row_cfg<1024, 32, 16, 1>(x, residual, dtab, addend, out, B, W, s);
  }
  return fail(CSMRI_E_SHAPE, "unsupported H=%d", H);
}

template <int H, int E, int CW, int MINB>
static int launch_strip_dense_cfg(const float* hyb, const float* k0, const float* mask, float* out,
                                  int B, int W, float sc, float nv, bool noisy, bool adj,
                                  cudaStream_t s) {
  typedef LineFFT<H, E, CW> L;
  const int nstrips = W / CW;
  const dim3 grid(B * nstrips), block(CW * L::T);
  constexpr int smem = L::kSmemBytes + L::kTwBytes;
#define CSMRI_DENSE(N_, A_, WT_)                                                      \
  {                                                                                   \
    auto kern = dc_strip_dense_kernel<H, E, CW, N_, A_, WT_, MINB>;                   \
    CSMRI_TRY(set_smem(kern, smem));                                                  \
    kern<<<grid, block, smem, s>>>(hyb, k0, mask, out, W, nstrips, sc, nv);           \
  }
  if (W == H) {
    if (noisy && adj) CSMRI_DENSE(true, true, H)
    else if (noisy) CSMRI_DENSE(true, false, H)
    else if (adj) CSMRI_DENSE(false, true, H)
    else CSMRI_DENSE(false, false, H)
  } else {
    if (noisy && adj) CSMRI_DENSE(true, true, 0)
    else if (noisy) CSMRI_DENSE(true, false, 0)
    else if (adj) CSMRI_DENSE(false, true, 0)
    else CSMRI_DENSE(false, false, 0)
  }
#undef CSMRI_DENSE
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

static int launch_strip_dense(const float* hyb, const float* k0, const float* mask, float* out,
                              int B, int H, int W, float sc, float nv, bool noisy, bool adj,
                              cudaStream_t s) {
  switch (H) {
    case 32: return launch_strip_dense_cfg<32, 8, 32, 1>(hyb, k0, mask, out, B, W, sc, nv, noisy, adj, s);
    case 64: return launch_strip_dense_cfg<64, 8, 32, 1>(hyb, k0, mask, out, B, W, sc, nv, noisy, adj, s);
    case 128: return launch_strip_dense_cfg<128, 16, 32, 2>(hyb, k0, mask, out, B, W, sc, nv, noisy, adj, s);
    case 256: return launch_strip_dense_cfg<256, 16, 16, 3>(hyb, k0, mask, out, B, W, sc, nv, noisy, adj, s);
    case 512: return launch_strip_dense_cfg<512, 32, 16, 1>(hyb, k0, mask, out, B, W, sc, nv, noisy, adj, s);
    case 1024: return launch_strip_dense_cfg<1024, 32, 16, 1>(hyb, k0, mask, out, B, W, sc, nv, noisy, adj, s);
    case 320: return launch_strip_dense_cfg<320, 40, 32, 1>(hyb, k0, mask, out, B, W, sc, nv, noisy, adj, s);
  }
  return fail(CSMRI_E_SHAPE, "unsupported H=%d", H);
}

template <int H, int E, int CW, int MINB>
static int launch_fft_strip_cfg(const float* in, float* out, int B, int W, float scale, bool inv,
                                const unsigned char* rows, cudaStream_t s) {
  typedef LineFFT<H, E, CW> L;
  const int nstrips = W / CW;
  constexpr int smem = L::kSmemBytes + L::kTwBytes;
#define CSMRI_FSTRIP(I_, WT_)                                                                \
  {                                                                                          \
    auto kern = fft_strip_kernel<H, E, CW, I_, WT_, MINB>;                                   \
    CSMRI_TRY(set_smem(kern, smem));                                                         \
    kern<<<B * nstrips, CW * L::T, smem, s>>>(in, out, W, nstrips, scale, rows);             \
  }
  if (W == H) {
    if (inv) CSMRI_FSTRIP(true, H) else CSMRI_FSTRIP(false, H)
  } else {
    if (inv) CSMRI_FSTRIP(true, 0) else CSMRI_FSTRIP(false, 0)
  }
#undef CSMRI_FSTRIP
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

static int launch_fft_strip(const float* in, float* out, int B, int H, int W, float scale, bool inv,
                            const unsigned char* rows, cudaStream_t s) {
  switch (H) {
    case 32: return launch_fft_strip_cfg<32, 8, 32, 1>(in, out, B, W, scale, inv, rows, s);
    case 64: return launch_fft_strip_cfg<64, 8, 32, 1>(in, out, B, W, scale, inv, rows, s);
    case 128: return launch_fft_strip_cfg<128, 16, 32, 2>(in, out, B, W, scale, inv, rows, s);
    case 256: return launch_fft_strip_cfg<256, 16, 16, 4>(in, out, B, W, scale, inv, rows, s);
    case 512: return launch_fft_strip_cfg<512, 32, 16, 1>(in, out, B, W, scale, inv, rows, s);
    case 1024: return launch_fft_strip_cfg<1024, 32, 16, 1>(in, out, B, W, scale, inv, rows, s);
    case 320: return launch_fft_strip_cfg<320, 40, 32, 1>(in, out, B, W, scale, inv, rows, s);
  }
  return fail(CSMRI_E_SHAPE, "unsupported H=%d", H);
}

template <int H, int E, int CW, int MINB>
static int launch_undersample_strip_cfg(const float* img, const unsigned char* rows, float* target,
                                        float* hyb, float* inp, int B, int W, cudaStream_t s) {
  typedef LineFFT<H, E, CW> L;
  const int nstrips = W / CW;
  constexpr int smem = L::kSmemBytes + L::kTwBytes;
  if (W == H) {
    auto kern = undersample_strip_kernel<H, E, CW, H, MINB>;
    CSMRI_TRY(set_smem(kern, smem));
    kern<<<B * nstrips, CW * L::T, smem, s>>>(img, rows, target, hyb, inp, W, nstrips);
  } else {
    auto kern = undersample_strip_kernel<H, E, CW, 0, MINB>;
    CSMRI_TRY(set_smem(kern, smem));
    kern<<<B * nstrips, CW * L::T, smem, s>>>(img, rows, target, hyb, inp, W, nstrips);
  }
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

// (E, T) per H must match strip_radix(H): hyb doubles as the DC plan's addend
static int launch_undersample_strip(const float* img, const unsigned char* rows, float* target,
                                    float* hyb, float* inp, int B, int H, int W, cudaStream_t s) {
  switch (H) {
    case 32: return launch_undersample_strip_cfg<32, 8, 32, 1>(img, rows, target, hyb, inp, B, W, s);
    case 64: return launch_undersample_strip_cfg<64, 8, 32, 1>(img, rows, target, hyb, inp, B, W, s);
    case 128: return launch_undersample_strip_cfg<128, 16, 32, 2>(img, rows, target, hyb, inp, B, W, s);
    case 256: return launch_undersample_strip_cfg<256, 16, 16, 3>(img, rows, target, hyb, inp, B, W, s);
    case 512: return launch_undersample_strip_cfg<512, 32, 16, 1>(img, rows, target, hyb, inp, B, W, s);
    case 1024: return launch_undersample_strip_cfg<1024, 32, 16, 1>(img, rows, target, hyb, inp, B, W, s);
    case 320: return launch_undersample_strip_cfg<320, 40, 32, 1>(img, rows, target, hyb, inp, B, W, s);
  }
  return fail(CSMRI_E_SHAPE, "unsupported H=%d", H);
}

// ---- row launches -----------------------------------------------------------
template <int W, int E, int R>
static int launch_fft_rows_cfg(const float* in, const float* aux, float* out, int B, int H,
                               float scale, float cmulv, bool inv, int pre, cudaStream_t s,
                               const unsigned char* rowsel = nullptr, float* mask_out = nullptr) {
  constexpr int T = W / E;
  constexpr int smem = (R * E * (T + 1) + W) * (int)sizeof(cf);
  if (H % R != 0) return fail(CSMRI_E_SHAPE, "H=%d is not a multiple of %d", H, R);
  const dim3 grid(B * (H / R)), block(R * T);
#define CSMRI_ROWS(I_, P_)                                              \
  {                                                                     \
    auto kern = fft_rows_kernel<W, E, R, I_, P_>;                       \
    CSMRI_TRY(set_smem(kern, smem));                                    \
    kern<<<grid, block, smem, s>>>(in, aux, out, H, scale, cmulv,       \
                                   rowsel, mask_out);                   \
  }
  if (!inv && pre == 0) CSMRI_ROWS(false, 0)
  else if (!inv && pre == 1) CSMRI_ROWS(false, 1)
  else if (!inv && pre == 3) CSMRI_ROWS(false, 3)
  else if (!inv && pre == 4) CSMRI_ROWS(false, 4)
  else if (inv && pre == 0) CSMRI_ROWS(true, 0)
  else if (inv && pre == 2) CSMRI_ROWS(true, 2)
  else if (inv && pre == 5) CSMRI_ROWS(true, 5)
  else return fail(CSMRI_E_ARG, "row kernel variant inv=%d pre=%d not instantiated", (int)inv, pre);
#undef CSMRI_ROWS
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

static int launch_fft_rows(const float* in, const float* aux, float* out, int B, int H, int W,
                           float scale, float cmulv, bool inv, int pre, cudaStream_t s,
                           const unsigned char* rowsel = nullptr, float* mask_out = nullptr) {
  if ((pre == 4) != (rowsel != nullptr && mask_out != nullptr) && pre != 5)
    return fail(CSMRI_E_ARG, "row selection needs pre == 4, a row table and a mask destination");
  if (pre == 5 && (rowsel == nullptr || mask_out != nullptr))
    return fail(CSMRI_E_ARG, "compact input needs a row table and no mask destination");
  switch (W) {
    case 32: return launch_fft_rows_cfg<32, 8, 32>(in, aux, out, B, H, scale, cmulv, inv, pre, s, rowsel, mask_out);   // T=4
    case 64: return launch_fft_rows_cfg<64, 8, 32>(in, aux, out, B, H, scale, cmulv, inv, pre, s, rowsel, mask_out);   // T=8
    case 128: return launch_fft_rows_cfg<128, 16, 32>(in, aux, out, B, H, scale, cmulv, inv, pre, s, rowsel, mask_out);  // T=8
    case 256: return launch_fft_rows_cfg<256, 16, 16>(in, aux, out, B, H, scale, cmulv, inv, pre, s, rowsel, mask_out);  // T=16
    case 512: return launch_fft_rows_cfg<512, 32, 8>(in, aux, out, B, H, scale, cmulv, inv, pre, s, rowsel, mask_out);   // T=16
    case 1024: return launch_fft_rows_cfg<1024, 32, 8>(in, aux, out, B, H, scale, cmulv, inv, pre, s, rowsel, mask_out);  // T=32
    case 320: return launch_fft_rows_cfg<320, 40, 16>(in, aux, out, B, H, scale, cmulv, inv, pre, s, rowsel, mask_out);  // T=8
  }
  return fail(CSMRI_E_SHAPE, "unsupported W=%d", W);
}

}  // namespace csmri

using namespace csmri;

extern "C" {

int csmri_version(void) { return 100; }

const char* csmri_last_error(void) { return g_err; }

int csmri_init(void) { return ensure_init(); }

// tuning knob for the benchmark harness (not part of the reference-facing ABI)
int csmri_set_variant(int v) {
  g_strip_variant = v;
  return CSMRI_OK;
}
int csmri_set_trace(void* device_buffer) {
  g_trace = (long long*)device_buffer;
  g_trace_launch = 0;
  return CSMRI_OK;
}
int csmri_set_tuning(int key, int value) {
  if (key == 0) g_strip_variant = value;
  else if (key == 4) g_use_pdl = value;
  else if (key == 5) g_wgrad_cot = value == 8 ? 8 : 4;
  else if (key == 6) g_tc_debug = value & 7;
  else if (key == 7) g_wgrad_tc = value != 0;
  else if (key == 8) g_thin_tma = value != 0;
  else if (key == 9) g_general_chunk_mib = value < 0 ? 0 : value;
  else if (key == 10) g_conv_pdl = value < 0 || value > 2 ? 0 : value;
  else return fail(CSMRI_E_ARG, "unknown tuning key %d", key);
  return CSMRI_OK;
}

size_t csmri_dc_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)B * 2 * H * W * sizeof(float);
}

int csmri_dc_prepare(const float* k0, const float* mask, int B, int H, int W, float noise_lvl,
                     float* dtab, float* addend, int* row_constant, void* scratch, void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  CSMRI_TRY(check_ptr(mask, "mask"));
  CSMRI_TRY(check_ptr(dtab, "dtab"));
  CSMRI_TRY(check_ptr(row_constant, "row_constant"));
  CSMRI_TRY(ensure_init());
  cudaStream_t s = (cudaStream_t)stream;
  const bool noisy = noise_lvl != 0.0f;
  set_flag_kernel<<<1, 1, 0, s>>>(row_constant, 1);
  {
    const int warps = B * H, threads = 256;
    const int blocks = (warps * 32 + threads - 1) / threads;
    mask_rows_kernel<<<blocks, threads, 0, s>>>(mask, B, H, W, noise_lvl, noisy ? 1 : 0,
                                                strip_radix(H), dtab, row_constant);
  }
  CSMRI_CUDA(cudaGetLastError());
  if (addend != nullptr) {
    // addend = iFFT_W(c*k0) / sqrt(H*W): the k0 term in hybrid (k_H, w) space.  The
    // strip kernels add it between their forward and inverse column passes, so
    // only the row transform is left to do here.
    CSMRI_TRY(check_ptr(k0, "k0"));
    const float sc = 1.0f / sqrtf((float)H * (float)W);
    if (noisy) {
      CSMRI_TRY(launch_fft_rows(k0, mask, addend, B, H, W, sc, noise_lvl / (1.0f + noise_lvl),
                                true, 2, s));
    } else {
      CSMRI_TRY(launch_fft_rows(k0, nullptr, addend, B, H, W, sc, 0.0f, true, 0, s));
    }
  }
  (void)scratch;
  return CSMRI_OK;
}

int csmri_dc_prepare_lines(const float* k0_lines, const unsigned char* rows, int B, int H, int W,
                           int L, float noise_lvl, float* dtab, float* addend, int* lines_ok,
                           void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  if (L <= 0 || L > H) return fail(CSMRI_E_SHAPE, "L must be in [1, H], got %d", L);
  CSMRI_TRY(check_ptr(k0_lines, "k0_lines"));
  if (rows == nullptr) return fail(CSMRI_E_NULLPTR, "rows is NULL");
  CSMRI_TRY(check_ptr(dtab, "dtab"));
  CSMRI_TRY(check_ptr(addend, "addend"));
  CSMRI_TRY(check_ptr(lines_ok, "lines_ok"));
  CSMRI_TRY(ensure_init());
  cudaStream_t s = (cudaStream_t)stream;
  const bool noisy = noise_lvl != 0.0f;
  set_flag_kernel<<<1, 1, 0, s>>>(lines_ok, 1);
  dtab_from_rows_checked_kernel<<<B, H < 256 ? H : 256, 0, s>>>(rows, H, strip_radix(H), L,
                                                               noise_lvl, noisy ? 1 : 0, dtab,
                                                               lines_ok);
  CSMRI_CUDA(cudaGetLastError());
  // addend = iFFT_W(c * k0) / sqrt(H*W) on the sampled rows, zero elsewhere; on a
  // sampled row c = 1 (noiseless) or v/(1+v) (noisy, m = 1), see csmri_dc_prepare
  float sc = 1.0f / sqrtf((float)H * (float)W);
  if (noisy) sc *= noise_lvl / (1.0f + noise_lvl);
  return launch_fft_rows(k0_lines, nullptr, addend, B, H, W, sc, (float)L, true, 5, s, rows,
                         nullptr);
}

int csmri_dc_forward_cartesian(const float* x, const float* residual, const float* dtab,
                               const float* addend, float* out, int B, int H, int W,
                               void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  if (H > 1024) return fail(CSMRI_E_SHAPE, "H too large");
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(dtab, "dtab"));
  CSMRI_TRY(check_ptr(out, "out"));
  if (out == x || out == residual || out == addend)
    return fail(CSMRI_E_ARG, "out must not alias an input");
  CSMRI_TRY(ensure_init());
  return launch_strip_row(x, residual, dtab, addend, out, B, H, W, (cudaStream_t)stream);
}

int csmri_dc_adjoint_cartesian(const float* grad_out, const float* dtab, float* grad_x, int B,
                               int H, int W, void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  CSMRI_TRY(check_ptr(grad_out, "grad_out"));
  CSMRI_TRY(check_ptr(dtab, "dtab"));
  CSMRI_TRY(check_ptr(grad_x, "grad_x"));
  if (grad_x == grad_out) return fail(CSMRI_E_ARG, "grad_x must not alias grad_out");
  CSMRI_TRY(ensure_init());
  return launch_strip_row(grad_out, nullptr, dtab, nullptr, grad_x, B, H, W,
                          (cudaStream_t)stream);
}

// slices per chunk of the general path: g_general_chunk_mib of hybrid scratch
static int general_chunk(int B, int H, int W) {
  if (g_general_chunk_mib <= 0) return B;
  const size_t per = (size_t)2 * H * W * sizeof(float);
  const size_t n = ((size_t)g_general_chunk_mib << 20) / per;
  return n < 1 ? 1 : (n > (size_t)B ? B : (int)n);
}

int csmri_dc_forward_general(const float* x, const float* residual, const float* k0,
                             const float* mask, float* out, int B, int H, int W, float noise_lvl,
                             void* scratch, void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(k0, "k0"));
  CSMRI_TRY(check_ptr(mask, "mask"));
  CSMRI_TRY(check_ptr(out, "out"));
  CSMRI_TRY(check_ptr(scratch, "scratch"));
  CSMRI_TRY(ensure_init());
  cudaStream_t s = (cudaStream_t)stream;
  float* hyb = (float*)scratch;
  const float sc = 1.0f / sqrtf((float)H * (float)W);
  const bool noisy = noise_lvl != 0.0f;
  // Tuning key 9 runs the three passes chunk by chunk so that a chunk's hybrid scratch could
  // stay in the 126 MB L2 between passes.  Measured (profiles/r2_general_chunk_sweep.json): every
  // chunk size is slower than whole-batch passes (256^2, B=256: 186 us unchunked, 190 us with two
  // 64 MiB chunks, 265 us at 16 MiB) - the extra kernel boundaries cost more than L2 hits save,
  // so the default is one chunk.
  const size_t slice = (size_t)2 * H * W;
  for (int b0 = 0, nb = general_chunk(B, H, W); b0 < B; b0 += nb) {
    const int n = B - b0 < nb ? B - b0 : nb;
    const size_t o = (size_t)b0 * slice;
    CSMRI_TRY(launch_fft_rows(x + o, residual ? residual + o : nullptr, hyb + o, n, H, W, 1.0f, 0.0f,
                              false, residual ? 1 : 0, s));
    CSMRI_TRY(launch_strip_dense(hyb + o, k0 + o, mask + o, hyb + o, n, H, W, sc, noise_lvl, noisy,
                                 false, s));
    CSMRI_TRY(launch_fft_rows(hyb + o, nullptr, out + o, n, H, W, 1.0f, 0.0f, true, 0, s));
  }
  return CSMRI_OK;
}

int csmri_dc_adjoint_general(const float* grad_out, const float* mask, float* grad_x, int B, int H,
                             int W, float noise_lvl, void* scratch, void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  CSMRI_TRY(check_ptr(grad_out, "grad_out"));
  CSMRI_TRY(check_ptr(mask, "mask"));
  CSMRI_TRY(check_ptr(grad_x, "grad_x"));
  CSMRI_TRY(check_ptr(scratch, "scratch"));
  CSMRI_TRY(ensure_init());
  cudaStream_t s = (cudaStream_t)stream;
  float* hyb = (float*)scratch;
  const float sc = 1.0f / sqrtf((float)H * (float)W);
  const bool noisy = noise_lvl != 0.0f;
  const size_t slice = (size_t)2 * H * W;
  for (int b0 = 0, nb = general_chunk(B, H, W); b0 < B; b0 += nb) {   // see csmri_dc_forward_general
    const int n = B - b0 < nb ? B - b0 : nb;
    const size_t o = (size_t)b0 * slice;
    CSMRI_TRY(launch_fft_rows(grad_out + o, nullptr, hyb + o, n, H, W, 1.0f, 0.0f, false, 0, s));
    CSMRI_TRY(launch_strip_dense(hyb + o, nullptr, mask + o, hyb + o, n, H, W, sc, noise_lvl, noisy,
                                 true, s));
    CSMRI_TRY(launch_fft_rows(hyb + o, nullptr, grad_x + o, n, H, W, 1.0f, 0.0f, true, 0, s));
  }
  return CSMRI_OK;
}

int csmri_dc_forward(const float* x, const float* residual, const float* k0, const float* mask,
                     const float* dtab, const float* addend, float* out, int B, int H, int W,
                     float noise_lvl, int mask_is_row_constant, void* scratch, void* stream) {
  if (mask_is_row_constant) {
    CSMRI_TRY(check_ptr(addend, "addend"));
    return csmri_dc_forward_cartesian(x, residual, dtab, addend, out, B, H, W, stream);
  }
  return csmri_dc_forward_general(x, residual, k0, mask, out, B, H, W, noise_lvl, scratch, stream);
}

int csmri_dc_adjoint(const float* grad_out, const float* mask, const float* dtab, float* grad_x,
                     int B, int H, int W, float noise_lvl, int mask_is_row_constant, void* scratch,
                     void* stream) {
  if (mask_is_row_constant)
    return csmri_dc_adjoint_cartesian(grad_out, dtab, grad_x, B, H, W, stream);
  return csmri_dc_adjoint_general(grad_out, mask, grad_x, B, H, W, noise_lvl, scratch, stream);
}

int csmri_undersample(const float* img, const unsigned char* rows, float* inp, float* kspace,
                      float* mask, float* target, float* dtab, float* addend, int B, int H, int W,
                      void* scratch, void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  CSMRI_TRY(check_ptr(img, "img"));
  if (rows == nullptr) return fail(CSMRI_E_NULLPTR, "rows is NULL");
  CSMRI_TRY(check_ptr(inp, "inp"));
  CSMRI_TRY(check_ptr(kspace, "kspace"));
  CSMRI_TRY(check_ptr(mask, "mask"));
  CSMRI_TRY(check_ptr(target, "target"));
  CSMRI_TRY(check_ptr(scratch, "scratch"));
  CSMRI_TRY(ensure_init());
  cudaStream_t s = (cudaStream_t)stream;
  if ((dtab == nullptr) != (addend == nullptr))
    return fail(CSMRI_E_ARG, "dtab and addend must be given together");
  if (dtab != nullptr) {
    CSMRI_TRY(check_ptr(dtab, "dtab"));
    CSMRI_TRY(check_ptr(addend, "addend"));
  }
  // x_f = fft2(x, ortho); x_fu = mask * x_f; x_u = ifft2(x_fu, ortho)
  // (compressed_sensing.py:509-511) for a row-constant mask, in two passes:
  // columns (target, masked hybrid rows, zero-filled image), then the sampled
  // rows only (k-space, dense mask).  The masked hybrid rows, scaled by 1/H,
  // are exactly the k0 term the DC strip kernels add (csmri_dc_prepare's
  // `addend`): when the caller wants the DC plan they are written there
  // instead of into scratch, for free.
  float* hyb = addend != nullptr ? addend : (float*)scratch;
  CSMRI_TRY(launch_undersample_strip(img, rows, target, hyb, inp, B, H, W, s));
  CSMRI_TRY(launch_fft_rows(hyb, nullptr, kspace, B, H, W, sqrtf((float)H) / sqrtf((float)W), 0.0f,
                            false, 4, s, rows, mask));
  if (dtab != nullptr)
    dtab_from_rows_kernel<<<(B * H + 255) / 256, 256, 0, s>>>(rows, B, H, strip_radix(H), dtab);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_fft2(const float* x, float* out, int B, int H, int W, int inverse, void* scratch,
               void* stream) {
  CSMRI_TRY(check_shape(B, H, W));
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(out, "out"));
  CSMRI_TRY(check_ptr(scratch, "scratch"));
  CSMRI_TRY(ensure_init());
  cudaStream_t s = (cudaStream_t)stream;
  const float sc = 1.0f / sqrtf((float)H * (float)W);
  CSMRI_TRY(launch_fft_rows(x, nullptr, (float*)scratch, B, H, W, 1.0f, 0.0f, inverse != 0, 0, s));
  CSMRI_TRY(launch_fft_strip((const float*)scratch, out, B, H, W, sc, inverse != 0, nullptr, s));
  return CSMRI_OK;
}

int csmri_magnitude_clamp(const float* x, float* out, int B, int H, int W, float lo, float hi,
                          void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return fail(CSMRI_E_SHAPE, "bad shape %dx%dx%d", B, H, W);
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(out, "out"));
  const size_t plane = (size_t)H * W, total = plane * B;
  const int blocks = (int)((total + 1023) / 1024 < 148 * 16 ? (total + 1023) / 1024 : 148 * 16);
  magnitude_clamp_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, out, plane, total, lo, hi);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_psnr_sum(const float* pred, const float* target, double* sum_sq, int B, int H, int W,
                   float lo, float hi, void* stream) {
  if (B <= 0 || H <= 0 || W <= 0) return fail(CSMRI_E_SHAPE, "bad shape %dx%dx%d", B, H, W);
  CSMRI_TRY(check_ptr(pred, "pred"));
  CSMRI_TRY(check_ptr(target, "target"));
  if (sum_sq == nullptr || ((uintptr_t)sum_sq & 7u) != 0)
    return fail(CSMRI_E_ALIGN, "sum_sq must be a non-NULL 8-byte aligned device double");
  const size_t plane = (size_t)H * W, total = plane * B;
  const int blocks = (int)((total + 1023) / 1024 < 148 * 8 ? (total + 1023) / 1024 : 148 * 8);
  CSMRI_CUDA(cudaMemsetAsync(sum_sq, 0, sizeof(double), (cudaStream_t)stream));
  psnr_sum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, sum_sq, plane, total, lo,
                                                           hi);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

// ---- refinement-path pointwise ops (refine_ops.cuh) ---------------------------
// ---- loader tail: index plumbing of CenterCropInKspace + max normalisation ----
int csmri_shift_crop(const float* in, float* out, int B, int in_ch, int IH, int IW, int out_ch,
                     int OH, int OW, int in_roll_y, int in_roll_x, int off_y, int off_x,
                     int out_roll_y, int out_roll_x, float* absmax, void* stream) {
  if (B <= 0 || B > 65535 || IH <= 0 || IW <= 0 || OH <= 0 || OH > 65535 || OW <= 0)
    return fail(CSMRI_E_SHAPE, "bad shape: B=%d in %dx%d out %dx%d", B, IH, IW, OH, OW);
  if ((in_ch != 1 && in_ch != 2) || (out_ch != 1 && out_ch != 2) || (out_ch == 1 && in_ch != 2))
    return fail(CSMRI_E_ARG, "channels must be 1 -> 2, 2 -> 2 or 2 -> 1 (got %d -> %d)", in_ch, out_ch);
  if (in_roll_y < 0 || in_roll_y >= IH || in_roll_x < 0 || in_roll_x >= IW || out_roll_y < 0 ||
      out_roll_y >= OH || out_roll_x < 0 || out_roll_x >= OW)
    return fail(CSMRI_E_ARG, "rolls must be reduced modulo the axis length");
  if (absmax != nullptr && out_ch != 1)
    return fail(CSMRI_E_ARG, "absmax needs the magnitude output (out_ch = 1)");
  CSMRI_TRY(check_ptr(in, "in"));
  CSMRI_TRY(check_ptr(out, "out"));
  if (in == out) return fail(CSMRI_E_ARG, "out must not alias in");
  cudaStream_t s = (cudaStream_t)stream;
  const ShiftCropAxis ay = {IH, OH, in_roll_y, off_y, out_roll_y};
  const ShiftCropAxis ax = {IW, OW, in_roll_x, off_x, out_roll_x};
  const dim3 grid((OW + 255) / 256, (OH + kShiftCropRows - 1) / kShiftCropRows, B);
  unsigned* km = reinterpret_cast<unsigned*>(absmax);
  if (km != nullptr) CSMRI_CUDA(cudaMemsetAsync(km, 0, (size_t)B * sizeof(unsigned), s));
  if (in_ch == 1) shift_crop_kernel<1, 2><<<grid, 256, 0, s>>>(in, out, ay, ax, nullptr);
  else if (out_ch == 2) shift_crop_kernel<2, 2><<<grid, 256, 0, s>>>(in, out, ay, ax, nullptr);
  else shift_crop_kernel<2, 1><<<grid, 256, 0, s>>>(in, out, ay, ax, km);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_plane_absmax(const float* x, float* absmax, int planes, int n, void* stream) {
  if (planes <= 0 || planes > 65535 || n <= 0)
    return fail(CSMRI_E_SHAPE, "bad plane shape: planes=%d n=%d", planes, n);
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(absmax, "absmax"));
  cudaStream_t s = (cudaStream_t)stream;
  CSMRI_CUDA(cudaMemsetAsync(absmax, 0, (size_t)planes * sizeof(float), s));
  const dim3 grid((n + 2047) / 2048 > 32 ? 32 : (n + 2047) / 2048, planes);
  plane_absmax_kernel<<<grid, 256, 0, s>>>(x, reinterpret_cast<unsigned*>(absmax), n);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_plane_divide(const float* x, const float* denom, float* out, int planes, int n,
                       void* stream) {
  if (planes <= 0 || planes > 65535 || n <= 0)
    return fail(CSMRI_E_SHAPE, "bad plane shape: planes=%d n=%d", planes, n);
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(denom, "denom"));
  CSMRI_TRY(check_ptr(out, "out"));
  const dim3 grid((n + 1023) / 1024 > 64 ? 64 : (n + 1023) / 1024, planes);
  if (n % 4 == 0 && (((uintptr_t)x | (uintptr_t)out) & 15u) == 0)
    plane_divide_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(x, denom, out, n);
  else
    plane_divide_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x, denom, out, n);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

static const int kRefinePartials = 32;

static int plane_chunks(int n) {
  const int c = (n + 2047) / 2048;
  return c < 1 ? 1 : (c > kRefinePartials ? kRefinePartials : c);
}

static int plane_minmax_impl(const float* x, float* minimum, float* maximum, int planes, int n,
                             long long pitch, cudaStream_t s) {
  unsigned* kmin = reinterpret_cast<unsigned*>(minimum);
  unsigned* kmax = reinterpret_cast<unsigned*>(maximum);
  minmax_init_kernel<<<(planes + 255) / 256, 256, 0, s>>>(kmin, kmax, planes);
  minmax_reduce_kernel<<<dim3(plane_chunks(n), planes), 256, 0, s>>>(x, kmin, kmax, n, pitch);
  minmax_finish_kernel<<<(planes + 255) / 256, 256, 0, s>>>(minimum, maximum, planes);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_plane_minmax(const float* x, float* minimum, float* maximum, int planes, int n,
                       long long pitch, void* stream) {
  if (planes <= 0 || planes > 65535 || n <= 0 || pitch < n)
    return fail(CSMRI_E_SHAPE, "bad plane shape: planes=%d n=%d pitch=%lld", planes, n, pitch);
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(minimum, "minimum"));
  CSMRI_TRY(check_ptr(maximum, "maximum"));
  return plane_minmax_impl(x, minimum, maximum, planes, n, pitch, (cudaStream_t)stream);
}

int csmri_plane_scale(const float* x, const float* minimum, const float* maximum, float* out,
                      int planes, int n, long long pitch_in, long long pitch_out, int mode,
                      void* stream) {
  if (planes <= 0 || planes > 65535 || n <= 0 || pitch_in < n || pitch_out < n)
    return fail(CSMRI_E_SHAPE, "bad plane shape: planes=%d n=%d pitch=%lld/%lld", planes, n,
                pitch_in, pitch_out);
  if (mode < 0 || mode > 2) return fail(CSMRI_E_ARG, "mode must be 0, 1 or 2 (got %d)", mode);
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(minimum, "minimum"));
  CSMRI_TRY(check_ptr(maximum, "maximum"));
  CSMRI_TRY(check_ptr(out, "out"));
  const dim3 grid((n + 1023) / 1024 > 64 ? 64 : (n + 1023) / 1024, planes);
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) plane_map_kernel<0><<<grid, 256, 0, s>>>(x, minimum, maximum, out, n, pitch_in, pitch_out);
  else if (mode == 1) plane_map_kernel<1><<<grid, 256, 0, s>>>(x, minimum, maximum, out, n, pitch_in, pitch_out);
  else plane_map_kernel<2><<<grid, 256, 0, s>>>(x, minimum, maximum, out, n, pitch_in, pitch_out);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_refine_partials(void) { return kRefinePartials; }

static int check_refine_shape(int B, int H, int W) {
  if (B <= 0 || B > 65535 || H <= 0 || W <= 0 || (long long)H * W > 0x3fffffff)
    return fail(CSMRI_E_SHAPE, "bad shape %dx%dx%d", B, H, W);
  return CSMRI_OK;
}

int csmri_refine_real_penalty_add(const float* pretrained, const float* learnable,
                                  const float* scale, float* pred, float* minimum, float* maximum,
                                  int B, int H, int W, void* stream) {
  CSMRI_TRY(check_refine_shape(B, H, W));
  CSMRI_TRY(check_ptr(pretrained, "pretrained"));
  CSMRI_TRY(check_ptr(learnable, "learnable"));
  CSMRI_TRY(check_ptr(scale, "scale"));
  CSMRI_TRY(check_ptr(pred, "pred"));
  CSMRI_TRY(check_ptr(minimum, "minimum"));
  CSMRI_TRY(check_ptr(maximum, "maximum"));
  const int n = H * W;
  cudaStream_t s = (cudaStream_t)stream;
  CSMRI_TRY(plane_minmax_impl(pretrained, minimum, maximum, B, n, 2LL * n, s));
  const dim3 grid((n + 1023) / 1024 > 64 ? 64 : (n + 1023) / 1024, B);
  refine_forward_kernel<<<grid, 256, 0, s>>>(pretrained, learnable, scale, minimum, maximum, pred, n);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_refine_real_penalty_add_backward(const float* grad_pred, const float* learnable,
                                           const float* scale, const float* maximum,
                                           float* grad_learnable, float* grad_scale_partial, int B,
                                           int H, int W, void* stream) {
  CSMRI_TRY(check_refine_shape(B, H, W));
  CSMRI_TRY(check_ptr(grad_pred, "grad_pred"));
  CSMRI_TRY(check_ptr(learnable, "learnable"));
  CSMRI_TRY(check_ptr(scale, "scale"));
  CSMRI_TRY(check_ptr(maximum, "maximum"));
  CSMRI_TRY(check_ptr(grad_learnable, "grad_learnable"));
  CSMRI_TRY(check_ptr(grad_scale_partial, "grad_scale_partial"));
  refine_backward_kernel<<<dim3(kRefinePartials, B), 256, 0, (cudaStream_t)stream>>>(
      grad_pred, learnable, scale, maximum, grad_learnable, grad_scale_partial, H * W);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

// ---- 3x3 convolution weight gradient (conv_wgrad.cuh) --------------------------
static const int kWgMaxCtas = 640;   // >= 4 resident CTAs on each of up to 160 SMs

static bool wgrad_thin(int CI, int CO) { return (CI == 2 && CO == 32) || (CI == 32 && CO == 2); }

static int wgrad_check_channels(int CI, int CO) {
  if (wgrad_thin(CI, CO)) return CSMRI_OK;
  if (CI <= 0 || CO <= 0 || CI % kWgC != 0 || CO % kWgC != 0 || CI > 1024 || CO > 1024)
    return fail(CSMRI_E_SHAPE,
                "conv3x3_wgrad needs CI and CO multiples of %d, or 2 -> 32 / 32 -> 2 (got %d, %d)",
                kWgC, CI, CO);
  return CSMRI_OK;
}

static int wgrad_parts(int CI, int CO, int cap) {
  const int blocks = (CI / kWgC) * (CO / kWgC);
  const int parts = cap / blocks;
  return parts < 1 ? 1 : parts;
}

size_t csmri_conv3x3_wgrad_workspace_bytes(int CI, int CO) {
  if (wgrad_check_channels(CI, CO) != CSMRI_OK) return 0;
  if (wgrad_thin(CI, CO)) return (size_t)kWgMaxCtas * CI * CO * 9 * sizeof(float);
  return (size_t)wgrad_parts(CI, CO, kWgMaxCtas) * (CI / kWgC) * (CO / kWgC) * kWgBlock *
         sizeof(float);
}

static int conv3x3_wgrad_impl(const float* x, const float* dy, float* dw, float* db, void* workspace, int N,
                              int CI, int CO, int H, int W, int pad, void* stream);

int csmri_conv3x3_wgrad(const float* x, const float* dy, float* dw, void* workspace, int N, int CI,
                        int CO, int H, int W, int pad, void* stream) {
  return conv3x3_wgrad_impl(x, dy, dw, nullptr, workspace, N, CI, CO, H, W, pad, stream);
}

int csmri_conv3x3_wgrad_bias(const float* x, const float* dy, float* dw, float* db, void* workspace,
                             int N, int H, int W, void* stream) {
  CSMRI_TRY(check_ptr(db, "db"));
  if (!g_wgrad_tc || H % kWtcRows != 0 || W % kWtcPx != 0 || ((uintptr_t)x & 15u) != 0)
    return fail(CSMRI_E_SHAPE,
                "conv3x3_wgrad_bias is the tensor-core path only: H %% %d == 0, W %% %d == 0, x 16-byte "
                "aligned (got %dx%dx%d)", kWtcRows, kWtcPx, N, H, W);
  return conv3x3_wgrad_impl(x, dy, dw, db, workspace, N, kWtcC, kWtcC, H, W, 1, stream);
}

int csmri_conv3x3_wgrad_thin_bias(const float* x, const float* dy, float* dw, float* db, void* workspace,
                                  int N, int H, int W, void* stream) {
  CSMRI_TRY(check_ptr(db, "db"));
  return conv3x3_wgrad_impl(x, dy, dw, db, workspace, N, 2, 32, H, W, 1, stream);
}

int csmri_conv3x3_wgrad_thin_in_bias(const float* x, const float* dy, float* dw, float* db, void* workspace,
                                     int N, int H, int W, void* stream) {
  CSMRI_TRY(check_ptr(db, "db"));
  return conv3x3_wgrad_impl(x, dy, dw, db, workspace, N, 32, 2, H, W, 1, stream);
}

static int conv3x3_wgrad_impl(const float* x, const float* dy, float* dw, float* db, void* workspace, int N,
                              int CI, int CO, int H, int W, int pad, void* stream) {
  CSMRI_TRY(wgrad_check_channels(CI, CO));
  const bool thin = wgrad_thin(CI, CO);
  const int th = thin ? kWtRows : kWgTH;
  if (N <= 0 || H <= 0 || W <= 0 || H % th != 0 || W % kWgTW != 0)
    return fail(CSMRI_E_SHAPE, "conv3x3_wgrad needs H %% %d == 0 and W %% %d == 0 (got %dx%dx%d)",
                th, kWgTW, N, H, W);
  if (pad != 0 && pad != 1) return fail(CSMRI_E_ARG, "pad must be 0 or 1 (got %d)", pad);
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(dy, "dy"));
  CSMRI_TRY(check_ptr(dw, "dw"));
  CSMRI_TRY(check_ptr(workspace, "workspace"));
  if (((uintptr_t)dy & 15u) != 0) return fail(CSMRI_E_ALIGN, "dy is not 16-byte aligned");
  const int tiles_x = W / kWgTW, tiles_y = H / th;
  const long long ntiles_ll = (long long)N * tiles_x * tiles_y;
  if (ntiles_ll > 0x7fffffffLL) return fail(CSMRI_E_SHAPE, "too many tiles");
  const int ntiles = (int)ntiles_ll;
  int cap = sm_count() * (g_wgrad_cot == 8 ? 3 : 2);   // resident CTAs per SM (conv_wgrad.cuh)
  if (cap > kWgMaxCtas) cap = kWgMaxCtas;
  int parts = wgrad_parts(CI, CO, cap);
  if (parts > ntiles) parts = ntiles;
  cudaStream_t s = (cudaStream_t)stream;
  if (thin) {
    int ctas = sm_count() * 2;
    if (ctas > kWgMaxCtas) ctas = kWgMaxCtas;
    if (ctas > ntiles) ctas = ntiles;
    const int HI = H + 2 - 2 * pad, WI = W + 2 - 2 * pad;
    if (CI == 2 && pad == 1 && ((uintptr_t)x & 15u) == 0) {
      const int ty8 = H / kThinInRows, nt8 = N * tiles_x * ty8;
      ctas = sm_count() * 2;
      if (ctas > kWgMaxCtas) ctas = kWgMaxCtas;
      if (ctas > nt8) ctas = nt8;
      float* bias_partial = db != nullptr ? (float*)workspace + (size_t)ctas * CI * CO * 9 : nullptr;
      conv3x3_wgrad_thin_out_staged_kernel<<<ctas, 256, 0, s>>>(x, dy, (float*)workspace, bias_partial,
                                                                H, W, tiles_x, ty8, nt8);
      if (db != nullptr)
        wgrad_thin_reduce_kernel<<<(CO + 31) / 32, 32 * kThinReduceGroups, 0, s>>>(bias_partial, db, CO, ctas);
    } else if (db != nullptr && !(CI == 32 && pad == 1 && ((uintptr_t)x & 15u) == 0 && g_thin_tma &&
                                  encode_tiled_fn() != nullptr)) {
      return fail(CSMRI_E_SHAPE, "conv3x3_wgrad_bias: shape not covered (the thin layers need pad 1 and a "
                                 "16-byte aligned x, 32 -> 2 also the TMA-staged kernel)");
    } else if (CI == 2) {
      conv3x3_wgrad_thin_kernel<2, 4, true><<<ctas, 256, 0, s>>>(
          x, dy, (float*)workspace, H, W, HI, WI, pad, tiles_x, tiles_y, ntiles);
    } else if (pad == 1 && ((uintptr_t)x & 15u) == 0) {
      // input tile staged through shared memory (tiles of 8 rows)
      const int ty8 = H / kThinInRows, nt8 = N * tiles_x * ty8;
      ctas = sm_count() * 2;
      if (ctas > kWgMaxCtas) ctas = kWgMaxCtas;
      if (ctas > nt8) ctas = nt8;
      constexpr int smem_staged = 2 * kThinInBuf * (int)sizeof(float);
      EncodeTiledFn enc = g_thin_tma ? encode_tiled_fn() : nullptr;
      if (enc != nullptr) {
        alignas(64) CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * 32};
        cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
        cuuint32_t box[3] = {(cuuint32_t)kThinTmaPC, (cuuint32_t)(kThinInRows + 2), 32};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(CSMRI_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        CSMRI_TRY(set_smem(conv3x3_wgrad_thin_staged_tma_kernel, kThinWgTmaSmem));
        float* bias_partial = db != nullptr ? (float*)workspace + (size_t)ctas * CI * CO * 9 : nullptr;
        conv3x3_wgrad_thin_staged_tma_kernel<<<ctas, 256, kThinWgTmaSmem, s>>>(
            tm, x, dy, (float*)workspace, bias_partial, H, W, tiles_x, ty8, nt8);
        if (db != nullptr) wgrad_thin_reduce_kernel<<<(CO + 31) / 32, 32 * kThinReduceGroups, 0, s>>>(bias_partial, db, CO, ctas);
      } else {
        CSMRI_TRY(set_smem(conv3x3_wgrad_thin_staged_kernel, smem_staged));
        conv3x3_wgrad_thin_staged_kernel<<<ctas, 256, smem_staged, s>>>(x, dy, (float*)workspace, H, W,
                                                                        tiles_x, ty8, nt8);
      }
    } else {
      conv3x3_wgrad_thin_kernel<4, 2, false><<<ctas, 256, 0, s>>>(
          x, dy, (float*)workspace, H, W, HI, WI, pad, tiles_x, tiles_y, ntiles);
    }
    wgrad_thin_reduce_kernel<<<(CI * CO * 9 + 31) / 32, 32 * kThinReduceGroups, 0, s>>>((const float*)workspace, dw,
                                                                    CI * CO * 9, ctas);
    CSMRI_CUDA(cudaGetLastError());
    return CSMRI_OK;
  }
  if (g_wgrad_tc && CI == kWtcC && CO == kWtcC && pad == 1 && H % kWtcRows == 0 && W % kWtcPx == 0 &&
      ((uintptr_t)x & 15u) == 0 && (long long)N * (W / kWtcPx) * (H / kWtcRows) <= 0x7fffffffLL) {
    // tensor-core path (conv_wgrad_tc.cuh): one partial 96 x 96 block per CTA, then a
    // fixed-order reduction; the workspace of the SIMT path (kWgMaxCtas blocks) is larger
    alignas(64) CUtensorMap tm_x, tm_dy;
    CSMRI_TRY(make_conv_map(&tm_x, x, N, CI, H, W));
    CSMRI_TRY(make_conv_map(&tm_dy, dy, N, CO, H, W));
    const int nitems = N * (W / kWtcPx) * (H / kWtcRows);
    int ctas = sm_count();
    if (ctas > kWgMaxCtas) ctas = kWgMaxCtas;
    if (ctas > nitems) ctas = nitems;
    CSMRI_TRY(set_smem(conv3x3_wgrad_tc_kernel, kWtcSmemBytes));
    float* bias_partial = db != nullptr ? (float*)workspace + (size_t)ctas * kWtcPartial : nullptr;
    CSMRI_CUDA(launch_pdl(conv3x3_wgrad_tc_kernel, dim3(ctas), dim3(kWtcThreads), kWtcSmemBytes, s, tm_x,
                          tm_dy, dy, (float*)workspace, bias_partial, H, W, nitems, 0));
    CSMRI_CUDA(launch_pdl(conv3x3_wgrad_tc_reduce_kernel, dim3(kWtcReduceBlocks), dim3(32 * kWtcReduceGroups),
                          0, s, (const float*)workspace, dw, (const float*)bias_partial, db, ctas));
    CSMRI_CUDA(cudaGetLastError());
    return CSMRI_OK;
  }
  if (db != nullptr) return fail(CSMRI_E_SHAPE, "conv3x3_wgrad_bias: shape not covered by the tensor-core path");
  constexpr int smem = kWgSmemFloats * (int)sizeof(float);
  const dim3 grid(parts, CO / kWgC, CI / kWgC);
  if (g_wgrad_cot == 8) {
    CSMRI_TRY(set_smem(conv3x3_wgrad_kernel<8>, smem));
    conv3x3_wgrad_kernel<8><<<grid, 128, smem, s>>>(x, dy, (float*)workspace, CI, CO, H, W,
                                                    H + 2 - 2 * pad, W + 2 - 2 * pad, pad, tiles_x,
                                                    tiles_y, ntiles);
  } else {
    CSMRI_TRY(set_smem(conv3x3_wgrad_kernel<4>, smem));
    conv3x3_wgrad_kernel<4><<<grid, 256, smem, s>>>(x, dy, (float*)workspace, CI, CO, H, W,
                                                    H + 2 - 2 * pad, W + 2 - 2 * pad, pad, tiles_x,
                                                    tiles_y, ntiles);
  }
  conv3x3_wgrad_reduce_kernel<<<dim3(kWgBlock / 64, CO / kWgC, CI / kWgC), dim3(64, 8), 0, s>>>(
      (const float*)workspace, dw, CI, parts);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

static int conv3x3_thin_impl(const float* x, const float* w, const float* bias, float* y, int N, int A,
                             int B, int H, int W, float slope, const unsigned* msigns, float mslope,
                             void* stream, int wtf = 0) {
  if (!wgrad_thin(A, B))
    return fail(CSMRI_E_SHAPE, "conv3x3_thin handles 2 -> 32 and 32 -> 2 channels (got %d -> %d)", A, B);
  if (N <= 0 || H <= 0 || W <= 0 || H % kThinOutRows != 0 || W % 32 != 0)
    return fail(CSMRI_E_SHAPE, "conv3x3_thin needs H %% %d == 0 and W %% 32 == 0 (got %dx%dx%d)",
                kThinOutRows, N, H, W);
  if (!(slope >= 0.0f)) return fail(CSMRI_E_ARG, "slope must be >= 0 (got %g)", slope);
  if (A == 32 && slope != 0.0f)
    return fail(CSMRI_E_ARG, "the 32 -> 2 layer has no activation (models/recnet.py:48)");
  CSMRI_TRY(check_ptr16(x, "x"));
  CSMRI_TRY(check_ptr(w, "w"));
  CSMRI_TRY(check_ptr(y, "y"));
  if (x == y) return fail(CSMRI_E_ARG, "y must not alias x");
  const int tiles_x = W / 32, tiles_y = H / (A == 2 ? kThinOutRows : kThinInRows);
  const long long ntiles_ll = (long long)N * tiles_x * tiles_y;
  if (ntiles_ll > 0x7fffffffLL) return fail(CSMRI_E_SHAPE, "too many tiles");
  int ctas = sm_count() * 2;
  if (ctas > ntiles_ll) ctas = (int)ntiles_ll;
  cudaStream_t s = (cudaStream_t)stream;
  if (A == 2) {
    conv3x3_thin_out_kernel<<<ctas, 256, 0, s>>>(x, w, bias, y, H, W, tiles_x, tiles_y,
                                                 (int)ntiles_ll, slope, msigns, mslope, wtf);
  } else {
    if (g_thin_tma) {
      // input tiles by TMA (box {36, 10, 32} from (x0 - 4, y0 - 1), zero-filled outside the image)
      EncodeTiledFn enc = encode_tiled_fn();
      if (enc == nullptr) return fail(CSMRI_E_CUDA, "cuTensorMapEncodeTiled is unavailable");
      alignas(64) CUtensorMap tm;
      cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * 32};
      cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
      cuuint32_t box[3] = {(cuuint32_t)kThinTmaPC, (cuuint32_t)(kThinInRows + 2), 32};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)x, dims, strides, box, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(CSMRI_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
      CSMRI_TRY(set_smem(conv3x3_thin_in_tma_kernel, kThinTmaSmem));
      conv3x3_thin_in_tma_kernel<<<ctas, 256, kThinTmaSmem, s>>>(tm, x, w, bias, y, H, W, tiles_x, tiles_y,
                                                                 (int)ntiles_ll, wtf);
    } else {
      CSMRI_TRY(set_smem(conv3x3_thin_in_kernel, kThinInSmem));
      conv3x3_thin_in_kernel<<<ctas, 256, kThinInSmem, s>>>(x, w, bias, y, H, W, tiles_x, tiles_y,
                                                            (int)ntiles_ll, wtf);
    }
  }
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_conv3x3_thin(const float* x, const float* w, const float* bias, float* y, int N, int A,
                       int B, int H, int W, float slope, void* stream) {
  return conv3x3_thin_impl(x, w, bias, y, N, A, B, H, W, slope, nullptr, 0.0f, stream);
}

int csmri_conv3x3_thin_masked(const float* x, const float* w, const unsigned* signs, float* y, int N,
                              int H, int W, float act_slope, void* stream) {
  CSMRI_TRY(check_ptr(signs, "signs"));
  if (!(act_slope >= 0.0f)) return fail(CSMRI_E_ARG, "act_slope must be >= 0 (got %g)", act_slope);
  return conv3x3_thin_impl(x, w, nullptr, y, N, 2, 32, H, W, 0.0f, signs, act_slope, stream);
}

int csmri_conv3x3_thin_dgrad(const float* dy, const float* w, const unsigned* signs, float* dx, int N,
                             int CI, int CO, int H, int W, float act_slope, void* stream) {
  // dx = conv(dy, w transposed and mirrored): the thin kernel of the opposite shape reading the
  // layer's own (CO, CI, 3, 3) weights through the transposed index map
  if (signs != nullptr && CI != 32)
    return fail(CSMRI_E_ARG, "signs (the derivative of the activation that produced x) need CI = 32");
  if (!(act_slope >= 0.0f)) return fail(CSMRI_E_ARG, "act_slope must be >= 0 (got %g)", act_slope);
  return conv3x3_thin_impl(dy, w, nullptr, dx, N, CO, CI, H, W, 0.0f, signs, act_slope, stream, 1);
}

// ---- 32 -> 32 convolution on the tensor cores, 3xTF32 (conv_tc.cuh) -------------
static int conv3x3_tc_launch(const float* x, const float* w, const float* bias, float* y, unsigned* signs,
                             unsigned* in_signs, int N, int C, int H, int W, float slope,
                             int transpose_flip, bool masked, void* stream) {
  if (C != kTcC) return fail(CSMRI_E_SHAPE, "conv3x3_tc handles %d -> %d channels (got %d)", kTcC, kTcC, C);
  if (N <= 0 || H <= 0 || W <= 0 || H % kTcRowBlock != 0 || W % kTcM != 0)
    return fail(CSMRI_E_SHAPE, "conv3x3_tc needs H %% %d == 0 and W %% %d == 0 (got %dx%dx%d)",
                kTcRowBlock, kTcM, N, H, W);
  if (!(slope >= 0.0f)) return fail(CSMRI_E_ARG, "slope must be >= 0 (got %g)", slope);
  CSMRI_TRY(check_ptr(x, "x"));
  CSMRI_TRY(check_ptr(w, "w"));
  CSMRI_TRY(check_ptr(y, "y"));
  if (masked) CSMRI_TRY(check_ptr(signs, "signs"));
  if (x == y) return fail(CSMRI_E_ARG, "y must not alias x");
  const long long nitems_ll = (long long)N * (W / kTcM) * (H / kTcRowBlock);
  if (nitems_ll > 0x7fffffffLL) return fail(CSMRI_E_SHAPE, "too many tiles");
  int grid = sm_count();
  if (grid > nitems_ll) grid = (int)nitems_ll;
  const int dbg = g_tc_debug | (g_conv_pdl == 2 ? 256 : 0);
  cudaStream_t s = (cudaStream_t)stream;
  if (masked) {
    CSMRI_TRY(set_smem(conv3x3_tc_kernel<true>, kTcSmemBytes));
    CSMRI_CUDA(launch_pdl(conv3x3_tc_kernel<true>, dim3(grid), dim3(kTcThreads), kTcSmemBytes, s, x, w,
                          (const float*)nullptr, y, signs, (uint32_t*)nullptr, H, W, (int)nitems_ll, slope,
                          (int)(transpose_flip != 0), dbg));
  } else {
    CSMRI_TRY(set_smem(conv3x3_tc_kernel<false>, kTcSmemBytes));
    CSMRI_CUDA(launch_pdl(conv3x3_tc_kernel<false>, dim3(grid), dim3(kTcThreads), kTcSmemBytes, s, x, w, bias,
                          y, signs, in_signs, H, W, (int)nitems_ll, slope, (int)(transpose_flip != 0), dbg));
  }
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_conv3x3_tc(const float* x, const float* w, const float* bias, float* y, int N, int C,
                     int H, int W, float slope, int transpose_flip, void* stream) {
  return conv3x3_tc_launch(x, w, bias, y, nullptr, nullptr, N, C, H, W, slope, transpose_flip, false, stream);
}

int csmri_conv3x3_tc_signs(const float* x, const float* w, const float* bias, float* y, unsigned* signs,
                           unsigned* in_signs, int N, int C, int H, int W, float slope, void* stream) {
  CSMRI_TRY(check_ptr(signs, "signs"));
  return conv3x3_tc_launch(x, w, bias, y, signs, in_signs, N, C, H, W, slope, 0, false, stream);
}

int csmri_conv3x3_tc_masked(const float* x, const float* w, const unsigned* signs, float* y, int N,
                            int C, int H, int W, float act_slope, int transpose_flip, void* stream) {
  return conv3x3_tc_launch(x, w, nullptr, y, const_cast<unsigned*>(signs), nullptr, N, C, H, W, act_slope,
                           transpose_flip, true, stream);
}

// ---- bias + LeakyReLU epilogue of the convolutions (conv_epilogue.cuh) ---------
static int check_epilogue(int N, int C, int H, int W, float slope) {
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || (long long)N * C > 65535 ||
      ((long long)H * W) % 4 != 0 || (long long)H * W > 0x7fffffffLL)
    return fail(CSMRI_E_SHAPE, "bias_lrelu: bad shape %dx%dx%dx%d (H*W %% 4, N*C <= 65535)", N, C,
                H, W);
  if (!(slope > 0.0f)) return fail(CSMRI_E_ARG, "bias_lrelu: slope must be > 0 (got %g)", slope);
  return CSMRI_OK;
}

int csmri_bias_lrelu(float* z, const float* bias, int N, int C, int H, int W, float slope,
                     void* stream) {
  CSMRI_TRY(check_epilogue(N, C, H, W, slope));
  CSMRI_TRY(check_ptr16(z, "z"));
  CSMRI_TRY(check_ptr(bias, "bias"));
  const int hw4 = H * W / 4;
  int chunks = (hw4 + 1023) / 1024;
  if (chunks > 16) chunks = 16;
  bias_lrelu_kernel<<<dim3(chunks, N * C), 256, 0, (cudaStream_t)stream>>>((float4*)z, bias, C, hw4,
                                                                          slope);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

int csmri_bias_lrelu_backward(const float* grad_y, const float* y, float* grad_z, float* grad_bias,
                              float* partial, int N, int C, int H, int W, float slope,
                              void* stream) {
  CSMRI_TRY(check_epilogue(N, C, H, W, slope));
  CSMRI_TRY(check_ptr16(grad_y, "grad_y"));
  CSMRI_TRY(check_ptr16(y, "y"));
  CSMRI_TRY(check_ptr16(grad_z, "grad_z"));
  CSMRI_TRY(check_ptr(grad_bias, "grad_bias"));
  CSMRI_TRY(check_ptr(partial, "partial"));
  cudaStream_t s = (cudaStream_t)stream;
  bias_lrelu_backward_kernel<<<dim3(kEpChunks, N * C), 256, 0, s>>>(
      (const float4*)grad_y, (const float4*)y, (float4*)grad_z, partial, H * W / 4, slope);
  bias_grad_reduce_kernel<<<C, 32, 0, s>>>(partial, grad_bias, N, C);
  CSMRI_CUDA(cudaGetLastError());
  return CSMRI_OK;
}

}  // extern "C"
