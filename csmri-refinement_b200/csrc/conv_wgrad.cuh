// Weight gradient of RecNet's 3x3 convolutions (models/recnet.py:37-48), fp32.
//
//   dW[co][ci][ky][kx] = sum_{n,y,x} dY[n][co][y][x] * X[n][ci][y + ky - p][x + kx - p]
//
// Why it is here: in the fp32 training step (training/runner.py:154-178) cuDNN's
// weight-gradient kernel for this shape (32 -> 32 channels, batch 32, 256^2) runs
// at ~15 TFLOP/s and is HALF of the whole step (profiles/r1_recnet_step_kernels.txt);
// the DC operator, the subject of this library, is 0.2 % of it.
//
// The product is a (CO x 9*CI) GEMM with a huge reduction dimension (all
// pixels), so the CTAs split the pixels: each one accumulates the complete
// 32 x 32 x 9 block of weight gradients for its tiles in registers (72 per
// thread: lane = input channel, warp = group of 8 output channels, 9 taps) and
// the partial blocks are summed by a second kernel in a fixed order
// (deterministic, no atomics).  Per pixel a thread issues 36 packed FFMA2
// (two output channels per instruction), fed by two broadcast LDS.128 of dY and
// three conflict-free LDS of the sliding 3x3 input window: the FP32 pipe is the
// bound, shared memory and issue slots stay below it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_regs.cuh"

namespace csmri {

constexpr int kWgC = 32;                    // channel block (input and output)
constexpr int kWgTW = 32, kWgTH = 4;        // output pixels per tile
constexpr int kWgPR = kWgTH + 2, kWgPC = kWgTW + 2;
constexpr int kWgXS = kWgC + 1;             // input patch [row][col][ci], ci-stride 1, col-stride 33
constexpr int kWgDS = 36;                   // dY tile [pixel][co], pixel-stride 36 (16-byte aligned rows)
constexpr int kWgSmemFloats = kWgPR * kWgPC * kWgXS + kWgTH * kWgTW * kWgDS;
constexpr int kWgBlock = kWgC * kWgC * 9;   // floats per partial block

// 4-byte asynchronous copy global -> shared (LDGSTS); !valid writes a zero.  All
// ~42 copies a thread makes for a tile are in flight together: the first version
// loaded through registers inside a loop and spent more time waiting for those
// serialised loads than computing (1.29 ms per layer, now see profiles/).
__device__ __forceinline__ void wg_copy4(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(valid ? 4 : 0)
               : "memory");
}

// Stage one tile: warp w copies input channels 4w .. 4w+3 (6 patch rows each) and
// dY channels 4w .. 4w+3 (4 rows each), lanes along x.  Everything but the per-tile
// base pointers and the edge predicates is a compile-time offset (the first version
// recomputed (channel, row) from a loop counter and spent 28 % of the kernel's
// instructions, on the same pipe as the FFMA2s, on address arithmetic).  Column /
// row indices are clamped into the image so that every source address is valid;
// out-of-image elements are copies of size 0 (zero fill).
template <int NW>   // NW warps stage the tile: warp w takes channels w*(32/NW) .. of x and of dY
__device__ __forceinline__ void wg_stage(float* X_s, float* D_s, const float* __restrict__ x,
                                         const float* __restrict__ dy, int tile, int tiles_x,
                                         int tiles_y, int CI, int CO, int ci0, int co0, int H, int W,
                                         int HI, int WI, int pad, int lane, int warp) {
  constexpr int CPW = kWgC / NW;   // channels per warp
  const int n = tile / (tiles_x * tiles_y);
  const int rem = tile - n * tiles_x * tiles_y;
  const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
  const int y0 = ty * kWgTH, x0 = tx * kWgTW;
  const int gx = x0 + lane - pad, gx2 = gx + kWgTW;
  const bool ok0 = gx >= 0 && gx < WI;
  const bool ok1 = gx2 >= 0 && gx2 < WI;
  const int cx0 = min(max(gx, 0), WI - 1), cx1 = min(max(gx2, 0), WI - 1);
  const size_t chs = (size_t)HI * WI;
  const float* src_c = x + ((size_t)n * CI + ci0 + warp * CPW) * chs;
  float* dst_c = X_s + lane * kWgXS + warp * CPW;
#pragma unroll
  for (int r = 0; r < kWgPR; ++r) {
    const int gy = y0 + r - pad;
    const bool row_ok = gy >= 0 && gy < HI;
    const float* src = src_c + (size_t)min(max(gy, 0), HI - 1) * WI;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      float* dst = dst_c + r * kWgPC * kWgXS + c;
      wg_copy4(dst, src + cx0, row_ok && ok0);
      if (lane < 2) wg_copy4(dst + kWgTW * kWgXS, src + cx1, row_ok && ok1);
      src += chs;
    }
  }
  const size_t ohs = (size_t)H * W;
  const float* dsrc = dy + (((size_t)n * CO + co0 + warp * CPW) * H + y0) * W + x0 + lane;
  float* ddst = D_s + lane * kWgDS + warp * CPW;
#pragma unroll
  for (int r = 0; r < kWgTH; ++r) {
    const float* src = dsrc + (size_t)r * W;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      wg_copy4(ddst + r * kWgTW * kWgDS + c, src, true);
      src += ohs;
    }
  }
}

// one staged tile into the thread's 9*COT accumulators: 9*COT/2 FFMA2 per pixel
template <int COT>
__device__ __forceinline__ void wg_compute(const float* X_s, const float* D_s,
                                           cf (&acc)[9][COT / 2], int lane, int warp) {
#pragma unroll 1
  for (int r = 0; r < kWgTH; ++r) {
    float w[3][3];   // sliding window
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      w[ky][1] = X_s[((r + ky) * kWgPC + 0) * kWgXS + lane];
      w[ky][2] = X_s[((r + ky) * kWgPC + 1) * kWgXS + lane];
    }
#pragma unroll
    for (int xx = 0; xx < kWgTW; ++xx) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        w[ky][0] = w[ky][1];
        w[ky][1] = w[ky][2];
        w[ky][2] = X_s[((r + ky) * kWgPC + xx + 2) * kWgXS + lane];
      }
      cf dp[COT / 2];
#pragma unroll
      for (int q = 0; q < COT / 4; ++q) {
        const float4 d = *reinterpret_cast<const float4*>(
            &D_s[(r * kWgTW + xx) * kWgDS + warp * COT + 4 * q]);
        dp[2 * q] = mk(d.x, d.y);
        dp[2 * q + 1] = mk(d.z, d.w);
      }
#pragma unroll
      for (int o = 0; o < COT / 2; ++o)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            // packed FFMA2 with a scalar-broadcast operand: two output channels per
            // instruction (scalar FFMAs measured 15 % slower: issue-bound)
            acc[ky * 3 + kx][o] = f2fma(dp[o], mk(w[ky][kx], w[ky][kx]), acc[ky * 3 + kx][o]);
    }
  }
}

// grid: (CTAs sharing the pixel tiles, CO / 32, CI / 32); block: 32 * (32 / COT).
// One tile buffer per CTA; four resident CTAs hide each other's staging (a
// double-buffered variant with two CTAs per SM measured the same).
// COT = output channels per thread: 4 (256 threads, 36 accumulators) or
// 8 (128 threads, 72 accumulators, half the shared-memory loads per FFMA2).
// Accumulation is two-level: a tile's 128 products per weight go into fresh
// registers, the tile sums into a second register set.  With one running sum per
// weight over all ~10^3 pixels of a CTA the fp32 rounding error of the strongly
// cancelling sum reached 1.8e-5 of |dW| at 512^2 (tests: 1-recnet.json against a
// float64 evaluation); the error of a two-level sum grows with sqrt(128 * tiles)
// instead of sqrt(128 * tiles^2 / 2).  The second set costs 9 * COT registers,
// hence three instead of four resident CTAs for COT = 8.
template <int COT>
__global__ void __launch_bounds__(32 * (kWgC / COT), COT == 8 ? 3 : 2)
    conv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                         float* __restrict__ partial, int CI, int CO, int H, int W, int HI, int WI,
                         int pad, int tiles_x, int tiles_y, int ntiles) {
  extern __shared__ __align__(16) float wg_smem[];
  float* X_s = wg_smem;
  float* D_s = wg_smem + kWgPR * kWgPC * kWgXS;
  const int lane = threadIdx.x & 31;        // input channel within the block
  const int warp = threadIdx.x >> 5;        // group of COT output channels
  const int co0 = blockIdx.y * kWgC, ci0 = blockIdx.z * kWgC;

  cf acc[9][COT / 2];   // sum over the CTA's tiles of the per-tile sums
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < COT / 2; ++o) acc[t][o] = mk(0.0f, 0.0f);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();   // everyone is done with the previous tile
    wg_stage<kWgC / COT>(X_s, D_s, x, dy, tile, tiles_x, tiles_y, CI, CO, ci0, co0, H, W, HI, WI,
                         pad, lane, warp);
    cf tacc[9][COT / 2];   // this tile
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < COT / 2; ++o) tacc[t][o] = mk(0.0f, 0.0f);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    wg_compute<COT>(X_s, D_s, tacc, lane, warp);
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < COT / 2; ++o) acc[t][o] = f2add(acc[t][o], tacc[t][o]);
  }
  // partial block in dW order: [co][ci][tap]
  float* dst = partial +
               ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * kWgBlock;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int o = 0; o < COT / 2; ++o) {
      dst[((warp * COT + 2 * o) * kWgC + lane) * 9 + t] = acc[t][o].x;
      dst[((warp * COT + 2 * o + 1) * kWgC + lane) * 9 + t] = acc[t][o].y;
    }
}

// dw[co][ci][tap] = sum over the `nparts` partial blocks of its channel-block pair.
// grid: (kWgBlock / 64, CO / 32, CI / 32); block (64, 8): eight threads share an
// element (partials p = g, g + 8, ...) and meet in shared memory in a fixed order.
__global__ void __launch_bounds__(512)
    conv3x3_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int CI,
                                int nparts) {
  __shared__ float s_part[8][64];
  const int e = blockIdx.x * 64 + threadIdx.x;          // element of the 32 x 32 x 9 block
  const float* src = partial + (size_t)(blockIdx.z * gridDim.y + blockIdx.y) * nparts * kWgBlock + e;
  float s0 = 0.0f, s1 = 0.0f;
  int p = threadIdx.y;
  for (; p + 8 < nparts; p += 16) {
    s0 += src[(size_t)p * kWgBlock];
    s1 += src[(size_t)(p + 8) * kWgBlock];
  }
  if (p < nparts) s0 += src[(size_t)p * kWgBlock];
  s_part[threadIdx.y][threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.y == 0) {
    float s = 0.0f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += s_part[g][threadIdx.x];
    const int co = e / (kWgC * 9), r = e - co * (kWgC * 9);
    const int ci = r / 9, t = r - ci * 9;
    dw[((size_t)(blockIdx.y * kWgC + co) * CI + blockIdx.z * kWgC + ci) * 9 + t] = s;
  }
}

// ---------------------------------------------------------------------------
// Thin layers: RecNet's first (2 -> 32) and last (32 -> 2) convolution of every
// block.  Only 576 weight gradients, so the work is streaming x and dY once
// (HBM-bound; cuDNN needs ~1 ms per layer, the traffic is worth ~50 us).
// Lanes own pixels (coalesced row segments straight from global memory, no
// staging), the 8 warps split the wide channel dimension, a thread keeps
// CIT * COT * 9 = 72 accumulators and a sliding 3 x 3 window per input channel;
// lanes are folded with shuffles once per CTA, CTAs by the reduce kernel.
//   WIDE_OUT: CI = CIT = 2,  CO = 8 * COT = 32   (warp = group of 4 output channels)
//  !WIDE_OUT: CI = 8 * CIT = 32, CO = COT = 2    (warp = group of 4 input channels)
// ---------------------------------------------------------------------------
constexpr int kWtRows = 16;                  // rows per tile (tile = 32 x 16 pixels)

template <int CIT, int COT, bool WIDE_OUT>
__global__ void __launch_bounds__(256, 2)
    conv3x3_wgrad_thin_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                              float* __restrict__ partial, int H, int W, int HI, int WI, int pad,
                              int tiles_x, int tiles_y, int ntiles) {
  constexpr int CI = WIDE_OUT ? CIT : 8 * CIT;
  constexpr int CO = WIDE_OUT ? 8 * COT : COT;
  static_assert(COT % 2 == 0, "output channels are processed in pairs");
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cib = WIDE_OUT ? 0 : warp * CIT;
  const int cob = WIDE_OUT ? warp * COT : 0;

  cf acc[CIT][9][COT / 2];
#pragma unroll
  for (int c = 0; c < CIT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < COT / 2; ++o) acc[c][t][o] = mk(0.0f, 0.0f);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int rem = tile - n * tiles_x * tiles_y;
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const int y0 = ty * kWtRows, gx0 = tx * 32 + lane - pad;
    const float* xn = x + ((size_t)n * CI + cib) * HI * WI;
    const float* dn = dy + (((size_t)n * CO + cob) * H + y0) * W + tx * 32 + lane;
    auto load_row = [&](int c, int gy, float* dst) {   // x[c][gy][gx0 .. gx0 + 2], zero outside
      const bool row_ok = gy >= 0 && gy < HI;
      const float* src = xn + ((size_t)c * HI + (row_ok ? gy : 0)) * WI;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int gx = gx0 + kx;
        dst[kx] = (row_ok && gx >= 0 && gx < WI) ? __ldg(src + gx) : 0.0f;
      }
    };
    // the next row of x and dY is requested before the current one is multiplied
    // (x only where the registers allow it: two input channels per thread)
    constexpr bool kPrefetchX = CIT <= 2;
    float win[CIT][3][3], nxt[kPrefetchX ? CIT : 1][3], dnx[COT];
#pragma unroll
    for (int c = 0; c < CIT; ++c) {
      load_row(c, y0 - pad, win[c][1]);
      load_row(c, y0 - pad + 1, win[c][2]);
      if (kPrefetchX) load_row(c, y0 - pad + 2, nxt[c]);
    }
#pragma unroll
    for (int o = 0; o < COT; ++o) dnx[o] = __ldg(dn + (size_t)o * H * W);
#pragma unroll 2
    for (int r = 0; r < kWtRows; ++r) {
      float d[COT];
#pragma unroll
      for (int o = 0; o < COT; ++o) {
        d[o] = dnx[o];
        if (r + 1 < kWtRows) dnx[o] = __ldg(dn + ((size_t)o * H + r + 1) * W);
      }
#pragma unroll
      for (int c = 0; c < CIT; ++c) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          if (kPrefetchX) win[c][2][kx] = nxt[c][kx];
        }
        if (kPrefetchX) load_row(c, y0 + r + 3 - pad, nxt[c]);
        else load_row(c, y0 + r + 2 - pad, win[c][2]);
      }
#pragma unroll
      for (int c = 0; c < CIT; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int o = 0; o < COT / 2; ++o)
              acc[c][ky * 3 + kx][o] = f2fma(mk(d[2 * o], d[2 * o + 1]),
                                             mk(win[c][ky][kx], win[c][ky][kx]),
                                             acc[c][ky * 3 + kx][o]);
    }
  }
  // fold the 32 pixels-lanes, then lane 0 writes the warp's 72 sums in dW order [co][ci][tap]
  float* dst = partial + (size_t)blockIdx.x * (CI * CO * 9);
#pragma unroll
  for (int c = 0; c < CIT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < COT / 2; ++o) {
        float a = acc[c][t][o].x, b = acc[c][t][o].y;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, m);
          b += __shfl_xor_sync(0xffffffffu, b, m);
        }
        if (lane == 0) {
          dst[((cob + 2 * o) * CI + cib + c) * 9 + t] = a;
          dst[((cob + 2 * o + 1) * CI + cib + c) * 9 + t] = b;
        }
      }
}

// dw[e] = sum_p partial[p][e], e < n_elem, in a fixed order: block = 32 elements x 8 groups of
// partial blocks (256 threads); group g adds the blocks p = g, g + 8, ... with coalesced
// 128-byte loads, the eight group sums meet in shared memory.  (One thread per element
// walking all ~300 blocks took 12 us per launch - 20 launches per D5C5 training step.)
constexpr int kThinReduceGroups = 8;

__global__ void __launch_bounds__(32 * kThinReduceGroups)
    wgrad_thin_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int n_elem,
                             int nparts) {
  __shared__ float part[kThinReduceGroups][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  float s0 = 0.0f, s1 = 0.0f;
  if (e < n_elem) {
    int p = g;
    for (; p + kThinReduceGroups < nparts; p += 2 * kThinReduceGroups) {
      s0 += partial[(size_t)p * n_elem + e];
      s1 += partial[(size_t)(p + kThinReduceGroups) * n_elem + e];
    }
    if (p < nparts) s0 += partial[(size_t)p * n_elem + e];
  }
  part[g][lane] = s0 + s1;
  __syncthreads();
  if (g == 0 && e < n_elem) {
    float s = part[0][lane];
#pragma unroll
    for (int k = 1; k < kThinReduceGroups; ++k) s += part[k][lane];
    dw[e] = s;
  }
}

}  // namespace csmri
