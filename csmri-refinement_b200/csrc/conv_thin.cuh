// RecNet's thin convolutions (models/recnet.py:45-48: the first layer of a block
// maps the 2-channel image to num_filters feature maps, the last one maps back),
// 3x3, stride 1, zero padding 1, fp32.  They are pure streaming work - one side
// is a (N,2,H,W) tensor, the other a (N,32,H,W) one - but cuDNN spends 0.2-0.4 ms
// on each (implicit_convolve_sgemm / magma_sgemmEx / dgrad engines in
// profiles/r1_recnet_step_kernels.txt) where the traffic is worth ~50 us.
//
//   thin_out : 2 -> 32 channels (forward of the first layer, with bias and the
//              LeakyReLU that follows it; data gradient of the last layer)
//   thin_in  : 32 -> 2 channels (forward of the last layer, with bias; data
//              gradient of the first layer)
// The data gradients are the same kernels run on flipped, transposed weights.
//
// Lanes own pixels (coalesced row segments straight from / to global memory),
// the 8 warps split the wide channel dimension, weights live in registers as
// (output pair) float2 so that every multiply-add is a packed FFMA2, and a
// thread slides a 3x3 window down its column.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_regs.cuh"

namespace csmri {

constexpr int kThinRows = 16;   // tile = 32 x 16 pixels

__device__ __forceinline__ void thin_load_row(const float* __restrict__ plane, int gy, int gx0, int H,
                                              int W, float* dst) {
  const bool row_ok = gy >= 0 && gy < H;
  const float* src = plane + (size_t)(row_ok ? gy : 0) * W;
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    const int gx = gx0 + kx;
    dst[kx] = (row_ok && gx >= 0 && gx < W) ? __ldg(src + gx) : 0.0f;
  }
}

// y[n][4*warp + o] = act(bias + sum_{c<2, taps} w[4*warp + o][c][tap] * x[n][c][..])
__global__ void __launch_bounds__(256, 2)
    conv3x3_thin_out_kernel(const float* __restrict__ x, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                            int tiles_x, int tiles_y, int ntiles, float slope) {
  constexpr int A = 2, B = 32, BT = 4;
  const int lane = threadIdx.x & 31;
  const int cob = (threadIdx.x >> 5) * BT;
  cf wp[A][9][BT / 2];
#pragma unroll
  for (int c = 0; c < A; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < BT / 2; ++o)
        wp[c][t][o] = mk(__ldg(w + ((cob + 2 * o) * A + c) * 9 + t),
                         __ldg(w + ((cob + 2 * o + 1) * A + c) * 9 + t));
  cf bp[BT / 2];
#pragma unroll
  for (int o = 0; o < BT / 2; ++o)
    bp[o] = bias != nullptr ? mk(__ldg(bias + cob + 2 * o), __ldg(bias + cob + 2 * o + 1))
                            : mk(0.0f, 0.0f);
  const size_t plane = (size_t)H * W;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int rem = tile - n * tiles_x * tiles_y;
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const int y0 = ty * kThinRows, gx0 = tx * 32 + lane - 1;
    const float* xn = x + (size_t)n * A * plane;
    float* yn = y + ((size_t)n * B + cob) * plane + (size_t)y0 * W + tx * 32 + lane;
    // the row after next is requested before the current row is multiplied:
    // two row loads per thread are always in flight
    float win[A][3][3], nxt[A][3];
#pragma unroll
    for (int c = 0; c < A; ++c) {
      thin_load_row(xn + c * plane, y0 - 1, gx0, H, W, win[c][1]);
      thin_load_row(xn + c * plane, y0, gx0, H, W, win[c][2]);
      thin_load_row(xn + c * plane, y0 + 1, gx0, H, W, nxt[c]);
    }
#pragma unroll 4
    for (int r = 0; r < kThinRows; ++r) {
#pragma unroll
      for (int c = 0; c < A; ++c) {
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          win[c][2][kx] = nxt[c][kx];
        }
        thin_load_row(xn + c * plane, y0 + r + 2, gx0, H, W, nxt[c]);
      }
      cf acc[BT / 2];
#pragma unroll
      for (int o = 0; o < BT / 2; ++o) acc[o] = bp[o];
#pragma unroll
      for (int c = 0; c < A; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int o = 0; o < BT / 2; ++o)
              acc[o] = f2fma(wp[c][ky * 3 + kx][o], mk(win[c][ky][kx], win[c][ky][kx]), acc[o]);
#pragma unroll
      for (int o = 0; o < BT / 2; ++o) {
        float a = acc[o].x, b = acc[o].y;
        if (slope > 0.0f) {
          a = a > 0.0f ? a : a * slope;
          b = b > 0.0f ? b : b * slope;
        }
        yn[(size_t)(2 * o) * plane + (size_t)r * W] = a;
        yn[(size_t)(2 * o + 1) * plane + (size_t)r * W] = b;
      }
    }
  }
}

// y[n][o] = bias[o] + sum_{c<32, taps} w[o][c][tap] * x[n][c][..],  o < 2.
// Warp g sums input channels 4g .. 4g+3, two at a time (so that weights, window
// and the prefetched row fit the register budget of two resident CTAs); the 8
// partial sums of a pixel meet in shared memory once per tile.
__global__ void __launch_bounds__(256, 2)
    conv3x3_thin_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                           const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                           int tiles_x, int tiles_y, int ntiles) {
  constexpr int A = 32, AT = 2;
  __shared__ cf part[8][kThinRows][32];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const cf bp = bias != nullptr ? mk(__ldg(bias), __ldg(bias + 1)) : mk(0.0f, 0.0f);
  const size_t plane = (size_t)H * W;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_x * tiles_y);
    const int rem = tile - n * tiles_x * tiles_y;
    const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
    const int y0 = ty * kThinRows, gx0 = tx * 32 + lane - 1;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int cib = warp * 4 + half * AT;
      cf wp[AT][9];
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int t = 0; t < 9; ++t)
          wp[c][t] = mk(__ldg(w + (cib + c) * 9 + t), __ldg(w + (A + cib + c) * 9 + t));
      const float* xn = x + ((size_t)n * A + cib) * plane;
      float win[AT][3][3], nxt[AT][3];
#pragma unroll
      for (int c = 0; c < AT; ++c) {
        thin_load_row(xn + c * plane, y0 - 1, gx0, H, W, win[c][1]);
        thin_load_row(xn + c * plane, y0, gx0, H, W, win[c][2]);
        thin_load_row(xn + c * plane, y0 + 1, gx0, H, W, nxt[c]);
      }
#pragma unroll 4
      for (int r = 0; r < kThinRows; ++r) {
#pragma unroll
        for (int c = 0; c < AT; ++c) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            win[c][0][kx] = win[c][1][kx];
            win[c][1][kx] = win[c][2][kx];
            win[c][2][kx] = nxt[c][kx];
          }
          thin_load_row(xn + c * plane, y0 + r + 2, gx0, H, W, nxt[c]);
        }
        cf acc = half == 0 ? mk(0.0f, 0.0f) : part[warp][r][lane];
#pragma unroll
        for (int c = 0; c < AT; ++c)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              acc = f2fma(wp[c][ky * 3 + kx], mk(win[c][ky][kx], win[c][ky][kx]), acc);
        part[warp][r][lane] = acc;
      }
    }
    __syncthreads();
    float* yn = y + (size_t)n * 2 * plane + (size_t)y0 * W + tx * 32;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int pix = threadIdx.x + 256 * k;          // 512 pixels of the tile
      const int r = pix >> 5, xx = pix & 31;
      cf s = bp;
#pragma unroll
      for (int g = 0; g < 8; ++g) s = f2add(s, part[g][r][xx]);
      yn[(size_t)r * W + xx] = s.x;
      yn[plane + (size_t)r * W + xx] = s.y;
    }
    __syncthreads();
  }
}

}  // namespace csmri
