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
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_regs.cuh"

namespace csmri {

// Staged input row (pitch 36 floats): [0..31] the 32 columns (16-byte aligned),
// [32] the right halo column; the LEFT halo column sits at [-1], i.e. in the
// unused tail [35] of the previous row slot (4 lead floats in front of the first),
// so that column x0 + j is always at offset j, j = -1 .. 32.
constexpr int kThinPC = 36;
constexpr int kThinLead = 4;
constexpr int kThinOutRows = 16;   // thin_out tile: 32 x 16 pixels (input tile 2 x 18 x 34)
constexpr int kThinInRows = 8;     // thin_in  tile: 32 x  8 pixels (input tile 32 x 10 x 34)

__device__ __forceinline__ void thin_copy4(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(valid ? 4 : 0)
               : "memory");
}

__device__ __forceinline__ void thin_copy16(float* dst, const float* src, bool valid) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(valid ? 16 : 0)
               : "memory");
}

// Four staged rows per call: quarter-warp q copies row `gy[q]` of `plane[q]`,
// columns x0 .. x0+31 as eight 16-byte pieces (x0 is a multiple of 32, so source
// and destination are aligned), then lanes 8q and 8q+1 add the two halo columns.
// (4-byte pieces for everything made the copy unit, not HBM, the bound of thin_in.)
__device__ __forceinline__ void thin_stage_row4(float* dst_row, const float* __restrict__ plane,
                                                int gy, int x0, int H, int W, int lane) {
  const int quad = lane & 7;
  const bool row_ok = gy >= 0 && gy < H;
  const float* src = plane + (size_t)min(max(gy, 0), H - 1) * W;
  thin_copy16(dst_row + quad * 4, src + x0 + quad * 4, row_ok);
  if (quad < 2) {
    const int gx = quad == 0 ? x0 - 1 : x0 + 32;
    thin_copy4(dst_row + (quad == 0 ? -1 : 32), src + min(max(gx, 0), W - 1),
               row_ok && gx >= 0 && gx < W);
  }
}

struct ThinTile {
  int n, y0, x0;
};
__device__ __forceinline__ ThinTile thin_tile(int tile, int tiles_x, int tiles_y, int rows) {
  ThinTile t;
  t.n = tile / (tiles_x * tiles_y);
  const int rem = tile - t.n * tiles_x * tiles_y;
  const int ty = rem / tiles_x;
  t.y0 = ty * rows;
  t.x0 = (rem - ty * tiles_x) * 32;
  return t;
}

// y[n][4*warp + o] = act(bias + sum_{c<2, taps} w[4*warp + o][c][tap] * x[n][c][..])
// The (tiny) input tile of the NEXT tile is staged with cp.async while the current
// one is multiplied: the row loop never waits for global memory.  (The first
// version loaded each row from global memory one step ahead and ran at a quarter
// of the write bandwidth: every step paid an L2 round trip.)
__global__ void __launch_bounds__(256, 2)
    conv3x3_thin_out_kernel(const float* __restrict__ x, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                            int tiles_x, int tiles_y, int ntiles, float slope,
                            const unsigned* __restrict__ msigns, float mslope, int wtf) {
  // msigns (optional; (N,H,W) uint32 sign words, see csmri_conv3x3_tc_signs): the result of
  // channel c is multiplied by (bit c ? 1 : mslope) - this kernel as the data gradient of the
  // 32 -> 2 layer, followed by the backward of the LeakyReLU in front of that layer
  constexpr int A = 2, B = 32, BT = 4, R = kThinOutRows;
  constexpr int kBuf = kThinLead + A * (R + 2) * kThinPC;
  __shared__ __align__(16) float stage[2][kBuf];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cob = warp * BT;
  cf wp[A][9][BT / 2];
#pragma unroll
  for (int c = 0; c < A; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < BT / 2; ++o)
        wp[c][t][o] = wtf ? mk(__ldg(w + (c * B + cob + 2 * o) * 9 + 8 - t),       // w is (A, B, 3, 3): the
                               __ldg(w + (c * B + cob + 2 * o + 1) * 9 + 8 - t))   // operator transposed + mirrored
                          : mk(__ldg(w + ((cob + 2 * o) * A + c) * 9 + t),
                               __ldg(w + ((cob + 2 * o + 1) * A + c) * 9 + t));
  cf bp[BT / 2];
#pragma unroll
  for (int o = 0; o < BT / 2; ++o)
    bp[o] = bias != nullptr ? mk(__ldg(bias + cob + 2 * o), __ldg(bias + cob + 2 * o + 1))
                            : mk(0.0f, 0.0f);
  const size_t plane = (size_t)H * W;
  auto stage_tile = [&](int tile, float* buf) {
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    static_assert(A * (R + 2) % 4 == 0, "rows are staged four at a time");
    for (int pair = warp * 4 + (lane >> 3); pair < A * (R + 2); pair += 32) {
      const int c = pair / (R + 2), r = pair - c * (R + 2);
      thin_stage_row4(buf + kThinLead + pair * kThinPC, x + ((size_t)t.n * A + c) * plane,
                      t.y0 - 1 + r, t.x0, H, W, lane);
    }
  };
  int tile = blockIdx.x;
  if (tile < ntiles) stage_tile(tile, stage[0]);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const float* cur = stage[it & 1] + kThinLead;
    const int next = tile + gridDim.x;
    if (next < ntiles) stage_tile(next, stage[(it + 1) & 1]);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    float* yn = y + ((size_t)t.n * B + cob) * plane + (size_t)t.y0 * W + t.x0 + lane;
    const unsigned* sg = msigns != nullptr ? msigns + ((size_t)t.n * H + t.y0) * W + t.x0 + lane : nullptr;
    unsigned sw[R];                    // the tile's sign words, all requested before the row loop
#pragma unroll
    for (int r = 0; r < R; ++r) sw[r] = sg != nullptr ? __ldg(sg + (size_t)r * W) : 0xffffffffu;
    float win[A][3][3];
#pragma unroll
    for (int c = 0; c < A; ++c)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        win[c][1][kx] = cur[(c * (R + 2) + 0) * kThinPC + lane + kx - 1];
        win[c][2][kx] = cur[(c * (R + 2) + 1) * kThinPC + lane + kx - 1];
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < A; ++c)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          win[c][2][kx] = cur[(c * (R + 2) + r + 2) * kThinPC + lane + kx - 1];
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
      const unsigned bits = sw[r] >> cob;
#pragma unroll
      for (int o = 0; o < BT / 2; ++o) {
        float a = acc[o].x, b = acc[o].y;
        if (slope > 0.0f) {
          a = a > 0.0f ? a : a * slope;
          b = b > 0.0f ? b : b * slope;
        }
        if (sg != nullptr) {
          a = ((bits >> (2 * o)) & 1u) ? a : a * mslope;
          b = ((bits >> (2 * o + 1)) & 1u) ? b : b * mslope;
        }
        yn[(size_t)(2 * o) * plane + (size_t)r * W] = a;
        yn[(size_t)(2 * o + 1) * plane + (size_t)r * W] = b;
      }
    }
    __syncthreads();   // `cur` is the staging target of the next iteration
  }
}

// y[n][o] = bias[o] + sum_{c<32, taps} w[o][c][tap] * x[n][c][..],  o < 2.
// Warp g sums input channels 4g .. 4g+3 from the cp.async-staged tile (next tile
// in flight while this one is multiplied); the 8 partial sums of a pixel meet in
// shared memory once per tile.  Dynamic shared memory: 2 tile buffers + partials.
constexpr int kThinInBuf = kThinLead + 32 * (kThinInRows + 2) * kThinPC;           // floats
constexpr int kThinInSmem = (2 * kThinInBuf + 8 * kThinInRows * 32 * 2) * 4;       // bytes

__global__ void __launch_bounds__(256, 2)
    conv3x3_thin_in_kernel(const float* __restrict__ x, const float* __restrict__ w,
                           const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                           int tiles_x, int tiles_y, int ntiles, int wtf) {
  constexpr int A = 32, AT = 4, R = kThinInRows;
  extern __shared__ __align__(16) float thin_smem[];
  cf* part = reinterpret_cast<cf*>(thin_smem + 2 * kThinInBuf);   // [8][R][32]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cib = warp * AT;
  cf wp[AT][9];
#pragma unroll
  for (int c = 0; c < AT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
      wp[c][t] = wtf ? mk(__ldg(w + ((cib + c) * 2 + 0) * 9 + 8 - t), __ldg(w + ((cib + c) * 2 + 1) * 9 + 8 - t))
                     : mk(__ldg(w + (cib + c) * 9 + t), __ldg(w + (A + cib + c) * 9 + t));
  const cf bp = bias != nullptr ? mk(__ldg(bias), __ldg(bias + 1)) : mk(0.0f, 0.0f);
  const size_t plane = (size_t)H * W;
  auto stage_tile = [&](int tile, float* buf) {   // warp g stages its own four channels
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    const float* xc = x + ((size_t)t.n * A + cib) * plane;
    static_assert(AT * (R + 2) % 4 == 0, "rows are staged four at a time");
#pragma unroll 2
    for (int pair = lane >> 3; pair < AT * (R + 2); pair += 4) {
      const int c = pair / (R + 2), r = pair - c * (R + 2);
      thin_stage_row4(buf + kThinLead + (cib * (R + 2) + pair) * kThinPC, xc + c * plane,
                      t.y0 - 1 + r, t.x0, H, W, lane);
    }
  };
  int tile = blockIdx.x;
  if (tile < ntiles) stage_tile(tile, thin_smem);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const float* cur = thin_smem + (it & 1) * kThinInBuf + kThinLead + cib * (R + 2) * kThinPC;
    const int next = tile + gridDim.x;
    if (next < ntiles) stage_tile(next, thin_smem + ((it + 1) & 1) * kThinInBuf);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();   // tile landed; previous tile's partial sums have been consumed
    float win[AT][3][3];
#pragma unroll
    for (int c = 0; c < AT; ++c)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        win[c][1][kx] = cur[(c * (R + 2) + 0) * kThinPC + lane + kx - 1];
        win[c][2][kx] = cur[(c * (R + 2) + 1) * kThinPC + lane + kx - 1];
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          win[c][2][kx] = cur[(c * (R + 2) + r + 2) * kThinPC + lane + kx - 1];
        }
      cf acc = mk(0.0f, 0.0f);
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            acc = f2fma(wp[c][ky * 3 + kx], mk(win[c][ky][kx], win[c][ky][kx]), acc);
      part[(warp * R + r) * 32 + lane] = acc;
    }
    __syncthreads();   // partial sums complete; everyone is done reading the tile
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    float* yn = y + (size_t)t.n * 2 * plane + (size_t)t.y0 * W + t.x0;
    {
      const int r = threadIdx.x >> 5;                 // 8 rows x 32 pixels = 256 threads
      cf sum = bp;
#pragma unroll
      for (int g = 0; g < 8; ++g) sum = f2add(sum, part[(g * R + r) * 32 + lane]);
      yn[(size_t)r * W + lane] = sum.x;
      yn[plane + (size_t)r * W + lane] = sum.y;
    }
  }
}

// The same 32 -> 2 convolution with the input tile staged by TMA instead of cp.async:
// one box {36 columns from x0 - 4, 10 rows from y0 - 1, 32 channels} per tile, zero-filled
// outside the image (= the convolution's padding), issued by one thread one tile ahead on an
// mbarrier.  (The innermost start coordinate of a TMA box must be 16-byte aligned - x0 - 1 is an
// illegal instruction, tools/tma_box_probe.cu - so the box starts 4 columns to the left and
// column x0 + j sits at offset j + 4.)  The box ends at column x0 + 31; the right halo column
// x0 + 32 of row r is read straight from global memory one tile ahead and dropped into the
// unused first slot of row r + 1, which is exactly where offset 32 + 4 of row r points.
// The cp.async version spends 40 % of its 1031 instructions per pixel on staging address
// arithmetic (profiles/r1_training_kernels_ncu.txt: 81 M warp instructions, issue-bound at 3x
// the HBM floor); here staging costs three instructions per tile plus two loads per thread.
constexpr int kThinTmaPC = 36;                                              // box width (floats)
constexpr int kThinTmaBox = 32 * (kThinInRows + 2) * kThinTmaPC;            // floats in one box
constexpr int kThinTmaBuf = kThinTmaBox + 32;                               // + the last row's halo slot; keeps buffer 1 128-byte aligned
constexpr int kThinTmaSmem = (2 * kThinTmaBuf + 8 * kThinInRows * 32 * 2) * 4 + 64 + 128;   // + partials, barriers, alignment slack

__global__ void __launch_bounds__(256, 2)
    conv3x3_thin_in_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ x,
                               const float* __restrict__ w, const float* __restrict__ bias,
                               float* __restrict__ y, int H, int W,
                               int tiles_x, int tiles_y, int ntiles, int wtf) {
  constexpr int A = 32, AT = 4, R = kThinInRows;
  extern __shared__ unsigned char thin_tma_raw[];
  // TMA destinations are 128-byte aligned (the dynamic shared-memory base is only 16)
  float* thin_smem = reinterpret_cast<float*>(
      thin_tma_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(thin_tma_raw) & 127u)) & 127u));
  cf* part = reinterpret_cast<cf*>(thin_smem + 2 * kThinTmaBuf);   // [8][R][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(thin_smem + 2 * kThinTmaBuf + 8 * R * 32 * 2);
  const uint32_t full = (uint32_t)__cvta_generic_to_shared(bars);  // [2] tile landed
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cib = warp * AT;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
  }
  cf wp[AT][9];
#pragma unroll
  for (int c = 0; c < AT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
      wp[c][t] = wtf ? mk(__ldg(w + ((cib + c) * 2 + 0) * 9 + 8 - t), __ldg(w + ((cib + c) * 2 + 1) * 9 + 8 - t))
                     : mk(__ldg(w + (cib + c) * 9 + t), __ldg(w + (A + cib + c) * 9 + t));
  const cf bp = bias != nullptr ? mk(__ldg(bias), __ldg(bias + 1)) : mk(0.0f, 0.0f);
  const size_t plane = (size_t)H * W;
  __syncthreads();
  // Tile coordinates are decoded once per tile (two runtime divisions) and shared by the TMA
  // issue, the halo loads and the output addresses; the halo entry (channel, row) a thread
  // owns does not depend on the tile.
  auto stage_tile = [&](const ThinTile t, int b) {   // one thread: expect the bytes, fire the box
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(thin_smem + b * kThinTmaBuf);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * b),
                 "r"((uint32_t)(kThinTmaBox * 4))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(&tm_x), "r"(full + 8 * b), "r"(t.x0 - 4), "r"(t.y0 - 1), "r"(t.n * A)
        : "memory");
  };
  // right halo column: 32 channels x 10 rows = 320 values per tile, thread i takes entries i and
  // i + 256 (entry e = channel * 10 + row), requested one tile ahead
  const int e1 = threadIdx.x + 256;
  const bool has1 = e1 < A * (R + 2);
  const int hc0 = threadIdx.x / (R + 2), hr0 = threadIdx.x - hc0 * (R + 2) - 1;
  const int hc1 = has1 ? e1 / (R + 2) : 0, hr1 = has1 ? e1 - hc1 * (R + 2) - 1 : 0;
  const size_t ho0 = (size_t)hc0 * plane, ho1 = (size_t)hc1 * plane;
  auto fetch_halo = [&](const ThinTile t, int hr, size_t ho, bool have) -> float {
    const int gy = t.y0 + hr, gx = t.x0 + 32;
    if (!have || gy < 0 || gy >= H || gx >= W) return 0.0f;
    return __ldg(x + (size_t)t.n * A * plane + ho + (size_t)gy * W + gx);
  };
  int tile = blockIdx.x;
  float h0 = 0.0f, h1 = 0.0f;
  ThinTile tc = thin_tile(tile < ntiles ? tile : 0, tiles_x, tiles_y, R);
  if (tile < ntiles) {
    if (threadIdx.x == 0) stage_tile(tc, 0);
    h0 = fetch_halo(tc, hr0, ho0, true);
    h1 = fetch_halo(tc, hr1, ho1, has1);
  }
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int b = it & 1;
    const float* cur = thin_smem + b * kThinTmaBuf + cib * (R + 2) * kThinTmaPC;
    const int next = tile + gridDim.x;
    const ThinTile tn = thin_tile(next < ntiles ? next : 0, tiles_x, tiles_y, R);
    // buffer b ^ 1 was last read in iteration it - 1, whose trailing __syncthreads has passed
    if (threadIdx.x == 0 && next < ntiles) stage_tile(tn, b ^ 1);
    {
      uint32_t done;
      do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;"
            " selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(full + 8 * b), "r"((uint32_t)((it >> 1) & 1))
            : "memory");
      } while (!done);
    }
    {   // the box has landed: drop this tile's right halo column into the slots behind the rows
      float* buf = thin_smem + b * kThinTmaBuf;
      buf[(threadIdx.x + 1) * kThinTmaPC] = h0;
      if (has1) buf[(e1 + 1) * kThinTmaPC] = h1;
    }
    if (next < ntiles) {
      h0 = fetch_halo(tn, hr0, ho0, true);
      h1 = fetch_halo(tn, hr1, ho1, has1);
    }
    __syncthreads();
    float win[AT][3][3];
#pragma unroll
    for (int c = 0; c < AT; ++c)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        win[c][1][kx] = cur[(c * (R + 2) + 0) * kThinTmaPC + lane + kx + 3];
        win[c][2][kx] = cur[(c * (R + 2) + 1) * kThinTmaPC + lane + kx + 3];
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          win[c][2][kx] = cur[(c * (R + 2) + r + 2) * kThinTmaPC + lane + kx + 3];
        }
      cf acc = mk(0.0f, 0.0f);
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            acc = f2fma(wp[c][ky * 3 + kx], mk(win[c][ky][kx], win[c][ky][kx]), acc);
      part[(warp * R + r) * 32 + lane] = acc;
    }
    __syncthreads();   // partial sums complete; everyone is done reading the tile
    float* yn = y + (size_t)tc.n * 2 * plane + (size_t)tc.y0 * W + tc.x0;
    {
      const int r = threadIdx.x >> 5;                 // 8 rows x 32 pixels = 256 threads
      cf sum = bp;
#pragma unroll
      for (int g = 0; g < 8; ++g) sum = f2add(sum, part[(g * R + r) * 32 + lane]);
      yn[(size_t)r * W + lane] = sum.x;
      yn[plane + (size_t)r * W + lane] = sum.y;
    }
    tc = tn;
    __syncthreads();   // partial sums consumed before the next tile overwrites them
  }
}

// ---------------------------------------------------------------------------
// Weight gradient of the 32 -> 2 layer with the input tile staged like thin_in
// (zero padding 1).  dw[o][c][tap] = sum dy[n][o][y][x] * x[n][c][y+ky-1][x+kx-1]:
// warp g owns input channels 4g .. 4g+3, lanes own pixels, 4 x 9 accumulator
// pairs (o = 0, 1) per thread; dY (2 channels) comes straight from global memory.
// The direct-from-global version (conv3x3_wgrad_thin_kernel) spent most of its
// 130 M instructions on predicated loads and waited on them (ncu: long_scoreboard
// 3.1 cycles per issue, FP32 pipe 20 %).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
    conv3x3_wgrad_thin_staged_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                     float* __restrict__ partial, int H, int W, int tiles_x,
                                     int tiles_y, int ntiles) {
  constexpr int A = 32, AT = 4, R = kThinInRows;
  extern __shared__ __align__(16) float thin_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cib = warp * AT;
  cf acc[AT][9];
#pragma unroll
  for (int c = 0; c < AT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[c][t] = mk(0.0f, 0.0f);
  const size_t plane = (size_t)H * W;
  auto stage_tile = [&](int tile, float* buf) {
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    const float* xc = x + ((size_t)t.n * A + cib) * plane;
#pragma unroll 2
    for (int pair = lane >> 3; pair < AT * (R + 2); pair += 4) {
      const int c = pair / (R + 2), r = pair - c * (R + 2);
      thin_stage_row4(buf + kThinLead + (cib * (R + 2) + pair) * kThinPC, xc + c * plane,
                      t.y0 - 1 + r, t.x0, H, W, lane);
    }
  };
  int tile = blockIdx.x;
  if (tile < ntiles) stage_tile(tile, thin_smem);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const float* cur = thin_smem + (it & 1) * kThinInBuf + kThinLead + cib * (R + 2) * kThinPC;
    const int next = tile + gridDim.x;
    if (next < ntiles) stage_tile(next, thin_smem + ((it + 1) & 1) * kThinInBuf);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    const float* dn = dy + ((size_t)t.n * 2 * H + t.y0) * W + t.x0 + lane;
    float d0[R], d1[R];   // this tile's dY column: in flight while the tile lands
#pragma unroll
    for (int r = 0; r < R; ++r) {
      d0[r] = __ldg(dn + (size_t)r * W);
      d1[r] = __ldg(dn + plane + (size_t)r * W);
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    float win[AT][3][3];
#pragma unroll
    for (int c = 0; c < AT; ++c)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        win[c][1][kx] = cur[(c * (R + 2) + 0) * kThinPC + lane + kx - 1];
        win[c][2][kx] = cur[(c * (R + 2) + 1) * kThinPC + lane + kx - 1];
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          win[c][2][kx] = cur[(c * (R + 2) + r + 2) * kThinPC + lane + kx - 1];
        }
      const cf d = mk(d0[r], d1[r]);
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            acc[c][ky * 3 + kx] =
                f2fma(d, mk(win[c][ky][kx], win[c][ky][kx]), acc[c][ky * 3 + kx]);
    }
    __syncthreads();   // `cur` is the staging target of the next iteration
  }
  // fold the 32 pixel-lanes; lane 0 writes the warp's sums in dW order [o][c][tap]
  float* dst = partial + (size_t)blockIdx.x * (A * 2 * 9);
#pragma unroll
  for (int c = 0; c < AT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float a = acc[c][t].x, b = acc[c][t].y;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, m);
        b += __shfl_xor_sync(0xffffffffu, b, m);
      }
      if (lane == 0) {
        dst[(0 * A + cib + c) * 9 + t] = a;
        dst[(1 * A + cib + c) * 9 + t] = b;
      }
    }
}

// The same weight gradient with the input tile staged by TMA (see conv3x3_thin_in_tma_kernel:
// box {36, 10, 32} from (x0 - 4, y0 - 1), right halo column dropped behind the rows).
constexpr int kThinWgTmaSmem = 2 * kThinTmaBuf * 4 + 64 + 128;   // two tile buffers + barriers + alignment slack

__global__ void __launch_bounds__(256, 2)
    conv3x3_wgrad_thin_staged_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const float* __restrict__ x,
                                         const float* __restrict__ dy, float* __restrict__ partial,
                                         float* __restrict__ bias_partial, int H, int W, int tiles_x,
                                         int tiles_y, int ntiles) {
  // bias_partial (optional, gridDim.x x 2 floats): per-CTA sums of dy = the layer's bias gradient;
  // every warp holds its pixels' two dY values anyway, warp 0 adds them up
  constexpr int A = 32, AT = 4, R = kThinInRows;
  extern __shared__ unsigned char thin_tma_raw[];
  float* thin_smem = reinterpret_cast<float*>(
      thin_tma_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(thin_tma_raw) & 127u)) & 127u));
  uint64_t* bars = reinterpret_cast<uint64_t*>(thin_smem + 2 * kThinTmaBuf);
  const uint32_t full = (uint32_t)__cvta_generic_to_shared(bars);  // [2] tile landed
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cib = warp * AT;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
  }
  cf acc[AT][9];
#pragma unroll
  for (int c = 0; c < AT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[c][t] = mk(0.0f, 0.0f);
  cf bsum = mk(0.0f, 0.0f);                  // warp 0: sum of dY over this lane's pixels
  const bool want_bias = bias_partial != nullptr && warp == 0;
  const size_t plane = (size_t)H * W;
  __syncthreads();
  auto stage_tile = [&](int tile, int b) {
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(thin_smem + b * kThinTmaBuf);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full + 8 * b),
                 "r"((uint32_t)(kThinTmaBox * 4))
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(&tm_x), "r"(full + 8 * b), "r"(t.x0 - 4), "r"(t.y0 - 1), "r"(t.n * A)
        : "memory");
  };
  auto fetch_halo = [&](int tile, int e) -> float {
    if (e >= A * (R + 2)) return 0.0f;
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    const int c = e / (R + 2), r = e - c * (R + 2);
    const int gy = t.y0 - 1 + r, gx = t.x0 + 32;
    if (gy < 0 || gy >= H || gx >= W) return 0.0f;
    return __ldg(x + ((size_t)t.n * A + c) * plane + (size_t)gy * W + gx);
  };
  int tile = blockIdx.x;
  float h0 = 0.0f, h1 = 0.0f;
  if (tile < ntiles) {
    if (threadIdx.x == 0) stage_tile(tile, 0);
    h0 = fetch_halo(tile, threadIdx.x);
    h1 = fetch_halo(tile, threadIdx.x + 256);
  }
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int b = it & 1;
    const float* cur = thin_smem + b * kThinTmaBuf + cib * (R + 2) * kThinTmaPC;
    const int next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < ntiles) stage_tile(next, b ^ 1);
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    const float* dn = dy + ((size_t)t.n * 2 * H + t.y0) * W + t.x0 + lane;
    float d0[R], d1[R];   // this tile's dY column: in flight while the tile lands
#pragma unroll
    for (int r = 0; r < R; ++r) {
      d0[r] = __ldg(dn + (size_t)r * W);
      d1[r] = __ldg(dn + plane + (size_t)r * W);
    }
    {
      uint32_t done;
      do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;"
            " selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(full + 8 * b), "r"((uint32_t)((it >> 1) & 1))
            : "memory");
      } while (!done);
    }
    {
      float* buf = thin_smem + b * kThinTmaBuf;
      buf[(threadIdx.x + 1) * kThinTmaPC] = h0;
      if (threadIdx.x + 256 < A * (R + 2)) buf[(threadIdx.x + 256 + 1) * kThinTmaPC] = h1;
    }
    if (next < ntiles) {
      h0 = fetch_halo(next, threadIdx.x);
      h1 = fetch_halo(next, threadIdx.x + 256);
    }
    __syncthreads();
    float win[AT][3][3];
#pragma unroll
    for (int c = 0; c < AT; ++c)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        win[c][1][kx] = cur[(c * (R + 2) + 0) * kThinTmaPC + lane + kx + 3];
        win[c][2][kx] = cur[(c * (R + 2) + 1) * kThinTmaPC + lane + kx + 3];
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          win[c][2][kx] = cur[(c * (R + 2) + r + 2) * kThinTmaPC + lane + kx + 3];
        }
      const cf d = mk(d0[r], d1[r]);
#pragma unroll
      for (int c = 0; c < AT; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            acc[c][ky * 3 + kx] =
                f2fma(d, mk(win[c][ky][kx], win[c][ky][kx]), acc[c][ky * 3 + kx]);
      if (want_bias) bsum = f2add(bsum, d);
    }
    __syncthreads();   // everyone is done with buffer b before it becomes a TMA target again
  }
  if (want_bias) {
    float a = bsum.x, b = bsum.y;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, m);
      b += __shfl_xor_sync(0xffffffffu, b, m);
    }
    if (lane == 0) {
      bias_partial[blockIdx.x * 2 + 0] = a;
      bias_partial[blockIdx.x * 2 + 1] = b;
    }
  }
  // fold the 32 pixel-lanes; lane 0 writes the warp's sums in dW order [o][c][tap]
  float* dst = partial + (size_t)blockIdx.x * (A * 2 * 9);
#pragma unroll
  for (int c = 0; c < AT; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float a = acc[c][t].x, b = acc[c][t].y;
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, m);
        b += __shfl_xor_sync(0xffffffffu, b, m);
      }
      if (lane == 0) {
        dst[(0 * A + cib + c) * 9 + t] = a;
        dst[(1 * A + cib + c) * 9 + t] = b;
      }
    }
}

// Weight gradient of the 2 -> 32 layer, same idea: the (tiny) 2-channel input tile
// is staged one tile ahead, warp g owns output channels 4g .. 4g+3 and streams its
// four dY rows straight from global memory (all 8 rows of a tile requested before
// the tile is multiplied).  dw[o][c][tap], o < 32, c < 2.
__global__ void __launch_bounds__(256, 2)
    conv3x3_wgrad_thin_out_staged_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                         float* __restrict__ partial, float* __restrict__ bias_partial,
                                         int H, int W, int tiles_x, int tiles_y, int ntiles) {
  // bias_partial (optional, gridDim.x x 32 floats): per-CTA sums of dy over all pixels = the bias
  // gradient of the layer, a by-product of every thread holding its pixel's dY values
  constexpr int A = 2, B = 32, BT = 4, R = kThinInRows;
  constexpr int kBuf = kThinLead + A * (R + 2) * kThinPC;
  __shared__ __align__(16) float stage[2][kBuf];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int cob = warp * BT;
  cf acc[A][9][BT / 2];
#pragma unroll
  for (int c = 0; c < A; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < BT / 2; ++o) acc[c][t][o] = mk(0.0f, 0.0f);
  float bsum[BT] = {0.0f, 0.0f, 0.0f, 0.0f};
  const size_t plane = (size_t)H * W;
  auto stage_tile = [&](int tile, float* buf) {
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    static_assert(A * (R + 2) % 4 == 0, "rows are staged four at a time");
    for (int pair = warp * 4 + (lane >> 3); pair < A * (R + 2); pair += 32) {
      const int c = pair / (R + 2), r = pair - c * (R + 2);
      thin_stage_row4(buf + kThinLead + pair * kThinPC, x + ((size_t)t.n * A + c) * plane,
                      t.y0 - 1 + r, t.x0, H, W, lane);
    }
  };
  int tile = blockIdx.x;
  if (tile < ntiles) stage_tile(tile, stage[0]);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const float* cur = stage[it & 1] + kThinLead;
    const int next = tile + gridDim.x;
    if (next < ntiles) stage_tile(next, stage[(it + 1) & 1]);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const ThinTile t = thin_tile(tile, tiles_x, tiles_y, R);
    const float* dn = dy + (((size_t)t.n * B + cob) * H + t.y0) * W + t.x0 + lane;
    float d[R][BT];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int o = 0; o < BT; ++o) d[r][o] = __ldg(dn + (size_t)o * plane + (size_t)r * W);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    if (bias_partial != nullptr) {          // two-level sum: tile (8 values) -> CTA
#pragma unroll
      for (int o = 0; o < BT; ++o) {
        float ts = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) ts += d[r][o];
        bsum[o] += ts;
      }
    }
    float win[A][3][3];
#pragma unroll
    for (int c = 0; c < A; ++c)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        win[c][1][kx] = cur[(c * (R + 2) + 0) * kThinPC + lane + kx - 1];
        win[c][2][kx] = cur[(c * (R + 2) + 1) * kThinPC + lane + kx - 1];
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int c = 0; c < A; ++c)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          win[c][0][kx] = win[c][1][kx];
          win[c][1][kx] = win[c][2][kx];
          win[c][2][kx] = cur[(c * (R + 2) + r + 2) * kThinPC + lane + kx - 1];
        }
#pragma unroll
      for (int c = 0; c < A; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int o = 0; o < BT / 2; ++o)
              acc[c][ky * 3 + kx][o] = f2fma(mk(d[r][2 * o], d[r][2 * o + 1]),
                                             mk(win[c][ky][kx], win[c][ky][kx]),
                                             acc[c][ky * 3 + kx][o]);
    }
    __syncthreads();   // `cur` is the staging target of the next iteration
  }
  float* dst = partial + (size_t)blockIdx.x * (A * B * 9);
#pragma unroll
  for (int c = 0; c < A; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int o = 0; o < BT / 2; ++o) {
        float a = acc[c][t][o].x, b = acc[c][t][o].y;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, m);
          b += __shfl_xor_sync(0xffffffffu, b, m);
        }
        if (lane == 0) {
          dst[((cob + 2 * o) * A + c) * 9 + t] = a;
          dst[((cob + 2 * o + 1) * A + c) * 9 + t] = b;
        }
      }
  if (bias_partial != nullptr) {
#pragma unroll
    for (int o = 0; o < BT; ++o) {
      float a = bsum[o];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) a += __shfl_xor_sync(0xffffffffu, a, m);
      if (lane == 0) bias_partial[blockIdx.x * B + cob + o] = a;
    }
  }
}

}  // namespace csmri
