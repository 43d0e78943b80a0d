// Bias + LeakyReLU around RecNet's convolutions (models/recnet.py:45-48), fused:
//   forward   y = lrelu(z + b[c])            in place, one pass (torch: bias add pass + activation pass)
//   backward  gz = gy * (y > 0 ? 1 : slope)  and  gb[c] = sum_{n,h,w} gz   in one pass
//             (torch: leaky_relu_backward pass + a reduction pass over gz)
// y > 0 <=> z + b > 0 for slope > 0, so the activation output is all the backward
// needs (what nn.LeakyReLU(inplace=True) keeps as well).  The bias gradient is
// summed in a fixed order: per (n, c, chunk) partials, then one warp per channel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csmri {

constexpr int kEpChunks = 8;   // partial sums per (n, c) plane

// grid (chunks, N*C); plane = hw4 float4s
__global__ void __launch_bounds__(256)
    bias_lrelu_kernel(float4* __restrict__ z, const float* __restrict__ bias, int C, int hw4,
                      float slope) {
  const int p = blockIdx.y;
  const float b = __ldg(bias + p % C);
  float4* row = z + (size_t)p * hw4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw4; i += gridDim.x * blockDim.x) {
    float4 v = row[i];
    v.x += b; v.y += b; v.z += b; v.w += b;
    v.x = v.x > 0.0f ? v.x : v.x * slope;
    v.y = v.y > 0.0f ? v.y : v.y * slope;
    v.z = v.z > 0.0f ? v.z : v.z * slope;
    v.w = v.w > 0.0f ? v.w : v.w * slope;
    row[i] = v;
  }
}

// grid (kEpChunks, N*C)
__global__ void __launch_bounds__(256)
    bias_lrelu_backward_kernel(const float4* __restrict__ gy, const float4* __restrict__ y,
                               float4* __restrict__ gz, float* __restrict__ partial, int hw4,
                               float slope) {
  const int p = blockIdx.y;
  const float4* g = gy + (size_t)p * hw4;
  const float4* a = y + (size_t)p * hw4;
  float4* o = gz + (size_t)p * hw4;
  float acc = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw4; i += gridDim.x * blockDim.x) {
    float4 v = __ldg(g + i);
    const float4 t = __ldg(a + i);
    v.x = t.x > 0.0f ? v.x : v.x * slope;
    v.y = t.y > 0.0f ? v.y : v.y * slope;
    v.z = t.z > 0.0f ? v.z : v.z * slope;
    v.w = t.w > 0.0f ? v.w : v.w * slope;
    o[i] = v;
    acc += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  __shared__ float s_acc[8];
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) acc += s_acc[w];
    partial[(size_t)p * kEpChunks + blockIdx.x] = acc;
  }
}

// one warp per channel: gb[c] = sum_n sum_chunk partial[(n*C + c)*kEpChunks + chunk]
__global__ void __launch_bounds__(32)
    bias_grad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ gb, int N, int C) {
  const int c = blockIdx.x;
  float acc = 0.0f;
  for (int i = threadIdx.x; i < N * kEpChunks; i += 32) {
    const int n = i / kEpChunks, k = i - n * kEpChunks;
    acc += partial[((size_t)n * C + c) * kEpChunks + k];
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if (threadIdx.x == 0) gb[c] = acc;
}

}  // namespace csmri
