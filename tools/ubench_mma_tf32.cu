// Throughput of the legacy warp-level tensor path on this GPU: mma.sync.m16n8k8 TF32
// (fp32 accumulate), register operands only, 8 independent accumulator tiles per warp.
// Go / no-go number for a 3xTF32 (fp32-accurate) implicit-GEMM path for RecNet's 32->32
// convolutions: it needs 3 MMAs per product, so it only pays if this rate / 3 clearly beats
// the ~55 TFLOP/s cuDNN's fp32 kernels reach on those layers.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma_tf32 tools/ubench_mma_tf32.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int TILES>
__global__ void __launch_bounds__(256) mma_loop(float* out, int iters) {
  float c[TILES][4];
#pragma unroll
  for (int t = 0; t < TILES; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f810000u, 0x3f820000u, 0x3f830000u};
  unsigned b[2] = {0x3f000000u + threadIdx.x, 0x3f010000u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES> static void run(int ctas_per_sm) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm, iters = 20000;
  float* out; cudaMalloc(&out, (size_t)grid * 256 * 4);
  mma_loop<TILES><<<grid, 256>>>(out, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mma_loop<TILES><<<grid, 256>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = (double)grid * 8 /*warps*/ * iters * TILES * 2.0 * 16 * 8 * 8;
  printf("  mma.sync m16n8k8 tf32: %d acc tiles/warp, %d CTAs/SM x 8 warps: %7.1f TFLOP/s (%s)\n", TILES,
         ctas_per_sm, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<4>(1); run<8>(1); run<8>(2); run<8>(4); run<16>(2);
  return 0;
}
