// Micro-benchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 issue throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench_fp32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
  float2 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // 2 scalar FFMA
        acc[i].x = fmaf(acc[i].x, a, b);
        acc[i].y = fmaf(acc[i].y, a, b);
      } else if (MODE == 1) {  // 1 packed FFMA2
        asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%2}; mov.b64 rc, {%3,%3};"
                     " fma.rn.f32x2 ra, ra, rb, rc; mov.b64 {%0,%1}, ra; }"
                     : "+f"(acc[i].x), "+f"(acc[i].y) : "f"(a), "f"(b));
      } else if (MODE == 2) {  // 2 scalar FADD
        acc[i].x = acc[i].x + b;
        acc[i].y = acc[i].y + a;
      } else {  // 1 packed FADD2
        asm volatile("{ .reg .b64 ra, rb; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3};"
                     " add.rn.f32x2 ra, ra, rb; mov.b64 {%0,%1}, ra; }"
                     : "+f"(acc[i].x), "+f"(acc[i].y) : "f"(b), "f"(a));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, float* d, int blocks) {
  const int iters = 8192;
  k<MODE><<<blocks, 256>>>(d, 64, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(d, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double lane_ops = (double)blocks * 256 * iters * 16;  // fp32 lane-operations
  printf("%-12s %8.3f ms  %8.2f T lane-ops/s  (%s)\n", name, ms, lane_ops / ms / 1e9,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SMs=%d clock=%d kHz\n", sms, clk);
  const int blocks = sms * 8;
  float* d; cudaMalloc(&d, (size_t)blocks * 256 * 4);
  run<0>("FFMA x2", d, blocks);
  run<1>("FFMA2", d, blocks);
  run<2>("FADD x2", d, blocks);
  run<3>("FADD2", d, blocks);
  return 0;
}
