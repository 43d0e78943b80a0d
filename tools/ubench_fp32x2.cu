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

// The weight-gradient inner product: 18 accumulator pairs, two dY pairs, nine scalar
// window values per "pixel"; MODE 0 packed FFMA2 with a scalar operand, MODE 1 scalar FFMA.
template <int MODE>
__global__ void __launch_bounds__(256, 4) kw(float* out, int iters, const float* in) {
  float2 acc[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) acc[i] = make_float2(0.f, 0.f);
  float2 d01 = make_float2(in[threadIdx.x], in[threadIdx.x + 1]);
  float2 d23 = make_float2(in[threadIdx.x + 2], in[threadIdx.x + 3]);
  float x[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) x[i] = in[threadIdx.x + 4 + i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {     // four "pixels" with rotating window registers
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float xv = x[(t + 3 * p) % 12];
        if (MODE == 0) {
          asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; mov.b64 rc, {%4,%4};"
                       " fma.rn.f32x2 ra, rb, rc, ra; mov.b64 {%0,%1}, ra; }"
                       : "+f"(acc[t].x), "+f"(acc[t].y) : "f"(d01.x), "f"(d01.y), "f"(xv));
          asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%0,%1}; mov.b64 rb, {%2,%3}; mov.b64 rc, {%4,%4};"
                       " fma.rn.f32x2 ra, rb, rc, ra; mov.b64 {%0,%1}, ra; }"
                       : "+f"(acc[9 + t].x), "+f"(acc[9 + t].y) : "f"(d23.x), "f"(d23.y), "f"(xv));
        } else {
          acc[t].x = fmaf(d01.x, xv, acc[t].x);
          acc[t].y = fmaf(d01.y, xv, acc[t].y);
          acc[9 + t].x = fmaf(d23.x, xv, acc[9 + t].x);
          acc[9 + t].y = fmaf(d23.y, xv, acc[9 + t].y);
        }
      }
      d01.x += 1e-9f;   // keep the operands loop-variant
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 18; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void runw(const char* name, float* d, const float* in, int blocks) {
  const int iters = 2048;
  kw<MODE><<<blocks, 256>>>(d, 8, in);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kw<MODE><<<blocks, 256>>>(d, iters, in);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double lane_ops = (double)blocks * 256 * iters * 4 * 36;  // fp32 lane-FMAs
  printf("%-28s %8.3f ms  %8.2f T lane-FMA/s  (%s)\n", name, ms, lane_ops / ms / 1e9,
         cudaGetErrorString(cudaGetLastError()));
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
  float* in; cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  runw<0>("wgrad pattern, FFMA2", d, in, sms * 4);
  runw<1>("wgrad pattern, FFMA", d, in, sms * 4);
  return 0;
}
