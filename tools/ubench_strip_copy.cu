// Memory-pattern ceiling for the column-strip access of the DC kernels:
// copy a (B,2,N,N) fp32 tensor (N = 256 by default) strip by strip (CW columns x 256 rows x 2
// planes per CTA, all loads issued before the stores, like the real kernel)
// with 32-bit or 128-bit accesses, against a linear copy.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_strip tools/ubench_strip_copy.cu
#include <cstdio>
#include <cuda_runtime.h>

#ifndef UB_N
#define UB_N 256
#define UB_B 256
#endif
constexpr int N = UB_N;   // -DUB_N=512 -DUB_B=64 for the other slice sizes
constexpr int B = UB_B;

__device__ __forceinline__ float ldf(const float* p) {
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }
__device__ __forceinline__ void stf(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ float4 ld4(const float* p) {
  float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ void st4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

// THREADS threads copy a tile of 512 rows (2 planes x 256) x CW columns
template <int CW, int VEC, int THREADS, int TWO>
__global__ void __launch_bounds__(THREADS) strip_copy(const float* __restrict__ a,
                                                      const float* __restrict__ b2,
                                                      float* __restrict__ out) {
  constexpr int nstrips = N / CW;
  const int b = blockIdx.x / nstrips, strip = blockIdx.x - b * nstrips;
  const size_t base = (size_t)b * 2 * N * N + (size_t)strip * CW;
  constexpr int LPR = CW / VEC;                 // lanes per row
  constexpr int RPP = THREADS / LPR;            // rows per pass
  constexpr int IT = 2 * N / RPP;               // passes
  const int c = (threadIdx.x % LPR) * VEC;
  const int r0 = threadIdx.x / LPR;
  if (VEC == 1) {
    float v[IT];
#pragma unroll
    for (int n = 0; n < IT; ++n) {
      v[n] = ldf(a + base + (size_t)(r0 + n * RPP) * N + c);
      if (TWO) v[n] += ldf(b2 + base + (size_t)(r0 + n * RPP) * N + c);
    }
#pragma unroll
    for (int n = 0; n < IT; ++n) stf(out + base + (size_t)(r0 + n * RPP) * N + c, v[n]);
  } else {
    float4 v[IT];
#pragma unroll
    for (int n = 0; n < IT; ++n) {
      v[n] = ld4(a + base + (size_t)(r0 + n * RPP) * N + c);
      if (TWO) { float4 w = ld4(b2 + base + (size_t)(r0 + n * RPP) * N + c);
                 v[n].x += w.x; v[n].y += w.y; v[n].z += w.z; v[n].w += w.w; }
    }
#pragma unroll
    for (int n = 0; n < IT; ++n) st4(out + base + (size_t)(r0 + n * RPP) * N + c, v[n]);
  }
}

template <int TWO>
__global__ void __launch_bounds__(256) linear_copy(const float* __restrict__ a,
                                                   const float* __restrict__ b2,
                                                   float* __restrict__ out) {
  size_t i = ((size_t)blockIdx.x * 2048 + threadIdx.x) * 4;
  float4 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    v[k] = ld4(a + i + k * 1024);
    if (TWO) { float4 w = ld4(b2 + i + k * 1024); v[k].x += w.x; v[k].y += w.y; v[k].z += w.z; v[k].w += w.w; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) st4(out + i + k * 1024, v[k]);
}

static float *A[2], *C[2], *O;
static size_t elems;

template <typename F> static float timeit(F launch, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch(i);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) launch(i);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

template <int CW, int VEC, int THREADS, int TWO> static void report(const char* name) {
  if (N % CW != 0 || (2 * N) % (THREADS / (CW / VEC)) != 0 ||
      2 * N / (THREADS / (CW / VEC)) * VEC > 200) return;   // shape does not tile / fit registers
  float ms = timeit([](int i) {
    strip_copy<CW, VEC, THREADS, TWO><<<B * (N / CW), THREADS>>>(A[i & 1], C[i & 1], O); }, 20);
  printf("  strip CW=%-3d %3d-bit %4d thr %-10s %8.2f us %6.0f GB/s (%s)\n", CW, VEC * 32, THREADS, name,
         ms * 1e3, (double)elems * 4 * (2 + TWO) / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}
template <int TWO> static void report_linear() {
  float ms = timeit([](int i) {
    linear_copy<TWO><<<(unsigned)(elems / 8192), 256>>>(A[i & 1], C[i & 1], O); }, 20);
  printf("  linear 128-bit                         %8.2f us %6.0f GB/s (%s)\n", ms * 1e3,
         (double)elems * 4 * (2 + TWO) / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  elems = (size_t)B * 2 * N * N;
  for (int i = 0; i < 2; ++i) { cudaMalloc(&A[i], elems * 4); cudaMalloc(&C[i], elems * 4);
    cudaMemset(A[i], 0, elems * 4); cudaMemset(C[i], 0, elems * 4); }
  cudaMalloc(&O, elems * 4);
  printf("adjoint-shaped traffic (1 read + 1 write stream), %zu MiB per stream\n", elems * 4 >> 20);
  report_linear<0>();
  report<16, 1, 256, 0>("");
  report<32, 1, 512, 0>("");
  report<32, 1, 256, 0>("");
  report<64, 1, 512, 0>("");
  report<16, 4, 256, 0>("");
  report<16, 4, 128, 0>("");
  report<32, 4, 256, 0>("");
  report<32, 4, 512, 0>("");
  report<64, 4, 256, 0>("");
  report<64, 4, 512, 0>("");
  report<128, 4, 512, 0>("");
  report<256, 4, 1024, 0>("");
  printf("forward-shaped traffic (2 read + 1 write streams)\n");
  report_linear<1>();
  report<16, 1, 256, 1>("");
  report<32, 1, 512, 1>("");
  report<16, 4, 256, 1>("");
  report<32, 4, 256, 1>("");
  report<64, 4, 512, 1>("");
  report<128, 4, 512, 1>("");
  report<256, 4, 1024, 1>("");
  return 0;
}
