// Unit probe for the tcgen05 path of the 32->32 convolution (kind::tf32, M=128, N=32, K=8 per
// instruction, operands in shared memory in the NON-swizzled K-major canonical layout, accumulator
// in TMEM).  Checks, against exact integer-valued references:
//   1. descriptor / instruction-descriptor encoding and the TMEM -> register read-back layout;
//   2. that starting the A descriptor 16 bytes (= one row of the canonical layout) further on
//      selects rows m+1 .. m+128: the property the implicit-GEMM convolution uses for its dx taps;
//   3. accumulation across several MMAs (K = 32 = 4 instructions, then a second operand pair).
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 32, K = 32, ROWS = 144;   // A holds ROWS >= M + shift rows

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
  return d;                                    // base offset 0, layout type 0 = no swizzle
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ a, const float* __restrict__ b,
                                             float* __restrict__ out, int shift, int second_pair) {
  // canonical no-swizzle K-major: [k/4][row][k%4]; rows 16 B apart, 8-row groups 128 B apart (dense)
  __shared__ __align__(128) float A_s[K / 4][ROWS][4];
  __shared__ __align__(128) float B_s[K / 4][N][4];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < ROWS * K; i += 128) { const int m = i / K, k = i % K; A_s[k / 4][m][k % 4] = a[i]; }
  for (int i = tid; i < N * K; i += 128) { const int n = i / K, k = i % K; B_s[k / 4][n][k % 4] = b[n * K + k]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(s32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic smem writes -> async proxy (tensor core)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t a0 = s32(&A_s[0][0][0]) + shift * 16, b0 = s32(&B_s[0][0][0]);
    const uint32_t a_lbo = ROWS * 16, b_lbo = N * 16;
    int first = 1;
    for (int rep = 0; rep < 1 + second_pair; ++rep)
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t da = make_desc(a0 + 2 * ks * a_lbo, a_lbo, 128);
        const uint64_t db = make_desc(b0 + 2 * ks * b_lbo, b_lbo, 128);
        const uint32_t acc = first ? 0u : 1u;
        asm volatile(
            "{ .reg .pred p; setp.ne.b32 p, %4, 0;"
            " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p; }" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u)
            : "memory");
        first = 0;
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(s32(&bar)), "r"(0u) : "memory");
  } while (!done);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < N; ++n) out[tid * N + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<float> a(ROWS * K), b(N * K), out(M * N);
  for (int m = 0; m < ROWS; ++m) for (int k = 0; k < K; ++k) a[m * K + k] = (float)((m * 7 + k * 3) % 17 - 8) * 0.125f;
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) b[n * K + k] = (float)((k * 5 + n * 11) % 13 - 6) * 0.25f;
  float *da, *db, *dout;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, out.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  int bad_total = 0;
  for (int test = 0; test < 4; ++test) {
    const int shift = (test == 1) ? 1 : (test == 2 ? 11 : 0), second = (test == 3);
    cudaMemset(dout, 0xff, out.size() * 4);
    probe<<<1, 128>>>(da, db, dout, shift, second);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double worst = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)a[(m + shift) * K + k] * b[n * K + k];
      ref *= (1 + second);
      const double d = fabs(out[m * N + n] - ref);
      if (d > worst) worst = d;
      if (d != 0) ++bad;
    }
    printf("test %d (row shift %2d, %d operand pairs): %s, %d / %d mismatches, worst |diff| %.3g; D[5][3]=%g\n", test, shift,
           1 + second, cudaGetErrorString(e), bad, M * N, worst, out[5 * N + 3]);
    bad_total += bad + (e != cudaSuccess);
  }
  printf(bad_total ? "UMMA PROBE FAILED\n" : "UMMA PROBE OK\n");
  return bad_total != 0;
}
