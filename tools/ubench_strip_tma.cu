// Memory-pattern ceiling for a strip pipeline that is TMA on BOTH sides: a (B,2,N,N) fp32
// tensor is copied strip by strip (box = CW columns x N rows x 2 planes) with
// cp.async.bulk.tensor loads into shared memory and cp.async.bulk.tensor STORES out of it
// (UTMALDG / UTMASTG), one elected thread per persistent CTA, NB buffers in flight.
// Question: would a TMA tile store lift the adjoint's 1R:1W ceiling
// above what 64-byte STG.64 row segments reach (tools/ubench_strip_copy.cu: 47.8 us)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_strip_tma tools/ubench_strip_tma.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#ifndef UB_N
#define UB_N 256
#define UB_B 256
#endif
constexpr int N = UB_N;
constexpr int B = UB_B;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CW, int NB, int LA>
__global__ void __launch_bounds__(32) tma_copy(const __grid_constant__ CUtensorMap tin,
                                               const __grid_constant__ CUtensorMap tout, int ntiles) {
  constexpr int kTile = 2 * N * CW * 4;
  constexpr int nstrips = N / CW;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NB * kTile);
  if (threadIdx.x != 0) return;
  for (int i = 0; i < NB; ++i)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const int n = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  for (int it = 0; it < n + LA; ++it) {
    if (it < n) {
      const int s = it % NB;
      if (it >= NB) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NB - LA - 1) : "memory");
      const int tile = blockIdx.x + it * gridDim.x;
      const int b = tile / nstrips, strip = tile - b * nstrips;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[s])), "r"(kTile) : "memory");
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(s32(smem + s * kTile)), "l"(&tin), "r"(s32(&bars[s])), "r"(strip * CW), "r"(0), "r"(b * 2) : "memory");
    }
    const int j = it - LA;
    if (j >= 0) {
      const int s = j % NB;
      const uint32_t parity = (j / NB) & 1;
      uint32_t done;
      do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(s32(&bars[s])), "r"(parity) : "memory");
      } while (!done);
      const int tile = blockIdx.x + j * gridDim.x;
      const int b = tile / nstrips, strip = tile - b * nstrips;
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                   ::"l"(&tout), "r"(s32(smem + s * kTile)), "r"(strip * CW), "r"(0), "r"(b * 2) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc;

static CUtensorMap make_map(float* ptr, int CW) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)2 * B};
  cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
  cuuint32_t box[3] = {(cuuint32_t)CW, (cuuint32_t)N, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
  return m;
}

static float *A[2], *O[2];

template <int CW, int NB, int LA> static void report(int ctas_per_sm) {
  static_assert(N <= 256, "one box covers all rows");
  constexpr int kTile = 2 * N * CW * 4;
  constexpr int smem = NB * kTile + 128;
  auto kern = tma_copy<CW, NB, LA>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int ntiles = B * (N / CW);
  int grid = sms * ctas_per_sm; if (grid > ntiles) grid = ntiles;
  CUtensorMap mi[2] = {make_map(A[0], CW), make_map(A[1], CW)};
  CUtensorMap mo[2] = {make_map(O[0], CW), make_map(O[1], CW)};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) kern<<<grid, 32, smem>>>(mi[i & 1], mo[i & 1], ntiles);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) kern<<<grid, 32, smem>>>(mi[i & 1], mo[i & 1], ntiles);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  // verify one element pattern
  printf("  TMA load + TMA store  CW=%-3d %d buffers (%d ahead) %d CTA/SM  %8.2f us %6.0f GB/s (%s)\n", CW, NB, LA,
         ctas_per_sm, ms * 1e3, (double)B * 2 * N * N * 4 * 2 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  enc = (EncodeTiledFn)p;
  const size_t elems = (size_t)B * 2 * N * N;
  for (int i = 0; i < 2; ++i) { cudaMalloc(&A[i], elems * 4); cudaMalloc(&O[i], elems * 4);
    cudaMemset(A[i], 1, elems * 4); cudaMemset(O[i], 0, elems * 4); }
  printf("adjoint-shaped traffic (1 read + 1 write stream) N=%d B=%d, TMA on both sides\n", N, B);
  report<16, 4, 2>(1);
  report<16, 6, 3>(1);
  report<16, 3, 1>(2);
  report<16, 2, 1>(3);
  report<32, 3, 1>(1);
  report<32, 2, 1>(1);
  report<64, 1, 0>(1);
  // correctness spot check of the last configuration's output buffer
  float h[4]; cudaMemcpy(h, O[1] + 12345, 16, cudaMemcpyDeviceToHost);
  unsigned u; memcpy(&u, &h[0], 4);
  printf("  check: out word = 0x%08x (expect 0x01010101)\n", u);
  return 0;
}
