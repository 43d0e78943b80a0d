// Standalone check + timing of csrc/conv_wgrad_tc.cuh (tcgen05 weight gradient, 32 -> 32 channels).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I csmri-refinement_b200/csrc -o tools/wgrad_tc_probe tools/wgrad_tc_probe.cu
#define CSMRI_TC_PROBE 1
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "conv_wgrad_tc.cuh"
using namespace csmri;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encoder() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return (EncodeTiledFn)p;
}
static void make_map(CUtensorMap* m, const float* ptr, int N, int H, int W) {
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N * 32};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  cuuint32_t box[3] = {32, 1, 32};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)ptr, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(3); }
}

static int g_sms = 0;
static void launch(const float* x, const float* dy, float* dw, float* ws, int N, int H, int W, int debug = 0) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWtcSmemBytes);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    attr = true;
  }
  alignas(64) CUtensorMap tx, td;
  make_map(&tx, x, N, H, W);
  make_map(&td, dy, N, H, W);
  const int nitems = N * (W / kWtcPx) * (H / kWtcRows);
  const int grid = nitems < g_sms ? nitems : g_sms;
  conv3x3_wgrad_tc_kernel<<<grid, kWtcThreads, kWtcSmemBytes>>>(tx, td, dy, ws, nullptr, H, W, nitems, debug);
  conv3x3_wgrad_tc_reduce_kernel<<<kWtcReduceBlocks, 32 * kWtcReduceGroups>>>(ws, dw, nullptr, nullptr, grid);
}

int main() {
  int fails = 0;
  const int shapes[4][3] = {{1, 16, 64}, {2, 32, 128}, {3, 16, 192}, {5, 64, 256}};
  for (int si = 0; si < 4; ++si) {
    const int N = shapes[si][0], H = shapes[si][1], W = shapes[si][2];
    const size_t ne = (size_t)N * 32 * H * W;
    std::vector<float> x(ne), dy(ne), out(9216);
    srand(si * 13 + 1);
    for (auto& v : x) v = rand() / (float)RAND_MAX * 2 - 1;
    for (auto& v : dy) v = (rand() / (float)RAND_MAX * 2 - 1) * 0.01f;
    std::vector<double> ref(9216, 0.0);
    for (int n = 0; n < N; ++n) for (int co = 0; co < 32; ++co) for (int ci = 0; ci < 32; ++ci)
      for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
        double s = 0;
        for (int yy = 0; yy < H; ++yy) {
          const int iy = yy + ky - 1;
          if (iy < 0 || iy >= H) continue;
          const float* dr = &dy[(((size_t)n * 32 + co) * H + yy) * W];
          const float* xr = &x[(((size_t)n * 32 + ci) * H + iy) * W];
          for (int xx = 0; xx < W; ++xx) {
            const int ix = xx + kx - 1;
            if (ix < 0 || ix >= W) continue;
            s += (double)dr[xx] * xr[ix];
          }
        }
        ref[((co * 32 + ci) * 3 + ky) * 3 + kx] += s;
      }
    float *dx, *ddy, *ddw, *ws;
    cudaMalloc(&dx, ne * 4); cudaMalloc(&ddy, ne * 4); cudaMalloc(&ddw, 9216 * 4); cudaMalloc(&ws, (size_t)160 * kWtcPartial * 4);
    cudaMemcpy(dx, x.data(), ne * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ddy, dy.data(), ne * 4, cudaMemcpyHostToDevice);
    cudaMemset(ddw, 0xff, 9216 * 4);
    launch(dx, ddy, ddw, ws, N, H, W, getenv("WTC_DEBUG") ? atoi(getenv("WTC_DEBUG")) : 0);   // skip bits, for sanitizer experiments
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), ddw, 9216 * 4, cudaMemcpyDeviceToHost);
    double num = 0, den = 0, worst = 0; int wi = 0;
    for (int i = 0; i < 9216; ++i) { const double d = out[i] - ref[i]; num += d * d; den += ref[i] * ref[i]; if (!(fabs(d) <= worst)) { worst = fabs(d); wi = i; } }
    const double rel = sqrt(num / den);
    printf("N=%d H=%d W=%d: %s  rel-L2 %.3e  worst |diff| %.3e at (co %d, ci %d, ky %d, kx %d): got %g want %g  %s\n", N, H, W,
           cudaGetErrorString(e), rel, worst, wi / 288, (wi / 9) % 32, (wi / 3) % 3, wi % 3, out[wi], ref[wi],
           (rel < 4e-6 && e == cudaSuccess) ? "ok" : "FAIL");
    fails += !(rel < 4e-6 && e == cudaSuccess);
    cudaFree(dx); cudaFree(ddy); cudaFree(ddw); cudaFree(ws);
    if (e != cudaSuccess) return 2;
  }
  const int tshape[2][3] = {{32, 256, 256}, {20, 512, 512}};
  for (int ti = 0; ti < 2; ++ti) {
    const int N = tshape[ti][0], H = tshape[ti][1], W = tshape[ti][2];
    const size_t ne = (size_t)N * 32 * H * W;
    float *dx, *ddy, *ddw, *ws;
    cudaMalloc(&dx, ne * 4); cudaMalloc(&ddy, ne * 4); cudaMalloc(&ddw, 9216 * 4); cudaMalloc(&ws, (size_t)160 * kWtcPartial * 4);
    cudaMemset(dx, 0, ne * 4); cudaMemset(ddy, 0, ne * 4);
    for (int i = 0; i < 3; ++i) launch(dx, ddy, ddw, ws, N, H, W);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) launch(dx, ddy, ddw, ws, N, H, W);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double flop = 2.0 * 9 * 32 * 32 * (double)N * H * W;
    printf("N=%d %dx%d: %.3f ms per layer  %.1f TFLOP/s (fp32-equivalent)  %.0f GB/s of x + dy traffic (%s)\n", N, H, W, ms,
           flop / ms / 1e9, 2.0 * ne * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    const int dbg[8] = {0, 32, 15, 15 | 32, 16, 31, 1, 1 | 32};
    for (int di = 0; di < 1; ++di) {
    launch(dx, ddy, ddw, ws, N, H, W, 128 | dbg[di]);
    cudaDeviceSynchronize();
    long long pr[8]; cudaMemcpyFromSymbol(pr, tc_prof, sizeof(pr));
    const double steps = (double)((N * (W / kWtcPx) * (H / kWtcRows) + g_sms - 1) / g_sms) * kWtcRows;
    printf("    [skip bits %2d: 1 staging, 2 drain, 4 splitter, 8 TMA, 16 MMA, 32 fence] CTA 0, cycles per step: staging waits A buffer %.0f, waits dy row %.0f | MMA thread fence %.0f | MMA waits drain %.0f, waits input rows %.0f, waits A %.0f, staging (kx 2) works %.0f, MMA thread issues commits %.0f\n",
           dbg[di], pr[0] / steps, pr[1] / steps, pr[2] / steps, pr[3] / steps, pr[4] / steps, pr[5] / steps, pr[6] / steps, pr[7] / steps);
    long long cc[256]; cudaMemcpyFromSymbol(cc, wtc_cta_cycles, sizeof(cc));
    long long mn = cc[0], mx = cc[0]; double av = 0;
    for (int i = 0; i < g_sms; ++i) { if (cc[i] < mn) mn = cc[i]; if (cc[i] > mx) mx = cc[i]; av += cc[i]; }
    printf("    MMA-warp lifetime over the %d CTAs (cycles): min %lld, mean %.0f, max %lld; CTA 0: %lld = %.0f per step\n", g_sms, mn, av / g_sms, mx, cc[0], cc[0] / steps);
    }
    cudaFree(dx); cudaFree(ddy); cudaFree(ddw); cudaFree(ws);
  }
  printf(fails ? "WGRAD TC PROBE FAILED\n" : "WGRAD TC PROBE OK\n");
  return fails != 0;
}
