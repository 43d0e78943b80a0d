// Standalone check + timing of csrc/conv_tc.cuh (tcgen05 3xTF32 32->32 convolution).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I csmri-refinement_b200/csrc -o tools/conv_tc_probe tools/conv_tc_probe.cu
#define CSMRI_TC_PROBE 1
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "conv_tc.cuh"
using namespace csmri;

static void reference(const std::vector<float>& x, const std::vector<float>& w, const std::vector<float>& b,
                      std::vector<double>& y, int N, int H, int W, float slope, int tf) {
  for (int n = 0; n < N; ++n) for (int co = 0; co < 32; ++co) for (int yy = 0; yy < H; ++yy) for (int xx = 0; xx < W; ++xx) {
    double s = b.empty() ? 0.0 : b[co];
    for (int ci = 0; ci < 32; ++ci) for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
      const int iy = yy + ky - 1, ix = xx + kx - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const double wv = tf ? w[((ci * 32 + co) * 3 + (2 - ky)) * 3 + (2 - kx)] : w[((co * 32 + ci) * 3 + ky) * 3 + kx];
      s += wv * x[(((size_t)n * 32 + ci) * H + iy) * W + ix];
    }
    if (slope > 0 && s < 0) s *= slope;
    y[(((size_t)n * 32 + co) * H + yy) * W + xx] = s;
  }
}

static uint32_t* g_signs = nullptr;   // != nullptr: time the MASKED kernel (data gradient x LeakyReLU derivative)
static int launch(const float* x, const float* w, const float* b, float* y, int N, int H, int W, float slope, int tf, int debug = 0) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv3x3_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    cudaFuncSetAttribute(conv3x3_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
    attr = true;
  }
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int nitems = N * (W / kTcM) * (H / kTcRowBlock);
  const int grid = nitems < sms ? nitems : sms;
  if (g_signs != nullptr)
    conv3x3_tc_kernel<true><<<grid, kTcThreads, kTcSmemBytes>>>(x, w, nullptr, y, g_signs, nullptr, H, W, nitems, slope, 1, debug);
  else
    conv3x3_tc_kernel<false><<<grid, kTcThreads, kTcSmemBytes>>>(x, w, b, y, nullptr, nullptr, H, W, nitems, slope, tf, debug);
  return 0;
}

int main() {
  int fails = 0;
  const int shapes[3][3] = {{1, 16, 128}, {2, 32, 256}, {3, 48, 128}};
  for (int si = 0; si < 3; ++si) for (int tf = 0; tf < 2; ++tf) {
    const int N = shapes[si][0], H = shapes[si][1], W = shapes[si][2];
    const size_t ne = (size_t)N * 32 * H * W;
    std::vector<float> x(ne), w(32 * 32 * 9), b(32), out(ne);
    srand(si * 7 + tf);
    for (auto& v : x) v = rand() / (float)RAND_MAX * 2 - 1;
    for (auto& v : w) v = (rand() / (float)RAND_MAX * 2 - 1) * 0.1f;
    for (auto& v : b) v = rand() / (float)RAND_MAX - 0.5f;
    std::vector<double> ref(ne);
    const float slope = tf ? 0.0f : 0.01f;
    std::vector<float> bb = tf ? std::vector<float>() : b;
    reference(x, w, bb, ref, N, H, W, slope, tf);
    float *dx, *dw, *db, *dy;
    cudaMalloc(&dx, ne * 4); cudaMalloc(&dw, w.size() * 4); cudaMalloc(&db, 128); cudaMalloc(&dy, ne * 4);
    cudaMemcpy(dx, x.data(), ne * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), 128, cudaMemcpyHostToDevice);
    cudaMemset(dy, 0xff, ne * 4);
    launch(dx, dw, tf ? nullptr : db, dy, N, H, W, slope, tf);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), dy, ne * 4, cudaMemcpyDeviceToHost);
    double num = 0, den = 0, worst = 0;
    for (size_t i = 0; i < ne; ++i) { const double d = out[i] - ref[i]; num += d * d; den += ref[i] * ref[i]; if (fabs(d) > worst) worst = fabs(d); }
    const double rel = sqrt(num / den);
    printf("N=%d H=%d W=%d transpose_flip=%d: %s  rel-L2 %.3e  worst |diff| %.3e  %s\n", N, H, W, tf, cudaGetErrorString(e), rel,
           worst, (rel < 4e-6 && e == cudaSuccess) ? "ok" : "FAIL");
    fails += !(rel < 4e-6 && e == cudaSuccess);
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dy);
    if (e != cudaSuccess) return 2;
  }
  // timing at the RecNet D5C5 layer shape and at the 1-recnet.json shape
  const int tshape[2][3] = {{32, 256, 256}, {20, 512, 512}};
  for (int ti = 0; ti < 2; ++ti) {
    const int N = tshape[ti][0], H = tshape[ti][1], W = tshape[ti][2];
    const size_t ne = (size_t)N * 32 * H * W;
    float *dx, *dw, *db, *dy;
    cudaMalloc(&dx, ne * 4); cudaMalloc(&dw, 9216 * 4); cudaMalloc(&db, 128); cudaMalloc(&dy, ne * 4);
    cudaMemset(dx, 0, ne * 4); cudaMemset(dw, 0, 9216 * 4); cudaMemset(db, 0, 128);
    const int dbg_list[9] = {0, 128, 128 | 55, 128 | 50, 128 | 7, 128 | 2, 128 | 1, 0, 128};
    for (int di = 0; di < 9; ++di) {
      const int debug = dbg_list[di];
      if (ti == 1 && debug) break;
      if (di == 7) {                        // the last two entries: the MASKED kernel
        cudaMalloc(&g_signs, (size_t)N * H * W * 4);
        cudaMemset(g_signs, 0x5a, (size_t)N * H * W * 4);
        printf("  -- MASKED kernel (data gradient x LeakyReLU derivative from sign words) --\n");
      }
      for (int i = 0; i < 3; ++i) launch(dx, dw, db, dy, N, H, W, 0.01f, 0, debug);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      const int reps = 20;
      for (int i = 0; i < reps; ++i) launch(dx, dw, db, dy, N, H, W, 0.01f, 0, debug);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
      const double flop = 2.0 * 9 * 32 * 32 * (double)N * H * W;
      printf("N=%d %dx%d [debug %2d: skip mma %d, loads %d, stores %d]: %.3f ms per layer  %.1f TFLOP/s (fp32-equivalent)  %.0f GB/s of in+out traffic (%s)\n",
             N, H, W, debug, debug & 1, (debug >> 1) & 1, (debug >> 2) & 1, ms, flop / ms / 1e9, 2.0 * ne * 4 / ms / 1e6,
             cudaGetErrorString(cudaGetLastError()));
      if (debug & 128) {
        long long pr[8]; cudaMemcpyFromSymbol(pr, tc_prof, sizeof(pr));
        const double rows = (double)((N * (W / kTcM) * (H / kTcRowBlock) + 147) / 148) * (kTcRowBlock + 2);
        printf("    CTA 0, cycles per staged row: producer waits slot %.0f | MMA waits item buffer %.0f, waits row %.0f | epilogue waits row %.0f | MMA thread total %.0f\n",
               2 * pr[0] / rows, pr[1] / rows, pr[2] / rows, pr[3] / rows, pr[5] / rows);
      }
    }
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dy);
    if (g_signs != nullptr) { cudaFree(g_signs); g_signs = nullptr; }
  }
  printf(fails ? "CONV TC PROBE FAILED\n" : "CONV TC PROBE OK\n");
  return fails != 0;
}
