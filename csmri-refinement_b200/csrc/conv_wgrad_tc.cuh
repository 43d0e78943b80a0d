// Weight gradient of RecNet's 32 -> 32 channel 3x3 convolutions (the backward-weight
// half of models/recnet.py:37-44 in Runner._train_step, training/runner.py:154-178) on
// the tcgen05 tensor cores, with the same error-compensated TF32 split as conv_tc.cuh:
//     dw[co][ci][ky][kx] = sum_{n,y,x} dy[n][co][y][x] * x[n][ci][y+ky-1][x+kx-1]
// is a (32 x 288) GEMM whose reduction dimension is every pixel of the batch.
//
// One tensor-core instruction covers all nine taps of an 8-pixel run of one row:
//     D[(kx, co) 96 of 128 lanes][(ky, ci) 96 columns] += A[(kx, co)][8 px] * B[(ky, ci)][8 px]^T
// * A = the dy row, three copies shifted by 1 - kx pixels, in TENSOR MEMORY (lane =
//   (kx, co), column = pixel; tcgen05.mma reads its A operand from TMEM): a shift
//   along the reduction dimension cannot be expressed in a shared-memory descriptor
//   (K advances in 32-byte steps), but the thread that writes lane (kx, co) simply
//   starts one register further left or right.  Three staging warps (kx = 0, 1, 2)
//   read the row from shared memory, split it into hi / lo and tcgen05.st it.
// * B = the three input rows y-1, y, y+1, UNshifted, stacked along N: each row
//   segment (32 channels x 32 pixels = one 128-byte-swizzled TMA box) is loaded
//   ONCE into a ring of row slots and serves as ky = 2, 1, 0 of three consecutive
//   steps; the three slots of a step are adjacent in shared memory, so one
//   descriptor with N = 96 spans them.  (Ring of 6 + the first two slots mirrored
//   behind the last, so a window never wraps.)  Image borders cost nothing: TMA
//   zero-fills the out-of-range rows, the staging threads zero the two halo pixels.
// * the raw fp32 box IS the hi operand (the tensor core ignores the 13 low mantissa
//   bits, tools/umma_probe2.cu); four splitter warps write lo = x - trunc(x) next to it.
// * 24 instructions (8 k-steps x {lo*hi, hi*lo, hi*hi}) of 128 x 96 x 8 per 64-pixel
//   row segment; the accumulator stays in tensor memory and is drained into fp32
//   registers after every row segment, double buffered (the tensor core rounds its
//   accumulator toward zero after every instruction: rel-L2 against float64 is 1.8e-6
//   with a drain every 4 segments, 4.7e-7 with one every segment, for 4 % of the time).
// * per CTA one partial 96 x 96 block in the workspace; conv3x3_wgrad_tc_reduce_kernel
//   sums the CTAs in a fixed order (deterministic) into dw's (co, ci, ky, kx) layout.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "conv_tc.cuh"

namespace csmri {

constexpr int kWtcC = 32;
constexpr int kWtcPx = 64;                 // pixels per step (two 32-pixel swizzle atoms)
constexpr int kWtcRows = 16;               // dy rows per work item (18 input rows loaded)
constexpr int kWtcRing = 6;                // input-row ring (+ 2 mirrored slots)
constexpr int kWtcBox = 32 * 128;          // one TMA box: 32 channels x 32 pixels fp32 (4 KiB)
constexpr int kWtcXAtom = (kWtcRing + 2) * kWtcBox;          // one atom column of the ring (32 KiB)
constexpr int kWtcXPart = 2 * kWtcXAtom;                     // hi or lo (64 KiB)
constexpr int kWtcDySlots = 3;
constexpr int kWtcDyBytes = kWtcDySlots * 2 * kWtcBox;       // 24 KiB
constexpr int kWtcSmemBytes = 2 * kWtcXPart + kWtcDyBytes + 1024 + 1024;   // + barriers + alignment slack
// 12 warps = 3 per SM sub-partition: 168 registers per thread (a 13th warp would cap every thread at 128)
constexpr int kWtcThreads = 384;           // 4 drain, 3 staging, 1 TMA, 2 splitter, 2 MMA warps
constexpr int kWtcMmaWarp = 10;            // and 11
constexpr int kWtcPartial = 96 * 96;       // floats per CTA in the workspace

__device__ __forceinline__ void wtc_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wtc_tma_box(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// A operand from tensor memory (TS form)
__device__ __forceinline__ void wtc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0;"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p; }" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
#define WTC_ST32(addr, v)                                                                                        \
  asm volatile(                                                                                                  \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "    \
      "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(    \
          addr),                                                                                                 \
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]),          \
      "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]),  \
      "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]),             \
      "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31])                                      \
      : "memory")
#define WTC_LD32(v, addr)                                                                                        \
  asm volatile(                                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "  \
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),           \
        "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),     \
        "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),   \
        "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])    \
      : "r"(addr))

__device__ long long wtc_cta_cycles[256];   // tuning probe (debug bit 7): MMA warp lifetime per CTA

struct WtcItem {
  int n, x0, y0;
};
__device__ __forceinline__ WtcItem wtc_item(int item, int xsegs, int yblocks) {
  WtcItem t;
  t.n = item / (xsegs * yblocks);
  const int rem = item - t.n * xsegs * yblocks;
  const int yb = rem / xsegs;
  t.y0 = yb * kWtcRows;
  t.x0 = (rem - yb * xsegs) * kWtcPx;
  return t;
}

// one dy row segment -> the A operand of one step: lane (KX, co) holds dy[co][x0 + p + 1 - KX], p = 0 .. 63,
// as hi (columns 0-63 of the buffer) and lo (columns 64-127).  Half a segment at a time (32 pixels of
// the row + the one pixel the shift pulls in from the neighbouring half or the halo) to stay inside
// the register budget of a 448-thread CTA.
template <int KX, bool BIAS>
__device__ __forceinline__ float wtc_stage_dy(const unsigned char* dy_slot, int co, float halo, uint32_t tmem_a) {
  const unsigned char* row = dy_slot + co * 128;
  const int sw = co & 7;
  float rowsum = 0.0f;            // KX == 1 only: sum of this channel's 64 dy values (bias gradient)
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float c[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 q = *reinterpret_cast<const float4*>(row + half * kWtcBox + ((k ^ sw) << 4));
      c[4 * k] = q.x;
      c[4 * k + 1] = q.y;
      c[4 * k + 2] = q.z;
      c[4 * k + 3] = q.w;
    }
    if (KX == 1 && BIAS) {
      float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int p = 0; p < 32; ++p) s4[p & 3] += c[p];
      rowsum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    float edge = halo;            // KX = 0: pixel 32 half + 32; KX = 2: pixel 32 half - 1
    if (KX == 0 && half == 0) edge = *reinterpret_cast<const float*>(row + kWtcBox + ((0 ^ sw) << 4));
    if (KX == 2 && half == 1) edge = *reinterpret_cast<const float*>(row + ((7 ^ sw) << 4) + 12);
    float hi[32], lo[32];
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      const float v = KX == 1 ? c[p] : KX == 0 ? (p < 31 ? c[p + 1 > 31 ? 31 : p + 1] : edge)
                                                : (p > 0 ? c[p - 1 < 0 ? 0 : p - 1] : edge);
      tc_split(v, hi[p], lo[p]);
    }
    WTC_ST32(tmem_a + 32 * half, hi);
    WTC_ST32(tmem_a + kWtcPx + 32 * half, lo);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  return rowsum;
}

// x, dy: (N, 32, H, W) fp32 as 3-D tensor maps {W, H, N * 32}, box {32, 1, 32}, 128-byte swizzle.
// H % kWtcRows == 0, W % kWtcPx == 0.  partial: gridDim.x blocks of 96 x 96 floats,
// [(kx, co)][(ky, ci)].  dy_ptr: the same dy tensor, for the two halo pixels of a segment.
// bias_partial (may be null): gridDim.x x 32 floats, the per-CTA sums of dy over all pixels
// (the bias gradient of the layer, a by-product of the kx = 1 staging warp reading every dy value).
__global__ void __launch_bounds__(kWtcThreads, 1)
    conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                            const float* __restrict__ dy_ptr, float* __restrict__ partial,
                            float* __restrict__ bias_partial, int H, int W, int nitems, int debug) {
  extern __shared__ unsigned char wtc_smem_raw[];
  // 128-byte-swizzled boxes are anchored at 1024-byte boundaries
  unsigned char* smem = wtc_smem_raw + ((1024u - (tc_s32(wtc_smem_raw) & 1023u)) & 1023u);
  unsigned char* X_hi = smem;                              // [atom 2][slot 8][32 ci][128 B]
  unsigned char* X_lo = smem + kWtcXPart;
  unsigned char* DY_s = smem + 2 * kWtcXPart;              // [slot 3][atom 2][32 co][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kWtcXPart + kWtcDyBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  const uint32_t x_full = tc_s32(&bars[0]);       // [6] TMA -> splitters (hi landed)
  const uint32_t xlo_full = tc_s32(&bars[6]);     // [6] splitters -> MMA
  const uint32_t x_free = tc_s32(&bars[12]);      // [6] MMA -> TMA
  const uint32_t dy_full = tc_s32(&bars[18]);     // [3] TMA -> staging
  const uint32_t dy_free = tc_s32(&bars[21]);     // [3] staging -> TMA
  const uint32_t a_full = tc_s32(&bars[24]);      // [2] staging -> MMA
  const uint32_t step_done = tc_s32(&bars[26]);   // [2] MMA -> staging (A buffer free) and drain (accumulator ready)
  const uint32_t d_free = tc_s32(&bars[30]);      // [2] drain -> MMA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xsegs = W / kWtcPx, yblocks = H / kWtcRows;
  long long prof[4] = {0, 0, 0, 0};        // TC_PROF_WAIT (debug bit 7): cycles inside barrier waits
  const long long t_start = clock64();

  if (tid == 0) {
    for (int i = 0; i < kWtcRing; ++i) {
      tc_mbar_init(x_full + 8 * i, 1);
      tc_mbar_init(xlo_full + 8 * i, 2);
      tc_mbar_init(x_free + 8 * i, 1);
    }
    for (int i = 0; i < kWtcDySlots; ++i) {
      tc_mbar_init(dy_full + 8 * i, 1);
      tc_mbar_init(dy_free + 8 * i, 3);
    }
    for (int i = 0; i < 2; ++i) {
      tc_mbar_init(a_full + 8 * i, 3);
      tc_mbar_init(step_done + 8 * i, 1);
      tc_mbar_init(d_free + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_dy) : "memory");
  }
  if (warp == kWtcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_s32(tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // programmatic dependent launch: the setup above overlaps the previous kernel's tail
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  // tensor memory: accumulators at columns 0 and 128 (96 used each), A buffers at 256 and 384
  // (64 hi + 64 lo columns each)

  if (warp == 7) {
    // ===== TMA producer: input rows (once each, + mirror) and dy rows, in the order the steps need them =====
    if (lane == 0) {
      uint32_t qx = 0, qd = 0;          // input rows / dy rows issued so far
      auto load_x = [&](const WtcItem& t, int i) {
        const uint32_t slot = qx % kWtcRing, use = qx / kWtcRing;
        tc_mbar_wait(x_free + 8 * slot, (use & 1) ^ 1);
        const bool mirror = slot < 2;
        if (kTcProbe && (debug & 8)) {
          tc_mbar_arrive(x_full + 8 * slot);
          ++qx;
          return;
        }
        wtc_expect_tx(x_full + 8 * slot, (mirror ? 4 : 2) * kWtcBox);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          wtc_tma_box(tc_s32(X_hi + a * kWtcXAtom + slot * kWtcBox), &tm_x, x_full + 8 * slot, t.x0 + 32 * a,
                      t.y0 - 1 + i, t.n * kWtcC);
          if (mirror)
            wtc_tma_box(tc_s32(X_hi + a * kWtcXAtom + (slot + kWtcRing) * kWtcBox), &tm_x, x_full + 8 * slot,
                        t.x0 + 32 * a, t.y0 - 1 + i, t.n * kWtcC);
        }
        ++qx;
      };
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const WtcItem t = wtc_item(item, xsegs, yblocks);
        load_x(t, 0);
        load_x(t, 1);
        for (int j = 0; j < kWtcRows; ++j, ++qd) {
          load_x(t, j + 2);
          const uint32_t slot = qd % kWtcDySlots, use = qd / kWtcDySlots;
          tc_mbar_wait(dy_free + 8 * slot, (use & 1) ^ 1);
          if (kTcProbe && (debug & 8)) {
            tc_mbar_arrive(dy_full + 8 * slot);
            continue;
          }
          wtc_expect_tx(dy_full + 8 * slot, 2 * kWtcBox);
#pragma unroll
          for (int a = 0; a < 2; ++a)
            wtc_tma_box(tc_s32(DY_s + (slot * 2 + a) * kWtcBox), &tm_dy, dy_full + 8 * slot, t.x0 + 32 * a,
                        t.y0 + j, t.n * kWtcC);
        }
      }
    }
  } else if (warp >= 8 && warp < 10) {
    // ===== splitters: lo = x - trunc_tf32(x), rounded, next to every landed input row =====
    const int s = tid - 256;                                  // 0 .. 63
    uint32_t qx = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      for (int i = 0; i < kWtcRows + 2; ++i, ++qx) {
        const uint32_t slot = qx % kWtcRing, use = qx / kWtcRing;
        TC_PROF_WAIT(0, x_full + 8 * slot, use & 1);
#pragma unroll
        for (int k = 0; k < ((kTcProbe && (debug & 4)) ? 0 : 8); ++k) {
          const int chunk = s + 64 * k;                      // 512 16-byte chunks: 2 atoms x 256
          const int off = (chunk >> 8) * kWtcXAtom + slot * kWtcBox + (chunk & 255) * 16;
          const float4 v = *reinterpret_cast<const float4*>(X_hi + off);
          float4 l;
          l.x = __uint_as_float(__float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u)) + 0x1000u);
          l.y = __uint_as_float(__float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u)) + 0x1000u);
          l.z = __uint_as_float(__float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u)) + 0x1000u);
          l.w = __uint_as_float(__float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u)) + 0x1000u);
          *reinterpret_cast<float4*>(X_lo + off) = l;
          if (slot < 2) *reinterpret_cast<float4*>(X_lo + off + kWtcRing * kWtcBox) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(xlo_full + 8 * slot);
      }
    }
  } else if (warp >= 4 && warp < 7) {
    // ===== dy staging: warp kx writes TMEM lanes 32 kx .. 32 kx + 31 =====
    const int kx = warp - 4, co = lane;
    const uint32_t lane_base = (uint32_t)(kx * 32) << 16;
    const size_t plane = (size_t)H * W;
    uint32_t sc = 0;                                          // steps staged so far
    float bsum = 0.0f;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const WtcItem t = wtc_item(item, xsegs, yblocks);
      float isum = 0.0f;                                      // two-level sum: segment -> item -> CTA
      const int hx = kx == 0 ? t.x0 + kWtcPx : t.x0 - 1;
      const bool has_halo = kx != 1 && hx >= 0 && hx < W;
      const float* hp = dy_ptr + ((size_t)t.n * kWtcC + co) * plane + (size_t)t.y0 * W + (has_halo ? hx : 0);
      for (int j = 0; j < kWtcRows; ++j, ++sc) {
        // the one halo pixel this lane needs (kx = 0: right of the segment, kx = 2: left),
        // straight from global memory, requested before the barrier waits
        const float halo = has_halo ? __ldg(hp + (size_t)j * W) : 0.0f;
        const uint32_t buf = sc & 1, slot = sc % kWtcDySlots;
        TC_PROF_WAIT(0, step_done + 8 * buf, ((sc >> 1) & 1) ^ 1);
        TC_PROF_WAIT(1, dy_full + 8 * slot, (sc / kWtcDySlots) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long ts0 = (kTcProbe && (debug & 128)) ? clock64() : 0;
        const unsigned char* src = DY_s + slot * 2 * kWtcBox;
        const uint32_t ta = tmem + lane_base + 256 + buf * 128;
        if (kTcProbe && (debug & 1)) {
        } else if (kx == 0) wtc_stage_dy<0, false>(src, co, halo, ta);
        else if (kx == 1 && bias_partial != nullptr) isum += wtc_stage_dy<1, true>(src, co, halo, ta);
        else if (kx == 1) wtc_stage_dy<1, false>(src, co, halo, ta);
        else wtc_stage_dy<2, false>(src, co, halo, ta);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tc_mbar_arrive(a_full + 8 * buf);
          tc_mbar_arrive(dy_free + 8 * slot);
        }
        if (kTcProbe && (debug & 128)) prof[2] += clock64() - ts0;
      }
      bsum += isum;
    }
    if (kx == 1 && bias_partial != nullptr) bias_partial[blockIdx.x * kWtcC + co] = bsum;
  } else if (warp >= kWtcMmaWarp) {
    // ===== MMA issuers: warp 10 issues the even steps, warp 11 the odd ones =====
    // The instruction queue of the tensor core is shallow and one thread needs ~500 cycles per
    // step for its barrier waits, fence and commits - time in which the pipe would run dry.
    // Even and odd steps use disjoint accumulator and A buffers, so two threads can issue them
    // independently: the bookkeeping of one overlaps the instructions of the other.
    // Each warp runs its loop with uniform control flow and one elected lane issues (descriptors
    // and barrier addresses then live in uniform registers).
    const uint32_t par = warp - kWtcMmaWarp;
    const bool leader = tc_elect();
    if (blockIdx.x < nitems) {
      // D fp32, A / B tf32, K-major, M = 128, N = 96
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // 128-byte swizzle, 8-row groups 1024 bytes apart
      const uint64_t db_hi0 = tc_desc(tc_s32(X_hi), 16, 1024) | ((uint64_t)2 << 61);
      const uint64_t db_lo0 = tc_desc(tc_s32(X_lo), 16, 1024) | ((uint64_t)2 << 61);
      const uint32_t total = (uint32_t)((nitems - blockIdx.x + gridDim.x - 1) / gridDim.x) * kWtcRows;
      long long t_commit = 0, t_fence = 0;
      auto ready = [&](uint32_t st, int part) {
        if (part == 0) {                                      // accumulator buffer drained
          TC_PROF_WAIT(0, d_free + 8 * (st & 1), ((st >> 1) & 1) ^ 1);
        } else if (part == 1) {                               // input rows j, j + 1, j + 2 of the item split
          const uint32_t it = st / kWtcRows, j = st - it * kWtcRows;
          for (uint32_t k = (j < 2 ? 0 : 1); k < 3; ++k) {    // this thread last waited two steps ago
            const uint32_t q = it * (kWtcRows + 2) + j + k;
            TC_PROF_WAIT(1, xlo_full + 8 * (q % kWtcRing), (q / kWtcRing) & 1);
          }
        } else {                                              // dy row staged in tensor memory
          TC_PROF_WAIT(2, a_full + 8 * (st & 1), (st >> 1) & 1);
        }
      };
      for (uint32_t st = par; st < total; st += 2) {
        ready(st, 0);
        ready(st, 1);
        ready(st, 2);
        const long long tf0 = (kTcProbe && (debug & 128)) ? clock64() : 0;
        if (!(kTcProbe && (debug & 32))) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long ti0 = (kTcProbe && (debug & 128)) ? clock64() : 0;
        t_fence += ti0 - tf0;
        const uint32_t it = st / kWtcRows, j = st - it * kWtcRows, qb = it * (kWtcRows + 2);
        const uint32_t wslot = (qb + j) % kWtcRing;           // window = slots wslot .. wslot + 2
        const uint32_t d_tmem = tmem + (st & 1) * 128;
        const uint32_t a_tmem = tmem + 256 + (st & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t boff = (uint32_t)(((ks >> 2) * kWtcXAtom + wslot * kWtcBox + (ks & 3) * 32) >> 4);
          if (leader && !(kTcProbe && (debug & 16))) {
            wtc_mma_ts(d_tmem, a_tmem + kWtcPx + 8 * ks, db_hi0 + boff, idesc, ks == 0 ? 0u : 1u);   // lo * hi
            wtc_mma_ts(d_tmem, a_tmem + 8 * ks, db_lo0 + boff, idesc, 1u);                            // hi * lo
            wtc_mma_ts(d_tmem, a_tmem + 8 * ks, db_hi0 + boff, idesc, 1u);                            // hi * hi
          }
        }
        const long long ti1 = (kTcProbe && (debug & 128)) ? clock64() : 0;
        if (leader) {
          tc_commit(step_done + 8 * (st & 1));                 // A buffer free, accumulator ready to drain
          tc_commit(x_free + 8 * wslot);                       // input row j of the item is not needed again
          if (j == kWtcRows - 1) {
            tc_commit(x_free + 8 * ((qb + j + 1) % kWtcRing));
            tc_commit(x_free + 8 * ((qb + j + 2) % kWtcRing));
          }
        }
        __syncwarp();
        if (kTcProbe && (debug & 128)) {
          prof[3] += ti1 - ti0;                               // issuing the 24 instructions
          t_commit += clock64() - ti1;                        // issuing the commits
        }
      }
      if ((kTcProbe && (debug & 128)) && lane == 0 && blockIdx.x == 0 && par == 0) { tc_prof[7] = t_commit; tc_prof[2] = t_fence; }
      if ((kTcProbe && (debug & 128)) && lane == 0 && par == 0 && blockIdx.x < 256) wtc_cta_cycles[blockIdx.x] = clock64() - t_start;
    }
  } else if (warp < 4) {
    // ===== drain: warp w owns TMEM lanes 32 w .. 32 w + 31; lanes 96-127 are never written =====
    float acc[96];
#pragma unroll
    for (int i = 0; i < 96; ++i) acc[i] = 0.0f;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t period = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      for (int pr = 0; pr < kWtcRows; ++pr, ++period) {     // one drain per step
        const uint32_t dbuf = period & 1;
        tc_mbar_wait(step_done + 8 * dbuf, (period >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (kTcProbe && (debug & 2)) {
          __syncwarp();
          if (lane == 0) tc_mbar_arrive(d_free + 8 * dbuf);
          continue;
        }
        float v[32];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          WTC_LD32(v, tmem + lane_base + dbuf * 128 + 32 * g);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (g == 2) {                                       // the buffer may be overwritten
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(d_free + 8 * dbuf);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[32 * g + i] += v[i];
        }
      }
    }
    if (warp < 3) {
      float* dst = partial + ((size_t)blockIdx.x * 96 + tid) * 96;
#pragma unroll
      for (int i = 0; i < 96; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }
  }
  if ((kTcProbe && (debug & 128)) && blockIdx.x == 0) {
    if (tid == 192) { tc_prof[0] = prof[0]; tc_prof[1] = prof[1]; tc_prof[6] = prof[2]; }   // staging warp kx = 2
    if (tid == kWtcMmaWarp * 32) { tc_prof[3] = prof[0]; tc_prof[4] = prof[1]; tc_prof[5] = prof[2]; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kWtcMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// dw[co][ci][ky][kx] = sum over CTAs, in a fixed order, of partial[cta][(kx, co)][(ky, ci)];
// db[co] (optional) = sum over CTAs of bias_partial[cta][co].
// Block = 32 elements x 4 groups of CTAs: group g adds the CTAs g, g + 4, ... (coalesced
// 128-byte loads, four independent chains per element instead of one 148-long one), the four
// group sums meet in shared memory.
constexpr int kWtcReduceGroups = 4;
constexpr int kWtcReduceBlocks = (kWtcPartial + kWtcC + 31) / 32;

__global__ void __launch_bounds__(32 * kWtcReduceGroups)
    conv3x3_wgrad_tc_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                   const float* __restrict__ bias_partial, float* __restrict__ db, int nctas) {
  __shared__ float part[kWtcReduceGroups][32];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;                        // (kx, co, ky, ci), then 32 bias entries
  asm volatile("griddepcontrol.wait;" ::: "memory");           // the partial blocks of the kernel before
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const bool is_bias = e >= kWtcPartial;
  const int co_b = e - kWtcPartial;
  float s = 0.0f;
  if (!is_bias) {
    for (int c = g; c < nctas; c += kWtcReduceGroups) s += partial[(size_t)c * kWtcPartial + e];
  } else if (co_b < kWtcC && db != nullptr) {
    for (int c = g; c < nctas; c += kWtcReduceGroups) s += bias_partial[c * kWtcC + co_b];
  }
  part[g][lane] = s;
  __syncthreads();
  if (g != 0) return;
  s = (part[0][lane] + part[1][lane]) + (part[2][lane] + part[3][lane]);
  if (is_bias) {
    if (co_b < kWtcC && db != nullptr) db[co_b] = s;
    return;
  }
  const int row = e / 96, col = e - row * 96;
  const int kx = row >> 5, co = row & 31, ky = col >> 5, ci = col & 31;
  dw[((co * kWtcC + ci) * 3 + ky) * 3 + kx] = s;
}

}  // namespace csmri
