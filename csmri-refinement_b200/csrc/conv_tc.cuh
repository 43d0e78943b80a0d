// RecNet's 32 -> 32 channel 3x3 convolutions (models/recnet.py:37-44: the inner
// layers of every ConvBlock) on the 5th-generation tensor cores - the one
// GEMM-shaped piece of the training step - with fp32 accuracy.
//
// The fp32 step spends 21 of its 40 ms in cuDNN's fp32 forward / data-gradient
// kernels for these layers (profiles/r1_recnet_step_kernels.txt; ~55 TFLOP/s
// effective).  The parity gate is stated for fp32 arithmetic (1e-5), so plain
// TF32 is out; the classic way to get fp32-accurate products out of TF32 tensor
// cores is the 3xTF32 split: a = a_hi + a_lo (a_hi = tf32(a), a_lo = tf32(a - a_hi)),
//     a * b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo          (error ~2^-21 per product)
// with fp32 accumulation.  The legacy warp-level path (mma.sync) runs TF32 at
// 278 TFLOP/s on a B200 (profiles/r2_mma_sync_tf32_rate.txt) - a third of that
// does not beat cuDNN - so this kernel uses tcgen05.mma (kind::tf32, ~1.1 PFLOP/s
// dense) with the accumulator in tensor memory.
//
// Implicit GEMM, one output row segment per tile:
//     D[m = 128 pixels, n = 32 co] = sum_{tap} A_tap[128 x 32 ci] * B_tap[32 ci x 32 co]
// * A lives in shared memory in the NON-swizzled K-major canonical layout
//   [ci/4][pixel][ci%4] (16-byte units, pixels contiguous).  In that layout
//   "start one pixel further right" is "descriptor start address + 16 bytes", so the
//   three horizontal taps read the SAME staged row at offsets 0 / 16 / 32 bytes and
//   the three vertical taps read three different staged rows: every input row
//   segment (130 pixels x 32 channels) is staged ONCE per output-row block and
//   used by 9 taps x 3 output rows.  (tools/umma_probe.cu checks the shifted-start
//   property on the hardware; swizzled layouts do not have it.)
// * staging = 4 producer warps: coalesced fp32 loads straight from the NCHW
//   tensor, hi / lo split in registers, two 16-byte shared stores per (pixel,
//   4 channels), fence.proxy.async, mbarrier arrive.  A ring of 4 row slots.
// * one thread issues the 72 MMAs of a tile (9 taps x 4 k-steps x 2), commits to
//   mbarriers that (a) hand the accumulator to the epilogue warps and (b) return
//   the oldest ring slot.  With both operands in shared memory a thin (N = 32)
//   MMA is bound by operand fetch, not by the tensor pipe (measured: 108 MMAs of
//   128 x 32 x 8 took ~68 cycles each, the pipe needs 16; A is 4 of the 5 KiB an MMA
//   reads).  So the B operand of a tap is [B_hi | B_lo] side by side (N = 64):
//   ONE MMA gives A_hi*B_hi in accumulator columns 0-31 and A_hi*B_lo in columns
//   32-63, a second (N = 32) adds A_lo*B_hi to columns 0-31 - A_hi is fetched once
//   instead of twice - and the epilogue adds the two column groups.
// * 4 epilogue warps: tcgen05.ld (lane = pixel, columns = output channels),
//   bias + LeakyReLU, coalesced 128-byte row stores per channel; two accumulator
//   buffers in TMEM so the next tile's MMAs overlap the epilogue.
// * B (the 9 x 32 x 32 weights, hi and lo) is split once per CTA into shared
//   memory; `transpose_flip` builds the data-gradient operator (ci <-> co swapped,
//   taps mirrored) from the same weight tensor, so backward-data is this kernel too.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csmri {

constexpr int kTcC = 32;                 // channels in and out
constexpr int kTcM = 128;                // pixels per tile (one row segment)
constexpr int kTcPA = 136;               // staged pixels per row slot (130 used), 16-byte units
constexpr int kTcPlane = kTcPA * 16;     // bytes between the 4-channel planes of a row slot
constexpr int kTcSlotPart = 8 * kTcPlane;            // one row, hi or lo
constexpr int kTcSlots = 4;
constexpr int kTcBTap = 8 * 2 * kTcC * 16;           // one tap of B: [ci/4][hi co | lo co][ci%4] (8 KiB)
constexpr int kTcBBytes = 9 * kTcBTap;               // 72 KiB
constexpr int kTcAccCols = 2 * kTcC;                 // TMEM columns per accumulator buffer
constexpr int kTcABytes = kTcSlots * 2 * kTcSlotPart;
constexpr int kTcSmemBytes = kTcBBytes + kTcABytes + 256;
constexpr int kTcThreads = 288;          // 4 epilogue warps, 4 producer warps, 1 MMA warp
constexpr int kTcRowBlock = 16;          // output rows per work item (18 staged rows)

__device__ __forceinline__ uint32_t tc_s32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;"
        " selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {   // arrives when all MMAs issued so far are done
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: rows 16 B apart, 8-row groups `sbo`
// bytes apart, the two 16-byte K chunks of one MMA `lbo` bytes apart (version 1 = sm_100)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0;"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p; }" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ float tc_tf32(float v) {     // round to nearest TF32, as a float
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

struct TcItem {
  int n, x0, y0;
};
__device__ __forceinline__ TcItem tc_item(int item, int xsegs, int yblocks) {
  TcItem t;
  t.n = item / (xsegs * yblocks);
  const int rem = item - t.n * xsegs * yblocks;
  const int yb = rem / xsegs;
  t.y0 = yb * kTcRowBlock;
  t.x0 = (rem - yb * xsegs) * kTcM;
  return t;
}

// y[n][co] = act(bias[co] + sum_{ci,ky,kx} wq[co][ci][ky][kx] * x[n][ci][. + ky - 1][. + kx - 1])
// with wq = w (forward) or, transpose_flip != 0, wq[co][ci][ky][kx] = w[ci][co][2-ky][2-kx]
// (the data gradient of the same layer).  H % kTcRowBlock == 0, W % 128 == 0.
// `debug` (tuning probes only): bit 0 skip the MMAs, bit 1 skip the global loads,
// bit 2 skip the global stores.
__global__ void __launch_bounds__(kTcThreads, 1)
    conv3x3_tc_kernel(const float* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, float* __restrict__ y, int H, int W, int nitems,
                      float slope, int transpose_flip, int debug) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char* B_s = tc_smem;                       // [tap][ci/4][hi co 0-31 | lo co 0-31][ci%4]
  unsigned char* A_s = tc_smem + kTcBBytes;           // [slot][part][ci/4][pixel][ci%4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + kTcBBytes + kTcABytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t row_full = tc_s32(&bars[0]);         // [4] producers -> MMA
  const uint32_t row_free = tc_s32(&bars[4]);         // [4] MMA -> producers
  const uint32_t acc_full = tc_s32(&bars[8]);         // [2] MMA -> epilogue
  const uint32_t acc_free = tc_s32(&bars[10]);        // [2] epilogue -> MMA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int xsegs = W / kTcM, yblocks = H / kTcRowBlock;
  const size_t plane = (size_t)H * W;

  // ---- one-time setup -----------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kTcSlots; ++i) {
      tc_mbar_init(row_full + 8 * i, 128);
      tc_mbar_init(row_free + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc_mbar_init(acc_full + 8 * i, 1);
      tc_mbar_init(acc_free + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(
                     tc_s32(tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights: split into hi / lo, laid out as the B operand of each tap
  for (int e = tid; e < 9 * kTcC * kTcC; e += kTcThreads) {
    const int tap = e / (kTcC * kTcC), r = e - tap * (kTcC * kTcC);
    const int co = r / kTcC, ci = r - co * kTcC;       // operator element (co, ci, tap)
    const int ky = tap / 3, kx = tap - 3 * ky;
    const float v = transpose_flip ? __ldg(w + ((ci * kTcC + co) * 3 + (2 - ky)) * 3 + (2 - kx))
                                   : __ldg(w + ((co * kTcC + ci) * 3 + ky) * 3 + kx);
    const float hi = tc_tf32(v), lo = tc_tf32(v - hi);
    const int off = tap * kTcBTap + (ci >> 2) * (2 * kTcC * 16) + co * 16 + (ci & 3) * 4;
    *reinterpret_cast<float*>(B_s + off) = hi;
    *reinterpret_cast<float*>(B_s + off + kTcC * 16) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ===== producers: one input row segment (130 pixels x 32 channels) per ring slot =====
    const int j = tid - 128;                       // 0 .. 127: pixel within the segment
    uint32_t q = 0;                                // ring rows produced by this CTA so far
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const TcItem t = tc_item(item, xsegs, yblocks);
      const float* xn = x + (size_t)t.n * kTcC * plane;
      for (int r = 0; r < kTcRowBlock + 2; ++r, ++q) {
        const int slot = q & 3;
        tc_mbar_wait(row_free + 8 * slot, ((q >> 2) & 1) ^ 1);
        const int gy = t.y0 - 1 + r;
        unsigned char* hi_s = A_s + (size_t)slot * 2 * kTcSlotPart;
        unsigned char* lo_s = hi_s + kTcSlotPart;
        const bool row_ok = gy >= 0 && gy < H;
        // pixel p of the slot is image column x0 - 1 + p; thread j stages p = j and,
        // for j < 2, the two right-most pixels p = 128 + j
        for (int pass = 0; pass < (j < 2 ? 2 : 1); ++pass) {
          const int p = j + 128 * pass;
          const int gx = t.x0 - 1 + p;
          const bool ok = row_ok && gx >= 0 && gx < W;
          const float* src = xn + (size_t)(ok ? gy : 0) * W + (ok ? gx : 0);
          float v[kTcC];
#pragma unroll
          for (int c = 0; c < kTcC; ++c) v[c] = (ok && !(debug & 2)) ? __ldg(src + (size_t)c * plane) : 0.0f;
#pragma unroll
          for (int kc = 0; kc < 8; ++kc) {
            float4 h, l;
            h.x = tc_tf32(v[4 * kc]);     l.x = tc_tf32(v[4 * kc] - h.x);
            h.y = tc_tf32(v[4 * kc + 1]); l.y = tc_tf32(v[4 * kc + 1] - h.y);
            h.z = tc_tf32(v[4 * kc + 2]); l.z = tc_tf32(v[4 * kc + 2] - h.z);
            h.w = tc_tf32(v[4 * kc + 3]); l.w = tc_tf32(v[4 * kc + 3] - h.w);
            *reinterpret_cast<float4*>(hi_s + kc * kTcPlane + p * 16) = h;
            *reinterpret_cast<float4*>(lo_s + kc * kTcPlane + p * 16) = l;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_mbar_arrive(row_full + 8 * slot);
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer: one elected thread =====
    if (lane == 0) {
      // D fp32, A / B tf32, both K-major, M = 128; N = 64 ([hi | lo] weights) and N = 32 (hi only)
      const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcC >> 3) << 17) |
                               ((uint32_t)(kTcM >> 4) << 24);
      const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(2 * kTcC >> 3) << 17) |
                               ((uint32_t)(kTcM >> 4) << 24);
      const uint64_t da_base = tc_desc(tc_s32(A_s), kTcPlane, 128);
      const uint64_t db_base = tc_desc(tc_s32(B_s), 2 * kTcC * 16, 128);
      uint32_t q = 0, tile = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        for (int r = 0; r < kTcRowBlock; ++r, ++q, ++tile) {
          // input rows of output row r: ring rows q, q + 1, q + 2
          if (r == 0) {
            tc_mbar_wait(row_full + 8 * (q & 3), (q >> 2) & 1);
            tc_mbar_wait(row_full + 8 * ((q + 1) & 3), ((q + 1) >> 2) & 1);
          }
          tc_mbar_wait(row_full + 8 * ((q + 2) & 3), ((q + 2) >> 2) & 1);
          const uint32_t buf = tile & 1;
          tc_mbar_wait(acc_free + 8 * buf, ((tile >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d_tmem = tmem + buf * kTcAccCols;
          uint32_t acc = 0;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint64_t da_row = da_base + (uint32_t)((((q + ky) & 3) * (2 * kTcSlotPart)) >> 4);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                // only the 14-bit start-address field (16-byte units) changes between MMAs
                const uint64_t dah = da_row + (uint32_t)((2 * ks * kTcPlane + kx * 16) >> 4);
                const uint64_t dal = dah + (uint32_t)(kTcSlotPart >> 4);
                const uint64_t db = db_base + (uint32_t)(((ky * 3 + kx) * kTcBTap + 2 * ks * (2 * kTcC * 16)) >> 4);
                if (!(debug & 1)) {
                  tc_mma(d_tmem, dah, db, idesc64, acc);    // cols 0-31 += hi*hi, cols 32-63 += hi*lo
                  tc_mma(d_tmem, dal, db, idesc32, 1u);     // cols 0-31 += lo*hi
                }
                acc = 1u;
              }
            }
          }
          tc_commit(acc_full + 8 * buf);
          tc_commit(row_free + 8 * (q & 3));              // ring row q is not needed again
          if (r == kTcRowBlock - 1) {                     // last output row of the block
            tc_commit(row_free + 8 * ((q + 1) & 3));
            tc_commit(row_free + 8 * ((q + 2) & 3));
          }
        }
        q += 2;   // the block consumed kTcRowBlock + 2 ring rows
      }
    }
  } else {
    // ===== epilogue: warp e owns TMEM lanes 32e .. 32e + 31 = pixels of the tile =====
    float bv[kTcC];
#pragma unroll
    for (int c = 0; c < kTcC; ++c) bv[c] = bias != nullptr ? __ldg(bias + c) : 0.0f;
    uint32_t tile = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const TcItem t = tc_item(item, xsegs, yblocks);
      float* yn = y + (size_t)t.n * kTcC * plane + t.x0 + tid;
      for (int r = 0; r < kTcRowBlock; ++r, ++tile) {
        const uint32_t buf = tile & 1;
        tc_mbar_wait(acc_full + 8 * buf, (tile >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[kTcC], u[kTcC];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + buf * kTcAccCols;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11,"
            " %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28,"
            " %29, %30, %31}, [%32];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]),
              "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]),
              "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]),
              "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]),
              "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
            : "r"(taddr + kTcC));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11,"
            " %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28,"
            " %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
              "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
              "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
              "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
              "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        tc_mbar_arrive(acc_free + 8 * buf);               // accumulator buffer may be overwritten
        float* dst = yn + (size_t)(t.y0 + r) * W;
#pragma unroll
        for (int c = 0; c < kTcC; ++c) {
          float o = (__uint_as_float(u[c]) + __uint_as_float(v[c])) + bv[c];   // small term first
          if (slope > 0.0f) o = o > 0.0f ? o : o * slope;
          if (!(debug & 4))
            asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(dst + (size_t)c * plane), "f"(o)
                         : "memory");
        }
      }
    }
  }
  // ---- teardown ------------------------------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 8)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

}  // namespace csmri
