// RecNet's 32 -> 32 channel 3x3 convolutions (models/recnet.py:37-44: the inner
// layers of every ConvBlock) on the 5th-generation tensor cores - the one
// GEMM-shaped piece of the training step - with fp32 accuracy.
//
// Round 1's fp32 step spent 21 of its 40 ms in cuDNN's fp32 forward / data-gradient
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
// Implicit GEMM over one staged input row segment at a time:
//     acc[out row r+1-ky][m = 128 pixels, n = 32 co] += A_r,kx[128 x 32 ci] * B_kx,ky[32 ci x 32 co]
// * A lives in shared memory in the NON-swizzled K-major canonical layout
//   [ci/4][pixel][ci%4] (16-byte units, pixels contiguous).  In that layout
//   "start one pixel further right" is "descriptor start address + 16 bytes", so the
//   three horizontal taps read the SAME staged row at offsets 0 / 16 / 32 bytes
//   (tools/umma_probe.cu checks the shifted-start property on the hardware;
//   swizzled layouts do not have it).
// * a tcgen05.mma of 128 x N x 8 costs max(44.5, N / 2) cycles whatever N is
//   (tools/umma_probe2.cu), so an N = 32 instruction runs the tensor pipe at a third
//   of its rate.  The three VERTICAL taps are therefore stacked along N: the B
//   operand of a horizontal tap is [ky = 0 co | ky = 1 co | ky = 2 co] (N = 96), and
//   one instruction computes input row r's contribution to output rows r+1, r, r-1
//   at once, into a 96-column accumulator block of its own.  An input row is staged
//   once and consumed by 36 instructions (3 horizontal taps x 4 k-steps x 3 split
//   terms).  Five blocks rotate through tensor memory; the epilogue adds the three
//   blocks that hold an output row's contributions in fp32 registers.
// * work item = 8 output rows x 128 pixels (10 staged rows).  At the top / bottom
//   of an item the N window shrinks to the taps that land inside it.
// * staging = two groups of 4 producer warps (even / odd ring rows, the next row
//   already in registers) + a halo warp: coalesced fp32 loads straight from the NCHW
//   tensor, hi / lo split in registers, two 16-byte shared stores per (pixel,
//   4 channels), fence.proxy.async, one mbarrier arrival per warp.  A ring of 4 row
//   slots, each released by a tcgen05.commit as soon as its 36 instructions have retired.
// * two MMA-issuing warps (even / odd staged rows; a row owns its accumulator block,
//   ring slot and barriers), each running its loop with warp-uniform control flow and
//   an elected lane around the instructions.
// * the two small split terms (lo*hi, hi*lo) of a row are issued before its hi*hi
//   term: the accumulator is rounded toward zero after every instruction, and that
//   rounding is relative to what the accumulator holds at the time.
// * 4 epilogue warps: tcgen05.ld (lane = pixel, columns = output channels, 16 at a
//   time), bias from shared memory + LeakyReLU, coalesced 128-byte row stores per
//   channel, one 32-bit sign word per pixel (bit c = output channel c > 0).  The
//   MASKED instantiation is the data gradient that also applies the derivative of the
//   LeakyReLU in front of the layer, selected by such a sign word.
// * B (the 9 x 32 x 32 weights, hi and lo) is split once per CTA into shared
//   memory; `transpose_flip` builds the data-gradient operator (ci <-> co swapped,
//   taps mirrored) from the same weight tensor, so backward-data is this kernel too.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// The tuning probes (tools/conv_tc_probe.cu, tools/wgrad_tc_probe.cu) compile the kernels with
// CSMRI_TC_PROBE=1: `debug` bits then skip roles / record clock64 accounting.  In the library the
// checks fold away (the kernels are instruction-cache sensitive: 12-14 warps run different code).
#ifndef CSMRI_TC_PROBE
#define CSMRI_TC_PROBE 0
#endif

namespace csmri {

constexpr bool kTcProbe = CSMRI_TC_PROBE != 0;
constexpr int kTcC = 32;                 // channels in and out
constexpr int kTcM = 128;                // pixels per tile (one row segment)
constexpr int kTcPA = 136;               // staged pixels per row slot (130 used), 16-byte units
constexpr int kTcPlane = kTcPA * 16;     // bytes between the 4-channel planes of a row slot
constexpr int kTcSlotPart = 8 * kTcPlane;            // one row, hi or lo
constexpr int kTcSlots = 4;
constexpr int kTcBPlane = 3 * kTcC * 16;            // B: bytes between 4-channel (ci) planes: 96 rows (ky, co)
constexpr int kTcBPart = 8 * kTcBPlane;             // one horizontal tap, hi or lo: [ci/4][ky co][ci%4] (12 KiB)
constexpr int kTcBBytes = 3 * 2 * kTcBPart;         // 72 KiB
constexpr int kTcRowBlock = 8;                      // output rows per work item (10 staged rows)
constexpr int kTcBlocks = 5;                        // accumulator blocks of 96 columns in tensor memory (one per staged row in flight)
static_assert((kTcRowBlock + 2) % kTcBlocks == 0, "the block of a staged row must not depend on the item");
constexpr int kTcABytes = kTcSlots * 2 * kTcSlotPart;
constexpr int kTcSmemBytes = kTcBBytes + kTcABytes + 512;
constexpr int kTcThreads = 480;          // 4 epilogue warps, 2 x 4 producer warps, 1 halo warp, 2 MMA warps
constexpr int kTcMmaWarp = 13;           // and 14

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
// one lane of a converged warp; everything around the guarded instructions stays warp-uniform,
// which lets the compiler keep descriptors and barrier addresses in uniform registers
__device__ __forceinline__ bool tc_elect() {
  uint32_t p;
  asm volatile("{ .reg .pred q; elect.sync _|q, 0xffffffff; selp.u32 %0, 1, 0, q; }" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ float tc_tf32(float v) {     // round to nearest TF32, as a float
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
// The same split for the streamed operand, 4 instructions per value instead of 9:
// hi = v rounded to nearest TF32 (integer add of half an ulp, mask; cvt.rna's Inf / NaN
// special case is dropped - a non-finite input gives a non-finite output either way),
// lo = v - hi exactly, rounded by the half-ulp add alone: the tensor core ignores the
// 13 low mantissa bits of its fp32 containers (tools/umma_probe2.cu), which is the mask.
__device__ __forceinline__ void tc_split(float v, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
  lo = __uint_as_float(__float_as_uint(v - hi) + 0x1000u);
}

// tuning probe (debug bit 7): cycles CTA 0 spent inside each kind of barrier wait
__device__ long long tc_prof[8];
#define TC_PROF_WAIT(slot_, bar_, par_)                                   \
  do {                                                                    \
    if (kTcProbe && (debug & 128)) {                                                    \
      const long long t0_ = clock64();                                    \
      tc_mbar_wait(bar_, par_);                                           \
      prof[slot_] += clock64() - t0_;                                     \
    } else {                                                              \
      tc_mbar_wait(bar_, par_);                                           \
    }                                                                     \
  } while (0)

#define TC_LD16(v, addr)                                                                                      \
  asm volatile(                                                                                               \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, " \
      "%15}, [%16];"                                                                                          \
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),        \
        "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])   \
      : "r"(addr))
#define TC_LD32(v, addr)                                                                                         \
  asm volatile(                                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "  \
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),           \
        "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),     \
        "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),   \
        "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])    \
      : "r"(addr))

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
// `in_signs` (optional output, MASKED = false): the same kind of sign word for the INPUT x,
// a by-product of the producer warps holding all 32 channels of a pixel - it lets the data
// gradient of this layer carry the derivative of the activation that produced x.
// `signs` (N,H,W) uint32, bit c of a pixel = (output channel c > 0):
// MASKED = false: optional OUTPUT of the forward pass (one 4-byte store per pixel);
// MASKED = true (the data gradient feeding a LeakyReLU layer): no bias, no activation, and
// the result is multiplied by the derivative of that LeakyReLU, (bit c ? 1 : slope), read from
// the `signs` its forward pass wrote - the backward pass of the activation
// (models/recnet.py:45-47: nn.LeakyReLU after every inner convolution) costs a 4-byte load
// per pixel here instead of a read-read-write pass over 32 channels of its own.
template <bool MASKED>
__global__ void __launch_bounds__(kTcThreads, 1)
    conv3x3_tc_kernel(const float* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, float* __restrict__ y, uint32_t* __restrict__ signs,
                      uint32_t* __restrict__ in_signs, int H, int W, int nitems,
                      float slope, int transpose_flip, int debug) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  unsigned char* B_s = tc_smem;                       // [kx][hi | lo][ci/4][ky co 0-95][ci%4]
  unsigned char* A_s = tc_smem + kTcBBytes;           // [slot][hi | lo][ci/4][pixel][ci%4]
  uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + kTcBBytes + kTcABytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  float* bias_s = reinterpret_cast<float*>(bars + 34);      // [32] the layer's bias (zeros without one)
  const uint32_t row_full = tc_s32(&bars[0]);         // [4] producers -> MMA
  const uint32_t row_free = tc_s32(&bars[4]);         // [4] MMA -> producers
  const uint32_t acc_free = tc_s32(&bars[8]);         // [5] epilogue -> MMA (one per accumulator block)
  const uint32_t acc_full = tc_s32(&bars[16]);        // [5] MMA -> epilogue
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long prof[4] = {0, 0, 0, 0};
  const long long t_start = clock64();
  const int xsegs = W / kTcM, yblocks = H / kTcRowBlock;
  const size_t plane = (size_t)H * W;

  // ---- one-time setup -----------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kTcSlots; ++i) {
      tc_mbar_init(row_full + 8 * i, 5);          // one arrival per producer warp of the row's group + the halo warp
      tc_mbar_init(row_free + 8 * i, 1);
    }
    for (int i = 0; i < kTcBlocks; ++i) {
      tc_mbar_init(acc_free + 8 * i, 4);          // one arrival per epilogue warp
      tc_mbar_init(acc_full + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                     tc_s32(tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Programmatic dependent launch: everything above overlaps the tail of the previous kernel
  // in the stream; nothing it wrote is read before this point.  debug bit 8 (tuning key 10 = 2)
  // moves the wait behind the weight staging - only valid when the previous kernel does not
  // write `w` or `bias`.
  const bool late_wait = (debug & 256) != 0;
  if (!late_wait) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  if (tid < kTcC) bias_s[tid] = (!MASKED && bias != nullptr) ? __ldg(bias + tid) : 0.0f;
  // weights: split into hi / lo, laid out as the B operand of each horizontal tap.  All of a
  // thread's loads are issued before the first is used (a rolled loop paid one global-memory
  // round trip per element: ~12 us of the ~20 us fixed cost of a launch).
  {
    constexpr int kPer = (9 * kTcC * kTcC + kTcThreads - 1) / kTcThreads;
    float wv[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int e = tid + i * kTcThreads;
      const int tap = e / (kTcC * kTcC), r = e - tap * (kTcC * kTcC);
      const int co = r / kTcC, ci = r - co * kTcC;     // operator element (co, ci, tap)
      const int ky = tap / 3, kx = tap - 3 * ky;
      wv[i] = e >= 9 * kTcC * kTcC ? 0.0f
              : transpose_flip     ? __ldg(w + ((ci * kTcC + co) * 3 + (2 - ky)) * 3 + (2 - kx))
                                   : __ldg(w + ((co * kTcC + ci) * 3 + ky) * 3 + kx);
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int e = tid + i * kTcThreads;
      if (e < 9 * kTcC * kTcC) {
        const int tap = e / (kTcC * kTcC), r = e - tap * (kTcC * kTcC);
        const int co = r / kTcC, ci = r - co * kTcC;
        const int ky = tap / 3, kx = tap - 3 * ky;
        const float hi = tc_tf32(wv[i]), lo = tc_tf32(wv[i] - hi);
        const int off = kx * (2 * kTcBPart) + (ci >> 2) * kTcBPlane + (ky * kTcC + co) * 16 + (ci & 3) * 4;
        *reinterpret_cast<float*>(B_s + off) = hi;
        *reinterpret_cast<float*>(B_s + off + kTcBPart) = lo;
      }
    }
  }
  if (late_wait) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp >= 4 && warp < 12) {
    // ===== producers: two groups of 4 warps; group g stages the ring rows q = g (mod 2),
    // pixels 1 .. 128 of the slot (image columns x0 .. x0 + 127), one pixel per thread.
    // The next row of the group is fetched into registers before the current one is
    // converted, so the global-load latency hides behind the split / store work and the
    // two groups overlap each other's barrier waits.
    const int g = (warp - 4) >> 2;
    const int j = (tid - 128) & 127;               // pixel within the segment
    int item = blockIdx.x, r = g;                  // row to fetch next
    uint32_t q = g;                                // its ring index
    TcItem t = tc_item(item < nitems ? item : 0, xsegs, yblocks);
    float v[kTcC], vn[kTcC];
    long long soff = -1, soff_next = -1;             // where the sign word of the row in v / vn goes (-1: nowhere)
    auto fetch = [&](float* dst, long long& so) {
      const int gy = t.y0 - 1 + r;
      const bool ok = item < nitems && gy >= 0 && gy < H && !(kTcProbe && (debug & 2));
      const float* src = x + ((size_t)t.n * kTcC * H + (ok ? gy : 0)) * W + t.x0 + j;
#pragma unroll
      for (int c = 0; c < kTcC; ++c) dst[c] = ok ? __ldg(src + (size_t)c * plane) : 0.0f;
      // rows 1 .. 8 of an item are its own (0 and 9 are halo rows another item owns)
      so = (!MASKED && in_signs != nullptr && ok && r >= 1 && r <= kTcRowBlock)
               ? (long long)(((size_t)t.n * H + gy) * W + t.x0 + j)
               : -1;
    };
    fetch(v, soff);
    while (item < nitems) {
      const uint32_t slot = q & 3;
      // advance the fetch position by two rows and get that row on its way
      r += 2;
      if (r >= kTcRowBlock + 2) {
        r -= kTcRowBlock + 2;
        item += gridDim.x;
        if (item < nitems) t = tc_item(item, xsegs, yblocks);
      }
      fetch(vn, soff_next);
      if (!MASKED && soff >= 0) {
        uint32_t bits = 0u;
#pragma unroll
        for (int c = 0; c < kTcC; ++c) bits |= (v[c] > 0.0f ? 1u : 0u) << c;
        in_signs[soff] = bits;
      }
      soff = soff_next;
      TC_PROF_WAIT(0, row_free + 8 * slot, ((q >> 2) & 1) ^ 1);
      unsigned char* hi_s = A_s + (size_t)slot * 2 * kTcSlotPart + (j + 1) * 16;
      unsigned char* lo_s = hi_s + kTcSlotPart;
      if (!(kTcProbe && (debug & 16))) {
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          float4 h, l;
          tc_split(v[4 * kc], h.x, l.x);
          tc_split(v[4 * kc + 1], h.y, l.y);
          tc_split(v[4 * kc + 2], h.z, l.z);
          tc_split(v[4 * kc + 3], h.w, l.w);
          *reinterpret_cast<float4*>(hi_s + kc * kTcPlane) = h;
          *reinterpret_cast<float4*>(lo_s + kc * kTcPlane) = l;
        }
      }
      // every thread publishes its own stores to the async proxy, then ONE lane per warp
      // arrives (an arrival per thread is serialised on the barrier word)
      if (!(kTcProbe && (debug & 8))) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(row_full + 8 * slot);
      q += 2;
#pragma unroll
      for (int c = 0; c < kTcC; ++c) v[c] = vn[c];
    }
  } else if (warp == 12) {
    // ===== halo warp: slot pixels 0 and 129 (image columns x0 - 1 and x0 + 128) of EVERY
    // ring row; lane = (side, channel pair), next row fetched one step ahead =====
    const int side = lane >> 4, c0 = (lane & 15) * 2;
    int item = blockIdx.x, r = 0;
    uint32_t q = 0;
    TcItem t = tc_item(item < nitems ? item : 0, xsegs, yblocks);
    float a0, a1, n0, n1;
    auto fetch = [&](float& o0, float& o1) {
      const int gy = t.y0 - 1 + r, gx = side ? t.x0 + kTcM : t.x0 - 1;
      const bool ok = item < nitems && gy >= 0 && gy < H && gx >= 0 && gx < W && !(kTcProbe && (debug & 2));
      const float* src = x + (((size_t)t.n * kTcC + c0) * H + (ok ? gy : 0)) * W + (ok ? gx : 0);
      o0 = ok ? __ldg(src) : 0.0f;
      o1 = ok ? __ldg(src + plane) : 0.0f;
    };
    fetch(a0, a1);
    while (item < nitems) {
      const uint32_t slot = q & 3;
      if (++r >= kTcRowBlock + 2) {
        r = 0;
        item += gridDim.x;
        if (item < nitems) t = tc_item(item, xsegs, yblocks);
      }
      fetch(n0, n1);
      tc_mbar_wait(row_free + 8 * slot, ((q >> 2) & 1) ^ 1);
      float* hi_s = reinterpret_cast<float*>(A_s + (size_t)slot * 2 * kTcSlotPart + (c0 >> 2) * kTcPlane +
                                             (side ? kTcM + 1 : 0) * 16) + (c0 & 3);
      float* lo_s = hi_s + kTcSlotPart / 4;
      float2 h, l;
      tc_split(a0, h.x, l.x);
      tc_split(a1, h.y, l.y);
      *reinterpret_cast<float2*>(hi_s) = h;
      *reinterpret_cast<float2*>(lo_s) = l;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(row_full + 8 * slot);
      ++q;
      a0 = n0;
      a1 = n1;
    }
  } else if (warp >= kTcMmaWarp) {
    // ===== MMA issuers: warp 13 issues the even staged rows, warp 14 the odd ones.  Every row
    // has an accumulator block, a ring slot and barriers of its own, so the two are independent;
    // one thread alone needs ~500 cycles per row for waits, fence and commits, during which the
    // (shallow) instruction queue of the tensor core would run dry.  Each warp runs its loop
    // with uniform control flow, one elected lane issues. =====
    const bool leader = tc_elect();
    const uint32_t par = warp - kTcMmaWarp;
    {
      // D fp32, A / B tf32, both K-major, M = 128, N = 32 x (vertical taps in the window)
      const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcM >> 4) << 24);
      const uint64_t da_base = tc_desc(tc_s32(A_s), kTcPlane, 128);
      const uint64_t db_base = tc_desc(tc_s32(B_s), kTcBPlane, 128);
      const uint32_t total = blockIdx.x < nitems
                                 ? (uint32_t)((nitems - blockIdx.x + gridDim.x - 1) / gridDim.x) * (kTcRowBlock + 2)
                                 : 0u;
      for (uint32_t q = par; q < total; q += 2) {
        {
          const uint32_t itemc = q / (kTcRowBlock + 2);
          const int i = (int)(q - itemc * (kTcRowBlock + 2));
          // staged row i is image row y0 - 1 + i; it feeds output row y0 + i - ky through vertical
          // tap ky, and only the taps ky_lo .. ky_hi land inside this item.  The row gets a FRESH
          // accumulator block (i % 5; the first instruction overwrites it): column group ky of
          // the block holds this row's contribution to output row i - ky, and the epilogue adds
          // the three contributions of an output row in fp32 registers.  The tensor core rounds
          // its accumulator toward zero after every instruction, so 36-instruction chains
          // instead of 108 are what keeps the result at fp32 level (rel-L2 5e-7 instead of 1.6e-6).
          const uint32_t slot = q & 3, blk = (uint32_t)i % kTcBlocks;
          const uint32_t use = itemc * ((kTcRowBlock + 2) / kTcBlocks) + (uint32_t)i / kTcBlocks;
          TC_PROF_WAIT(1, acc_free + 8 * blk, (use & 1) ^ 1);
          TC_PROF_WAIT(2, row_full + 8 * slot, (q >> 2) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int ky_lo = i > kTcRowBlock - 1 ? i - (kTcRowBlock - 1) : 0;
          const int ky_hi = i < 2 ? i : 2;
          const int nky = ky_hi - ky_lo + 1;
          const uint32_t d_win = tmem + blk * (3 * kTcC) + (uint32_t)ky_lo * kTcC;
          const uint32_t idesc = idesc0 | ((uint32_t)(nky * kTcC >> 3) << 17);
          const uint64_t da_hi = da_base + (uint32_t)((slot * (2 * kTcSlotPart)) >> 4);
          const uint64_t da_lo = da_hi + (uint32_t)(kTcSlotPart >> 4);
          const uint64_t db_hi = db_base + (uint32_t)((ky_lo * kTcC * 16) >> 4);
          const uint64_t db_lo = db_hi + (uint32_t)(kTcBPart >> 4);
          if (leader && !(kTcProbe && (debug & 1))) {
#pragma unroll
            for (int term = 0; term < 3; ++term) {          // lo*hi, hi*lo, then hi*hi
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  // only the 14-bit start-address field (16-byte units) changes between MMAs
                  const uint64_t da = (term == 0 ? da_lo : da_hi) + (uint32_t)((2 * ks * kTcPlane + kx * 16) >> 4);
                  const uint64_t db = (term == 1 ? db_lo : db_hi) +
                                      (uint32_t)((kx * (2 * kTcBPart) + 2 * ks * kTcBPlane) >> 4);
                  tc_mma(d_win, da, db, idesc, (term == 0 && kx == 0 && ks == 0) ? 0u : 1u);
                }
              }
            }
          }
          if (leader) {
            tc_commit(row_free + 8 * slot);                 // the ring slot may be refilled
            tc_commit(acc_full + 8 * blk);                  // this row's contributions are complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: warp e owns TMEM lanes 32e .. 32e + 31 = pixels of the segment =====
    uint32_t positive = 0u;         // bit c = (output channel c of this pixel > 0)
    uint32_t sign_next = 0u;        // MASKED only: the sign word of the NEXT row, in flight
    if (MASKED && blockIdx.x < nitems) {
      const TcItem t0 = tc_item(blockIdx.x, xsegs, yblocks);
      sign_next = __ldg(signs + ((size_t)t0.n * H + t0.y0) * W + t0.x0 + tid);
    }
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t itemc = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++itemc) {
      const TcItem t = tc_item(item, xsegs, yblocks);
      float* yn = y + (size_t)t.n * kTcC * plane + t.x0 + tid;
      for (int r = 0; r < kTcRowBlock; ++r) {
        if (MASKED) {
          // this row's sign word was requested one row ago; request the next one (same item,
          // or the first row of this CTA's next item)
          positive = sign_next;
          size_t nx = ((size_t)t.n * H + (t.y0 + r + 1)) * W + t.x0 + tid;
          bool more = true;
          if (r == kTcRowBlock - 1) {
            const int nitem = item + gridDim.x;
            more = nitem < nitems;
            const TcItem tn = tc_item(more ? nitem : item, xsegs, yblocks);
            nx = ((size_t)tn.n * H + tn.y0) * W + tn.x0 + tid;
          }
          if (more) sign_next = __ldg(signs + nx);
        }
        // output row r = staged rows r (tap 0), r + 1 (tap 1), r + 2 (tap 2), issued by two
        // independent threads: wait for all three blocks (the first two are normally long done)
        const uint32_t i2 = (uint32_t)r + 2;
#pragma unroll
        for (uint32_t k = 0; k < 3; ++k) {
          const uint32_t ik = (uint32_t)r + k, usek = itemc * ((kTcRowBlock + 2) / kTcBlocks) + ik / kTcBlocks;
          TC_PROF_WAIT(3, acc_full + 8 * (ik % kTcBlocks), usek & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* dst = yn + (size_t)(t.y0 + r) * W;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                       // 16 channels at a time (register budget)
          float v0[16], v1[16], v2[16];
          if (!(kTcProbe && (debug & 32))) {
            TC_LD16(v0, tmem + lane_base + ((uint32_t)r % kTcBlocks) * (3 * kTcC) + 16 * h);
            TC_LD16(v1, tmem + lane_base + (((uint32_t)r + 1) % kTcBlocks) * (3 * kTcC) + kTcC + 16 * h);
            TC_LD16(v2, tmem + lane_base + (i2 % kTcBlocks) * (3 * kTcC) + 2 * kTcC + 16 * h);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          if (h == 1) {
            // staged row r is read for the last time here (the last row of an item also
            // retires the two rows below it): their accumulator blocks may be overwritten
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
              tc_mbar_arrive(acc_free + 8 * ((uint32_t)r % kTcBlocks));
              if (r == kTcRowBlock - 1) {
                tc_mbar_arrive(acc_free + 8 * (((uint32_t)r + 1) % kTcBlocks));
                tc_mbar_arrive(acc_free + 8 * (i2 % kTcBlocks));
              }
            }
          }
          float* dp = dst + (size_t)(16 * h) * plane;      // one live pointer, advanced per channel
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float o = (kTcProbe && (debug & 32)) ? 0.0f : (v0[c] + v2[c]) + v1[c];
            float ov;
            if (MASKED) {
              ov = ((positive >> (16 * h + c)) & 1u) ? o : o * slope;
            } else {
              ov = o + bias_s[16 * h + c];
              positive = (h == 0 && c == 0 ? 0u : positive) | ((ov > 0.0f ? 1u : 0u) << (16 * h + c));
              if (slope > 0.0f) ov = ov > 0.0f ? ov : ov * slope;
            }
            if (!(kTcProbe && (debug & 4)))
              asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(dp), "f"(ov) : "memory");
            asm volatile("add.u64 %0, %0, %1;" : "+l"(dp) : "l"((unsigned long long)plane * 4ull));
          }
        }
        if (!MASKED && signs != nullptr) signs[((size_t)t.n * H + (t.y0 + r)) * W + t.x0 + tid] = positive;
      }
    }
  }
  if ((kTcProbe && (debug & 128)) && blockIdx.x == 0) {
    if (tid == 128) tc_prof[0] = prof[0];                      // producer group 0: waiting for a free ring slot
    if (tid == kTcMmaWarp * 32) { tc_prof[1] = 2 * prof[1]; tc_prof[2] = 2 * prof[2]; tc_prof[5] = clock64() - t_start; }
    if (tid == 0) tc_prof[3] = prof[3];                        // epilogue: waiting for a finished row
  }
  // ---- teardown ------------------------------------------------------------------
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kTcMmaWarp)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace csmri
