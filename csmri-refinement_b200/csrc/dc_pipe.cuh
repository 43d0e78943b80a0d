// TMA-fed, persistent version of the Cartesian DC strip kernel (sm_100a), one
// image column per thread.  Also home of the mbarrier / TMA PTX wrappers the
// two-column kernel (dc_pipev.cuh, the default for 64..256) shares.
//
// The direct kernel (dc_strip_row_kernel) stalls on global-load latency: ncu
// shows long_scoreboard as the dominant stall with HBM at ~50 %.  Here the
// loads are taken off the instruction stream:
//
//   * one elected thread issues cp.async.bulk.tensor (TMA) box loads
//     {CW columns x H rows x 2 planes} of x and of the hybrid-space k0 term
//     ("addend") into shared memory, completion signalled on an mbarrier
//     (complete_tx::bytes);
//   * the CTA is persistent (grid = resident CTAs); the load of its next tile
//     is issued as soon as the current one has been pulled into registers, so
//     HBM latency hides behind the register FFTs;
//   * tiles are handed out by an atomic counter one tile ahead (CTAs drift by
//     microseconds; a static round-robin leaves the fast ones idle at the end);
//   * the D row of the next slice rides along as a 1-D bulk copy on the same
//     mbarrier; the addend buffer is released with an mbarrier arrive instead
//     of a CTA-wide barrier;
//   * launched with programmatic stream serialization: the barrier-init
//     prologue overlaps the previous kernel's tail (griddepcontrol.wait);
//   * results go straight from registers to HBM (coalesced row segments).
//
// Shared memory per CTA: exchange H*CW*8 + x tile H*CW*8 + addend tile
// H*CW*8 + 2 D rows + twiddle table.
//
// INPL ("in place") trades the one-tile-ahead prefetch for occupancy: the x
// tile lands in the exchange buffer itself (one extra CTA barrier between
// "everyone has its column in registers" and the first exchange store) and the
// next tile is requested once the inverse half has read its last exchange slot.
// The load latency is then hidden by a second resident CTA instead of by the
// CTA's own pipeline.  For the tall sizes (320, 512, 1024), where a column strip
// is 40-64 KiB, that is the difference between one and two CTAs per SM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dc_core.cuh"

namespace csmri {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void st_stream_f32(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <int H, int E, int CW, bool ADD = true, bool INPL = false>
struct PipeSmem {
  static constexpr int kRowChunks = (H + 255) / 256;          // TMA box dims are <= 256
  static constexpr int kTileFloats = 2 * H * CW;
  static constexpr int kTileBytes = kTileFloats * 4;
  static constexpr int kExchBytes = (LineFFT<H, E, CW>::kSmemBytes + 127) / 128 * 128;
  static constexpr int kDBytes = 2 * H * 4;
  static constexpr int kTwBytes = LineFFT<H, E, CW>::kTwBytes;
  // byte offsets: exchange | x tile (aliases the exchange when INPL) | addend | D | twiddles
  static constexpr int kXOff = INPL ? 0 : kExchBytes;
  static constexpr int kAOff = INPL ? (kExchBytes > kTileBytes ? kExchBytes : kTileBytes)
                                    : kExchBytes + kTileBytes;
  static constexpr int kDOff = kAOff + (ADD ? kTileBytes : 0);
  static constexpr int kBytes = kDOff + kDBytes + kTwBytes + 64;
};

template <int H, int E, int CW, int MINB, int WT, bool ADD, bool INPL>
__global__ void __launch_bounds__(CW*(H / E), MINB)
    dc_strip_pipe_kernel(const __grid_constant__ CUtensorMap tm_x,
                         const __grid_constant__ CUtensorMap tm_add,
                         const float* __restrict__ residual, const float* __restrict__ dtab,
                         float* __restrict__ out, int W_rt, int nstrips_rt, int ntiles,
                         unsigned* __restrict__ sched) {
  // WT != 0: compile-time row pitch (square slices) -> immediate store offsets
  const int W = WT ? WT : W_rt;
  const int nstrips = WT ? WT / CW : nstrips_rt;
  typedef LineFFT<H, E, CW> L;
  typedef PipeSmem<H, E, CW, ADD, INPL> S;
  constexpr int T = L::T;
  constexpr int NT = CW * T;
  // NB: no integer arithmetic on this pointer - it would demote every access
  // below from LDS/STS to generic LD/ST (seen in the first ncu source page).
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  cf* sm = reinterpret_cast<cf*>(smem_dyn);
  float* xbuf = reinterpret_cast<float*>(smem_dyn + S::kXOff);
  float* abuf = reinterpret_cast<float*>(smem_dyn + S::kAOff);   // absent when !ADD
  float* dbuf = reinterpret_cast<float*>(smem_dyn + S::kDOff);   // [2][H]
  cf* tw_s = reinterpret_cast<cf*>(dbuf + 2 * H);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tw_s + L::T * L::kTwPitch);
  const uint32_t bar_x = smem_u32(&bars[0]);    // x tile (+ D row) landed
  const uint32_t bar_a = smem_u32(&bars[1]);    // addend tile landed
  const uint32_t bar_ae = smem_u32(&bars[2]);   // addend tile consumed by all NT threads
  const uint32_t bar_f = smem_u32(&bars[3]);    // INPL: exchange buffer free for the next x tile

  const int lane = threadIdx.x % CW;
  const int j = threadIdx.x / CW;
  const size_t plane = (size_t)H * W;

  auto issue_x = [&](int tile, int slot) {
    const int b = tile / nstrips, strip = tile - b * nstrips;
    mbar_expect_tx(bar_x, S::kTileBytes + H * 4);
    if (S::kRowChunks == 1)
      tma_load_3d(smem_u32(xbuf), &tm_x, bar_x, strip * CW, 0, b * 2);
    else
      tma_load_4d(smem_u32(xbuf), &tm_x, bar_x, strip * CW, 0, 0, b * 2);
    bulk_load_1d(smem_u32(dbuf + slot * H), dtab + (size_t)b * H, H * 4, bar_x);
  };
  auto issue_a = [&](int tile) {
    const int b = tile / nstrips, strip = tile - b * nstrips;
    mbar_expect_tx(bar_a, S::kTileBytes);
    if (S::kRowChunks == 1)
      tma_load_3d(smem_u32(abuf), &tm_add, bar_a, strip * CW, 0, b * 2);
    else
      tma_load_4d(smem_u32(abuf), &tm_add, bar_a, strip * CW, 0, 0, b * 2);
  };

  // thread 0 gets the first tile moving before anything else; the twiddle
  // table fill and the barrier-visibility sync overlap its HBM latency
  int tile = blockIdx.x;
  int* next_tile = reinterpret_cast<int*>(&bars[4]);   // [2], written by thread 0
  if (threadIdx.x == 0) {
    mbar_init(bar_x, 1);
    mbar_init(bar_a, 1);
    mbar_init(bar_ae, NT);
    mbar_init(bar_f, NT);
    fence_barrier_init();
    prefetch_tensormap(&tm_x);
    if (ADD) prefetch_tensormap(&tm_add);
  }
  L::fill_twiddles(tw_s, threadIdx.x, NT);   // immutable table: safe before the wait
  // programmatic dependent launch (see dc_pipev.cuh): everything above overlaps
  // the previous kernel's tail, its results are visible after the wait
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0 && tile < ntiles) {
    issue_x(tile, 0);
    if (ADD) issue_a(tile);
  }
  __syncthreads();   // barriers initialised, twiddle table filled

  uint32_t phase = 0;
  // dynamic tile scheduler, see dc_pipev.cuh
  for (int it = 0; tile < ntiles; ++it, phase ^= 1) {
    const int slot = it & 1;
    const int b = tile / nstrips, strip = tile - b * nstrips;
    const size_t gbase = (size_t)b * 2 * plane + (size_t)strip * CW + lane;

    // the addend tile of the previous iteration has been read by everyone
    // (bar_ae): refill it now, it is needed again only at the end of this tile
    if (ADD && it > 0 && threadIdx.x == 0) {
      mbar_wait(bar_ae, phase ^ 1);
      issue_a(tile);
    }

    cf v[E];
    mbar_wait(bar_x, phase);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int h = j + T * i;
      v[i] = mk(xbuf[h * CW + lane], xbuf[(H + h) * CW + lane]);
    }
    if (residual != nullptr) {
      const float* pr = residual + gbase;
      const float* pi = pr + plane;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const size_t o = (size_t)(j + T * i) * W;
        v[i] = cadd(v[i], mk(ld_stream_f32(pr + o), ld_stream_f32(pi + o)));
      }
    }

    if (threadIdx.x == 0)
      next_tile[slot ^ 1] = (int)atomicAdd(&sched[0], 1u) + (int)gridDim.x;
    if (INPL) __syncthreads();   // the exchange stores below overwrite the x tile
    L::template a_front<false>(v, sm, tw_s, j, lane);
    __syncthreads();  // exchange written; x tile consumed by everyone; next_tile visible
    const int next = next_tile[slot ^ 1];
    if (!INPL && threadIdx.x == 0 && next < ntiles) issue_x(next, slot ^ 1);

    L::template a_back<false>(v, sm, j, lane);
    L::apply_dtab(v, dbuf + slot * H + j * E);
    if (ADD) {   // + hybrid-space k0 term (rows of the tile in k-layout order)
      mbar_wait(bar_a, phase);
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int k = L::k_index(j, r);
        v[r] = cadd(v[r], mk(abuf[k * CW + lane], abuf[(H + k) * CW + lane]));
      }
      mbar_arrive(bar_ae);
    }
    L::template b_front<true>(v, sm, j, lane);
    __syncthreads();
    if (INPL) {
      L::b_back_load(v, sm, j, lane);
      // generic-proxy accesses to the buffer are ordered before the TMA write
      // of the next tile by fence + arrive (all NT threads) / wait (thread 0)
      fence_proxy_async_smem();
      mbar_arrive(bar_f);
      if (threadIdx.x == 0 && next < ntiles) {
        mbar_wait(bar_f, phase);
        issue_x(next, slot ^ 1);
      }
      L::template b_back_compute<true>(v, tw_s, j);
    } else {
      L::template b_back<true>(v, sm, tw_s, j, lane);
    }
    {
      float* pr = out + gbase;
      float* pi = pr + plane;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const size_t o = (size_t)(j + T * i) * W;
        st_stream_f32(pr + o, v[i].x);
        st_stream_f32(pi + o, v[i].y);
      }
    }
    tile = next;
  }
  if (threadIdx.x == 0) {
    if (atomicAdd(&sched[1], 1u) == gridDim.x - 1) {   // last CTA out re-arms the slot
      sched[0] = 0;
      sched[1] = 0;
      __threadfence();
    }
  }
}

}  // namespace csmri
