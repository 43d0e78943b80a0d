// Persistent TMA-fed strip kernel, TWO image columns per thread (cf2 / LineFFTV).
//
// ncu on the one-column kernel put `mio_throttle` on top: ~150-190 of its ~900
// instructions per thread and tile are shared/global memory instructions moving
// 4 or 8 bytes each.  With two adjacent columns per thread in SoA form
// ({re_w, re_w+1}, {im_w, im_w+1}) every access doubles in width for free -
// LDS.64 from the TMA tile, STS/LDS.128 for the FFT exchange, STG.64 for the
// result - while the packed FADD2/FFMA2 count per element is unchanged and the
// shared twiddle / D factors are loaded once for both columns.  Memory
// instructions per element drop from ~10 to ~4.5, and each thread carries two
// independent butterfly streams (more ILP for the same number of warps).
#pragma once
#include "dc_pipe.cuh"

namespace csmri {

__device__ __forceinline__ void st_stream_v2(float* p, float2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y)
               : "memory");
}
__device__ __forceinline__ float2 ld_stream_v2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];"
               : "=f"(v.x), "=f"(v.y)
               : "l"(p));
  return v;
}

template <int H, int E, int CW, bool ADD>
struct PipeVSmem {
  static constexpr int LW = CW / 2;
  static constexpr int kRowChunks = (H + 255) / 256;
  static constexpr int kTileFloats = 2 * H * CW;
  static constexpr int kTileBytes = kTileFloats * 4;
  static constexpr int kExchBytes = LineFFTV<H, E, LW>::kSmemBytes;   // = H*CW*8, 128-B multiple
  static constexpr int kTwBytes = LineFFTV<H, E, LW>::kTwBytes;
  static constexpr int kBytes = kExchBytes + (ADD ? 2 : 1) * kTileBytes + 2 * H * 4 + kTwBytes + 64;
};

template <int H, int E, int CW, int MINB, int WT, bool ADD>
__global__ void __launch_bounds__((CW / 2) * (H / E), MINB)
    dc_strip_pipev_kernel(const __grid_constant__ CUtensorMap tm_x,
                          const __grid_constant__ CUtensorMap tm_add,
                          const float* __restrict__ residual, const float* __restrict__ dtab,
                          float* __restrict__ out, int W_rt, int nstrips_rt, int ntiles,
                          long long* __restrict__ trace, unsigned* __restrict__ sched) {
  const int W = WT ? WT : W_rt;
  const int nstrips = WT ? WT / CW : nstrips_rt;
  constexpr int LW = CW / 2;
  typedef LineFFTV<H, E, LW> L;
  typedef PipeVSmem<H, E, CW, ADD> S;
  constexpr int T = L::T;
  constexpr int NT = LW * T;
  if (trace != nullptr && threadIdx.x == 0) {   // tuning probe: kernel entry time
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(size_t)blockIdx.x * 40 + 39] = t;
  }
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  float4* sm = reinterpret_cast<float4*>(smem_dyn);
  float* xbuf = reinterpret_cast<float*>(smem_dyn + S::kExchBytes);   // [2][H][CW]
  float* abuf = xbuf + S::kTileFloats;                                // absent when !ADD
  float* dbuf = xbuf + (ADD ? 2 : 1) * S::kTileFloats;                // [2][H]
  cf* tw_s = reinterpret_cast<cf*>(dbuf + 2 * H);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tw_s + L::Base::T * L::Base::kTwPitch);
  const uint32_t bar_x = smem_u32(&bars[0]);
  const uint32_t bar_a = smem_u32(&bars[1]);
  const uint32_t bar_ae = smem_u32(&bars[2]);

  const int lane = threadIdx.x % LW;
  const int j = threadIdx.x / LW;
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
    fence_barrier_init();
    prefetch_tensormap(&tm_x);
    if (ADD) prefetch_tensormap(&tm_add);
  }
  L::Base::fill_twiddles(tw_s, threadIdx.x, NT);   // immutable table: safe before the wait
  // Programmatic dependent launch: everything above overlapped the tail of the
  // previous kernel in the stream; its results are visible after this wait, and
  // the kernel after us may start its own prologue as our CTAs retire.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0 && tile < ntiles) {
    issue_x(tile, 0);
    if (ADD) issue_a(tile);
  }
  __syncthreads();

  const float2* xs_re = reinterpret_cast<const float2*>(xbuf);
  const float2* xs_im = reinterpret_cast<const float2*>(xbuf + H * CW);
  const float2* as_re = reinterpret_cast<const float2*>(abuf);
  const float2* as_im = reinterpret_cast<const float2*>(abuf + H * CW);

  // tuning probe: per-CTA timeline (globaltimer ns) - [0] start, [1+it] end of tile it
  auto stamp = [&](int slot_) {
    if (trace != nullptr && threadIdx.x == 0 && slot_ < 40) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trace[(size_t)blockIdx.x * 40 + slot_] = t;
    }
  };
  stamp(0);

  // Tiles are handed out dynamically (one atomic per tile, taken by thread 0
  // one tile ahead, when it issues that tile's TMA load): CTAs drift by several
  // microseconds over a launch, and a static round-robin leaves the fast ones
  // idle at the end.  sched[0] = tiles handed out beyond the first wave,
  // sched[1] = retired CTAs; the last CTA to retire re-arms both.
  uint32_t phase = 0;
  for (int it = 0; tile < ntiles; ++it, phase ^= 1) {
    const int slot = it & 1;
    const int b = tile / nstrips, strip = tile - b * nstrips;
    const size_t gbase = (size_t)b * 2 * plane + (size_t)strip * CW + 2 * lane;

    if (ADD && it > 0 && threadIdx.x == 0) {
      mbar_wait(bar_ae, phase ^ 1);
      issue_a(tile);
    }

    cf2 v[E];
    mbar_wait(bar_x, phase);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int h = j + T * i;
      v[i] = mk2(xs_re[h * LW + lane], xs_im[h * LW + lane]);
    }
    if (residual != nullptr) {
      const float* pr = residual + gbase;
      const float* pi = pr + plane;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const size_t o = (size_t)(j + T * i) * W;
        v[i] = cadd(v[i], mk2(ld_stream_v2(pr + o), ld_stream_v2(pi + o)));
      }
    }

    L::template a_front<false>(v, sm, tw_s, j, lane);
    if (threadIdx.x == 0)
      next_tile[slot ^ 1] = (int)atomicAdd(&sched[0], 1u) + (int)gridDim.x;
    __syncthreads();  // exchange written; x tile consumed by everyone; next_tile visible
    const int next = next_tile[slot ^ 1];
    if (threadIdx.x == 0 && next < ntiles) issue_x(next, slot ^ 1);

    L::template a_back<false>(v, sm, j, lane);
    L::apply_dtab(v, dbuf + slot * H + j * E);
    if (ADD) {   // + hybrid-space k0 term (rows of the tile in k-layout order)
      mbar_wait(bar_a, phase);
#pragma unroll
      for (int r = 0; r < E; ++r) {
        const int k = L::Base::k_index(j, r);
        v[r] = cadd(v[r], mk2(as_re[k * LW + lane], as_im[k * LW + lane]));
      }
      mbar_arrive(bar_ae);
    }
    L::template b_front<true>(v, sm, j, lane);
    __syncthreads();
    L::template b_back<true>(v, sm, tw_s, j, lane);
    {
      float* pr = out + gbase;
      float* pi = pr + plane;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const size_t o = (size_t)(j + T * i) * W;
        st_stream_v2(pr + o, v[i].re);
        st_stream_v2(pi + o, v[i].im);
      }
    }
    stamp(1 + it);
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
