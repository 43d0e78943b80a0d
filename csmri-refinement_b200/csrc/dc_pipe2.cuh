// Second-generation persistent strip kernel: more bytes in flight per SM.
//
// Measurements behind it (profiles/): with the FFT removed the
// first pipelined kernel ran no faster, while a plain strip copy with 256 KiB
// of loads in flight per SM reaches the HBM ceiling - the limit was the amount
// of landing space (shared memory) the loads could be in flight into.  So:
//   * the x tile is multi-buffered (XS stages): the TMA load of tile i+XS is
//     issued the moment tile i has been pulled into registers;
//   * the addend no longer lands in shared memory at all: each thread loads its
//     own 2*E values straight into registers right after the first barrier of
//     the tile, where they have the whole blend + inverse FFT to arrive;
//   * the freed shared memory pays for the extra x stage(s).
// 256^2, CW=16: exchange 32 KiB + 2 x 32 KiB + tables -> two 256-thread CTAs/SM,
// up to 2*(64 + 32) KiB of loads in flight per SM.
#pragma once
#include "dc_pipe.cuh"

namespace csmri {

template <int H, int E, int CW, int XS>
struct Pipe2Smem {
  static constexpr int kRowChunks = (H + 255) / 256;
  static constexpr int kTileFloats = 2 * H * CW;
  static constexpr int kTileBytes = kTileFloats * 4;
  static constexpr int kExchBytes = (LineFFT<H, E, CW>::kSmemBytes + 127) / 128 * 128;
  static constexpr int kDSlots = XS + 1;
  static constexpr int kTwBytes = LineFFT<H, E, CW>::kTwBytes;
  static constexpr int kBytes = kExchBytes + XS * kTileBytes + kDSlots * H * 4 + kTwBytes + 8 * XS + 64;
};

template <int H, int E, int CW, int MINB, int WT, bool ADD, int XS>
__global__ void __launch_bounds__(CW*(H / E), MINB)
    dc_strip_pipe2_kernel(const __grid_constant__ CUtensorMap tm_x,
                          const float* __restrict__ addend, const float* __restrict__ residual,
                          const float* __restrict__ dtab, float* __restrict__ out, int W_rt,
                          int nstrips_rt, int ntiles) {
  const int W = WT ? WT : W_rt;
  const int nstrips = WT ? WT / CW : nstrips_rt;
  typedef LineFFT<H, E, CW> L;
  typedef Pipe2Smem<H, E, CW, XS> S;
  constexpr int T = L::T;
  constexpr int NT = CW * T;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  cf* sm = reinterpret_cast<cf*>(smem_dyn);
  float* xbuf = reinterpret_cast<float*>(smem_dyn + S::kExchBytes);   // [XS][2][H][CW]
  float* dbuf = xbuf + XS * S::kTileFloats;                           // [XS+1][H]
  cf* tw_s = reinterpret_cast<cf*>(dbuf + S::kDSlots * H);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tw_s + H);
  L::fill_twiddles(tw_s, threadIdx.x, NT);

  const int lane = threadIdx.x % CW;
  const int j = threadIdx.x / CW;
  const size_t plane = (size_t)H * W;

  auto issue_x = [&](int tile, int stage, int dslot) {
    const int b = tile / nstrips, strip = tile - b * nstrips;
    const uint32_t bar = smem_u32(&bars[stage]);
    mbar_expect_tx(bar, S::kTileBytes + H * 4);
    const uint32_t dst = smem_u32(xbuf + stage * S::kTileFloats);
    if (S::kRowChunks == 1)
      tma_load_3d(dst, &tm_x, bar, strip * CW, 0, b * 2);
    else
      tma_load_4d(dst, &tm_x, bar, strip * CW, 0, 0, b * 2);
    bulk_load_1d(smem_u32(dbuf + dslot * H), dtab + (size_t)b * H, H * 4, bar);
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < XS; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncthreads();   // barriers initialised, twiddle table filled
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < XS; ++s) {
      const int t = blockIdx.x + s * gridDim.x;
      if (t < ntiles) issue_x(t, s, s);
    }
  }

  int stage = 0, dslot = 0;
  uint32_t parity = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / nstrips, strip = tile - b * nstrips;
    const size_t gbase = (size_t)b * 2 * plane + (size_t)strip * CW + lane;
    const float* xs = xbuf + stage * S::kTileFloats;

    cf v[E];
    mbar_wait(smem_u32(&bars[stage]), parity);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int h = j + T * i;
      v[i] = mk(xs[h * CW + lane], xs[(H + h) * CW + lane]);
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

    L::template a_front<false>(v, sm, tw_s, j, lane);
    __syncthreads();  // exchange written; this x stage consumed by every thread
    {
      const int nxt = tile + XS * gridDim.x;
      int nd = dslot + XS;                       // D slot of tile it+XS in a ring of XS+1
      if (nd >= S::kDSlots) nd -= S::kDSlots;
      if (threadIdx.x == 0 && nxt < ntiles) issue_x(nxt, stage, nd);
    }
    // addend straight into registers; it has a_back .. b_back to arrive
    cf ad[ADD ? E : 1];
    if (ADD) {
      const float* pr = addend + gbase;
      const float* pi = pr + plane;
#pragma unroll
      for (int i = 0; i < E; ++i) {
        const size_t o = (size_t)(j + T * i) * W;
        ad[i] = mk(ld_stream_f32(pr + o), ld_stream_f32(pi + o));
      }
    }

    L::template a_back<false>(v, sm, j, lane);
    L::apply_dtab(v, dbuf + dslot * H + j * E);
    L::template b_front<true>(v, sm, j, lane);
    __syncthreads();
    L::template b_back<true>(v, sm, tw_s, j, lane);

    if (ADD) {
#pragma unroll
      for (int i = 0; i < E; ++i) v[i] = cadd(v[i], ad[i]);
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
    if (++stage == XS) { stage = 0; parity ^= 1; }
    if (++dslot == S::kDSlots) dslot = 0;
  }
}

}  // namespace csmri
