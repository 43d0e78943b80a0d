// Pointwise ops of the refinement path, one pass each over HBM:
//   plane_minmax / plane_scale / plane_unscale
//       models/refinement_wrapper.py:51-92 (_scale, _unscale) and the min-max
//       step of utils/tensor_transforms.py:78-99 (magnitude_image)
//   refine_real_penalty_add (+ backward)
//       models/refinement_wrapper.py:173-197 with the pretrained output detached
// Every arithmetic step is rounded separately (__f*_rn, IEEE division) in the
// order the reference's tensor expressions evaluate, so results are bit-identical
// to the op-by-op torch evaluation.  Inputs are assumed finite (torch's min/max
// propagate NaN, fminf/fmaxf do not).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csmri {

// order-preserving map float -> uint32 (so that atomicMin/Max work on floats)
__device__ __forceinline__ unsigned float_key(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void minmax_init_kernel(unsigned* kmin, unsigned* kmax, int planes) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < planes) {
    kmin[p] = 0xffffffffu;
    kmax[p] = 0u;
  }
}

// grid (chunks, planes); plane p = n floats at x + p*pitch
__global__ void __launch_bounds__(256)
    minmax_reduce_kernel(const float* __restrict__ x, unsigned* __restrict__ kmin,
                         unsigned* __restrict__ kmax, int n, long long pitch) {
  const float* row = x + (size_t)blockIdx.y * (size_t)pitch;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = __ldg(row + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_lo[8], s_hi[8];
  if ((threadIdx.x & 31) == 0) {
    s_lo[threadIdx.x >> 5] = lo;
    s_hi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      lo = fminf(lo, s_lo[w]);
      hi = fmaxf(hi, s_hi[w]);
    }
    atomicMin(kmin + blockIdx.y, float_key(lo));
    atomicMax(kmax + blockIdx.y, float_key(hi));
  }
}

// minimum = min x; maximum = max (x - min) = fl(max x - min): fl(a - m) is monotone in a
__global__ void minmax_finish_kernel(float* minimum, float* maximum, int planes) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < planes) {
    const float lo = key_float(reinterpret_cast<unsigned*>(minimum)[p]);
    const float hi = key_float(reinterpret_cast<unsigned*>(maximum)[p]);
    minimum[p] = lo;
    maximum[p] = __fsub_rn(hi, lo);
  }
}

__device__ __forceinline__ float scale01(float v, float lo, float range) {
  return __fdiv_rn(__fsub_rn(v, lo), range);
}
__device__ __forceinline__ float scale_pm1(float v, float lo, float range) {
  return __fsub_rn(__fmul_rn(scale01(v, lo, range), 2.0f), 1.0f);
}
__device__ __forceinline__ float unscale_pm1(float t, float lo, float range) {
  return __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(t, 1.0f), 0.5f), range), lo);
}

// MODE 0: (x - min) / max     1: ... * 2 - 1     2: ((x + 1) / 2) * max + min
template <int MODE>
__global__ void __launch_bounds__(256)
    plane_map_kernel(const float* __restrict__ x, const float* __restrict__ minimum,
                     const float* __restrict__ maximum, float* __restrict__ out, int n,
                     long long pitch_in, long long pitch_out) {
  const int p = blockIdx.y;
  const float lo = minimum[p], range = maximum[p];
  const float* src = x + (size_t)p * (size_t)pitch_in;
  float* dst = out + (size_t)p * (size_t)pitch_out;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = __ldg(src + i);
    dst[i] = MODE == 0 ? scale01(v, lo, range)
                       : (MODE == 1 ? scale_pm1(v, lo, range) : unscale_pm1(v, lo, range));
  }
}

// pred[b,0] = unscale(scale(pre[b,0]) + s * learn[b]),  pred[b,1] = pre[b,1]
__global__ void __launch_bounds__(256)
    refine_forward_kernel(const float* __restrict__ pre, const float* __restrict__ learn,
                          const float* __restrict__ scale, const float* __restrict__ minimum,
                          const float* __restrict__ maximum, float* __restrict__ pred, int n) {
  const int b = blockIdx.y;
  const float lo = minimum[b], range = maximum[b], s = __ldg(scale);
  const float* re = pre + (size_t)b * 2 * n;
  const float* l = learn + (size_t)b * n;
  float* o = pred + (size_t)b * 2 * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float refined = __fadd_rn(scale_pm1(__ldg(re + i), lo, range), __fmul_rn(s, __ldg(l + i)));
    o[i] = unscale_pm1(refined, lo, range);
    o[n + i] = __ldg(re + n + i);
  }
}

// d pred_real / d refined = max / 2;  grad_learn = that * s;
// grad_scale = sum(that * learn), one partial per CTA at partial[b * gridDim.x + blockIdx.x]
__global__ void __launch_bounds__(256)
    refine_backward_kernel(const float* __restrict__ grad_pred, const float* __restrict__ learn,
                           const float* __restrict__ scale, const float* __restrict__ maximum,
                           float* __restrict__ grad_learn, float* __restrict__ partial, int n) {
  const int b = blockIdx.y;
  const float range = maximum[b], s = __ldg(scale);
  const float* g = grad_pred + (size_t)b * 2 * n;
  const float* l = learn + (size_t)b * n;
  float* gl = grad_learn + (size_t)b * n;
  float acc = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float gr = __fmul_rn(__fmul_rn(__ldg(g + i), range), 0.5f);
    gl[i] = __fmul_rn(gr, s);
    acc = fmaf(gr, __ldg(l + i), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s_acc[8];
  if ((threadIdx.x & 31) == 0) s_acc[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) acc += s_acc[w];
    partial[(size_t)b * gridDim.x + blockIdx.x] = acc;
  }
}

}  // namespace csmri
