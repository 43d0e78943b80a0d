// Index plumbing of the loader tail: CenterCropInKspace and the max
// normalisation of data/reconstruction/rec_transforms.py:40-47,62-65.
//
// CenterCropInKspace (myImageTransformations.py:935-954) is
//     abs(ifft2c(crop_image_at(fft2c(x), nx/2, ny/2, sx, sy)))
// with fft2c = fftshift . fft2 . ifftshift (deep_med_lib/utils/mymath.py:18-29).  Every
// shift, the crop and its zero padding are index maps, so each stage between two FFTs is ONE
// pass that gathers through the composed map instead of a chain of roll / slice / pad / roll
// copies:
//     out[b, c, i, j] = in[b, c, sy(i), sx(j)]          (0 where the crop box leaves the input)
//     u = (i + out_roll) mod O      undo the roll applied to the output
//     v = u + off                   crop box origin, may be negative (zero padding)
//     s = (v + in_roll) mod I       undo the roll applied to the input
// A real input (in_ch = 1) gets a zero imaginary plane; out_ch = 1 writes the magnitude
// (same rounding as magnitude_clamp_kernel) and can fold max |.| per slice into `absmax`
// (values are >= 0, so their bit patterns order like unsigned integers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace csmri {

struct ShiftCropAxis {
  int in_n, out_n, in_roll, off, out_roll;
};

__device__ __forceinline__ int shift_crop_src(const ShiftCropAxis a, int i) {
  int u = i + a.out_roll;
  if (u >= a.out_n) u -= a.out_n;
  const int v = u + a.off;
  if (v < 0 || v >= a.in_n) return -1;
  int s = v + a.in_roll;
  if (s >= a.in_n) s -= a.in_n;
  return s;
}

// One thread per output pixel pair (re, im) of kShiftCropRows rows; threads of a warp walk
// along a row, and the source of a row is at most two contiguous runs (the roll wraps once),
// so loads coalesce.  The loads of all rows are issued before the first store.
constexpr int kShiftCropRows = 4;

template <int IN_CH, int OUT_CH>
__global__ void __launch_bounds__(256)
    shift_crop_kernel(const float* __restrict__ in, float* __restrict__ out, ShiftCropAxis ay,
                      ShiftCropAxis ax, unsigned* __restrict__ absmax) {
  const int b = blockIdx.z;
  const int i0 = blockIdx.y * kShiftCropRows;
  const size_t ip = (size_t)ay.in_n * ax.in_n, op = (size_t)ay.out_n * ax.out_n;
  const float* src = in + (size_t)b * IN_CH * ip;
  float* dst = out + (size_t)b * OUT_CH * op;
  int sy[kShiftCropRows];
#pragma unroll
  for (int r = 0; r < kShiftCropRows; ++r) sy[r] = i0 + r < ay.out_n ? shift_crop_src(ay, i0 + r) : -1;
  float m = 0.0f;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ax.out_n; j += gridDim.x * blockDim.x) {
    const int sx = shift_crop_src(ax, j);
    float re[kShiftCropRows], im[kShiftCropRows];
#pragma unroll
    for (int r = 0; r < kShiftCropRows; ++r) {
      re[r] = im[r] = 0.0f;
      if (sx >= 0 && sy[r] >= 0) {
        const size_t o = (size_t)sy[r] * ax.in_n + sx;
        re[r] = __ldg(src + o);
        if (IN_CH == 2) im[r] = __ldg(src + ip + o);
      }
    }
#pragma unroll
    for (int r = 0; r < kShiftCropRows; ++r) {
      if (i0 + r >= ay.out_n) break;
      float* d = dst + (size_t)(i0 + r) * ax.out_n + j;
      if (OUT_CH == 2) {
        d[0] = re[r];
        d[op] = im[r];
      } else {
        const float v = __fsqrt_rn(__fadd_rn(__fmul_rn(re[r], re[r]), __fmul_rn(im[r], im[r])));
        d[0] = v;
        m = fmaxf(m, v);
        if (v != v) m = v;                                // NaN propagates like torch.amax
      }
    }
  }
  if (OUT_CH == 1 && absmax != nullptr) {
    unsigned k = __float_as_uint(m);
    for (int s = 16; s > 0; s >>= 1) {
      const unsigned o = __shfl_xor_sync(0xffffffffu, k, s);
      k = o > k ? o : k;
    }
    __shared__ unsigned part[8];                          // one same-address atomic per block
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) k = part[w] > k ? part[w] : k;
      if (k != 0u) atomicMax(absmax + b, k);
    }
  }
}

// max |x| per plane (bit pattern of a non-negative float; NaN sorts above everything)
__global__ void __launch_bounds__(256)
    plane_absmax_kernel(const float* __restrict__ x, unsigned* __restrict__ absmax, int n) {
  const int p = blockIdx.y;
  const float* src = x + (size_t)p * n;
  unsigned k = 0u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned v = __float_as_uint(fabsf(__ldg(src + i)));
    k = v > k ? v : k;
  }
  for (int s = 16; s > 0; s >>= 1) {
    const unsigned o = __shfl_xor_sync(0xffffffffu, k, s);
    k = o > k ? o : k;
  }
  __shared__ unsigned part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) k = part[w] > k ? part[w] : k;
    if (k != 0u) atomicMax(absmax + p, k);
  }
}

// out = x / denom[plane], IEEE division (what torch's `x / x.abs().amax(...)` evaluates);
// VEC = 4 when n % 4 == 0 and both pointers are 16-byte aligned
template <int VEC>
__global__ void __launch_bounds__(256)
    plane_divide_kernel(const float* __restrict__ x, const float* __restrict__ denom,
                        float* __restrict__ out, int n) {
  const int p = blockIdx.y;
  const float d = denom[p];
  const float* src = x + (size_t)p * n;
  float* dst = out + (size_t)p * n;
  if (VEC == 4) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += gridDim.x * blockDim.x) {
      float4 v = __ldg(s4 + i);
      v.x = __fdiv_rn(v.x, d);
      v.y = __fdiv_rn(v.y, d);
      v.z = __fdiv_rn(v.z, d);
      v.w = __fdiv_rn(v.w, d);
      d4[i] = v;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      dst[i] = __fdiv_rn(__ldg(src + i), d);
  }
}

}  // namespace csmri
