// Operand preparation for the fp32-accurate ("bf16x2") mode: every feature value x is carried as a bf16 pair
// h = bf16(x), l = bf16(x - h) (|x - h - l| <= 2^-18 |x|), laid out so that the UNCHANGED bf16 tensor-core
// kernels compute x.y ~= xh.yh + xh.yl + xl.yh by contracting over K-concatenated rows of width 3 d:
//   row operand     X' = (h | h | l)        column operand  Y' = (h | l | h)
// (the gradient GEMM reads Yh / Yl as the column ranges [0, d) / [d, 2 d) of Y', MN-major).  HBM-bound, coalesced.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "scl_kernels.h"

namespace scl {
namespace {

template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void ld4<__half>(const __half* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

__device__ __forceinline__ uint2 pack4(const float (&v)[4]) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]);
  const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2], v[3]);
  uint2 p;
  p.x = *reinterpret_cast<const uint32_t*>(&lo);
  p.y = *reinterpret_cast<const uint32_t*>(&hi);
  return p;
}

// one thread = 4 consecutive elements of one row
template <typename T>
__global__ void __launch_bounds__(256) split_cast_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ rows_out,
                                                         __nv_bfloat16* __restrict__ cols_out, int rows, int d) {
  const int qpr = d >> 2;  // quads per row
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= static_cast<size_t>(rows) * qpr) return;
  const int row = static_cast<int>(t / qpr);
  const int c = static_cast<int>(t - static_cast<size_t>(row) * qpr) * 4;
  float v[4];
  ld4(x + static_cast<size_t>(row) * d + c, v);
  const uint2 ph = pack4(v);
  float r[4];
  r[0] = v[0] - __uint_as_float(ph.x << 16);
  r[1] = v[1] - __uint_as_float(ph.x & 0xffff0000u);
  r[2] = v[2] - __uint_as_float(ph.y << 16);
  r[3] = v[3] - __uint_as_float(ph.y & 0xffff0000u);
  const uint2 pl = pack4(r);
  const size_t base = static_cast<size_t>(row) * 3 * d + c;
  if (rows_out != nullptr) {  // (h | h | l)
    *reinterpret_cast<uint2*>(rows_out + base) = ph;
    *reinterpret_cast<uint2*>(rows_out + base + d) = ph;
    *reinterpret_cast<uint2*>(rows_out + base + 2 * d) = pl;
  }
  if (cols_out != nullptr) {  // (h | l | h)
    *reinterpret_cast<uint2*>(cols_out + base) = ph;
    *reinterpret_cast<uint2*>(cols_out + base + d) = pl;
    *reinterpret_cast<uint2*>(cols_out + base + 2 * d) = ph;
  }
}

}  // namespace

cudaError_t launch_split_cast(const void* x, int src_dtype, void* rows_out, void* cols_out, int rows, int d,
                              cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  const size_t quads = static_cast<size_t>(rows) * (d / 4);
  const unsigned grid = static_cast<unsigned>((quads + 255) / 256);
  auto ro = static_cast<__nv_bfloat16*>(rows_out);
  auto co = static_cast<__nv_bfloat16*>(cols_out);
  if (src_dtype == 0)
    split_cast_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), ro, co, rows, d);
  else if (src_dtype == 1)
    split_cast_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ro, co, rows, d);
  else
    split_cast_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(x), ro, co, rows, d);
  return cudaGetLastError();
}

}  // namespace scl
