// In-pass retrieval ranks: side kernels of fwd_rowstats_pair_kernel<1>.
//   diag_z[i]  = <x_i, y_{first_col + i}>   the similarity of local row i with its own pair
//   ranks[i]   = sum over partial slots of the per-chunk counts written by the tensor-core pass
#include <cuda_bf16.h>

#include "scl_kernels.h"

namespace scl {
namespace {

// warp per row, 16-byte loads
__global__ void __launch_bounds__(256) retrieval_diag_kernel(const __nv_bfloat16* __restrict__ x_rows, int m_rows,
                                                             const __nv_bfloat16* __restrict__ y_cols, int d,
                                                             int first_col, float* __restrict__ diag_z) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m_rows) return;
  const __nv_bfloat16* a = x_rows + static_cast<size_t>(row) * d;
  const __nv_bfloat16* b = y_cols + static_cast<size_t>(first_col + row) * d;
  float acc = 0.f;
  for (int c = lane * 8; c < d; c += 256) {
    const uint4 ua = *reinterpret_cast<const uint4*>(a + c);
    const uint4 ub = *reinterpret_cast<const uint4*>(b + c);
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w};
    const uint32_t wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wa[k]));
      const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wb[k]));
      acc = fmaf(fa.x, fb.x, acc);
      acc = fmaf(fa.y, fb.y, acc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) diag_z[row] = acc;
}

__global__ void __launch_bounds__(256) retrieval_rank_sum_kernel(const int* __restrict__ rank_part, int n_slots,
                                                                 int m_pad, int m_rows, int* __restrict__ ranks) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m_rows) return;
  int acc = 0;
  for (int s = 0; s < n_slots; ++s) acc += rank_part[static_cast<size_t>(s) * m_pad + row];
  ranks[row] = acc;
}

}  // namespace

cudaError_t launch_retrieval_diag(const void* x_rows, int m_rows, const void* y_cols, int d, int first_col,
                                  float* diag_z, cudaStream_t stream) {
  if (m_rows <= 0) return cudaSuccess;
  retrieval_diag_kernel<<<(m_rows + 7) / 8, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x_rows), m_rows,
                                                              static_cast<const __nv_bfloat16*>(y_cols), d, first_col,
                                                              diag_z);
  return cudaGetLastError();
}

cudaError_t launch_retrieval_rank_sum(const int* rank_part, int n_slots, int m_pad, int m_rows, int* ranks,
                                      cudaStream_t stream) {
  if (m_rows <= 0) return cudaSuccess;
  retrieval_rank_sum_kernel<<<(m_rows + 255) / 256, 256, 0, stream>>>(rank_part, n_slots, m_pad, m_rows, ranks);
  return cudaGetLastError();
}

}  // namespace scl
