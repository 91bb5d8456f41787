// Internal launcher declarations shared by the .cu files and the C-ABI layer (scl_api.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace scl {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) unless `stream` is being captured into a CUDA graph
template <typename Kernel>
inline cudaError_t set_max_dynamic_smem(Kernel kernel, int bytes, cudaStream_t stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) return cudaSuccess;
  cudaGetLastError();
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// ---- tensor-core kernels (CTA pairs, tcgen05 cta_group::2)
// Column chunking for a grid of `units` row blocks (CTA pairs) over `slots` concurrently resident units: choose the
// chunk count that minimises  waves * (tiles_per_chunk + per-CTA overhead) + chunks * chunk_cost  -- the launch
// lasts `waves` rounds of the longest chunk, and every chunk adds a partial-result slab over all rows that a later
// pass reads back (`chunk_cost`, in tile-steps) -- breaking ties towards fewer chunks.
inline int pick_chunks_balanced(int units, int n_tiles, int slots, int min_tiles, double chunk_cost,
                                int* tiles_per_chunk) {
  int best_c = 1, best_tpc = n_tiles;
  double best_cost = 1e30;
  const int max_c = n_tiles / min_tiles > 1 ? n_tiles / min_tiles : 1;
  for (int c = 1; c <= max_c; ++c) {
    const int tpc = (n_tiles + c - 1) / c;
    const int real_c = (n_tiles + tpc - 1) / tpc;
    const long long ctas = static_cast<long long>(units) * real_c;
    const long long waves = (ctas + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (tpc + 0.35) + chunk_cost * real_c;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best_c = real_c;
      best_tpc = tpc;
    }
  }
  *tiles_per_chunk = best_tpc;
  return best_c;
}
size_t fwd_pair_smem_bytes(int d);
int fwd_pair_pick_chunks(int m_rows, int n_cols, int num_sms, int* tiles_per_chunk);
cudaError_t launch_fwd_rowstats_pair(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows, int n_cols,
                                     int d, int chunks, int tiles_per_chunk, int m_pad, const float* scale_log2,
                                     float4* partial, float* dbg_z, int dbg_ld, cudaStream_t stream);
cudaError_t launch_fwd_rowstats_pair_ranks(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows,
                                           int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad,
                                           const float* scale_log2, float4* partial, const float* diag_z, int loc_lo,
                                           int loc_hi, int* rank_part, cudaStream_t stream);
size_t bwd_pair_smem_bytes(int d, int split);
int bwd_pair_pick_chunks(int m_rows, int n_cols, int d, int num_sms, int* tiles_per_chunk);
int bwd_pair_d_slices(int d);
cudaError_t launch_bwd_rows_pair(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, const CUtensorMap& tm_cols_mn,
                                 int m_rows, int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad, int diag0,
                                 const float* scale_log2, const float4* row_coef, const float4* col_coef,
                                 float* dx_partial, int split, cudaStream_t stream);

// ---- HBM-bound side passes (scl_aux.cu)
cudaError_t launch_cast_bf16(const void* x, int src_dtype, void* y, int rows, int d, int normalize,
                             cudaStream_t stream);
cudaError_t launch_prep_scalars(const float* logit_scale, float cap, float* scalars, cudaStream_t stream);
cudaError_t launch_build_positives(const int64_t* all_ids, int n_global, const int64_t* nbr_ids,
                                   const float* nbr_alpha, int b_local, int k, float alpha_scale, int rank,
                                   void* hash_ws, size_t hash_ws_bytes, int32_t* pos_col, float* pos_w, float* pos_q,
                                   const int* differ_flag, const int32_t* src_col, const float* src_w,
                                   const float* src_q, cudaStream_t stream);
cudaError_t launch_ids_differ(const int64_t* a, const int64_t* b, int n, int* flag, cudaStream_t stream);
size_t positives_hash_bytes(int n_global);
cudaError_t launch_row_finalize(const float4* partial, int n_slots, int m_pad, int m_rows, int d, const void* x_rows,
                                const void* y_all, const int32_t* pos_col, const float* pos_q, int kp1,
                                float4* row_stats, cudaStream_t stream);
cudaError_t launch_reduce_rows(const float4* stats_a, const float4* stats_b, int m_rows, const float* scalars,
                               float* sums6, float c, float w, float* out4, cudaStream_t stream);
cudaError_t launch_prepare(const void* image, const void* text, int src_dtype, int rows, int d, void* image_bf16,
                           void* text_bf16, const float* logit_scale, float cap, float* scalars, cudaStream_t stream);
cudaError_t launch_loss_scalars(const float* sums6, const float* scalars, float c, float w, float* out4,
                                cudaStream_t stream);
cudaError_t launch_bwd_coeffs(const float4* row_stats, int m_rows, int m_pad, const float4* col_stats, int n_cols,
                              int n_pad, int b_local, int rank, const float* gaps, const float* scalars,
                              const float* grad_out, float c, float w, float mult, int col_mode, const float* pos_q,
                              const float* opp_q_local, int kp1, float4* row_coef, float4* col_coef,
                              cudaStream_t stream);
cudaError_t launch_bwd_finish(const float* dx_partial, int chunks, int m_pad, int m_rows, int d, const void* y_all,
                              const int32_t* pos_col, const float* pos_q, int kp1, const int32_t* opp_col_all,
                              const float* opp_q_all, int n_global, int b_local, int rank, const float* gaps,
                              const float* scalars, const float* grad_out, float c, float w, float mult, int col_mode,
                              int split, void* workspace, size_t workspace_bytes, void* dx_out, int out_dtype,
                              cudaStream_t stream);
size_t bwd_finish_workspace_bytes(int n_global, int b_local, int kp1);
cudaError_t launch_check_positives(const int32_t* col_in, const float* q_in, int b_local, int kp1, int n_global,
                                   int rank, int32_t* col_out, float* q_out, int* flag, cudaStream_t stream);

// ---- fp32-accurate ("bf16x2") operand preparation (scl_split.cu)
cudaError_t launch_split_cast(const void* x, int src_dtype, void* rows_out, void* cols_out, int rows, int d,
                              cudaStream_t stream);

// ---- in-pass retrieval ranks (scl_rank.cu)
cudaError_t launch_retrieval_diag(const void* x_rows, int m_rows, const void* y_cols, int d, int first_col,
                                  float* diag_z, cudaStream_t stream);
cudaError_t launch_retrieval_rank_sum(const int* rank_part, int n_slots, int m_pad, int m_rows, int* ranks,
                                      cudaStream_t stream);

cudaError_t launch_unpack_records(const float* in, int world, int rec_floats, int n_comp, float* const* outs,
                                  const int* offs, const int* lens, cudaStream_t stream);

}  // namespace scl
