// Internal launcher declarations shared by the .cu files and the C-ABI layer (scl_api.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

namespace scl {

// ---- tensor-core kernels
size_t fwd_smem_bytes(int d);
int fwd_pick_chunks(int m_rows, int n_cols, int num_sms, int* tiles_per_chunk);
cudaError_t launch_fwd_rowstats(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows, int n_cols, int d,
                                int chunks, int tiles_per_chunk, int m_pad, const float* scale_log2, float4* partial,
                                float* dbg_z, int dbg_ld, cudaStream_t stream);

size_t bwd_smem_bytes();
void bwd_pick_split(int d, int* n_dsplit, int* dn);
int bwd_pick_chunks(int m_rows, int n_cols, int d, int num_sms, int* tiles_per_chunk);
cudaError_t launch_bwd_rows(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, const CUtensorMap& tm_cols_t,
                            int m_rows, int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad, int diag0,
                            const float* scale_log2, const float4* row_coef, const float4* col_coef,
                            float* dx_partial, cudaStream_t stream);

// ---- CTA-pair (cta_group::2) variants
// Column chunking for a grid of `units` row blocks (CTAs or CTA pairs) over `slots` concurrently resident
// units: choose the chunk count that minimises  waves * tiles_per_chunk  (the launch lasts `waves` rounds of
// the longest chunk), breaking ties towards fewer chunks (each chunk adds a partial-result slab).
inline int pick_chunks_balanced(int units, int n_tiles, int slots, int min_tiles, int* tiles_per_chunk) {
  int best_c = 1, best_tpc = n_tiles;
  double best_cost = 1e30;
  const int max_c = n_tiles / min_tiles > 1 ? n_tiles / min_tiles : 1;
  for (int c = 1; c <= max_c; ++c) {
    const int tpc = (n_tiles + c - 1) / c;
    const int real_c = (n_tiles + tpc - 1) / tpc;
    const long long ctas = static_cast<long long>(units) * real_c;
    const long long waves = (ctas + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (tpc + 0.35) + 0.02 * real_c;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best_c = real_c;
      best_tpc = tpc;
    }
  }
  *tiles_per_chunk = best_tpc;
  return best_c;
}
// Developer knob: SCL_FWD_CHUNKS / SCL_BWD_CHUNKS = c > 0 overrides the picker of the CTA-pair kernels with (about) c
// column chunks (plans and work-area sizes follow, they all go through the pickers).  Unset: the balanced choice.
inline int chunks_override(const char* env_name, int n_tiles, int* tiles_per_chunk) {
  const char* e = std::getenv(env_name);
  if (e == nullptr || e[0] == 0) return 0;
  const int c = std::atoi(e);
  if (c <= 0) return 0;
  const int cc = c < n_tiles ? c : n_tiles;
  const int tpc = (n_tiles + cc - 1) / cc;
  *tiles_per_chunk = tpc;
  return (n_tiles + tpc - 1) / tpc;
}
size_t fwd_pair_smem_bytes(int d);
int fwd_pair_pick_chunks(int m_rows, int n_cols, int num_sms, int* tiles_per_chunk);
cudaError_t launch_fwd_rowstats_pair(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows, int n_cols,
                                     int d, int chunks, int tiles_per_chunk, int m_pad, const float* scale_log2,
                                     float4* partial, float* dbg_z, int dbg_ld, long long* dbg_t,
                                     cudaStream_t stream);
cudaError_t launch_fwd_rowstats_pair_ranks(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows,
                                           int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad,
                                           const float* scale_log2, float4* partial, const float* diag_z, int loc_lo,
                                           int loc_hi, int* rank_part, cudaStream_t stream);
size_t bwd_pair_smem_bytes(int d, int split);
int bwd_pair_pick_chunks(int m_rows, int n_cols, int d, int num_sms, int* tiles_per_chunk);
int bwd_pair_d_slices(int d);
bool bwd_pair_mn_major();
cudaError_t launch_bwd_rows_pair(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, const CUtensorMap& tm_cols_t,
                                 int m_rows, int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad, int diag0,
                                 const float* scale_log2, const float4* row_coef, const float4* col_coef,
                                 float* dx_partial, long long* dbg_t, int split, cudaStream_t stream);

// ---- HBM-bound side passes (scl_aux.cu)
cudaError_t launch_cast_bf16(const void* x, int src_dtype, void* y, void* y_t, int rows, int d, int ld_t,
                             int normalize, cudaStream_t stream);
cudaError_t launch_prep_scalars(const float* logit_scale, float cap, float* scalars, cudaStream_t stream);
cudaError_t launch_build_positives(const int64_t* all_ids, int n_global, const int64_t* nbr_ids,
                                   const float* nbr_alpha, int b_local, int k, float alpha_scale, int rank,
                                   void* hash_ws, size_t hash_ws_bytes, int32_t* pos_col, float* pos_w, float* pos_q,
                                   cudaStream_t stream);
size_t positives_hash_bytes(int n_global);
cudaError_t launch_row_finalize(const float4* partial, int n_slots, int m_pad, int m_rows, int d, const void* x_rows,
                                const void* y_all, const int32_t* pos_col, const float* pos_q, int kp1,
                                float4* row_stats, cudaStream_t stream);
cudaError_t launch_reduce_rows(const float4* stats_a, const float4* stats_b, int m_rows, const float* scalars,
                               float* sums6, cudaStream_t stream);
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
                              int split, float* dx32, void* dx_out, int out_dtype, cudaStream_t stream);

// ---- fp32-accurate ("bf16x2") operand preparation (scl_split.cu)
cudaError_t launch_split_cast(const void* x, int src_dtype, void* rows_out, void* cols_out, int rows, int d,
                              cudaStream_t stream);
cudaError_t launch_transpose_split(const void* cols_all, int n_rows, int d, int ld_t, void* out_t, cudaStream_t stream);

// ---- in-pass retrieval ranks (scl_rank.cu)
cudaError_t launch_retrieval_diag(const void* x_rows, int m_rows, const void* y_cols, int d, int first_col,
                                  float* diag_z, cudaStream_t stream);
cudaError_t launch_retrieval_rank_sum(const int* rank_part, int n_slots, int m_pad, int m_rows, int* ranks,
                                      cudaStream_t stream);

cudaError_t launch_unpack_records(const float* in, int world, int rec_floats, int n_comp, float* const* outs,
                                  const int* offs, const int* lens, cudaStream_t stream);

}  // namespace scl
