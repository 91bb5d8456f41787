// extern "C" boundary of libscl_b200.so (see include/scl_b200.h).  Plain pointers and sizes only,
// no torch types, no allocation, no stream synchronisation, no mutable global state.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/scl_b200.h"
#include "scl_kernels.h"

namespace {

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? SCL_OK : -1000 - static_cast<int>(e); }

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  // resolved through the runtime so the library has no link-time dependency on libcuda
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return static_cast<EncodeTiledFn>(nullptr);
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, inner] matrix with `ld` elements between rows;
// box = {64 inner elements (128 B, one swizzle row), box_rows}, SWIZZLE_128B, zero fill out of bounds.
int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return SCL_ERR_NO_DRIVER_ENTRY;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0) return SCL_ERR_INVALID_ARG;
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SCL_OK : SCL_ERR_TENSOR_MAP;
}

int num_sms_or_default() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 148;
  }
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    return 148;
  }
  return sms;
}

// Output width D of the backward: D % 64 == 0 up to 512 (one pass), or up to 1536 with an equal cut into slices of at
// most 512 columns (640, 768, 1024, 1152, 1280, 1536: one pass per slice).
bool shape_ok(int m_rows, int n_cols, int d) {
  if (m_rows < 1 || n_cols < 1 || d < 64 || d % 64 != 0) return false;
  return d <= 512 || (d <= 1536 && scl::bwd_pair_d_slices(d) > 0);
}
// Contraction length of the forward pass: as above, plus the K-concatenated operands of the fp32-accurate mode
// (3 D, so up to 4608) -- beyond 512 the kernel streams X with Y and any multiple of 64 works.
bool fwd_shape_ok(int m_rows, int n_cols, int d) {
  return m_rows >= 1 && n_cols >= 1 && d >= 64 && d % 64 == 0 && d <= 4608;
}

}  // namespace

extern "C" {

int scl_abi_version(void) { return SCL_ABI_VERSION; }

const char* scl_error_string(int code) {
  switch (code) {
    case SCL_OK: return "ok";
    case SCL_ERR_INVALID_ARG: return "invalid argument (null / misaligned pointer or bad size)";
    case SCL_ERR_UNSUPPORTED_SHAPE:
      return "unsupported shape (need rows >= 1 and D % 64 == 0 up to 512, or up to 1536 with an equal cut into slices "
             "of at most 512 columns, each a multiple of 64)";
    case SCL_ERR_NO_DRIVER_ENTRY: return "cuTensorMapEncodeTiled not available from the CUDA driver";
    case SCL_ERR_TENSOR_MAP: return "cuTensorMapEncodeTiled rejected the tensor map";
    case SCL_ERR_NOT_SM100: return "device is not compute capability 10.x (kernels are sm_100a only)";
    default: break;
  }
  if (code <= -1000) return cudaGetErrorString(static_cast<cudaError_t>(-code - 1000));
  return "unknown scl error";
}

int scl_check_device(int* num_sms) {
  int dev = 0, major = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_rc(e);
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_rc(e);
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_rc(e);
  if (num_sms) *num_sms = sms;
  return major == 10 ? SCL_OK : SCL_ERR_NOT_SM100;
}

int scl_fwd_plan(int m_rows, int n_cols, int d, scl_plan* plan) {
  if (plan == nullptr) return SCL_ERR_INVALID_ARG;
  if (!fwd_shape_ok(m_rows, n_cols, d)) return SCL_ERR_UNSUPPORTED_SHAPE;
  int tpc = 0;
  plan->split = 0;
  plan->chunks = scl::fwd_pair_pick_chunks(m_rows, n_cols, num_sms_or_default(), &tpc);
  plan->m_pad = (m_rows + 255) / 256 * 256;
  plan->tiles_per_chunk = tpc;
  plan->n_slots = 4 * plan->chunks;
  plan->n_pad = (n_cols + 255) / 256 * 256;
  plan->d_split = 1;
  return SCL_OK;
}

int scl_bwd_plan(int m_rows, int n_cols, int d, int split, scl_plan* plan) {
  if (plan == nullptr) return SCL_ERR_INVALID_ARG;
  if (!shape_ok(m_rows, n_cols, d)) return SCL_ERR_UNSUPPORTED_SHAPE;
  int tpc = 0;
  plan->split = split ? 1 : 0;
  plan->m_pad = (m_rows + 127) / 128 * 128;
  plan->n_slots = 0;
  plan->chunks = scl::bwd_pair_pick_chunks(m_rows, n_cols, d, num_sms_or_default(), &tpc);
  plan->n_pad = (n_cols + 255) / 256 * 256;
  plan->d_split = scl::bwd_pair_d_slices(d);
  plan->tiles_per_chunk = tpc;
  return SCL_OK;
}

int scl_cast_bf16(const void* x, int src_dtype, void* y, int rows, int d, int normalize, void* stream) {
  if (x == nullptr || y == nullptr || rows < 0 || d <= 0 || d % 64 != 0 || src_dtype < 0 || src_dtype > 2 ||
      (reinterpret_cast<uintptr_t>(x) & 15) != 0 || (reinterpret_cast<uintptr_t>(y) & 15) != 0)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_cast_bf16(x, src_dtype, y, rows, d, normalize, static_cast<cudaStream_t>(stream)));
}

int scl_prep_scalars(const float* logit_scale, float cap, float* scalars3, void* stream) {
  if (logit_scale == nullptr || scalars3 == nullptr) return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_prep_scalars(logit_scale, cap, scalars3, static_cast<cudaStream_t>(stream)));
}

size_t scl_positives_workspace_bytes(int n_global) { return scl::positives_hash_bytes(n_global); }

int scl_build_positives(const int64_t* all_ids, int n_global, const int64_t* nbr_ids, const float* nbr_alpha,
                        int b_local, int k, float alpha_scale, int rank, void* workspace, size_t workspace_bytes,
                        int32_t* pos_col, float* pos_w, float* pos_q, void* stream) {
  if (b_local < 1 || n_global < b_local || k < 0 || rank < 0 || pos_col == nullptr || pos_w == nullptr ||
      pos_q == nullptr)
    return SCL_ERR_INVALID_ARG;
  if (k > 0 && (all_ids == nullptr || nbr_ids == nullptr || nbr_alpha == nullptr || workspace == nullptr))
    return SCL_ERR_INVALID_ARG;
  if (k > 31) return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_build_positives(all_ids, n_global, nbr_ids, nbr_alpha, b_local, k, alpha_scale, rank,
                                             workspace, workspace_bytes, pos_col, pos_w, pos_q, nullptr, nullptr,
                                             nullptr, nullptr, static_cast<cudaStream_t>(stream)));
}

int scl_fwd_rowstats(const void* x_rows, int m_rows, const void* y_cols, int n_cols, int d, const float* scalars3,
                     const scl_plan* plan, void* partial, float* dbg_z, int dbg_ld, void* stream) {
  if (x_rows == nullptr || y_cols == nullptr || scalars3 == nullptr || plan == nullptr || partial == nullptr)
    return SCL_ERR_INVALID_ARG;
  if (!fwd_shape_ok(m_rows, n_cols, d)) return SCL_ERR_UNSUPPORTED_SHAPE;
  CUtensorMap tm_rows, tm_cols;
  int rc = make_map(&tm_rows, x_rows, d, m_rows, d, 128);
  if (rc != SCL_OK) return rc;
  rc = make_map(&tm_cols, y_cols, d, n_cols, d, 128);
  if (rc != SCL_OK) return rc;
  return cuda_rc(scl::launch_fwd_rowstats_pair(tm_rows, tm_cols, m_rows, n_cols, d, plan->chunks,
                                               plan->tiles_per_chunk, plan->m_pad, scalars3 + 1,
                                               static_cast<float4*>(partial), dbg_z, dbg_ld,
                                               static_cast<cudaStream_t>(stream)));
}

int scl_fwd_rowstats_ranks(const void* x_rows, int m_rows, const void* y_cols, int n_cols, int d, const float* scalars3,
                           const scl_plan* plan, void* partial, int first_col, float* diag_z, int32_t* rank_partial,
                           int32_t* ranks, void* stream) {
  if (x_rows == nullptr || y_cols == nullptr || scalars3 == nullptr || plan == nullptr || partial == nullptr ||
      diag_z == nullptr || rank_partial == nullptr || ranks == nullptr || first_col < 0 || first_col + m_rows > n_cols)
    return SCL_ERR_INVALID_ARG;
  if (!fwd_shape_ok(m_rows, n_cols, d)) return SCL_ERR_UNSUPPORTED_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = cuda_rc(scl::launch_retrieval_diag(x_rows, m_rows, y_cols, d, first_col, diag_z, st));
  if (rc != SCL_OK) return rc;
  CUtensorMap tm_rows, tm_cols;
  rc = make_map(&tm_rows, x_rows, d, m_rows, d, 128);
  if (rc != SCL_OK) return rc;
  rc = make_map(&tm_cols, y_cols, d, n_cols, d, 128);
  if (rc != SCL_OK) return rc;
  rc = cuda_rc(scl::launch_fwd_rowstats_pair_ranks(tm_rows, tm_cols, m_rows, n_cols, d, plan->chunks,
                                                   plan->tiles_per_chunk, plan->m_pad, scalars3 + 1,
                                                   static_cast<float4*>(partial), diag_z, first_col, first_col + m_rows,
                                                   rank_partial, st));
  if (rc != SCL_OK) return rc;
  return cuda_rc(scl::launch_retrieval_rank_sum(rank_partial, plan->n_slots, plan->m_pad, m_rows, ranks, st));
}

int scl_row_finalize(const void* partial, const scl_plan* plan, int m_rows, int d, const void* x_rows,
                     const void* y_all, const int32_t* pos_col, const float* pos_q, int k_plus_1, void* row_stats,
                     void* stream) {
  if (partial == nullptr || plan == nullptr || x_rows == nullptr || y_all == nullptr || pos_col == nullptr ||
      pos_q == nullptr || row_stats == nullptr || k_plus_1 < 1)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_row_finalize(static_cast<const float4*>(partial), plan->n_slots, plan->m_pad, m_rows, d,
                                          x_rows, y_all, pos_col, pos_q, k_plus_1, static_cast<float4*>(row_stats),
                                          static_cast<cudaStream_t>(stream)));
}

int scl_check_positives(const int32_t* col_in, const float* q_in, int b_local, int k_plus_1, int n_global, int rank,
                        int32_t* col_out, float* q_out, int* flag, void* stream) {
  if (col_in == nullptr || q_in == nullptr || col_out == nullptr || q_out == nullptr || flag == nullptr ||
      b_local < 1 || k_plus_1 < 1 || n_global < b_local || rank < 0)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_check_positives(col_in, q_in, b_local, k_plus_1, n_global, rank, col_out, q_out, flag,
                                             static_cast<cudaStream_t>(stream)));
}

int scl_reduce_rows(const void* stats_img, const void* stats_txt, int m_rows, const float* scalars3, float* sums6,
                    void* stream) {
  if (stats_img == nullptr || stats_txt == nullptr || scalars3 == nullptr || sums6 == nullptr || m_rows < 1)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_reduce_rows(static_cast<const float4*>(stats_img), static_cast<const float4*>(stats_txt),
                                         m_rows, scalars3, sums6, 0.f, 0.f, nullptr, static_cast<cudaStream_t>(stream)));
}

int scl_loss_scalars(const float* sums6, const float* scalars3, float c, float w, float* out4, void* stream) {
  if (sums6 == nullptr || scalars3 == nullptr || out4 == nullptr) return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_loss_scalars(sums6, scalars3, c, w, out4, static_cast<cudaStream_t>(stream)));
}

int scl_bwd_coeffs(const void* row_stats, int m_rows, const void* col_stats, int n_cols, const scl_plan* plan,
                   int b_local, int rank, const float* gaps, const float* scalars3, const float* grad_out, float c,
                   float w, float mult, int col_mode, const float* pos_q, const float* opp_q_local, int k_plus_1,
                   void* row_coef, void* col_coef, void* stream) {
  if (row_stats == nullptr || col_stats == nullptr || plan == nullptr || gaps == nullptr || scalars3 == nullptr ||
      grad_out == nullptr || row_coef == nullptr || col_coef == nullptr || b_local < 1 || col_mode < 0 ||
      col_mode > 2 || pos_q == nullptr || opp_q_local == nullptr || k_plus_1 < 1 || m_rows > b_local)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_bwd_coeffs(static_cast<const float4*>(row_stats), m_rows, plan->m_pad,
                                        static_cast<const float4*>(col_stats), n_cols, plan->n_pad, b_local, rank, gaps,
                                        scalars3, grad_out, c, w, mult, col_mode, pos_q, opp_q_local, k_plus_1,
                                        static_cast<float4*>(row_coef), static_cast<float4*>(col_coef),
                                        static_cast<cudaStream_t>(stream)));
}

int scl_bwd_rows(const void* x_rows, int m_rows, const void* y_cols, int n_cols, int d, int diag_col0,
                 const float* scalars3, const scl_plan* plan, const void* row_coef, const void* col_coef,
                 float* dx_partial, void* stream) {
  if (x_rows == nullptr || y_cols == nullptr || scalars3 == nullptr || plan == nullptr || row_coef == nullptr ||
      col_coef == nullptr || dx_partial == nullptr)
    return SCL_ERR_INVALID_ARG;
  if (!shape_ok(m_rows, n_cols, d)) return SCL_ERR_UNSUPPORTED_SHAPE;
  // fp32-accurate mode: x_rows / y_cols are the K-concatenated operands [., 3 d]
  const int kd = plan->split ? 3 * d : d;
  CUtensorMap tm_rows, tm_cols, tm_cols_mn;
  int rc = make_map(&tm_rows, x_rows, kd, m_rows, kd, 64);
  if (rc != SCL_OK) return rc;
  rc = make_map(&tm_cols, y_cols, kd, n_cols, kd, 128);
  if (rc != SCL_OK) return rc;
  rc = make_map(&tm_cols_mn, y_cols, kd, n_cols, kd, 64);  // the same matrix in {64 d, 64 j} boxes (MN-major operand)
  if (rc != SCL_OK) return rc;
  return cuda_rc(scl::launch_bwd_rows_pair(tm_rows, tm_cols, tm_cols_mn, m_rows, n_cols, d, plan->chunks,
                                           plan->tiles_per_chunk, plan->m_pad, diag_col0, scalars3 + 1,
                                           static_cast<const float4*>(row_coef), static_cast<const float4*>(col_coef),
                                           dx_partial, plan->split, static_cast<cudaStream_t>(stream)));
}

size_t scl_bwd_finish_workspace_bytes(int n_global, int b_local, int k_plus_1) {
  return scl::bwd_finish_workspace_bytes(n_global, b_local, k_plus_1);
}

int scl_bwd_finish(const float* dx_partial, const scl_plan* plan, int m_rows, int d, const void* y_all,
                   const int32_t* pos_col, const float* pos_q, int k_plus_1, const int32_t* opp_col_all,
                   const float* opp_q_all, int n_global, int b_local, int rank, const float* gaps,
                   const float* scalars3, const float* grad_out, float c, float w, float mult, int col_mode,
                   void* workspace, size_t workspace_bytes, void* dx_out, int out_dtype, void* stream) {
  if (dx_partial == nullptr || plan == nullptr || y_all == nullptr || pos_col == nullptr || pos_q == nullptr ||
      gaps == nullptr || scalars3 == nullptr || grad_out == nullptr || dx_out == nullptr || k_plus_1 < 1 ||
      k_plus_1 > 32 || out_dtype < 0 || out_dtype > 2 ||
      (col_mode != 0 && (opp_col_all == nullptr || opp_q_all == nullptr || workspace == nullptr ||
                         workspace_bytes < scl::bwd_finish_workspace_bytes(n_global, b_local, k_plus_1))))
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_bwd_finish(dx_partial, plan->chunks, plan->m_pad, m_rows, d, y_all, pos_col, pos_q,
                                        k_plus_1, opp_col_all, opp_q_all, n_global, b_local, rank, gaps, scalars3,
                                        grad_out, c, w, mult, col_mode, plan->split, workspace, workspace_bytes, dx_out,
                                        out_dtype, static_cast<cudaStream_t>(stream)));
}

namespace {
inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
}  // namespace

int scl_prepare(const scl_prepare_args* a, void* stream) {
  if (a == nullptr || a->image == nullptr || a->text == nullptr || a->image_bf16 == nullptr ||
      a->text_bf16 == nullptr || a->scalars3 == nullptr || a->logit_scale == nullptr || a->rows < 1 || a->d < 64 ||
      a->d % 64 != 0 || a->src_dtype < 0 || a->src_dtype > 2 ||
      ((reinterpret_cast<uintptr_t>(a->image) | reinterpret_cast<uintptr_t>(a->text) |
        reinterpret_cast<uintptr_t>(a->image_bf16) | reinterpret_cast<uintptr_t>(a->text_bf16)) & 15) != 0)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_prepare(a->image, a->text, a->src_dtype, a->rows, a->d, a->image_bf16, a->text_bf16,
                                     a->logit_scale, a->cap, a->scalars3, static_cast<cudaStream_t>(stream)));
}

size_t scl_fwd_workspace_bytes(int b_local, int n_global, int d, int k) {
  scl_plan p;
  if (scl_fwd_plan(b_local, n_global, d, &p) != SCL_OK) return 0;
  const size_t partial = align256(static_cast<size_t>(p.n_slots) * p.m_pad * 16);
  const size_t hash = k > 0 ? align256(scl::positives_hash_bytes(n_global)) + 256 : 0;  // + the ids-differ flag
  // + the retrieval-rank work areas (rank partials, own-pair similarities); small, so always included
  const size_t ranks = align256(static_cast<size_t>(p.n_slots) * p.m_pad * 4) + align256(static_cast<size_t>(p.m_pad) * 4);
  return 2 * partial + hash + ranks;
}

int scl_fwd_all(const scl_fwd_args* a, void* stream) {
  if (a == nullptr || a->workspace == nullptr) return SCL_ERR_INVALID_ARG;
  scl_plan p;
  int rc = scl_fwd_plan(a->b_local, a->n_global, a->d, &p);
  if (rc != SCL_OK) return rc;
  if (a->workspace_bytes < scl_fwd_workspace_bytes(a->b_local, a->n_global, a->d, a->k))
    return SCL_ERR_INVALID_ARG;
  const size_t partial_bytes = align256(static_cast<size_t>(p.n_slots) * p.m_pad * 16);
  char* ws = static_cast<char*>(a->workspace);
  void* part_i = ws;
  void* part_t = ws + partial_bytes;
  void* hash = ws + 2 * partial_bytes;
  const size_t hash_bytes = a->k > 0 ? scl::positives_hash_bytes(a->n_global) : 0;
  char* rank_ws = ws + 2 * partial_bytes + (a->k > 0 ? align256(hash_bytes) + 256 : 0);
  int32_t* rank_partial = reinterpret_cast<int32_t*>(rank_ws);
  float* diag_z = reinterpret_cast<float*>(rank_ws + align256(static_cast<size_t>(p.n_slots) * p.m_pad * 4));
  // phases (0 = everything): 1 soft targets (needs the gathered ids), 2 image-rows pass (needs txt_all),
  // 4 text-rows pass + reductions (needs img_all) -- separate calls let the caller overlap the exchanges
  const int phases = a->phases == 0 ? 7 : a->phases;
  if (phases & 1) {
    // soft targets: image rows use the text-id map, text rows the image-id map (losses.py:102-108)
    rc = scl_build_positives(a->txt_ids_all, a->n_global, a->nbr_ids, a->nbr_alpha, a->b_local, a->k, a->alpha_scale,
                             a->rank, a->k > 0 ? hash : nullptr, hash_bytes, a->col_it, a->w_it, a->q_it, stream);
    if (rc != SCL_OK) return rc;
    if (!a->same_ids && a->k > 0) {
      // text rows resolve their neighbours in the IMAGE id map.  The reference's loader makes the two id vectors
      // equal (as separate tensors), so a device flag decides: identical vectors -> copy the lists just built
      // (the hash kernels return at once), otherwise build them.  No host synchronisation either way.
      if (a->img_ids_all == nullptr || a->txt_ids_all == nullptr) return SCL_ERR_INVALID_ARG;
      int* differ = reinterpret_cast<int*>(static_cast<char*>(hash) + align256(hash_bytes));
      rc = cuda_rc(scl::launch_ids_differ(a->img_ids_all, a->txt_ids_all, a->n_global, differ,
                                          static_cast<cudaStream_t>(stream)));
      if (rc != SCL_OK) return rc;
      rc = cuda_rc(scl::launch_build_positives(a->img_ids_all, a->n_global, a->nbr_ids, a->nbr_alpha, a->b_local, a->k,
                                               a->alpha_scale, a->rank, hash, hash_bytes, a->col_ti, a->w_ti, a->q_ti,
                                               differ, a->col_it, a->w_it, a->q_it, static_cast<cudaStream_t>(stream)));
      if (rc != SCL_OK) return rc;
    }
  }
  if (phases & 2) {
    if (a->ranks_out != nullptr)  // image -> gene retrieval ranks within the local block, counted in the same pass
      rc = scl_fwd_rowstats_ranks(a->img_l, a->b_local, a->txt_all, a->n_global, a->d, a->scalars3, &p, part_i,
                                  a->rank * a->b_local, diag_z, rank_partial, a->ranks_out, stream);
    else
      rc = scl_fwd_rowstats(a->img_l, a->b_local, a->txt_all, a->n_global, a->d, a->scalars3, &p, part_i, nullptr, 0,
                            stream);
    if (rc != SCL_OK) return rc;
    rc = scl_row_finalize(part_i, &p, a->b_local, a->d, a->img_l, a->txt_all, a->col_it, a->q_it, a->k + 1,
                          a->stats_i, stream);
    if (rc != SCL_OK) return rc;
  }
  if (!(phases & 4)) return SCL_OK;
  rc = scl_fwd_rowstats(a->txt_l, a->b_local, a->img_all, a->n_global, a->d, a->scalars3, &p, part_t, nullptr, 0,
                        stream);
  if (rc != SCL_OK) return rc;
  rc = scl_row_finalize(part_t, &p, a->b_local, a->d, a->txt_l, a->img_all, a->col_ti, a->q_ti, a->k + 1, a->stats_t,
                        stream);
  if (rc != SCL_OK) return rc;
  if (a->stats_i == nullptr || a->stats_t == nullptr || a->sums6 == nullptr || (a->finalize_scalars && a->out4 == nullptr))
    return SCL_ERR_INVALID_ARG;
  // row reductions (+ the loss scalars in the same launch)
  return cuda_rc(scl::launch_reduce_rows(static_cast<const float4*>(a->stats_i), static_cast<const float4*>(a->stats_t),
                                         a->b_local, a->scalars3, a->sums6, a->c, a->w,
                                         a->finalize_scalars ? a->out4 : nullptr, static_cast<cudaStream_t>(stream)));
}

size_t scl_bwd_workspace_bytes(int b_local, int n_global, int d, int k_plus_1) {
  scl_plan p;
  if (scl_bwd_plan(b_local, n_global, d, 0, &p) != SCL_OK) return 0;
  return align256(static_cast<size_t>(p.m_pad) * 16) + align256(static_cast<size_t>(p.n_pad) * 16) +
         align256(static_cast<size_t>(p.chunks) * p.m_pad * d * 4) +
         align256(scl::bwd_finish_workspace_bytes(n_global, b_local, k_plus_1));
}

int scl_bwd_dir(const scl_bwd_args* a, void* stream) {
  if (a == nullptr || a->workspace == nullptr || a->dx_out == nullptr) return SCL_ERR_INVALID_ARG;
  scl_plan p;
  int rc = scl_bwd_plan(a->b_local, a->n_global, a->d, a->split, &p);
  if (rc != SCL_OK) return rc;
  if (a->workspace_bytes < scl_bwd_workspace_bytes(a->b_local, a->n_global, a->d, a->k_plus_1))
    return SCL_ERR_INVALID_ARG;
  char* ws = static_cast<char*>(a->workspace);
  void* row_coef = ws;
  ws += align256(static_cast<size_t>(p.m_pad) * 16);
  void* col_coef = ws;
  ws += align256(static_cast<size_t>(p.n_pad) * 16);
  float* partial = reinterpret_cast<float*>(ws);
  ws += align256(static_cast<size_t>(p.chunks) * p.m_pad * a->d * 4);
  const size_t fin_bytes = scl::bwd_finish_workspace_bytes(a->n_global, a->b_local, a->k_plus_1);
  rc = scl_bwd_coeffs(a->row_stats, a->b_local, a->col_stats_all, a->n_global, &p, a->b_local, a->rank, a->gaps,
                      a->scalars3, a->grad_out, a->c, a->w, a->mult, a->col_mode, a->pos_q, a->opp_q_local,
                      a->k_plus_1, row_coef, col_coef, stream);
  if (rc != SCL_OK) return rc;
  rc = scl_bwd_rows(a->x_rows, a->b_local, a->y_all, a->n_global, a->d, a->rank * a->b_local, a->scalars3, &p,
                    row_coef, col_coef, partial, stream);
  if (rc != SCL_OK) return rc;
  return scl_bwd_finish(partial, &p, a->b_local, a->d, a->y_all, a->pos_col, a->pos_q, a->k_plus_1, a->opp_col_all,
                        a->opp_q_all, a->n_global, a->b_local, a->rank, a->gaps, a->scalars3, a->grad_out, a->c, a->w,
                        a->mult, a->col_mode, ws, fin_bytes, a->dx_out, a->out_dtype, stream);
}

int scl_split_bf16(const void* x, int src_dtype, void* rows_out, void* cols_out, int rows, int d, void* stream) {
  if (x == nullptr || (rows_out == nullptr && cols_out == nullptr) || rows < 0 || d <= 0 || d % 64 != 0 ||
      src_dtype < 0 || src_dtype > 2 || (reinterpret_cast<uintptr_t>(x) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(rows_out) & 15) != 0 || (reinterpret_cast<uintptr_t>(cols_out) & 15) != 0)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_split_cast(x, src_dtype, rows_out, cols_out, rows, d, static_cast<cudaStream_t>(stream)));
}

int scl_unpack_records(const float* gathered, int world, int rec_floats, int n_comp, float* const* outs,
                       const int* offs, const int* lens, void* stream) {
  if (gathered == nullptr || outs == nullptr || offs == nullptr || lens == nullptr || world < 1 || rec_floats < 1)
    return SCL_ERR_INVALID_ARG;
  return cuda_rc(scl::launch_unpack_records(gathered, world, rec_floats, n_comp, outs, offs, lens,
                                            static_cast<cudaStream_t>(stream)));
}

}  // extern "C"
