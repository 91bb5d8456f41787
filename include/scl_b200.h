/*
 * scl_b200.h -- C ABI of the B200-native contrastive-loss hot path (libscl_b200.so).
 *
 * The reference (Biogod2020/Spatial-Clip) is pure Python: it has no FFI of its own.  The boundary that
 * a maintainer binds is its loss-module API,
 *     SpatialLoss.forward   /root/reference/src/models/components/losses.py:44-124
 *     ClipLoss.forward      /root/reference/src/models/components/losses.py:131-141
 *                           /root/reference/src/open_clip/loss.py:132-155
 *     gather_features       /root/reference/src/open_clip/loss.py:21-65
 * and every entry point below replaces one group of torch ops inside those functions (cited per
 * function).  spatial_clip_b200/losses.py binds these with ctypes and re-assembles the two forwards
 * with the reference's exact parameter names; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless named host_*; `stream` is a cudaStream_t passed as void*
 *   - no function allocates, synchronises the stream, or keeps mutable global state; work areas are
 *     caller-provided (sizes from the *_plan / *_bytes queries)
 *   - kernels launch on the calling thread's current CUDA device
 *   - return 0 on success; SCL_ERR_* (<0) on bad arguments / unsupported shapes;
 *     -1000 - cudaError_t on a CUDA runtime failure; scl_error_string() decodes all of them
 *   - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16
 *   - feature matrices are row-major [rows, D]; D % 64 == 0 up to 512, or D % 64 == 0 up to 1536 with an equal cut
 *     into slices of at most 512 columns (640, 768, 1024, 1152, 1280, 1536: the backward then runs one pass per
 *     D slice and recomputes the similarity tile per slice)
 *   - the tensor-core kernels run on CTA pairs (tcgen05 cta_group::2); no transposed operand copies exist: the
 *     gradient GEMM reads the row-major column operand as an MN-major UMMA operand
 *   - fp32-accurate mode ("split"): operands are bf16 hi/lo pairs laid out by
 *     scl_split_bf16 as K-concatenated rows of width 3 D -- (h|h|l) for row operands, (h|l|h) for column operands --
 *     so that the same bf16 tensor-core kernels contract xh.yh + xh.yl + xl.yh; the forward entry points are then
 *     simply called with d = 3 D, the backward ones with plan.split = 1 and d = D
 */
#ifndef SCL_B200_H_
#define SCL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCL_ABI_VERSION 7
#define SCL_OK 0
#define SCL_ERR_INVALID_ARG (-1)
#define SCL_ERR_UNSUPPORTED_SHAPE (-2)
#define SCL_ERR_NO_DRIVER_ENTRY (-3)
#define SCL_ERR_TENSOR_MAP (-4)
#define SCL_ERR_NOT_SM100 (-5)

int scl_abi_version(void);
const char* scl_error_string(int code);
/* compute capability must be 10.x; returns SCL_ERR_NOT_SM100 otherwise */
int scl_check_device(int* num_sms);

/* ---- plans: tile/chunk decomposition and padded extents, pure host arithmetic --------------------- */
typedef struct scl_plan {
  int chunks;          /* column chunks (grid.y for fwd, grid.z for bwd)            */
  int tiles_per_chunk; /* column tiles per chunk                                    */
  int n_slots;         /* fwd: partial-statistics slots per row (2 or 4 per chunk)  */
  int m_pad;           /* rows padded to the 128-row MMA tile                       */
  int n_pad;           /* columns padded to the column tile                         */
  int d_split;         /* bwd: number of D slices (D > 512)                                          */
  int split;           /* bwd: 1 = fp32-accurate mode (operands are bf16 hi/lo pairs, see above)      */
} scl_plan;
int scl_fwd_plan(int m_rows, int n_cols, int d, scl_plan* plan);
/* split != 0: the fp32-accurate mode */
int scl_bwd_plan(int m_rows, int n_cols, int d, int split, scl_plan* plan);

/* ---- HBM-bound producer pass --------------------------------------------------------------------
 * y[rows,d] (bf16) from x; normalize != 0 applies F.normalize(x, dim=-1) first
 * (/root/reference/src/open_clip/model.py:326-345). */
int scl_cast_bf16(const void* x, int src_dtype, void* y, int rows, int d, int normalize, void* stream);

/* fp32-accurate mode: x[rows,d] (any dtype code) -> bf16 pairs h = bf16(x), l = bf16(x - h), written as
 * rows_out[rows, 3d] = (h|h|l) and/or cols_out[rows, 3d] = (h|l|h) (either may be NULL).  Replaces the same
 * reference lines as scl_cast_bf16 when float32_logits-grade accuracy (loss rel 1e-5) is wanted. */
int scl_split_bf16(const void* x, int src_dtype, void* rows_out, void* cols_out, int rows, int d, void* stream);

/* scalars[0] = min(logit_scale, cap) (cap <= 0: no cap), [1] = scalars[0]*log2(e), [2] = logit_scale
 * -- forward value of the straight-through cap, losses.py:73-76 */
int scl_prep_scalars(const float* logit_scale, float cap, float* scalars3, void* stream);

/* ---- soft-target builder (integer path; replaces the dict + Python double loop, losses.py:91-111) ----
 * all_ids[n_global]: gathered tile ids of the COLUMN modality; nbr_ids/nbr_alpha [b_local, k].
 * Outputs are ELL lists [b_local, k+1]: pos_col (global column or -1), pos_w (un-normalised fp32
 * weight, bit-exact vs the reference's dense labels), pos_q (pos_w / max(row sum, 1e-12)).
 * k == 0 (plain ClipLoss, loss.py:91-102) needs no ids and no work area. */
size_t scl_positives_workspace_bytes(int n_global);
int scl_build_positives(const int64_t* all_ids, int n_global, const int64_t* nbr_ids, const float* nbr_alpha,
                        int b_local, int k, float alpha_scale, int rank, void* workspace, size_t workspace_bytes,
                        int32_t* pos_col, float* pos_w, float* pos_q, void* stream);

/* ---- fused similarity GEMM + online row log-sum-exp (tcgen05) ------------------------------------
 * x_rows[m_rows,d] x y_cols[n_cols,d]^T, both bf16. partial is float4[plan.n_slots * plan.m_pad].
 * Replaces matmul + logit_scale multiply + log_softmax/softmax passes: losses.py:78-89,113-121,
 * loss.py:117-124,150-153.  dbg_z (optional, float[m_rows, dbg_ld]) receives the raw similarities. */
int scl_fwd_rowstats(const void* x_rows, int m_rows, const void* y_cols, int n_cols, int d, const float* scalars3,
                     const scl_plan* plan, void* partial, float* dbg_z, int dbg_ld, void* stream);

/* ---- the same pass + in-pass retrieval ranks -------------------------------------------------
 * ranks[i] = number of columns j in the LOCAL block [first_col, first_col + m_rows), j != first_col + i, whose
 * similarity with row i exceeds that of the row's own pair: the position of the matching profile in the in-batch
 * image -> gene retrieval.  Replaces the [B_l, B_l] logits matmul + topk the LightningModule runs every step only to
 * find that position (/root/reference/src/models/spatial_clip_module.py:68,106,113,127;
 * src/models/components/metrics.py:22-36): Recall@k = mean(ranks < k).  diag_z (float[m_rows]) and rank_partial
 * (int32[plan.n_slots * plan.m_pad]) are work areas. */
int scl_fwd_rowstats_ranks(const void* x_rows, int m_rows, const void* y_cols, int n_cols, int d, const float* scalars3,
                           const scl_plan* plan, void* partial, int first_col, float* diag_z, int32_t* rank_partial,
                           int32_t* ranks, void* stream);

/* merge partials and add the positive logits: row_stats[i] = {LSE_i/ln2, E_p[z], Var_p[z], sum_k q_k z_ik} */
int scl_row_finalize(const void* partial, const scl_plan* plan, int m_rows, int d, const void* x_rows,
                     const void* y_all, const int32_t* pos_col, const float* pos_q, int k_plus_1, void* row_stats,
                     void* stream);
/* caller-resolved soft targets (SpatialLossFromColumns; producer: spatial_clip_b200/positives.py, replacing
 * losses.py:91-111 on the data side): copy col_in / q_in [b_local, k_plus_1] to col_out / q_out with out-of-range
 * columns turned into unused slots and unused slots given q = 0, and OR into *flag (int, caller-zeroed):
 * bit 0 = a column outside [-1, n_global), bit 1 = slot 0 is not the row's own column rank*b_local + i,
 * bit 2 = an unused slot carried weight.  The kernels below index y_all + col*d for every col >= 0 and assume
 * slot 0 == own column, so lists from outside the library must pass through this call. */
int scl_check_positives(const int32_t* col_in, const float* q_in, int b_local, int k_plus_1, int n_global, int rank,
                        int32_t* col_out, float* q_out, int* flag, void* stream);
/* sums6 = per-rank sums feeding the loss / gap / d-scale (deterministic single-CTA tree) */
int scl_reduce_rows(const void* stats_img, const void* stats_txt, int m_rows, const float* scalars3, float* sums6,
                    void* stream);
/* out4 = {loss, gap, dloss/dlogit_scale, 2*w*gap}; c = 0.5 / rows-in-the-mean, w = temp_reg_weight
 * (losses.py:113-122) */
int scl_loss_scalars(const float* sums6, const float* scalars3, float c, float w, float* out4, void* stream);

/* ---- backward ---------------------------------------------------------------------------------- */
/* coefficient vectors of dL/dz = P(u_i + v_i z) + Pc(u'_j + v'_j z)  (SURVEY.md 8a closed forms);
 * col_mode: 0 = no column-direction terms, 1 = only this rank's columns, 2 = all columns.
 * pos_q / opp_q_local [b_local, k_plus_1]: the local rows' soft-target weights in the row direction and in the
 * opposite direction; slot 0 (the row's own column) goes into row_coef[i].w and is subtracted on chip */
int scl_bwd_coeffs(const void* row_stats, int m_rows, const void* col_stats, int n_cols, const scl_plan* plan,
                   int b_local, int rank, const float* gaps, const float* scalars3, const float* grad_out, float c,
                   float w, float mult, int col_mode, const float* pos_q, const float* opp_q_local, int k_plus_1,
                   void* row_coef, void* col_coef, void* stream);
/* fused recompute + dL/dz + second GEMM (tcgen05): dx_partial float[plan.chunks, plan.m_pad, d];
 * diag_col0 = global column of local row 0 (rank * b_local, the ground-truth offset of losses.py:94 /
 * loss.py:95-96).  fp32-accurate mode (plan.split): x_rows [m_rows, 3d], y_cols [n_cols, 3d] */
int scl_bwd_rows(const void* x_rows, int m_rows, const void* y_cols, int n_cols, int d, int diag_col0,
                 const float* scalars3, const scl_plan* plan, const void* row_coef, const void* col_coef,
                 float* dx_partial, void* stream);
/* sum the chunk partials, add the neighbour (slot >= 1) soft-target terms of both directions, cast: dx_out[m_rows, d]
 * in out_dtype.  Deterministic (no atomics on dx): the opposite-direction entries are bucketed per local row in
 * `workspace` (>= scl_bwd_finish_workspace_bytes; only touched when col_mode != 0) and added in a fixed order. */
size_t scl_bwd_finish_workspace_bytes(int n_global, int b_local, int k_plus_1);
int scl_bwd_finish(const float* dx_partial, const scl_plan* plan, int m_rows, int d, const void* y_all,
                   const int32_t* pos_col, const float* pos_q, int k_plus_1, const int32_t* opp_col_all,
                   const float* opp_q_all, int n_global, int b_local, int rank, const float* gaps,
                   const float* scalars3, const float* grad_out, float c, float w, float mult, int col_mode,
                   void* workspace, size_t workspace_bytes, void* dx_out, int out_dtype, void* stream);

/* ---- composite entry points: one host call per phase ----------------------------------------------
 * The per-step host cost (ctypes calls, work-area allocations) matters once the batch is sharded over 8 GPUs and a
 * rank's kernels take ~1 ms; these chain the launches above on the given stream.  All pointers as above. */
typedef struct scl_prepare_args {
  const void* image; const void* text; int src_dtype;     /* [rows, d] inputs of the loss modules          */
  const float* logit_scale; float cap;                     /* cap <= 0: none                                */
  int rows, d;
  void* image_bf16; void* text_bf16;                       /* [rows, d] bf16 (required)                     */
  float* scalars3;
} scl_prepare_args;
int scl_prepare(const scl_prepare_args* a, void* stream);

typedef struct scl_fwd_args {
  const void* img_l; const void* txt_l;                    /* local rows  [b_local, d] bf16                 */
  const void* img_all; const void* txt_all;                /* gathered    [n_global, d] bf16                */
  int b_local, n_global, d, rank;
  const float* scalars3;
  const int64_t* img_ids_all; const int64_t* txt_ids_all;  /* gathered tile ids (k > 0)                     */
  const int64_t* nbr_ids; const float* nbr_alpha;          /* [b_local, k]                                  */
  int k; float alpha_scale; int same_ids;
  float c, w; int finalize_scalars;                        /* 0: stop after sums6 (caller exchanges them)   */
  int32_t* col_it; float* w_it; float* q_it;               /* image rows -> text columns  [b_local, k+1]    */
  int32_t* col_ti; float* w_ti; float* q_ti;               /* text rows -> image columns (may alias *_it)   */
  void* stats_i; void* stats_t;                            /* float4[b_local]                               */
  float* sums6; float* out4;
  void* workspace; size_t workspace_bytes;                 /* >= scl_fwd_workspace_bytes(...)               */
  int32_t* ranks_out;                                      /* NULL, or int32[b_local]: image -> gene retrieval ranks
                                                              in the local block (scl_fwd_rowstats_ranks)      */
  int phases;                                              /* 0 = all; bit 0 soft targets (needs the gathered ids),
                                                              bit 1 image-rows pass (needs txt_all), bit 2 text-rows
                                                              pass + reductions (needs img_all): one call per phase
                                                              lets the caller overlap the three all-gathers;
                                                              6 (no bit 0): col_* / q_* hold soft targets the
                                                              caller resolved itself (data-side lists)             */
} scl_fwd_args;
size_t scl_fwd_workspace_bytes(int b_local, int n_global, int d, int k);
int scl_fwd_all(const scl_fwd_args* a, void* stream);

typedef struct scl_bwd_args {
  const void* x_rows; const void* y_all;
  int b_local, n_global, d, rank;
  const void* row_stats; const void* col_stats_all;        /* float4[b_local], float4[n_global]             */
  const int32_t* pos_col; const float* pos_q; const float* opp_q_local;
  const int32_t* opp_col_all; const float* opp_q_all; int k_plus_1;
  const float* gaps; const float* scalars3; const float* grad_out;
  float c, w, mult; int col_mode;
  void* dx_out; int out_dtype;                             /* [b_local, d]                                  */
  void* workspace; size_t workspace_bytes;                 /* >= scl_bwd_workspace_bytes(...)               */
  int split;                                               /* 1: x_rows [b_local,3d], y_all [n,3d]          */
} scl_bwd_args;
size_t scl_bwd_workspace_bytes(int b_local, int n_global, int d, int k_plus_1);
int scl_bwd_dir(const scl_bwd_args* a, void* stream);

/* ---- statistics exchange helper (replaces torch.distributed.nn's reduce-scatter in backward) ----
 * gathered: float[world][rec_floats], the all-gathered flat per-rank records (host arrays outs/offs/lens of
 * n_comp <= 8 entries): component k of every rank is copied to outs[k][rank * lens[k] ...], one launch. */
int scl_unpack_records(const float* gathered, int world, int rec_floats, int n_comp, float* const* outs,
                       const int* offs, const int* lens, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCL_B200_H_ */
