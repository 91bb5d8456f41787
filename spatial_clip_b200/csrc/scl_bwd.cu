// Backward row-gradient kernel of the contrastive loss (sm_100a, tcgen05 + TMEM + TMA).
//
//   dX[i, :] = sum_j G_ij * Y[j, :]          for the local rows i of X, all global columns j of Y
//   G_ij     = P_ij  * (u_i  + v_i  * z_ij)   row-direction softmax term   (P  = 2^(z*s2 - Lr_i))
//            + Pc_ij * (u'_j + v'_j * z_ij)   column-direction term        (Pc = 2^(z*s2 - Lc_j))
// z = X . Y^T is recomputed tile by tile in tensor memory, G is formed in registers, rounded to bf16
// into a swizzled shared-memory operand tile and fed straight back to the tensor core for the
// second GEMM, so neither the logits nor dL/dlogits ever touch HBM.  The sparse soft-target part
// (-Q terms, <= K+1 entries per row) is added by scl_bwd_finish.
// Replaces autograd's backward through losses.py:78-122 / loss.py:117-153 of the reference:
// two softmax-backward passes over materialised [B_l, N] matrices plus four cuBLAS GEMMs.
//
// (u, v, u', v') carry every scalar of the closed form: 0.5/B_l, s_eff, the temperature
// regulariser 2*w*gap of the rank owning the row / the column, the upstream gradient, and which
// gathered operands carry gradient (local_loss / gather_with_grad).  Called twice per step with the
// roles of image and gene embeddings swapped.
//
// CTA = (128-row block, D split, chunk of 128-column tiles).  TMEM: dX accumulator [128 x DN] fp32 in
// columns [0, DN), the recomputed z tile double buffered in columns [256, 512).  One 32 KB x 5 smem
// ring carries, in issue order, the K-chunks of (X, Y) for z and the Y^T chunks for the second GEMM.
#include "scl_kernels.h"
#include "scl_ptx.cuh"

namespace scl {

constexpr int kBwdBM = 128;
constexpr int kBwdBN = 128;
constexpr int kBwdBK = 64;
constexpr int kBwdStages = 5;
constexpr int kBwdStageBytes = 32768;
constexpr int kBwdGBytes = kBwdBM * kBwdBN * 2;  // 32 KB, two [128 x 64] K-major sub-tiles
constexpr int kBwdThreads = 384;
constexpr int kBwdZCol = 256;  // TMEM column of z buffer 0

struct BwdSmemBars {
  uint64_t full[kBwdStages];
  uint64_t empty[kBwdStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t g_full[2];
  uint64_t g_empty[2];
  uint64_t acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
bwd_rows_kernel(const __grid_constant__ CUtensorMap tm_rows,    // X   [M, D]  box {64, 128}
                const __grid_constant__ CUtensorMap tm_cols,    // Y   [N, D]  box {64, 128}
                const __grid_constant__ CUtensorMap tm_cols_t,  // Y^T [D, N]  box {64, DN}
                int m_rows, int n_cols, int d, int dn, int n_tiles, int tiles_per_chunk, int m_pad, int diag0,
                const float* __restrict__ scale_log2_ptr, const float4* __restrict__ row_coef,
                const float4* __restrict__ col_coef, float* __restrict__ dx_partial) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ BwdSmemBars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_ring = smem;
  uint8_t* smem_g = smem + kBwdStages * kBwdStageBytes;  // 2 x 32 KB

  const int nk = d / kBwdBK;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kBwdBM;
  const int dsplit = blockIdx.y;  // which DN-wide slice of D this CTA accumulates
  const int t_begin = blockIdx.z * tiles_per_chunk;
  const int t_end = min(t_begin + tiles_per_chunk, n_tiles);
  const int n_my = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_rows);
    tma_prefetch_desc(&tm_cols);
    tma_prefetch_desc(&tm_cols_t);
    for (int s = 0; s < kBwdStages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.tmem_full[b], 1);
      mbar_init(&bars.tmem_empty[b], 8);
      mbar_init(&bars.g_full[b], 8);
      mbar_init(&bars.g_empty[b], 1);
    }
    mbar_init(&bars.acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&bars.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      auto push_z = [&](int lt) {
        const int col0 = (t_begin + lt) * kBwdBN;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % kBwdStages;
          mbar_wait(&bars.empty[s], ((it / kBwdStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.full[s], kBwdStageBytes);
          tma_load_2d(smem_ring + s * kBwdStageBytes, &tm_rows, &bars.full[s], kc * kBwdBK, row0);
          tma_load_2d(smem_ring + s * kBwdStageBytes + 16384, &tm_cols, &bars.full[s], kc * kBwdBK, col0);
        }
      };
      auto push_yt = [&](int lt) {
        const int col0 = (t_begin + lt) * kBwdBN;
        for (int js = 0; js < 2; ++js, ++it) {
          const int s = it % kBwdStages;
          mbar_wait(&bars.empty[s], ((it / kBwdStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.full[s], static_cast<uint32_t>(dn * 128));
          tma_load_2d(smem_ring + s * kBwdStageBytes, &tm_cols_t, &bars.full[s], col0 + js * 64, dsplit * dn);
        }
      };
      push_z(0);
      for (int lt = 0; lt < n_my; ++lt) {
        if (lt + 1 < n_my) push_z(lt + 1);
        push_yt(lt);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc_z = umma_idesc_bf16(kBwdBM, kBwdBN);
      const uint32_t idesc_acc = umma_idesc_bf16(kBwdBM, dn);
      int it = 0;
      auto issue_z = [&](int lt) {
        const int buf = lt & 1;
        mbar_wait(&bars.tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kBwdZCol + buf * kBwdBN;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % kBwdStages;
          mbar_wait(&bars.full[s], (it / kBwdStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_ring + s * kBwdStageBytes);
          const uint32_t b_addr = a_addr + 16384;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(d_tmem, umma_desc_kmajor_sw128(a_addr + k * 32), umma_desc_kmajor_sw128(b_addr + k * 32),
                        idesc_z, (kc | k) != 0 ? 1u : 0u);
          tc_commit(&bars.empty[s]);
        }
        tc_commit(&bars.tmem_full[buf]);
      };
      auto issue_acc = [&](int lt) {
        const int gbuf = lt & 1;
        mbar_wait(&bars.g_full[gbuf], (lt >> 1) & 1);
        tc_fence_after();
        for (int js = 0; js < 2; ++js, ++it) {
          const int s = it % kBwdStages;
          mbar_wait(&bars.full[s], (it / kBwdStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_g + gbuf * kBwdGBytes + js * 16384);
          const uint32_t b_addr = smem_u32(smem_ring + s * kBwdStageBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16(tmem_base, umma_desc_kmajor_sw128(a_addr + k * 32), umma_desc_kmajor_sw128(b_addr + k * 32),
                        idesc_acc, (lt | js | k) != 0 ? 1u : 0u);
          tc_commit(&bars.empty[s]);
        }
        tc_commit(&bars.g_empty[gbuf]);
      };
      issue_z(0);
      for (int lt = 0; lt < n_my; ++lt) {
        if (lt + 1 < n_my) issue_z(lt + 1);
        issue_acc(lt);
      }
      tc_commit(&bars.acc_full);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: z -> G (bf16, swizzled smem)
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;  // 64-column half of the tile == G sub-tile index
    const int r_loc = q * 32 + lane;
    const float s2 = __ldg(scale_log2_ptr);
    const float4 rc = __ldg(&row_coef[row0 + r_loc]);  // {Lr (log2 units), u, v, diagonal soft-target term}
    // The row's own column (weight >= 1/2 of the soft target, P ~ 1 near convergence) is subtracted here, in
    // fp32 BEFORE the bf16 rounding of G: done later in fp32 against the rounded dense value it would cancel
    // catastrophically (net (P - 1) vs an absolute rounding error of 2^-9 * P).
    const int diag_col = diag0 + row0 + r_loc;
    const int warp_diag_lo = diag0 + row0 + q * 32;
    for (int lt = 0; lt < n_my; ++lt) {
      const int buf = lt & 1;
      mbar_wait(&bars.tmem_full[buf], (lt >> 1) & 1);
      mbar_wait(&bars.g_empty[buf], ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      uint8_t* g_tile = smem_g + buf * kBwdGBytes + h * 16384;
      uint8_t* g_row = g_tile + (r_loc >> 3) * 1024 + (r_loc & 7) * 128;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = (t_begin + lt) * kBwdBN + h * 64 + c * 32;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kBwdZCol + buf * kBwdBN + h * 64 + c * 32, r);
        tmem_ld_wait();
        uint32_t packed[16];
        const bool has_diag = (col0 + 32 > warp_diag_lo) && (col0 < warp_diag_lo + 32);  // warp-uniform
        const int di = has_diag ? diag_col - col0 : -1;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float g2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float z = __uint_as_float(r[j + e]);
            const float4 cc = __ldg(&col_coef[col0 + j + e]);  // {Lc, u', v', -}; warp-uniform address
            const float y = z * s2;
            const float p = ex2_approx(y - rc.x);
            const float pc = ex2_approx(y - cc.x);
            float g = p * fmaf(rc.z, z, rc.y);
            g = fmaf(pc, fmaf(cc.z, z, cc.y), g);
            if (has_diag) g -= (j + e == di) ? rc.w : 0.f;
            g2[e] = (col0 + j + e < n_cols) ? g : 0.f;
          }
          packed[j >> 1] = pack_bf16x2(g2[0], g2[1]);
        }
        // K-major SWIZZLE_128B operand layout: 16-byte chunk index XOR (row % 8)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (c * 4 + ch) ^ (r_loc & 7);
          *reinterpret_cast<uint4*>(g_row + chunk * 16) =
              make_uint4(packed[ch * 4 + 0], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.tmem_empty[buf]);
        mbar_arrive(&bars.g_full[buf]);
      }
    }
    // ---- drain the dX accumulator
    mbar_wait(&bars.acc_full, 0);
    tc_fence_after();
    const int n_ch = dn / 32;
    const int c_begin = h == 0 ? 0 : n_ch / 2;
    const int c_end = h == 0 ? n_ch / 2 : n_ch;
    float* out_row = dx_partial + (static_cast<size_t>(blockIdx.z) * m_pad + row0 + r_loc) * d + dsplit * dn;
    for (int c = c_begin; c < c_end; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(out_row + c * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

size_t bwd_smem_bytes() { return 1024 + kBwdStages * kBwdStageBytes + 2 * kBwdGBytes; }

void bwd_pick_split(int d, int* n_dsplit, int* dn) {
  *n_dsplit = d > 256 ? 2 : 1;
  *dn = d / *n_dsplit;
}

int bwd_pick_chunks(int m_rows, int n_cols, int d, int num_sms, int* tiles_per_chunk) {
  int n_dsplit, dn;
  bwd_pick_split(d, &n_dsplit, &dn);
  const int row_blocks = (m_rows + kBwdBM - 1) / kBwdBM;
  const int n_tiles = (n_cols + kBwdBN - 1) / kBwdBN;
  int chunks = (6 * num_sms + row_blocks * n_dsplit - 1) / (row_blocks * n_dsplit);
  chunks = max(1, min(chunks, max(1, n_tiles / 8)));
  int tpc = (n_tiles + chunks - 1) / chunks;
  chunks = (n_tiles + tpc - 1) / tpc;
  *tiles_per_chunk = tpc;
  return chunks;
}

cudaError_t launch_bwd_rows(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, const CUtensorMap& tm_cols_t,
                            int m_rows, int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad, int diag0,
                            const float* scale_log2, const float4* row_coef, const float4* col_coef,
                            float* dx_partial, cudaStream_t stream) {
  int n_dsplit, dn;
  bwd_pick_split(d, &n_dsplit, &dn);
  const size_t smem = bwd_smem_bytes();
  // opt in to > 48 KB dynamic shared memory once per device (the attribute is sticky; 227 KB covers every D)
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t err = cudaFuncSetAttribute(bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 231424);
    if (err != cudaSuccess) return err;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int row_blocks = (m_rows + kBwdBM - 1) / kBwdBM;
  const int n_tiles = (n_cols + kBwdBN - 1) / kBwdBN;
  dim3 grid(row_blocks, n_dsplit, chunks);
  bwd_rows_kernel<<<grid, kBwdThreads, smem, stream>>>(tm_rows, tm_cols, tm_cols_t, m_rows, n_cols, d, dn, n_tiles,
                                                       tiles_per_chunk, m_pad, diag0, scale_log2, row_coef, col_coef,
                                                       dx_partial);
  return cudaGetLastError();
}

}  // namespace scl
