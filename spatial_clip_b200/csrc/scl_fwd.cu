// Forward row-statistics kernel of the contrastive loss (sm_100a, tcgen05 + TMEM + TMA).
//
// For a block of local rows X[M, D] (image OR gene embeddings, bf16) against all global columns
// Y[N, D] it streams the similarity tiles  z = X . Y^T  through tensor memory and reduces, per row i,
//   m_i   = max_j  z_ij * s2                      (s2 = s_eff * log2(e))
//   S0_i  = sum_j 2^(z_ij*s2 - m_i)               -> LSE_i = (m_i + log2 S0_i) * ln2
//   S1_i  = sum_j 2^(...) * z_ij                  -> mu_i  = E_p[z]   (temperature regulariser, d/ds)
//   S2_i  = sum_j 2^(...) * z_ij^2                -> Var_p[z]         (d gap / d s)
// with an online (flash-style) running max, so the [M, N] logits never exist in HBM.
// Replaces: torch.matmul + F.log_softmax/F.cross_entropy/F.softmax passes of
//   /root/reference/src/models/components/losses.py:78-89,113-121 and
//   /root/reference/src/open_clip/loss.py:117-124,150-153.
//
// CTA = (128-row block, chunk of 256-column tiles).  Warp roles: w0 TMA producer, w1 MMA issuer,
// w2 TMEM allocator, w4..w11 epilogue (TMEM lane quadrant = warp % 4, column half = (warp-4) / 4).
// X block is smem-stationary (D <= 512), Y streams through a 3-stage 32 KB ring, the fp32 accumulator
// is double buffered in TMEM (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of t+1.
#include "scl_kernels.h"
#include "scl_ptx.cuh"

namespace scl {

constexpr int kFwdBM = 128;
constexpr int kFwdBN = 256;
constexpr int kBK = 64;
constexpr int kFwdStages = 3;
constexpr int kFwdThreads = 384;
constexpr int kFwdAChunkBytes = kFwdBM * kBK * 2;  // 16 KB
constexpr int kFwdBStageBytes = kFwdBN * kBK * 2;  // 32 KB

struct FwdSmemBars {
  uint64_t a_full;
  uint64_t full[kFwdStages];
  uint64_t empty[kFwdStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kFwdThreads, 1)
fwd_rowstats_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_cols,
                    int m_rows, int n_cols, int d, int n_tiles, int tiles_per_chunk, int m_pad,
                    const float* __restrict__ scale_log2_ptr, float4* __restrict__ partial,
                    float* __restrict__ dbg_z, int dbg_ld) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ FwdSmemBars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nk = d / kBK;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + nk * kFwdAChunkBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kFwdBM;
  const int t_begin = blockIdx.y * tiles_per_chunk;
  const int t_end = min(t_begin + tiles_per_chunk, n_tiles);
  const int n_my = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_rows);
    tma_prefetch_desc(&tm_cols);
    mbar_init(&bars.a_full, 1);
    for (int s = 0; s < kFwdStages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.tmem_full[b], 1);
      mbar_init(&bars.tmem_empty[b], 8);  // one arrive per epilogue warp
    }
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
      mbar_arrive_expect_tx(&bars.a_full, static_cast<uint32_t>(nk * kFwdAChunkBytes));
      for (int kc = 0; kc < nk; ++kc)
        tma_load_2d(smem_a + kc * kFwdAChunkBytes, &tm_rows, &bars.a_full, kc * kBK, row0);
      int it = 0;
      for (int lt = 0; lt < n_my; ++lt) {
        const int col0 = (t_begin + lt) * kFwdBN;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % kFwdStages;
          const uint32_t ph = (it / kFwdStages) & 1;
          mbar_wait(&bars.empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&bars.full[s], kFwdBStageBytes);
          tma_load_2d(smem_b + s * kFwdBStageBytes, &tm_cols, &bars.full[s], kc * kBK, col0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kFwdBM, kFwdBN);
      mbar_wait(&bars.a_full, 0);
      tc_fence_after();
      int it = 0;
      for (int lt = 0; lt < n_my; ++lt) {
        const int buf = lt & 1;
        mbar_wait(&bars.tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kFwdBN;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % kFwdStages;
          const uint32_t ph = (it / kFwdStages) & 1;
          mbar_wait(&bars.full[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + kc * kFwdAChunkBytes);
          const uint32_t b_addr = smem_u32(smem_b + s * kFwdBStageBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            tc_mma_bf16(d_tmem, umma_desc_kmajor_sw128(a_addr + k * 32), umma_desc_kmajor_sw128(b_addr + k * 32),
                        idesc, (kc | k) != 0 ? 1u : 0u);
          }
          tc_commit(&bars.empty[s]);  // smem stage reusable once these MMAs retire
        }
        tc_commit(&bars.tmem_full[buf]);  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: online softmax statistics
    const int q = warp & 3;         // TMEM lane quadrant this warp may access
    const int h = (warp - 4) >> 2;  // which 128-column half of the 256-wide tile
    const int row = row0 + q * 32 + lane;
    const float s2 = __ldg(scale_log2_ptr);
    float m = -INFINITY, s_e = 0.f, s_ez = 0.f, s_ezz = 0.f;
    for (int lt = 0; lt < n_my; ++lt) {
      const int buf = lt & 1;
      mbar_wait(&bars.tmem_full[buf], (lt >> 1) & 1);
      tc_fence_after();
      const int tile_col0 = (t_begin + lt) * kFwdBN + h * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const int col0 = tile_col0 + c * 32;
        if (col0 >= n_cols) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kFwdBN + h * 128 + c * 32, r);
        tmem_ld_wait();
        const int n_valid = min(32, n_cols - col0);
        float y[32];
        float cm = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[j]);
          y[j] = (j < n_valid) ? x * s2 : -INFINITY;
          cm = fmaxf(cm, y[j]);
        }
        const float m_new = fmaxf(m, cm);
        const float sc = ex2_approx(m - m_new);
        s_e *= sc;
        s_ez *= sc;
        s_ezz *= sc;
        m = m_new;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[j]);
          const float e = ex2_approx(y[j] - m);
          s_e += e;
          const float t = e * x;
          s_ez += t;
          s_ezz = fmaf(t, x, s_ezz);
        }
        if (dbg_z != nullptr && row < m_rows) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < n_valid) dbg_z[static_cast<size_t>(row) * dbg_ld + col0 + j] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.tmem_empty[buf]);
    }
    const int slot = blockIdx.y * 2 + h;
    partial[static_cast<size_t>(slot) * m_pad + row] = make_float4(m, s_e, s_ez, s_ezz);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

size_t fwd_smem_bytes(int d) { return 1024 + static_cast<size_t>(d / kBK) * kFwdAChunkBytes + kFwdStages * kFwdBStageBytes; }

int fwd_pick_chunks(int m_rows, int n_cols, int num_sms, int* tiles_per_chunk) {
  const int row_blocks = (m_rows + kFwdBM - 1) / kFwdBM;
  const int n_tiles = (n_cols + kFwdBN - 1) / kFwdBN;
  // aim for >= ~6 CTAs per SM in total so the last wave is a small fraction, but keep >= 4 tiles per
  // CTA so the stationary X block load is amortised
  int chunks = (6 * num_sms + row_blocks - 1) / row_blocks;
  chunks = max(1, min(chunks, max(1, n_tiles / 4)));
  int tpc = (n_tiles + chunks - 1) / chunks;
  chunks = (n_tiles + tpc - 1) / tpc;
  *tiles_per_chunk = tpc;
  return chunks;
}

cudaError_t launch_fwd_rowstats(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows, int n_cols, int d,
                                int chunks, int tiles_per_chunk, int m_pad, const float* scale_log2, float4* partial,
                                float* dbg_z, int dbg_ld, cudaStream_t stream) {
  const size_t smem = fwd_smem_bytes(d);
  // opt in to > 48 KB dynamic shared memory once per device (the attribute is sticky; 227 KB covers every D)
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t err = cudaFuncSetAttribute(fwd_rowstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 231424);
    if (err != cudaSuccess) return err;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int row_blocks = (m_rows + kFwdBM - 1) / kFwdBM;
  const int n_tiles = (n_cols + kFwdBN - 1) / kFwdBN;
  dim3 grid(row_blocks, chunks);
  fwd_rowstats_kernel<<<grid, kFwdThreads, smem, stream>>>(tm_rows, tm_cols, m_rows, n_cols, d, n_tiles,
                                                           tiles_per_chunk, m_pad, scale_log2, partial, dbg_z, dbg_ld);
  return cudaGetLastError();
}

}  // namespace scl
