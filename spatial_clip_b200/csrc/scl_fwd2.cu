// Forward row-statistics kernel (tcgen05 cta_group::2, cluster of 2 CTAs).
//
// Online row max / sum / first and second moments of z = X . Y^T.  One UMMA instruction spans two SMs: the pair
// owns 256 rows (128 per CTA), each CTA loads only ITS half of every 256-column Y tile (16 KB per K-chunk), and the
// leader CTA's single MMA thread issues 256x256x16 instructions that read both CTAs' shared memory.  The X block
// stays resident (128 KB at D = 512) next to a 6-deep TMA ring.
//
// Barrier protocol (all barriers live at the same smem offset in both CTAs):
//   a_full, full[s]   waited on by the leader only; both CTAs' TMA loads credit the LEADER's barrier
//   empty[s]          each CTA's producer waits on its own copy; the leader's tcgen05.commit multicasts
//   tmem_full[b]      each CTA's epilogue waits on its own copy (multicast commit)
//   tmem_empty[b]     leader only, 16 arrivals: the 8 epilogue warps of both CTAs (peer arrives remotely)
#include "scl_kernels.h"
#include "scl_ptx.cuh"

namespace scl {

constexpr int kF2Rows = 128;        // rows per CTA (256 per pair)
constexpr int kF2TileN = 256;       // columns per tile (128 loaded by each CTA)
constexpr int kF2BK = 64;
constexpr int kF2Stages = 6;
constexpr int kF2EpiWarps = 16;
constexpr int kF2Threads = (kF2EpiWarps + 3) * 32;  // 608
// Warp roles: 0..15 epilogue, 16 TMA producer, 17 MMA issuer, 18 TMEM allocator.  The single-thread roles get
// the HIGHEST warp ids on purpose: the SM's warp arbiter favours high warp ids, and the MMA issuer shares its
// scheduler with four busy epilogue warps -- at a low id it was starved of issue slots (measured: ~130 cycles
// per tcgen05.mma issue against 65 cycles of execution).
constexpr int kF2ProducerWarp = kF2EpiWarps;
constexpr int kF2MmaWarp = kF2EpiWarps + 1;
constexpr int kF2AllocWarp = kF2EpiWarps + 2;
constexpr int kF2AChunkBytes = kF2Rows * kF2BK * 2;  // 16 KB
constexpr int kF2BStageBytes = 128 * kF2BK * 2;      // 16 KB (this CTA's half of the tile)

struct F2Bars {
  uint64_t a_full;
  uint64_t full[kF2Stages];
  uint64_t empty[kF2Stages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

// kRank = 1 (in-pass retrieval ranks): every row additionally counts the columns of the LOCAL block
// [loc_lo, loc_hi) whose similarity exceeds the row's own pair diag_z[row] (own column loc_lo + row excluded), i.e.
// the position of the matching gene profile in the image -> gene retrieval that the reference finds with a
// [B_l, B_l] matmul + topk (spatial_clip_module.py:68, metrics.py:22-36).  rank_part mirrors `partial`'s slots.
template <int kRank>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kF2Threads, 1)
fwd_rowstats_pair_kernel(const __grid_constant__ CUtensorMap tm_rows,  // X [M, D], box {64, 128}
                         const __grid_constant__ CUtensorMap tm_cols,  // Y [N, D], box {64, 128}
                         int m_rows, int n_cols, int d, int n_tiles, int tiles_per_chunk, int m_pad,
                         const float* __restrict__ scale_log2_ptr, float4* __restrict__ partial,
                         float* __restrict__ dbg_z, int dbg_ld,
                         const float* __restrict__ diag_z, int loc_lo, int loc_hi, int* __restrict__ rank_part) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ F2Bars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nk = d / kF2BK;
  // D <= 512: the X block stays resident (nk x 16 KB) and the ring carries only Y chunks (16 KB stages).
  // D  > 512: X no longer fits next to a useful ring, so X chunks stream with the Y chunks (32 KB stages).
  const bool stream_x = d > 512;
  const int stage_bytes = stream_x ? 2 * kF2BStageBytes : kF2BStageBytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (stream_x ? 0 : nk * kF2AChunkBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int row0 = (blockIdx.x >> 1) * 256 + static_cast<int>(cta) * kF2Rows;
  const int t_begin = blockIdx.y * tiles_per_chunk;
  const int t_end = min(t_begin + tiles_per_chunk, n_tiles);
  const int n_my = t_end - t_begin;

  if (warp == kF2ProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_rows);
    tma_prefetch_desc(&tm_cols);
    mbar_init(&bars.a_full, 1);
    for (int s = 0; s < kF2Stages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.tmem_full[b], 1);
      mbar_init(&bars.tmem_empty[b], 2 * kF2EpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kF2AllocWarp) {
    tmem_alloc_pair(&bars.tmem_base, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers initialised and TMEM allocated before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp == kF2ProducerWarp) {
    // ------------------------------------------------------------ TMA producer (one thread per CTA)
    if (lane == 0) {
      if (!stream_x) {
        if (leader) mbar_arrive_expect_tx(&bars.a_full, static_cast<uint32_t>(2 * nk * kF2AChunkBytes));
        for (int kc = 0; kc < nk; ++kc)
          tma_load_2d_pair(smem_a + kc * kF2AChunkBytes, &tm_rows, &bars.a_full, kc * kF2BK, row0);
      }
      int it = 0;
      for (int lt = 0; lt < n_my; ++lt) {
        const int col0 = (t_begin + lt) * kF2TileN + static_cast<int>(cta) * 128;
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const int s = it % kF2Stages;
          mbar_wait(&bars.empty[s], ((it / kF2Stages) & 1) ^ 1);
          if (leader) mbar_arrive_expect_tx(&bars.full[s], static_cast<uint32_t>(2 * stage_bytes));
          tma_load_2d_pair(smem_b + s * stage_bytes, &tm_cols, &bars.full[s], kc * kF2BK, col0);
          if (stream_x)
            tma_load_2d_pair(smem_b + s * stage_bytes + kF2BStageBytes, &tm_rows, &bars.full[s], kc * kF2BK, row0);
        }
      }
    }
  } else if (warp == kF2MmaWarp) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    // The whole warp walks the loop converged (all lanes poll the barriers); one elected lane issues the
    // tcgen05 instructions, so the compiler emits them straight-line instead of per-active-lane loops.
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, kF2TileN);
      if (!stream_x) mbar_wait_warp(&bars.a_full, 0);
      tc_fence_after();
      // Descriptors and barrier addresses of a batch are computed BEFORE the wait for its data (base + offset in the
      // 16-byte address field) and pinned in registers, so that only the MMAs themselves follow the wait: with the
      // descriptor arithmetic behind the barrier the tensor pipe drained ~150 cycles per batch of four MMAs.
      const uint64_t a_desc0 = umma_desc_kmajor_sw128(smem_u32(smem_a));
      const uint64_t b_desc0 = umma_desc_kmajor_sw128(smem_u32(smem_b));
      const uint64_t stage_units = static_cast<uint64_t>(stage_bytes >> 4);
      int s = 0;
      uint32_t ph = 0;
      for (int lt = 0; lt < n_my; ++lt) {
        const int buf = lt & 1;
        mbar_wait_warp(&bars.tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kF2TileN;
        for (int kc = 0; kc < nk; ++kc) {
          uint64_t b_desc = b_desc0 + static_cast<uint64_t>(s) * stage_units;
          uint64_t a_desc = stream_x ? b_desc + (kF2BStageBytes >> 4)
                                     : a_desc0 + static_cast<uint64_t>(kc) * (kF2AChunkBytes >> 4);
          uint32_t bar_full = smem_u32(&bars.full[s]), bar_empty = smem_u32(&bars.empty[s]);
          asm volatile("" : "+l"(a_desc), "+l"(b_desc), "+r"(bar_full), "+r"(bar_empty));
          mbar_wait_warp_u32(bar_full, ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kF2BK / 16; ++k)  // +32 B per K step == +2 in the descriptor's 16-byte address field
              tc_mma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0 ? 1u : 0u);
            tc_commit_pair_u32(bar_empty);
            if (kc == nk - 1) tc_commit_pair(&bars.tmem_full[buf]);
          }
          __syncwarp();
          if (++s == kF2Stages) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp < kF2EpiWarps) {
    // ------------------------------------------------------------ epilogue (each CTA: its own 128 rows)
    // 16 warps: TMEM lane quadrant q = warp % 4, 64-column slice hh = (warp - 4) / 4 of the 256-wide tile.
    // Thread = row: running max m (log2 units) and the three sums stay in registers for the whole chunk.
    const int q = warp & 3;
    const int hh = warp >> 2;
    const int row = row0 + q * 32 + lane;
    const float s2 = __ldg(scale_log2_ptr);
    float m = -INFINITY, s_e = 0.f, s_ez = 0.f, s_ezz = 0.f;
    int above = 0;  // kRank: local columns scoring above this row's own pair
    float zd = 0.f;
    if constexpr (kRank != 0) zd = row < m_rows ? __ldg(diag_z + row) : INFINITY;
    for (int lt = 0; lt < n_my; ++lt) {
      const int buf = lt & 1;
      mbar_wait_warp(&bars.tmem_full[buf], (lt >> 1) & 1);
      tc_fence_after();
      const int tile_col0 = (t_begin + lt) * kF2TileN + hh * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = tile_col0 + c * 32;
        if (col0 >= n_cols) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kF2TileN + hh * 64 + c * 32, r);
        tmem_ld_wait();
        if (col0 + 32 <= n_cols && s2 > 0.f) {
          // ---- fast path (every full tile): no masks, max over the raw similarities, fused scale-and-shift
          float cm0 = __uint_as_float(r[0]), cm1 = __uint_as_float(r[1]);
#pragma unroll
          for (int j = 2; j < 32; j += 2) {
            cm0 = fmaxf(cm0, __uint_as_float(r[j]));
            cm1 = fmaxf(cm1, __uint_as_float(r[j + 1]));
          }
          const float m_new = fmaxf(m, fmaxf(cm0, cm1) * s2);
          if (m_new > m) {  // rare after the first few chunks
            const float sc = ex2_approx(m - m_new);
            s_e *= sc;
            s_ez *= sc;
            s_ezz *= sc;
            m = m_new;
          }
          const float neg_m = -m;
          float e0 = 0.f, e1 = 0.f, z0 = 0.f, z1 = 0.f, w0 = 0.f, w1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float xa = __uint_as_float(r[j]), xb = __uint_as_float(r[j + 1]);
            const float ea = ex2_approx(fmaf(xa, s2, neg_m));
            const float eb = ex2_approx(fmaf(xb, s2, neg_m));
            e0 += ea;
            e1 += eb;
            const float ta = ea * xa, tb = eb * xb;
            z0 += ta;
            z1 += tb;
            w0 = fmaf(ta, xa, w0);
            w1 = fmaf(tb, xb, w1);
          }
          s_e += e0 + e1;
          s_ez += z0 + z1;
          s_ezz += w0 + w1;
        } else {
          // ---- general path: ragged last tile (masked columns) or non-positive scale
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
        }
        if constexpr (kRank != 0) {
          if (col0 < loc_hi && col0 + 32 > loc_lo) {  // warp-uniform: this chunk touches the local block
            const int own = loc_lo + row;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int cidx = col0 + j;
              above += (cidx >= loc_lo && cidx < loc_hi && cidx != own && __uint_as_float(r[j]) > zd) ? 1 : 0;
            }
          }
        }
        if (dbg_z != nullptr && row < m_rows) {
          const int n_valid = min(32, n_cols - col0);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < n_valid) dbg_z[static_cast<size_t>(row) * dbg_ld + col0 + j] = __uint_as_float(r[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&bars.tmem_empty[buf]);
        else mbar_arrive_remote(&bars.tmem_empty[buf], 0);
      }
    }
    const int slot = blockIdx.y * 4 + hh;
    partial[static_cast<size_t>(slot) * m_pad + row] = make_float4(m, s_e, s_ez, s_ezz);
    if constexpr (kRank != 0) rank_part[static_cast<size_t>(slot) * m_pad + row] = above;
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair still uses its smem / barriers
  if (warp == kF2AllocWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

size_t fwd_pair_smem_bytes(int d) {
  if (d > 512) return 1024 + static_cast<size_t>(kF2Stages) * 2 * kF2BStageBytes;  // streamed X: 6 x 32 KB
  return 1024 + static_cast<size_t>(d / kF2BK) * kF2AChunkBytes + kF2Stages * kF2BStageBytes;
}

int fwd_pair_pick_chunks(int m_rows, int n_cols, int num_sms, int* tiles_per_chunk) {
  const int pairs = (m_rows + 255) / 256;
  const int n_tiles = (n_cols + kF2TileN - 1) / kF2TileN;
  return pick_chunks_balanced(pairs, n_tiles, num_sms / 2, 2, 0.02, tiles_per_chunk);
}

template <int kRank>
static cudaError_t launch_fwd_rowstats_pair_t(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows,
                                              int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad,
                                              const float* scale_log2, float4* partial, float* dbg_z, int dbg_ld,
                                              const float* diag_z, int loc_lo, int loc_hi, int* rank_part,
                                              cudaStream_t stream) {
  const size_t smem = fwd_pair_smem_bytes(d);
  // opt in to > 48 KB dynamic shared memory (sticky per device; set on every launch so the library keeps no state --
  // except under stream capture, where the eager warm-up launches have already set it)
  cudaError_t err = set_max_dynamic_smem(fwd_rowstats_pair_kernel<kRank>, 231424, stream);
  if (err != cudaSuccess) return err;
  const int pairs = (m_rows + 255) / 256;
  const int n_tiles = (n_cols + kF2TileN - 1) / kF2TileN;
  dim3 grid(2 * pairs, chunks);
  fwd_rowstats_pair_kernel<kRank><<<grid, kF2Threads, smem, stream>>>(tm_rows, tm_cols, m_rows, n_cols, d, n_tiles,
                                                                      tiles_per_chunk, m_pad, scale_log2, partial,
                                                                      dbg_z, dbg_ld, diag_z, loc_lo, loc_hi, rank_part);
  return cudaGetLastError();
}

cudaError_t launch_fwd_rowstats_pair(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows, int n_cols,
                                     int d, int chunks, int tiles_per_chunk, int m_pad, const float* scale_log2,
                                     float4* partial, float* dbg_z, int dbg_ld, cudaStream_t stream) {
  return launch_fwd_rowstats_pair_t<0>(tm_rows, tm_cols, m_rows, n_cols, d, chunks, tiles_per_chunk, m_pad, scale_log2,
                                       partial, dbg_z, dbg_ld, nullptr, 0, 0, nullptr, stream);
}

cudaError_t launch_fwd_rowstats_pair_ranks(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, int m_rows,
                                           int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad,
                                           const float* scale_log2, float4* partial, const float* diag_z, int loc_lo,
                                           int loc_hi, int* rank_part, cudaStream_t stream) {
  return launch_fwd_rowstats_pair_t<1>(tm_rows, tm_cols, m_rows, n_cols, d, chunks, tiles_per_chunk, m_pad, scale_log2,
                                       partial, nullptr, 0, diag_z, loc_lo, loc_hi, rank_part, stream);
}

}  // namespace scl
