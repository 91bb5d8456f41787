// Backward row-gradient kernel (tcgen05 cta_group::2, cluster of 2 CTAs).
//
//   dX[i,:] = sum_j G_ij Y[j,:],
//   G_ij = P_ij (u_i + v_i z_ij) + Pc_ij (u'_j + v'_j z_ij) - [j == own column] t_i,
// organised so that the similarity tile is recomputed exactly ONCE per (row block, column tile):
// a CTA pair owns 128 rows (64 per CTA).  With UMMA M = 128 across two SMs each SM keeps a 64-row slice
// of every accumulator in the "2x2" TMEM layout (64 rows x N as 128 lanes x N/2 columns: lanes 64..127
// hold the upper half of N), so per SM
//   dX accumulator  [64 x 512] fp32 = 2 groups x 128 columns     (TMEM columns   0..255)
//   z tile          [64 x 256] fp32 =            128 columns x 2 (TMEM columns 256..511, double buffered)
// fit together.  G is written once per CTA ([64 x 256] bf16) and both GEMMs are issued by the leader CTA's MMA
// thread with cta_group::2.
//
// The gradient GEMM reads Y itself as an MN-major B operand: tm_cols_mn is a {64 d, 64 j}-box map over the row-major
// Y [N, D] (the same bytes as tm_cols, a different box), whose SWIZZLE_128B image is exactly the canonical MN-major
// UMMA layout (umma_desc_mnmajor_sw128) -- no transposed copy of Y exists anywhere.
//
// Barriers (same smem offset in both CTAs): x_full, full[s], tmem_empty[b], g_full are waited on by
// the leader (TMA bytes / remote arrivals from the peer are credited to the leader's copy); empty[s],
// tmem_full[b], g_empty, acc_full are multicast by the leader's tcgen05.commit to both CTAs.
#include "scl_kernels.h"
#include "scl_ptx.cuh"

namespace scl {

constexpr int kB2Rows = 64;     // rows per CTA (128 per pair)
constexpr int kB2TileN = 256;   // columns per step (each CTA loads 128 of them for z)
constexpr int kB2BK = 64;
// ring stages, each two 16 KB TMA boxes behind ONE full/empty barrier pair.  Measured (profiles/r2_bwd_lab.md): 2 stages
// 2.29 ms, 3 stages 1.93 ms per launch at N = 32768; 3.5 stages (seven per-slot barriers) 1.97 ms -- more does not fit
constexpr int kB2Stages = 3;
constexpr int kB2SlotBytes = 16384;
constexpr int kB2StageBytes = 2 * kB2SlotBytes;
constexpr int kB2XChunkBytes = kB2Rows * kB2BK * 2;   // 8 KB
constexpr int kB2GSubBytes = kB2Rows * 64 * 2;        // 8 KB: [64 rows x 64 cols] bf16
constexpr int kB2GBytes = 4 * kB2GSubBytes;           // 32 KB per buffer
constexpr int kB2EpiWarps = 16;
constexpr int kB2Threads = (kB2EpiWarps + 3) * 32;  // 608
// Warp roles: 0..15 epilogue, 16 TMA producer, 17 MMA issuer, 18 TMEM allocator.  The single-thread roles get
// the HIGHEST warp ids on purpose: the SM's warp arbiter favours high warp ids, and the MMA issuer shares its
// scheduler with four busy epilogue warps -- at a low id it was starved of issue slots (measured: ~130 cycles
// per tcgen05.mma issue against 65 cycles of execution).
constexpr int kB2ProducerWarp = kB2EpiWarps;
constexpr int kB2MmaWarp = kB2EpiWarps + 1;
constexpr int kB2AllocWarp = kB2EpiWarps + 2;
constexpr int kB2CoefBytes = kB2TileN * 16;           // 4 KB: float4 per column of the step
constexpr int kB2ZCol = 256;

struct B2Bars {
  uint64_t x_full;
  uint64_t full[kB2Stages];
  uint64_t empty[kB2Stages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t g_full;
  uint64_t g_empty;
  uint64_t acc_full;
  uint64_t coef_full[2];
  uint64_t coef_empty[2];
  uint32_t tmem_base;
};

// kSplit = 1 is the fp32-accurate ("bf16x2") mode: every operand is a bf16 hi + lo pair.  The similarity is
// contracted over the K-concatenated rows X' = (h|h|l), Y' = (h|l|h) of width 3 d (x.y ~= xh.yh + xh.yl + xl.yh,
// the dropped terms are O(2^-18)), G is written as two bf16 tiles G1 + G2 and the gradient GEMM runs three passes
// G1.Yh + G1.Yl + G2.Yh, reading Yh / Yl as the column ranges [0, d) / [d, 2 d) of Y'.  X is always streamed.
template <int kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kB2Threads, 1)
bwd_rows_pair_kernel(const __grid_constant__ CUtensorMap tm_rows,     // X [M, kd]  box {64, 64}
                     const __grid_constant__ CUtensorMap tm_cols,     // Y [N, kd]  box {64, 128}
                     const __grid_constant__ CUtensorMap tm_cols_mn,  // Y [N, kd]  box {64, 64}
                     int m_rows, int n_cols, int d, int d_slices, int n_tiles, int tiles_per_chunk, int m_pad,
                     int diag0, const float* __restrict__ scale_log2_ptr, const float4* __restrict__ row_coef,
                     const float4* __restrict__ col_coef, float* __restrict__ dx_partial) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ B2Bars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nk = (kSplit ? 3 * d : d) / kB2BK;  // K chunks of the similarity contraction
  // D <= 512: one D slice, X block resident.  D > 512 (TMEM cannot hold dX[64 x D]): blockIdx.z selects a
  // slice of ds = D / d_slices output columns; z is still contracted over all of D, with the X chunks
  // streamed through the ring next to the Y chunks.
  const bool stream_x = kSplit != 0 || d_slices > 1;
  const int ds = d / d_slices;
  const int d0 = static_cast<int>(blockIdx.z) * ds;
  uint8_t* smem_x = smem;                                // nk x 8 KB, stationary (absent when streamed)
  uint8_t* smem_g = smem_x + (stream_x ? 0 : nk * kB2XChunkBytes);  // 32 KB, single buffer
  uint8_t* smem_g2 = smem_g + kB2GBytes;                 // split mode only: the low-order tile G2
  uint8_t* smem_ring = smem_g + (kSplit ? 2 : 1) * kB2GBytes;  // kB2Stages x 32 KB
  uint8_t* smem_coef = smem_ring + kB2Stages * kB2StageBytes;  // 2 x 4 KB column coefficients

  const int ng = (ds + 255) / 256;  // accumulator groups of up to 256 output columns
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;
  const int pair_row0 = (blockIdx.x >> 1) * 128;
  const int row0 = pair_row0 + static_cast<int>(cta) * kB2Rows;
  const int t_begin = blockIdx.y * tiles_per_chunk;
  const int t_end = min(t_begin + tiles_per_chunk, n_tiles);
  const int n_my = t_end - t_begin;

  if (warp == kB2ProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_rows);
    tma_prefetch_desc(&tm_cols);
    tma_prefetch_desc(&tm_cols_mn);
    mbar_init(&bars.x_full, 1);
    for (int s = 0; s < kB2Stages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars.tmem_full[b], 1);
      mbar_init(&bars.tmem_empty[b], 2 * kB2EpiWarps);
      mbar_init(&bars.coef_full[b], 1);
      mbar_init(&bars.coef_empty[b], kB2EpiWarps);
    }
    mbar_init(&bars.g_full, 2 * kB2EpiWarps);
    mbar_init(&bars.g_empty, 1);
    mbar_init(&bars.acc_full, 1);
    fence_mbar_init();
  }
  if (warp == kB2AllocWarp) {
    tmem_alloc_pair(&bars.tmem_base, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp == kB2ProducerWarp) {
    // ------------------------------------------------------------ TMA producer (one thread per CTA)
    if (lane == 0) {
      if (!stream_x) {
        if (leader) mbar_arrive_expect_tx(&bars.x_full, static_cast<uint32_t>(2 * nk * kB2XChunkBytes));
        for (int kc = 0; kc < nk; ++kc)
          tma_load_2d_pair(smem_x + kc * kB2XChunkBytes, &tm_rows, &bars.x_full, kc * kB2BK, row0);
      }
      int ring_s = 0;
      uint32_t ring_ph = 0;
      // one ring stage = up to two 16 KB boxes from BOTH CTAs, all credited to the leader's full[s]
      auto acquire = [&](int bytes_per_cta) {
        const int s = ring_s;
        mbar_wait(&bars.empty[s], ring_ph ^ 1);
        if (leader) mbar_arrive_expect_tx(&bars.full[s], static_cast<uint32_t>(2 * bytes_per_cta));
        if (++ring_s == kB2Stages) {
          ring_s = 0;
          ring_ph ^= 1;
        }
        return s;
      };
      auto push_z = [&](int lt) {
        // this step's column coefficients {Lc, u', v', -} (4 KB) into this CTA's smem, one bulk copy
        const int cb = lt & 1;
        mbar_wait(&bars.coef_empty[cb], ((lt >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars.coef_full[cb], kB2CoefBytes);
        bulk_load_1d(smem_coef + cb * kB2CoefBytes, col_coef + static_cast<size_t>(t_begin + lt) * kB2TileN,
                     kB2CoefBytes, &bars.coef_full[cb]);
        const int col0 = (t_begin + lt) * kB2TileN + static_cast<int>(cta) * 128;
        if (stream_x) {  // one K chunk per stage: Y chunk in slot 0, X chunk (8 KB) in slot 1
          for (int kc = 0; kc < nk; ++kc) {
            const int s = acquire(kB2SlotBytes + kB2XChunkBytes);
            tma_load_2d_pair(smem_ring + s * kB2StageBytes, &tm_cols, &bars.full[s], kc * kB2BK, col0);
            tma_load_2d_pair(smem_ring + s * kB2StageBytes + kB2SlotBytes, &tm_rows, &bars.full[s], kc * kB2BK, row0);
          }
        } else {
          for (int kc = 0; kc < nk; kc += 2) {
            const int nb = min(2, nk - kc);
            const int s = acquire(nb * kB2SlotBytes);
            for (int b = 0; b < nb; ++b)
              tma_load_2d_pair(smem_ring + s * kB2StageBytes + b * kB2SlotBytes, &tm_cols, &bars.full[s],
                               (kc + b) * kB2BK, col0);
          }
        }
      };
      auto push_y = [&](int lt) {
        const int col0 = (t_begin + lt) * kB2TileN;
        const int n_units = 4 * ng;  // unit u = (64-column sub-tile js = u / ng, accumulator group g = u % ng)
        for (int p = 0; p < (kSplit ? 3 : 1); ++p) {  // split passes: (G1, Yh), (G1, Yl), (G2, Yh)
          const int p_d0 = (p == 1) ? d : 0;          // Yl sits at columns [d, 2 d) of Y' = (h | l | h)
          for (int u = 0; u < n_units; u += 2) {
            const int nb = min(2, n_units - u);
            const int s = acquire(nb * kB2SlotBytes);
            for (int b = 0; b < nb; ++b) {
              const int js = (u + b) / ng, g = (u + b) % ng;
              const int n_g = min(256, ds - 256 * g);
              // two {64 d, 64 j} boxes of the row-major Y: this CTA's d range of group g, 64 columns j of the step
              const int dbase = p_d0 + d0 + 256 * g + static_cast<int>(cta) * (n_g / 2);
              uint8_t* slot = smem_ring + s * kB2StageBytes + b * kB2SlotBytes;
              tma_load_2d_pair(slot, &tm_cols_mn, &bars.full[s], dbase, col0 + js * 64);
              tma_load_2d_pair(slot + kB2SlotBytes / 2, &tm_cols_mn, &bars.full[s], dbase + 64, col0 + js * 64);
            }
          }
        }
      };
      push_z(0);
      for (int lt = 0; lt < n_my; ++lt) {
        if (lt + 1 < n_my) push_z(lt + 1);
        push_y(lt);
      }
    }
  } else if (warp == kB2MmaWarp) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    // Whole warp converged, one elected lane issues (see scl_fwd2.cu).  Each barrier wait releases up to
    // 8 MMAs (two 16 KB boxes): a wait + commit round trip costs the issuing thread ~250 cycles, which at 4
    // MMAs of 65 cycles per wait left the tensor pipe under-fed.
    if (leader) {
      constexpr uint32_t idesc_z = umma_idesc_bf16(128, kB2TileN);
      if (!stream_x) mbar_wait_warp(&bars.x_full, 0);
      tc_fence_after();
      // Everything an MMA batch needs besides the data is computed BEFORE the wait for that data: the shared-memory
      // descriptors are base + offset in the 16-byte address field (all operands live below 256 KB), pinned in
      // registers by an empty asm so that the compiler cannot sink the arithmetic behind the barrier.  With the
      // arithmetic (~35 dependent uniform-datapath instructions) between the wait and the first UTCHMMA the tensor
      // pipe drained for ~150 cycles per batch (measured: 77 % pipe activity with all operands resident).
      const uint64_t x_desc0 = umma_desc_kmajor_sw128(smem_u32(smem_x));
      const uint64_t ring_k0 = umma_desc_kmajor_sw128(smem_u32(smem_ring));
      const uint64_t ring_mn0 = umma_desc_mnmajor_sw128(smem_u32(smem_ring), kB2SlotBytes / 2);
      const uint64_t g_desc0 = umma_desc_kmajor_sw128(smem_u32(smem_g));
      constexpr uint64_t kSlotUnits = kB2SlotBytes >> 4, kStageUnits = kB2StageBytes >> 4;
      const int ng_shift = ng - 1;  // ng is 1 or 2
      const uint32_t idesc_acc0 = umma_idesc_bf16(128, min(256, ds)) | kUmmaIdescBMnMajor;
      const uint32_t idesc_acc1 = umma_idesc_bf16(128, max(16, min(256, ds - 256))) | kUmmaIdescBMnMajor;
      int ring_s = 0;
      uint32_t ring_ph = 0;
      auto advance = [&]() {
        if (++ring_s == kB2Stages) {
          ring_s = 0;
          ring_ph ^= 1;
        }
      };
      auto issue_z = [&](int lt) {
        const int buf = lt & 1;
        mbar_wait_warp(&bars.tmem_empty[buf], ((lt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + kB2ZCol + buf * 128;
        const int kstep = stream_x ? 1 : 2;
        for (int kc = 0; kc < nk; kc += kstep, advance()) {
          const int nb = min(kstep, nk - kc);
          const int s = ring_s;
          uint64_t b_desc = ring_k0 + static_cast<uint64_t>(s) * kStageUnits;
          uint64_t a_desc = stream_x ? b_desc + kSlotUnits : x_desc0 + static_cast<uint64_t>(kc) * (kB2XChunkBytes >> 4);
          uint32_t bar_full = smem_u32(&bars.full[s]), bar_empty = smem_u32(&bars.empty[s]);
          asm volatile("" : "+l"(a_desc), "+l"(b_desc), "+r"(bar_full), "+r"(bar_empty));
          mbar_wait_warp_u32(bar_full, ring_ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_z, (kc | k) != 0 ? 1u : 0u);
            if (nb > 1) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_bf16_pair(d_tmem, a_desc + (kB2XChunkBytes >> 4) + 2 * k, b_desc + kSlotUnits + 2 * k, idesc_z, 1u);
            }
            tc_commit_pair_u32(bar_empty);
            if (kc + nb >= nk) tc_commit_pair(&bars.tmem_full[buf]);
          }
          __syncwarp();
        }
      };
      auto issue_acc = [&](int lt) {
        mbar_wait_warp(&bars.g_full, lt & 1);
        tc_fence_after();
        const int n_units = 4 * ng;
        constexpr int n_pass = kSplit ? 3 : 1;
        for (int p = 0; p < n_pass; ++p) {
          const uint64_t g_base = (kSplit && p == 2) ? g_desc0 + (kB2GBytes >> 4) : g_desc0;
          for (int u = 0; u < n_units; u += 2, advance()) {
            const int nb = min(2, n_units - u);
            const int s = ring_s;
            uint64_t b_desc = ring_mn0 + static_cast<uint64_t>(s) * kStageUnits;
            // unit u -> (64-column G sub-tile js, accumulator group g); u is even, so unit u + 1 is (js, 1) when
            // there are two groups and (js + 1, 0) when there is one
            const int js0 = u >> ng_shift;
            uint64_t a_desc0 = g_base + static_cast<uint64_t>(js0) * (kB2GSubBytes >> 4);
            uint64_t a_desc1 = g_base + static_cast<uint64_t>((u + 1) >> ng_shift) * (kB2GSubBytes >> 4);
            uint32_t bar_full = smem_u32(&bars.full[s]), bar_empty = smem_u32(&bars.empty[s]);
            asm volatile("" : "+l"(a_desc0), "+l"(a_desc1), "+l"(b_desc), "+r"(bar_full), "+r"(bar_empty));
            mbar_wait_warp_u32(bar_full, ring_ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t acc_first = (lt | p | js0) != 0 ? 1u : 0u;
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_bf16_pair(tmem_base, a_desc0 + 2 * k, b_desc + 128 * k, idesc_acc0, (acc_first | k) != 0 ? 1u : 0u);
              if (nb > 1) {
                // second unit: group 1 of the same sub-tile (fresh accumulator at js == 0) or group 0 of the next one
                const uint32_t t1 = tmem_base + (ng > 1 ? 128u : 0u);
                const uint32_t id1 = ng > 1 ? idesc_acc1 : idesc_acc0;
                const uint32_t acc1 = ng > 1 ? acc_first : 1u;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  tc_mma_bf16_pair(t1, a_desc1 + 2 * k, b_desc + kSlotUnits + 128 * k, id1, (acc1 | k) != 0 ? 1u : 0u);
              }
              tc_commit_pair_u32(bar_empty);
              if (p == n_pass - 1 && u + nb >= n_units) tc_commit_pair(&bars.g_empty);
            }
            __syncwarp();
          }
        }
      };
      issue_z(0);
      for (int lt = 0; lt < n_my; ++lt) {
        if (lt + 1 < n_my) issue_z(lt + 1);
        issue_acc(lt);
      }
      if (elect_one()) tc_commit_pair(&bars.acc_full);
      __syncwarp();
    }
  } else if (warp < kB2EpiWarps) {
    // ------------------------------------------------------------ epilogue: z -> G (bf16, swizzled smem)
    // 16 warps; each owns one 32-column chunk of the step: TMEM lane quadrant q = warp % 4 (2x2 layout:
    // lanes 64..127 hold the upper 128 columns), chunk hh = (warp - 4) / 4 of that half's 128 columns.
    const int q = warp & 3;
    const int hh = warp >> 2;
    const int r_loc = (q & 1) * 32 + lane;  // row within this CTA's 64
    const int n_half = q >> 1;
    const int col_in_step = n_half * 128 + hh * 32;  // first of this warp's 32 columns within the 256-wide step
    const int js = col_in_step >> 6;                 // 64-column G sub-tile
    const int c16 = (col_in_step & 63) >> 3;         // first 16-byte chunk inside the sub-tile row (0 or 4)
    const float s2 = __ldg(scale_log2_ptr);
    const float4 rc = __ldg(&row_coef[row0 + r_loc]);  // {Lr, u, v, own-column soft-target term}
    const float neg_lr = -rc.x;
    const int diag_col = diag0 + row0 + r_loc;
    const int warp_diag_lo = diag0 + row0 + (q & 1) * 32;
    const uint32_t coef_u32 = smem_u32(smem_coef);
    const uint32_t g_u32 = smem_u32(smem_g);
    const uint32_t g_off = static_cast<uint32_t>(js * kB2GSubBytes + (r_loc >> 3) * 1024 + (r_loc & 7) * 128);
    for (int lt = 0; lt < n_my; ++lt) {
      const int buf = lt & 1;
      const uint32_t par = (lt >> 1) & 1;
      mbar_wait_warp(&bars.coef_full[buf], par);
      mbar_wait_warp(&bars.tmem_full[buf], par);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kB2ZCol + buf * 128 + hh * 32, r);
      const uint32_t cf = coef_u32 + static_cast<uint32_t>(buf * kB2CoefBytes + col_in_step * 16);
      const int col0 = (t_begin + lt) * kB2TileN + col_in_step;
      const bool has_diag = (col0 + 32 > warp_diag_lo) && (col0 < warp_diag_lo + 32);  // warp-uniform
      const bool ragged = col0 + 32 > n_cols;                                           // warp-uniform
      tmem_ld_wait();
      // dL/dz of one element (fp32): row term + column term, own-column soft target, ragged-edge mask
      auto g_of = [&](int j) {
        const float z = __uint_as_float(r[j]);
        const float4 cc = lds_v4(cf + static_cast<uint32_t>(j * 16));  // smem broadcast (same address across the warp)
        const float p = ex2_approx(fmaf(z, s2, neg_lr));
        const float pc = ex2_approx(fmaf(z, s2, -cc.x));
        float g = fmaf(pc, fmaf(cc.z, z, cc.y), p * fmaf(rc.z, z, rc.y));
        if (has_diag) g -= (j == diag_col - col0) ? rc.w : 0.f;
        if (ragged) g = (col0 + j < n_cols) ? g : 0.f;
        return g;
      };
      if constexpr (kSplit != 0) {
        // fp32-accurate mode: G = G1 + G2 (two bf16 tiles).  z already sits in registers, so TMEM goes back first;
        // the arithmetic then runs 8 columns at a time straight into the two 16-byte stores (low register pressure;
        // the step is dominated by its 3x MMA work, so not overlapping the math with the g_empty wait costs nothing).
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&bars.tmem_empty[buf]);
          else mbar_arrive_remote(&bars.tmem_empty[buf], 0);
        }
        mbar_wait_warp(&bars.g_empty, (lt & 1) ^ 1);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const float g0 = g_of(ch * 8 + jj * 2), g1 = g_of(ch * 8 + jj * 2 + 1);
            hi[jj] = pack_bf16x2(g0, g1);
            lo[jj] = pack_bf16x2(g0 - __uint_as_float(hi[jj] << 16), g1 - __uint_as_float(hi[jj] & 0xffff0000u));
          }
          const uint32_t off = g_off + static_cast<uint32_t>(((c16 + ch) ^ (r_loc & 7)) * 16);
          sts_v4(g_u32 + off, hi[0], hi[1], hi[2], hi[3]);
          sts_v4(g_u32 + kB2GBytes + off, lo[0], lo[1], lo[2], lo[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.coef_empty[buf]);
      } else {
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) packed[j >> 1] = pack_bf16x2(g_of(j), g_of(j + 1));
        // z is in registers now: hand the TMEM buffer back before the (possibly waiting) G write
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars.coef_empty[buf]);
          if (leader) mbar_arrive(&bars.tmem_empty[buf]);
          else mbar_arrive_remote(&bars.tmem_empty[buf], 0);
        }
        // single G buffer: the second GEMM of the previous step must have consumed it
        mbar_wait_warp(&bars.g_empty, (lt & 1) ^ 1);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)  // K-major SWIZZLE_128B: 16-byte chunk XOR (row % 8)
          sts_v4(g_u32 + g_off + static_cast<uint32_t>(((c16 + ch) ^ (r_loc & 7)) * 16), packed[ch * 4 + 0],
                 packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&bars.g_full);
        else mbar_arrive_remote(&bars.g_full, 0);
      }
    }
    // ---- drain this CTA's 64-row slice of the dX accumulators (warp hh takes chunk hh of each group)
    mbar_wait_warp(&bars.acc_full, 0);
    tc_fence_after();
    float* out_row = dx_partial + (static_cast<size_t>(blockIdx.y) * m_pad + row0 + r_loc) * d;
    for (int g = 0; g < ng; ++g) {
      const int n_g = min(256, ds - 256 * g);
      if (hh * 32 < n_g / 2) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 128 + hh * 32, r);
        tmem_ld_wait();
        float* dst = out_row + d0 + 256 * g + n_half * (n_g / 2) + hh * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dst + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == kB2AllocWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// D slices of the backward (TMEM holds dX[64 x <= 512] per SM): the smallest count that cuts D into equal slices of
// at most 512 columns, each a multiple of 64 (768 -> 2 x 384, 1024 -> 2 x 512, 640 -> 2 x 320, 1152 -> 3 x 384,
// 1280 -> 4 x 320, 1536 -> 3 x 512); 0 = no such cut
int bwd_pair_d_slices(int d) {
  if (d <= 512) return 1;
  for (int s = 2; s <= 8; ++s)
    if (d % s == 0 && d / s <= 512 && (d / s) % 64 == 0) return s;
  return 0;
}

size_t bwd_pair_smem_bytes(int d, int split) {
  const size_t x_block = (split || d > 512) ? 0 : static_cast<size_t>(d / kB2BK) * kB2XChunkBytes;
  return 1024 + x_block + (split ? 2 : 1) * kB2GBytes + kB2Stages * kB2StageBytes + 2 * kB2CoefBytes;
}

// Column chunks of the backward grid.  Every chunk adds one [m_rows, D] fp32 slab that the kernel writes and
// bwd_gather reads back (~4 us at 4096 x 512 = 0.9 of a tile-step): charged per chunk, scaled by the slab size.
int bwd_pair_pick_chunks(int m_rows, int n_cols, int d, int num_sms, int* tiles_per_chunk) {
  const int pairs = (m_rows + 127) / 128 * max(1, bwd_pair_d_slices(d));
  const int n_tiles = (n_cols + kB2TileN - 1) / kB2TileN;
  const double slab_cost = 0.9 * (static_cast<double>(m_rows) / 4096.0) * (static_cast<double>(d) / 512.0);
  return pick_chunks_balanced(pairs, n_tiles, num_sms / 2, 2, slab_cost, tiles_per_chunk);
}

template <int kSplit>
static cudaError_t launch_bwd_rows_pair_t(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols,
                                          const CUtensorMap& tm_cols_mn, int m_rows, int n_cols, int d, int chunks,
                                          int tiles_per_chunk, int m_pad, int diag0, const float* scale_log2,
                                          const float4* row_coef, const float4* col_coef, float* dx_partial,
                                          cudaStream_t stream) {
  const size_t smem = bwd_pair_smem_bytes(d, kSplit);
  // opt in to > 48 KB dynamic shared memory (sticky per device; set on every launch so the library keeps no state --
  // except under stream capture, where the eager warm-up launches have already set it)
  cudaError_t err = set_max_dynamic_smem(bwd_rows_pair_kernel<kSplit>, 231424, stream);
  if (err != cudaSuccess) return err;
  const int pairs = (m_rows + 127) / 128;
  const int n_tiles = (n_cols + kB2TileN - 1) / kB2TileN;
  const int d_slices = bwd_pair_d_slices(d);
  dim3 grid(2 * pairs, chunks, d_slices);
  bwd_rows_pair_kernel<kSplit><<<grid, kB2Threads, smem, stream>>>(tm_rows, tm_cols, tm_cols_mn, m_rows, n_cols, d,
                                                                   d_slices, n_tiles, tiles_per_chunk, m_pad, diag0,
                                                                   scale_log2, row_coef, col_coef, dx_partial);
  return cudaGetLastError();
}

cudaError_t launch_bwd_rows_pair(const CUtensorMap& tm_rows, const CUtensorMap& tm_cols, const CUtensorMap& tm_cols_mn,
                                 int m_rows, int n_cols, int d, int chunks, int tiles_per_chunk, int m_pad, int diag0,
                                 const float* scale_log2, const float4* row_coef, const float4* col_coef,
                                 float* dx_partial, int split, cudaStream_t stream) {
  if (split)
    return launch_bwd_rows_pair_t<1>(tm_rows, tm_cols, tm_cols_mn, m_rows, n_cols, d, chunks, tiles_per_chunk, m_pad,
                                     diag0, scale_log2, row_coef, col_coef, dx_partial, stream);
  return launch_bwd_rows_pair_t<0>(tm_rows, tm_cols, tm_cols_mn, m_rows, n_cols, d, chunks, tiles_per_chunk, m_pad, diag0,
                                   scale_log2, row_coef, col_coef, dx_partial, stream);
}

}  // namespace scl
