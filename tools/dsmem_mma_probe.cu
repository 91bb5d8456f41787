// Dev probe (not part of the product): may a tcgen05.mma read a shared-memory operand that ANOTHER SM of the cluster
// wrote with st.shared::cluster, and which fences does the hand-over need?  Gate for the split-role backward
// (profiles/ROUND2_PLAN.md): there the SM that forms G = dL/dz stores the bf16 tile straight into the shared memory
// of the SM that runs the gradient GEMM.
//
// Cluster of 2 CTAs.  CTA 0 (producer): 128 threads, thread = row, write A[128 x 64] bf16 (K-major, SWIZZLE_128B
// layout) into CTA 1's shared memory, fence (variant), remote mbarrier.arrive.release.cluster.  CTA 1 (consumer): one
// thread waits, fence (variant), issues 4 x tcgen05.mma (cta_group::1, M = 128, N = 64, K = 16) against a local B tile,
// commits; 128 threads read the accumulator back and compare with the exact integer result.  The data change every
// round, so a stale or partial tile shows up as a mismatch.
//   variant 0: producer fence.proxy.async (all state spaces)                         <- what the design assumes
//   variant 1: producer fence.proxy.async.shared::cluster
//   variant 2: no producer fence, consumer fence.proxy.async after the wait
//   variant 3: producer fence (as 0) + consumer fence
//   variant 4: no proxy fence at all (negative control; may pass by luck)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spatial_clip_b200/csrc tools/dsmem_mma_probe.cu -o /tmp/dsmem_mma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "scl_ptx.cuh"
using namespace scl;

constexpr int kRows = 128, kN = 64, kK = 64;
constexpr int kABytes = kRows * kK * 2;  // 16 KB
constexpr int kBBytes = kN * kK * 2;     // 8 KB
constexpr int kThreads = 160;            // warps 0..3: data, warp 4: MMA issuer / TMEM allocator

struct Bars {
  uint64_t a_full;   // consumer: 4 remote arrivals (producer warps)
  uint64_t a_empty;  // producer: 1 remote arrival (consumer, after the accumulator has been read)
  uint64_t mma_done; // consumer: tcgen05.commit
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_cluster() {
  asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
}
__device__ __forceinline__ int a_val(int row, int k, int round) { return ((row + k + round) % 7) - 3; }
__device__ __forceinline__ int b_val(int n, int k) { return ((3 * n + k) % 5) - 2; }
// byte offset of the 16-byte chunk `chunk` (8 bf16) of row `r` in a K-major SWIZZLE_128B tile of 64-element rows
__device__ __forceinline__ uint32_t sw128_off(int r, int chunk) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) * 16));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
probe(int rounds, int variant, int* mismatches, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;            // consumer: written remotely
  uint8_t* smem_b = smem + kABytes;  // consumer: written locally, once
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const bool consumer = cta == 1;
  if (threadIdx.x == 0) {
    mbar_init(&bars.a_full, 4);
    mbar_init(&bars.a_empty, 1);
    mbar_init(&bars.mma_done, 1);
    fence_mbar_init();
  }
  if (consumer && warp == 4) {
    tmem_alloc(&bars.tmem_base, 64);
    tmem_relinquish();
  }
  if (consumer && threadIdx.x < kN) {  // B[n][k], one row per thread, same swizzled layout
    const int n = threadIdx.x;
    for (int ch = 0; ch < 8; ++ch) {
      uint32_t w[4];
      for (int e = 0; e < 4; ++e)
        w[e] = pack_bf16x2(static_cast<float>(b_val(n, ch * 8 + 2 * e)), static_cast<float>(b_val(n, ch * 8 + 2 * e + 1)));
      *reinterpret_cast<uint4*>(smem_b + sw128_off(n, ch)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;
  const long long t0 = clock64();

  if (!consumer) {
    if (warp < 4) {
      const int row = threadIdx.x;
      const uint32_t a_remote = map_to_cta(smem_u32(smem_a), 1);
      for (int r = 0; r < rounds; ++r) {
        if (r > 0) mbar_wait(&bars.a_empty, (r - 1) & 1);
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t w[4];
          for (int e = 0; e < 4; ++e)
            w[e] = pack_bf16x2(static_cast<float>(a_val(row, ch * 8 + 2 * e, r)),
                               static_cast<float>(a_val(row, ch * 8 + 2 * e + 1, r)));
          st_cluster_v4(a_remote + sw128_off(row, ch), w[0], w[1], w[2], w[3]);
        }
        if (variant == 0 || variant == 3) fence_proxy_async_all();
        if (variant == 1) fence_proxy_async_cluster();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&bars.a_full, 1);
      }
    }
  } else if (warp == 4) {
    constexpr uint32_t idesc = umma_idesc_bf16(kRows, kN);
    for (int r = 0; r < rounds; ++r) {
      mbar_wait(&bars.a_full, r & 1);
      if (variant == 2 || variant == 3) fence_proxy_async_all();
      tc_fence_after();
      if (elect_one()) {
        const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(smem_a));
        const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(smem_b));
#pragma unroll
        for (int k = 0; k < kK / 16; ++k) tc_mma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0 ? 1u : 0u);
        tc_commit(&bars.mma_done);
      }
      __syncwarp();
      named_barrier_sync(1, kThreads);  // data warps have read the accumulator: the next MMA may overwrite it
    }
  } else {
    const int row = threadIdx.x;  // TMEM lane == row (M = 128, cta_group::1)
    int bad = 0;
    for (int r = 0; r < rounds; ++r) {
      mbar_wait(&bars.mma_done, r & 1);
      tc_fence_after();
      uint32_t acc[2][32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), acc[0]);
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + 32, acc[1]);
      tmem_ld_wait();
      for (int n = 0; n < kN; ++n) {
        int want = 0;
        for (int k = 0; k < kK; ++k) want += a_val(row, k, r) * b_val(n, k);
        bad += (__uint_as_float(acc[n >> 5][n & 31]) != static_cast<float>(want)) ? 1 : 0;
      }
      tc_fence_before();
      named_barrier_sync(1, kThreads);
      if (threadIdx.x == 0) mbar_arrive_remote(&bars.a_empty, 0);  // the producer may overwrite A
    }
    if (bad) atomicAdd(mismatches, bad);
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (consumer && warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

int main() {
  const int smem = kABytes + kBBytes + 2048, rounds = 2000;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int* mism;
  long long* cyc;
  cudaMalloc(&mism, sizeof(int));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const char* names[5] = {"producer fence.proxy.async", "producer fence.proxy.async.shared::cluster",
                          "consumer fence.proxy.async only", "producer + consumer fences", "no proxy fence (control)"};
  for (int grid : {2, 148}) {
    for (int variant = 0; variant < 5; ++variant) {
      cudaMemset(mism, 0, sizeof(int));
      probe<<<grid, kThreads, smem>>>(rounds, variant, mism, cyc);
      const cudaError_t e = cudaDeviceSynchronize();
      int h = -1;
      long long hc[148] = {0};
      cudaMemcpy(&h, mism, sizeof(int), cudaMemcpyDeviceToHost);
      cudaMemcpy(hc, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      printf("grid=%3d variant %d (%-44s) err=%d (%s)  mismatching elements: %d of %lld  round trip %.0f clk\n", grid, variant,
             names[variant], int(e), cudaGetErrorString(e), h, 1LL * (grid / 2) * rounds * kRows * kN,
             double(hc[0]) / rounds);
      if (e != cudaSuccess) return 1;  // a trapped kernel poisons the context
    }
  }
  cudaFree(mism);
  cudaFree(cyc);
  return 0;
}
