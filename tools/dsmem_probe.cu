// Dev probe (not part of the product): how fast can one SM of a CTA pair hand a 64 KB tile to its peer's shared
// memory?  Decides whether the split-role backward (one SM keeps z, the other dX, G crosses through DSMEM --
// profiles/r1_smem_port_accounting.md) is feasible: it needs 64 KB per ~5600 cycles (>= 12 B/clk) with headroom.
//   A: remote vector stores  st.shared::cluster.v4.b32   (16 warps, 16 B per lane)
//   B: bulk copies           cp.async.bulk.shared::cluster.shared::cta  (2 x 32 KB per round, mbarrier completion)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spatial_clip_b200/csrc tools/dsmem_probe.cu -o /tmp/dsmem_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "scl_ptx.cuh"
using namespace scl;

constexpr int kTile = 64 * 1024;
constexpr int kThreads = 512;

__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bulk_smem_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(bar_cluster) : "memory");
}

struct Bars { uint64_t rx; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
probe(int rounds, int both_ways, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* src = smem;          // 64 KB, local source tile
  uint8_t* dst = smem + kTile;  // 64 KB, written by the peer
  const uint32_t cta = cluster_ctarank();
  const uint32_t peer = cta ^ 1;
  for (int i = threadIdx.x; i < 2 * kTile / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(i, 1, 2, 3);
  if (threadIdx.x == 0) { mbar_init(&bars.rx, 1); fence_mbar_init(); }
  __syncthreads();
  cluster_sync_all();
  const bool sender = cta == 0 || both_ways != 0;

  // ---- A: remote vector stores
  const uint32_t dst_remote = map_to_cta(smem_u32(dst), peer);
  long long t0 = clock64();
  if (sender) {
    for (int r = 0; r < rounds; ++r)
      for (int off = threadIdx.x * 16; off < kTile; off += kThreads * 16)
        st_cluster_v4(dst_remote + off, r, off, 2, 3);
  }
  cluster_sync_all();  // release/acquire at cluster scope: all remote stores have landed
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x * 4 + 0] = t1 - t0;

  // ---- B: bulk copies, completion on the receiver's mbarrier
  const bool receiver = cta == 1 || both_ways != 0;
  const uint32_t bar_remote = map_to_cta(smem_u32(&bars.rx), peer);
  fence_proxy_async();
  cluster_sync_all();
  t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if (receiver && threadIdx.x == 0) mbar_arrive_expect_tx(&bars.rx, kTile);
    cluster_sync_all();  // the receiver's expect_tx is armed before the sender issues (keeps the probe simple)
    if (sender && threadIdx.x == 0) {
      bulk_smem_to_peer(dst_remote, smem_u32(src), kTile / 2, bar_remote);
      bulk_smem_to_peer(dst_remote + kTile / 2, smem_u32(src) + kTile / 2, kTile / 2, bar_remote);
    }
    if (receiver) mbar_wait(&bars.rx, r & 1);
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x * 4 + 1] = t1 - t0;
  // cost of the per-round cluster barrier alone, to subtract
  t0 = clock64();
  for (int r = 0; r < rounds; ++r) cluster_sync_all();
  t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x * 4 + 2] = t1 - t0;
  cluster_sync_all();
}

// how many clusters of `csize` CTAs (608 threads, 227 KB dynamic smem each: the pair kernels' footprint) fit at once
__global__ void __launch_bounds__(608, 1) footprint_kernel(int* p) {
  extern __shared__ uint8_t s[];
  if (p != nullptr && threadIdx.x == 0) p[blockIdx.x] = s[0];
}
static int max_clusters(int csize) {
  const int smem_bytes = 231424;
  cudaFuncSetAttribute(footprint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(footprint_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize * 64);
  cfg.blockDim = dim3(608);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = csize;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, footprint_kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return n;
}

int main() {
  for (int cs : {1, 2, 4, 8})
    printf("co-resident clusters of %d CTAs (608 threads, 226 KB smem each): %d  -> %d SMs busy\n", cs, max_clusters(cs),
           cs * max_clusters(cs));
  const int rounds = 64, smem = 2 * kTile + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMalloc(&out, 148 * 4 * sizeof(long long));
  for (int grid : {2, 148}) {
    for (int both : {0, 1}) {
      cudaMemset(out, 0, 148 * 4 * sizeof(long long));
      for (int it = 0; it < 2; ++it) probe<<<grid, kThreads, smem>>>(rounds, both, out);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148 * 4];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      long long a = 0, b = 0, c = 0;
      for (int i = 0; i < grid; ++i) { a = h[i * 4] > a ? h[i * 4] : a; b = h[i * 4 + 1] > b ? h[i * 4 + 1] : b; c = h[i * 4 + 2] > c ? h[i * 4 + 2] : c; }
      printf("grid=%3d both_ways=%d err=%d | A remote st.v4: %8.0f clk / 64 KB = %5.1f B/clk | B bulk copy: %8.0f clk / 64 KB "
             "(minus barrier %6.0f) = %5.1f B/clk\n", grid, both, int(e), double(a) / rounds, kTile * double(rounds) / a,
             double(b) / rounds, double(c) / rounds, kTile * double(rounds) / (b - c > 0 ? b - c : 1));
    }
  }
  cudaFree(out);
  return 0;
}
