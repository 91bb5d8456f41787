// Dev probe (not part of the product): issue-rate of tcgen05.mma shapes with operands resident in smem.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I spatial_clip_b200/csrc tools/umma_probe.cu -o /tmp/umma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "scl_ptx.cuh"
using namespace scl;

struct Bars { uint64_t done; uint32_t tmem_base; };

template <int CG>
__global__ void __launch_bounds__(128, 1) probe(int m, int n, int reps, int kblocks, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ Bars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta = 0;
  if (CG == 2) cta = cluster_ctarank();
  // zero the operand area so values are finite
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && lane == 0) { mbar_init(&bars.done, 1); fence_mbar_init(); }
  if (warp == 1) { if (CG == 2) { tmem_alloc_pair(&bars.tmem_base, 512); tmem_relinquish_pair(); } else { tmem_alloc(&bars.tmem_base, 512); tmem_relinquish(); } }
  fence_proxy_async();
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tb = bars.tmem_base;
  if (warp == 0 && lane == 0 && cta == 0) {
    const uint32_t idesc = umma_idesc_bf16(m, n);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t a = a0 + (kb % 4) * 16384, b = b0 + (kb % 4) * 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (CG == 2) tc_mma_bf16_pair(tb, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), idesc, 1u);
          else tc_mma_bf16(tb, umma_desc_kmajor_sw128(a + k * 32), umma_desc_kmajor_sw128(b + k * 32), idesc, 1u);
        }
      }
    }
    if (CG == 2) tc_commit_pair(&bars.done); else tc_commit(&bars.done);
    mbar_wait(&bars.done, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  if (CG == 2 && cta == 1 && warp == 0 && lane == 0) mbar_wait(&bars.done, 0);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) { tc_fence_after(); if (CG == 2) tmem_dealloc_pair(tb, 512); else tmem_dealloc(tb, 512); }
}

template <int CG>
void run(const char* name, int m, int n, int grid) {
  long long* out; cudaMalloc(&out, grid * sizeof(long long)); cudaMemset(out, 0, grid * sizeof(long long));
  const int smem = 200 * 1024, reps = 64, kblocks = 8;
  cudaFuncSetAttribute(probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {unsigned(CG), 1, 1};
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) cudaLaunchKernelEx(&cfg, probe<CG>, m, n, reps, kblocks, out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[512]; cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; i += CG) mx = h[i] > mx ? h[i] : mx;
  const double ninstr = double(reps) * kblocks * 4;
  const double macs_per_sm = double(m) * n * 16 / CG;
  printf("%-34s grid=%3d err=%d  clk/instr=%7.1f  MAC/clk/SM=%7.1f\n", name, grid, int(e), mx / ninstr, macs_per_sm / (mx / ninstr));
  cudaFree(out);
}

int main() {
  for (int grid : {2, 148}) {
    run<1>("cta1 M=128 N=256", 128, 256, grid);
    run<1>("cta1 M=128 N=128", 128, 128, grid);
    run<1>("cta1 M=64  N=256", 64, 256, grid);
    run<2>("cta2 M=256 N=256", 256, 256, grid);
    run<2>("cta2 M=256 N=128", 256, 128, grid);
    run<2>("cta2 M=128 N=256", 128, 256, grid);
    run<2>("cta2 M=128 N=128", 128, 128, grid);
  }
  return 0;
}
