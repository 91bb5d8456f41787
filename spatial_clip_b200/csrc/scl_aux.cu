// HBM-bound / integer side passes of the contrastive-loss path (plain CUDA, coalesced, no tensor cores):
//   * cast (+ optional L2-normalise) to bf16, row-major and transposed copies
//   * soft-target ("positives") builder: tile-id -> global column hash, ELL lists per row
//   * row finalisation: merge online-softmax partials, positive logits  sum_k q_k <x_i, y_col_k>
//   * scalar reductions (loss, gap, d/ds) and backward coefficient vectors
//   * backward finish: sum column-chunk partials, sparse -Q terms, cast
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "scl_kernels.h"
#include "scl_ptx.cuh"

namespace scl {

// =====================================================================================
// cast / normalise / transpose
// =====================================================================================
template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, size_t i) { return __half2float(p[i]); }

// One CTA handles a [64 rows x d] slab in [64 x 64] tiles: 16-byte global loads (4 fp32 / 8 bf16 per thread),
// optional per-row 1/||x||, 8-byte row-major bf16 stores, and the transposed copy y_t[d][ld_t] written as
// 16-byte vectors of 8 consecutive rows assembled from a padded smem tile.
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename T>
__global__ void __launch_bounds__(256) cast_bf16_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        __nv_bfloat16* __restrict__ y_t, int rows, int d, int ld_t,
                                                        int normalize) {
  __shared__ float inv_norm[64];
  // [row][col + 8 * (row / 8)]: 240-byte pitch plus a per-8-row-group skew makes the 8-row column gathers of the
  // transposed store hit 8 different banks
  __shared__ __nv_bfloat16 tile[64][120];
  const int r0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (normalize) {
    for (int rr = warp; rr < 64; rr += 8) {
      const int r = r0 + rr;
      float ss = 0.f;
      if (r < rows) {
        for (int c = lane * 4; c < d; c += 128) {
          float v[4];
          load4(x + static_cast<size_t>(r) * d + c, v);
          ss = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], fmaf(v[3], v[3], ss))));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) inv_norm[rr] = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (open_clip model.py:328,345)
    }
  } else if (threadIdx.x < 64) {
    inv_norm[threadIdx.x] = 1.f;
  }
  __syncthreads();
  // thread -> (row rr = tid / 16 + 16 * pass, 4 columns at 4 * (tid % 16))
  const int tc = (threadIdx.x & 15) * 4;
  const int tr = threadIdx.x >> 4;
  // transposed stores: thread -> (column cc = tid / 8 + 32 * pass, 8 rows at 8 * (tid % 8))
  const int sr = (threadIdx.x & 7) * 8;
  const int sc = threadIdx.x >> 3;
  const bool full_rows = r0 + 64 <= rows && (ld_t % 8) == 0;
  for (int c0 = 0; c0 < d; c0 += 64) {
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int rr = tr + 16 * pass;
      const int r = r0 + rr;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < rows) load4(x + static_cast<size_t>(r) * d + c0 + tc, v);
      const float s = inv_norm[rr];
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0] * s, v[1] * s);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2] * s, v[3] * s);
      uint2 packed;
      packed.x = *reinterpret_cast<const uint32_t*>(&lo);
      packed.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(&tile[rr][tc + 8 * (rr >> 3)]) = packed;
      if (y != nullptr && r < rows) *reinterpret_cast<uint2*>(y + static_cast<size_t>(r) * d + c0 + tc) = packed;
    }
    if (y_t != nullptr) {
      __syncthreads();
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int cc = sc + 32 * pass;
        __nv_bfloat16 col[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) col[k] = tile[sr + k][cc + sr];
        __nv_bfloat16* dst = y_t + static_cast<size_t>(c0 + cc) * ld_t + r0 + sr;
        if (full_rows) {
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(col);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (r0 + sr + k < rows) dst[k] = col[k];
        }
      }
      __syncthreads();
    }
  }
}

cudaError_t launch_cast_bf16(const void* x, int src_dtype, void* y, void* y_t, int rows, int d, int ld_t,
                             int normalize, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  const int grid = (rows + 63) / 64;
  auto yb = static_cast<__nv_bfloat16*>(y);
  auto ytb = static_cast<__nv_bfloat16*>(y_t);
  if (src_dtype == 0)
    cast_bf16_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), yb, ytb, rows, d, ld_t, normalize);
  else if (src_dtype == 1)
    cast_bf16_kernel<__nv_bfloat16>
        <<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), yb, ytb, rows, d, ld_t, normalize);
  else
    cast_bf16_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(x), yb, ytb, rows, d, ld_t, normalize);
  return cudaGetLastError();
}

// scalars[0] = s_eff = min(s, cap)  (forward value of the straight-through cap, losses.py:73-76)
// scalars[1] = s_eff * log2(e), scalars[2] = s
__global__ void prep_scalars_kernel(const float* __restrict__ logit_scale, float cap, float* __restrict__ scalars) {
  const float s = logit_scale[0];
  const float s_eff = cap > 0.f ? fminf(s, cap) : s;
  scalars[0] = s_eff;
  scalars[1] = s_eff * kLog2e;
  scalars[2] = s;
}
cudaError_t launch_prep_scalars(const float* logit_scale, float cap, float* scalars, cudaStream_t stream) {
  prep_scalars_kernel<<<1, 1, 0, stream>>>(logit_scale, cap, scalars);
  return cudaGetLastError();
}

// =====================================================================================
// positives builder (integer path, bit-exact vs losses.py:91-108)
// =====================================================================================
constexpr long long kEmptyKey = static_cast<long long>(0x8000000000000000ull);

__host__ __device__ inline uint32_t hash_capacity(int n) {
  uint32_t cap = 64;
  while (cap < 2u * static_cast<uint32_t>(n)) cap <<= 1;
  return cap;
}
size_t positives_hash_bytes(int n_global) {
  const size_t cap = hash_capacity(n_global);
  return (cap + 1) * sizeof(long long) + (cap + 1) * sizeof(int);
}
__device__ __forceinline__ uint32_t mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return static_cast<uint32_t>(k);
}
__global__ void hash_clear_kernel(long long* keys, int* vals, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= cap) {
    keys[i] = kEmptyKey;
    vals[i] = -1;
  }
}
// "last index wins" (dict comprehension, losses.py:92-93) == max index per key
__global__ void hash_insert_kernel(const long long* __restrict__ ids, int n, long long* keys, int* vals, uint32_t cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long id = ids[i];
  if (id == kEmptyKey) {  // the one key that collides with the sentinel lives in the extra slot
    atomicMax(&vals[cap], i);
    return;
  }
  uint32_t slot = mix64(static_cast<unsigned long long>(id)) & (cap - 1);
  while (true) {
    const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[slot]),
                                              static_cast<unsigned long long>(kEmptyKey),
                                              static_cast<unsigned long long>(id));
    if (prev == static_cast<unsigned long long>(kEmptyKey) || prev == static_cast<unsigned long long>(id)) {
      atomicMax(&vals[slot], i);
      return;
    }
    slot = (slot + 1) & (cap - 1);
  }
}
__device__ __forceinline__ int hash_lookup(long long id, const long long* __restrict__ keys,
                                           const int* __restrict__ vals, uint32_t cap) {
  if (id == kEmptyKey) return vals[cap];
  uint32_t slot = mix64(static_cast<unsigned long long>(id)) & (cap - 1);
  while (true) {
    const long long k = keys[slot];
    if (k == id) return vals[slot];
    if (k == kEmptyKey) return -1;
    slot = (slot + 1) & (cap - 1);
  }
}
// One thread per local row.  Slot 0 is the row's own column rank*B_l + i with weight 1 (losses.py:94-98);
// neighbour slots are visited in k order, skipped when alpha*scale <= 0 or the id is not in the global
// batch, and merged (fp32 +=) into the slot already holding the same column (losses.py:100-108).
__global__ void build_ell_kernel(const long long* __restrict__ nbr_ids, const float* __restrict__ nbr_alpha,
                                 int b_local, int k, float alpha_scale, int rank, const long long* __restrict__ keys,
                                 const int* __restrict__ vals, uint32_t cap, int* __restrict__ pos_col,
                                 float* __restrict__ pos_w, float* __restrict__ pos_q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b_local) return;
  const int kp1 = k + 1;
  int* col = pos_col + static_cast<size_t>(i) * kp1;
  float* w = pos_w + static_cast<size_t>(i) * kp1;
  float* q = pos_q + static_cast<size_t>(i) * kp1;
  col[0] = rank * b_local + i;
  w[0] = 1.0f;
  int cnt = 1;
  for (int s = 0; s < k; ++s) {
    float a = __fmul_rn(nbr_alpha[static_cast<size_t>(i) * k + s], alpha_scale);
    a = fmaxf(a, 0.f);
    if (!(a > 0.f)) continue;
    const int c = hash_lookup(nbr_ids[static_cast<size_t>(i) * k + s], keys, vals, cap);
    if (c < 0) continue;
    int hit = -1;
    for (int t = 0; t < cnt; ++t)
      if (col[t] == c) hit = t;
    if (hit >= 0) {
      w[hit] = __fadd_rn(w[hit], a);
    } else {
      col[cnt] = c;
      w[cnt] = a;
      ++cnt;
    }
  }
  float tot = 0.f;
  for (int t = 0; t < cnt; ++t) tot = __fadd_rn(tot, w[t]);
  const float inv = 1.f / fmaxf(tot, 1e-12f);  // F.normalize(p=1) eps, losses.py:110-111
  for (int t = 0; t < cnt; ++t) q[t] = w[t] * inv;
  for (int t = cnt; t < kp1; ++t) {
    col[t] = -1;
    w[t] = 0.f;
    q[t] = 0.f;
  }
}

cudaError_t launch_build_positives(const int64_t* all_ids, int n_global, const int64_t* nbr_ids,
                                   const float* nbr_alpha, int b_local, int k, float alpha_scale, int rank,
                                   void* hash_ws, size_t hash_ws_bytes, int32_t* pos_col, float* pos_w, float* pos_q,
                                   cudaStream_t stream) {
  const uint32_t cap = hash_capacity(n_global);
  long long* keys = static_cast<long long*>(hash_ws);
  int* vals = reinterpret_cast<int*>(keys + cap + 1);
  if (k > 0) {
    if (hash_ws == nullptr || hash_ws_bytes < positives_hash_bytes(n_global)) return cudaErrorInvalidValue;
    hash_clear_kernel<<<(cap + 1 + 255) / 256, 256, 0, stream>>>(keys, vals, cap);
    hash_insert_kernel<<<(n_global + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const long long*>(all_ids),
                                                                   n_global, keys, vals, cap);
  }
  build_ell_kernel<<<(b_local + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const long long*>(nbr_ids), nbr_alpha,
                                                             b_local, k, alpha_scale, rank, keys, vals, cap, pos_col,
                                                             pos_w, pos_q);
  return cudaGetLastError();
}

// =====================================================================================
// row finalisation
// =====================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dot_bf16_row(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                              int d, int lane) {
  float acc = 0.f;
  for (int c = lane * 2; c < d; c += 64) {
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + c));
    const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(b + c));
    acc = fmaf(fa.x, fb.x, acc);
    acc = fmaf(fa.y, fb.y, acc);
  }
  return warp_sum(acc);
}
// warp per local row: stats = {L2 = m + log2(S0)  (LSE in log2 units), mu = S1/S0, var = S2/S0 - mu^2,
//                              zq = sum_k q_k <x_i, y_col_k>}
__global__ void __launch_bounds__(256) row_finalize_kernel(const float4* __restrict__ partial, int n_slots, int m_pad,
                                                           int m_rows, int d, const __nv_bfloat16* __restrict__ x_rows,
                                                           const __nv_bfloat16* __restrict__ y_all,
                                                           const int* __restrict__ pos_col,
                                                           const float* __restrict__ pos_q, int kp1,
                                                           float4* __restrict__ row_stats) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m_rows) return;
  // lanes split the slots; the merge is a max-rescaled sum, combined across lanes in a fixed butterfly order
  float m = -INFINITY;
  for (int s = lane; s < n_slots; s += 32) m = fmaxf(m, partial[static_cast<size_t>(s) * m_pad + row].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int s = lane; s < n_slots; s += 32) {
    const float4 p = partial[static_cast<size_t>(s) * m_pad + row];
    const float w = (p.x == -INFINITY) ? 0.f : exp2f(p.x - m);
    s0 = fmaf(p.y, w, s0);
    s1 = fmaf(p.z, w, s1);
    s2 = fmaf(p.w, w, s2);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float mu = s1 / s0;
  const float var = s2 / s0 - mu * mu;
  float zq = 0.f;
  for (int t = 0; t < kp1; ++t) {
    const int c = pos_col[static_cast<size_t>(row) * kp1 + t];
    if (c < 0) continue;  // warp-uniform
    const float z = dot_bf16_row(x_rows + static_cast<size_t>(row) * d, y_all + static_cast<size_t>(c) * d, d, lane);
    zq = fmaf(pos_q[static_cast<size_t>(row) * kp1 + t], z, zq);
  }
  if (lane == 0) row_stats[row] = make_float4(m + log2f(s0), mu, var, zq);
}
// Developer variant (SCL_AUX_V2=1; the kernel above measured 24 % of the copy bandwidth, profiles/r1_hbm_passes.md,
// because every dot product walks its two 1 KB rows in eight dependent 4-byte steps per lane): the same contract
// with 16-byte loads, all loads of a dot product in flight together and the slot list of the row read once.
// Summation order inside a dot product differs from the kernel above (last-bit differences in zq).
__device__ __forceinline__ float dot8(const uint4& ua, const uint4& ub, float acc) {
  const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w};
  const uint32_t wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wa[k]));
    const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wb[k]));
    acc = fmaf(fa.x, fb.x, acc);
    acc = fmaf(fa.y, fb.y, acc);
  }
  return acc;
}
__global__ void __launch_bounds__(256) row_finalize_v2_kernel(const float4* __restrict__ partial, int n_slots,
                                                              int m_pad, int m_rows, int d,
                                                              const __nv_bfloat16* __restrict__ x_rows,
                                                              const __nv_bfloat16* __restrict__ y_all,
                                                              const int* __restrict__ pos_col,
                                                              const float* __restrict__ pos_q, int kp1,
                                                              float4* __restrict__ row_stats) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m_rows) return;
  // this row's slots: lane t holds slot t (kp1 <= 32 on this path)
  int my_col = -1;
  float my_q = 0.f;
  if (lane < kp1) {
    my_col = pos_col[static_cast<size_t>(row) * kp1 + lane];
    my_q = pos_q[static_cast<size_t>(row) * kp1 + lane];
  }
  float m = -INFINITY;
  for (int s = lane; s < n_slots; s += 32) m = fmaxf(m, partial[static_cast<size_t>(s) * m_pad + row].x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int s = lane; s < n_slots; s += 32) {
    const float4 p = partial[static_cast<size_t>(s) * m_pad + row];
    const float w = (p.x == -INFINITY) ? 0.f : exp2f(p.x - m);
    s0 = fmaf(p.y, w, s0);
    s1 = fmaf(p.z, w, s1);
    s2 = fmaf(p.w, w, s2);
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float mu = s1 / s0;
  const float var = s2 / s0 - mu * mu;
  const __nv_bfloat16* xr = x_rows + static_cast<size_t>(row) * d;
  float zq = 0.f;
  for (int t = 0; t < kp1; ++t) {
    const int c = __shfl_sync(0xffffffffu, my_col, t);
    if (c < 0) continue;  // warp-uniform
    const float q = __shfl_sync(0xffffffffu, my_q, t);
    const __nv_bfloat16* yr = y_all + static_cast<size_t>(c) * d;
    float acc = 0.f;
    int cb = lane * 8;
    for (; cb + 768 < d; cb += 1024) {  // four 16-byte pairs in flight per lane
      const uint4 a0 = *reinterpret_cast<const uint4*>(xr + cb), b0 = *reinterpret_cast<const uint4*>(yr + cb);
      const uint4 a1 = *reinterpret_cast<const uint4*>(xr + cb + 256), b1 = *reinterpret_cast<const uint4*>(yr + cb + 256);
      const uint4 a2 = *reinterpret_cast<const uint4*>(xr + cb + 512), b2 = *reinterpret_cast<const uint4*>(yr + cb + 512);
      const uint4 a3 = *reinterpret_cast<const uint4*>(xr + cb + 768), b3 = *reinterpret_cast<const uint4*>(yr + cb + 768);
      acc = dot8(a3, b3, dot8(a2, b2, dot8(a1, b1, dot8(a0, b0, acc))));
    }
    for (; cb + 256 < d; cb += 512) {  // two pairs
      const uint4 a0 = *reinterpret_cast<const uint4*>(xr + cb), b0 = *reinterpret_cast<const uint4*>(yr + cb);
      const uint4 a1 = *reinterpret_cast<const uint4*>(xr + cb + 256), b1 = *reinterpret_cast<const uint4*>(yr + cb + 256);
      acc = dot8(a1, b1, dot8(a0, b0, acc));
    }
    for (; cb < d; cb += 256) {
      const uint4 a0 = *reinterpret_cast<const uint4*>(xr + cb), b0 = *reinterpret_cast<const uint4*>(yr + cb);
      acc = dot8(a0, b0, acc);
    }
    zq = fmaf(q, warp_sum(acc), zq);
  }
  if (lane == 0) row_stats[row] = make_float4(m + log2f(s0), mu, var, zq);
}
static bool aux_v2() {
  static const bool on = [] {
    const char* e = std::getenv("SCL_AUX_V2");
    return e != nullptr && e[0] == '1' && e[1] == 0;
  }();
  return on;
}

cudaError_t launch_row_finalize(const float4* partial, int n_slots, int m_pad, int m_rows, int d, const void* x_rows,
                                const void* y_all, const int32_t* pos_col, const float* pos_q, int kp1,
                                float4* row_stats, cudaStream_t stream) {
  if (aux_v2() && kp1 <= 32 && d % 8 == 0) {
    row_finalize_v2_kernel<<<(m_rows + 7) / 8, 256, 0, stream>>>(partial, n_slots, m_pad, m_rows, d,
                                                                 static_cast<const __nv_bfloat16*>(x_rows),
                                                                 static_cast<const __nv_bfloat16*>(y_all), pos_col,
                                                                 pos_q, kp1, row_stats);
    return cudaGetLastError();
  }
  row_finalize_kernel<<<(m_rows + 7) / 8, 256, 0, stream>>>(partial, n_slots, m_pad, m_rows, d,
                                                            static_cast<const __nv_bfloat16*>(x_rows),
                                                            static_cast<const __nv_bfloat16*>(y_all), pos_col, pos_q,
                                                            kp1, row_stats);
  return cudaGetLastError();
}

// single CTA, fixed-order tree: sums6 = { sum_A (LSE - s_eff zq), sum_B (...), sum_A (mu - zq), sum_B (...),
//                                         sum_A var, sum_B var }
__global__ void __launch_bounds__(512) reduce_rows_kernel(const float4* __restrict__ stats_a,
                                                           const float4* __restrict__ stats_b, int m_rows,
                                                           const float* __restrict__ scalars,
                                                           float* __restrict__ sums6) {
  __shared__ double sh[6][512];
  const float s_eff = scalars[0];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < m_rows; i += 512) {
    const float4 a = stats_a[i];
    const float4 b = stats_b[i];
    acc[0] += static_cast<double>(a.x) * kLn2 - static_cast<double>(s_eff) * a.w;
    acc[1] += static_cast<double>(b.x) * kLn2 - static_cast<double>(s_eff) * b.w;
    acc[2] += static_cast<double>(a.y) - a.w;
    acc[3] += static_cast<double>(b.y) - b.w;
    acc[4] += a.z;
    acc[5] += b.z;
  }
  for (int k = 0; k < 6; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 6; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 6) sums6[threadIdx.x] = static_cast<float>(sh[threadIdx.x][0]);
}
cudaError_t launch_reduce_rows(const float4* stats_a, const float4* stats_b, int m_rows, const float* scalars,
                               float* sums6, cudaStream_t stream) {
  reduce_rows_kernel<<<1, 512, 0, stream>>>(stats_a, stats_b, m_rows, scalars, sums6);
  return cudaGetLastError();
}

// out4 = { loss, gap, d loss / d logit_scale, 2*w*gap }   (losses.py:113-122; SURVEY §8a closed forms)
__global__ void loss_scalars_kernel(const float* __restrict__ sums6, const float* __restrict__ scalars, float c,
                                    float w, float* __restrict__ out4) {
  const float gsum = c * (sums6[2] + sums6[3]);
  const float gap = w > 0.f ? gsum : 0.f;
  const float loss = c * (sums6[0] + sums6[1]) + w * gap * gap;
  const float ds = gsum + 2.f * w * gap * c * (sums6[4] + sums6[5]);
  out4[0] = loss;
  out4[1] = gap;
  out4[2] = ds;
  out4[3] = 2.f * w * gap;
}
cudaError_t launch_loss_scalars(const float* sums6, const float* scalars, float c, float w, float* out4,
                                cudaStream_t stream) {
  loss_scalars_kernel<<<1, 1, 0, stream>>>(sums6, scalars, c, w, out4);
  return cudaGetLastError();
}

// =====================================================================================
// backward coefficient vectors
// =====================================================================================
// row_coef[i] = {Lr, u_i, v_i, t_i},  col_coef[j] = {Lc, u'_j, v'_j, 0}  with
//   t_i = g * c * (s_eff + k2) * (q_own[i][0] + q_opp[i][0]): the soft-target weight on the row's own column
//         (slot 0 of both ELL lists), subtracted inside the tensor-core kernel before bf16 rounding
//   u = g * c * (s_eff + k2 * (1 - s_eff * mu)),  v = g * c * k2 * s_eff,  k2 = 2 * w * gap(owner rank)
// g = upstream grad * mult.  col_mode: 0 none, 1 only columns owned by `rank`, 2 all columns.
__global__ void bwd_coeffs_kernel(const float4* __restrict__ row_stats, int m_rows, int m_pad,
                                  const float4* __restrict__ col_stats, int n_cols, int n_pad, int b_local, int rank,
                                  const float* __restrict__ gaps, const float* __restrict__ scalars,
                                  const float* __restrict__ grad_out, float c, float w, float mult, int col_mode,
                                  const float* __restrict__ pos_q, const float* __restrict__ opp_q_local, int kp1,
                                  float4* __restrict__ row_coef, float4* __restrict__ col_coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float s_eff = scalars[0];
  const float g = grad_out[0] * mult * c;
  if (i < m_pad) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < m_rows) {
      const float4 st = row_stats[i];
      const float k2 = 2.f * w * gaps[rank];
      float qd = pos_q[static_cast<size_t>(i) * kp1];
      if (col_mode != 0) qd += opp_q_local[static_cast<size_t>(i) * kp1];
      o = make_float4(st.x, g * (s_eff + k2 * (1.f - s_eff * st.y)), g * k2 * s_eff, g * (s_eff + k2) * qd);
    }
    row_coef[i] = o;
  }
  if (i < n_pad) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n_cols) {
      const float4 st = col_stats[i];
      const int owner = i / b_local;
      const bool on = col_mode == 2 || (col_mode == 1 && owner == rank);
      const float k2 = 2.f * w * gaps[owner];
      const float gg = on ? g : 0.f;
      o = make_float4(st.x, gg * (s_eff + k2 * (1.f - s_eff * st.y)), gg * k2 * s_eff, 0.f);
    }
    col_coef[i] = o;
  }
}
cudaError_t launch_bwd_coeffs(const float4* row_stats, int m_rows, int m_pad, const float4* col_stats, int n_cols,
                              int n_pad, int b_local, int rank, const float* gaps, const float* scalars,
                              const float* grad_out, float c, float w, float mult, int col_mode, const float* pos_q,
                              const float* opp_q_local, int kp1, float4* row_coef, float4* col_coef,
                              cudaStream_t stream) {
  const int n = max(m_pad, n_pad);
  bwd_coeffs_kernel<<<(n + 255) / 256, 256, 0, stream>>>(row_stats, m_rows, m_pad, col_stats, n_cols, n_pad, b_local,
                                                         rank, gaps, scalars, grad_out, c, w, mult, col_mode, pos_q,
                                                         opp_q_local, kp1, row_coef, col_coef);
  return cudaGetLastError();
}

// =====================================================================================
// backward finish: chunk partial sums + sparse soft-target terms + cast
// =====================================================================================
// warp per local row:  dx32[i,:] = sum_chunks partial  -  g*c*(s_eff + k2_rank) * sum_{k>=1} q_ik * Y[col_ik,:]
__global__ void __launch_bounds__(256) bwd_gather_kernel(const float* __restrict__ dx_partial, int chunks, int m_pad,
                                                         int m_rows, int d, const __nv_bfloat16* __restrict__ y_all,
                                                         const int* __restrict__ pos_col,
                                                         const float* __restrict__ pos_q, int kp1, int rank,
                                                         const float* __restrict__ gaps,
                                                         const float* __restrict__ scalars,
                                                         const float* __restrict__ grad_out, float c, float w,
                                                         float mult, int y_ld, int y_lo,
                                                         float* __restrict__ dx32) {
  // y_ld: row pitch of y_all in elements; y_lo >= 0 (fp32-accurate mode): offset of the low-order bf16 part of
  // every row, added to the high-order part
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m_rows) return;
  const float coef = grad_out[0] * mult * c * (scalars[0] + 2.f * w * gaps[rank]);
  for (int c0 = lane * 4; c0 < d; c0 += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ch = 0; ch < chunks; ++ch) {
      const float4 p =
          *reinterpret_cast<const float4*>(dx_partial + (static_cast<size_t>(ch) * m_pad + row) * d + c0);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    for (int t = 1; t < kp1; ++t) {  // slot 0 (own column) is handled inside bwd_rows_kernel
      const int col = pos_col[static_cast<size_t>(row) * kp1 + t];
      if (col < 0) continue;
      const float qc = -coef * pos_q[static_cast<size_t>(row) * kp1 + t];
      const uint2 raw = *reinterpret_cast<const uint2*>(y_all + static_cast<size_t>(col) * y_ld + c0);
      float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
      if (y_lo >= 0) {
        const uint2 raw2 = *reinterpret_cast<const uint2*>(y_all + static_cast<size_t>(col) * y_ld + y_lo + c0);
        const float2 lo2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw2.x));
        const float2 hi2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw2.y));
        lo.x += lo2.x; lo.y += lo2.y; hi.x += hi2.x; hi.y += hi2.y;
      }
      acc.x = fmaf(qc, lo.x, acc.x); acc.y = fmaf(qc, lo.y, acc.y);
      acc.z = fmaf(qc, hi.x, acc.z); acc.w = fmaf(qc, hi.y, acc.w);
    }
    *reinterpret_cast<float4*>(dx32 + static_cast<size_t>(row) * d + c0) = acc;
  }
}
// Scan of the gathered opposite-direction lists: lane = one (row j, slot t) entry (coalesced), entries that
// list one of OUR rows as a positive are then processed by the whole warp, one hit at a time:
//   dx32[col - rank*B_l, :] -= g*c*(s_eff + k2_owner(j)) * q_jt * Y[j,:]        (fp32 atomics)
__global__ void __launch_bounds__(256) bwd_scatter_kernel(const int* __restrict__ opp_col_all,
                                                          const float* __restrict__ opp_q_all, int n_global, int kp1,
                                                          int b_local, int rank, int d,
                                                          const __nv_bfloat16* __restrict__ y_all,
                                                          const float* __restrict__ gaps,
                                                          const float* __restrict__ scalars,
                                                          const float* __restrict__ grad_out, float c, float w,
                                                          float mult, int col_mode, int y_ld, int y_lo,
                                                          float* __restrict__ dx32) {
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(n_global) * kp1;
  const long long e = (static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * 32 + lane;
  int col = -1, j = 0;
  float qv = 0.f;
  if (e < total) {
    j = static_cast<int>(e / kp1);
    const int t = static_cast<int>(e - static_cast<long long>(j) * kp1);
    const int cc = opp_col_all[e];
    const int owner = j / b_local;
    const bool mode_ok = col_mode == 2 || (col_mode == 1 && owner == rank);
    // slot 0 (own column) is handled inside the tensor-core kernel
    if (t != 0 && mode_ok && cc >= rank * b_local && cc < (rank + 1) * b_local) {
      col = cc;
      qv = opp_q_all[e];
    }
  }
  unsigned hits = __ballot_sync(0xffffffffu, col >= 0);
  const float base = -grad_out[0] * mult * c;
  const float s_eff = scalars[0];
  while (hits) {
    const int src = __ffs(hits) - 1;
    hits &= hits - 1;
    const int hc = __shfl_sync(0xffffffffu, col, src);
    const int hj = __shfl_sync(0xffffffffu, j, src);
    const float hq = __shfl_sync(0xffffffffu, qv, src);
    const float qc = base * (s_eff + 2.f * w * gaps[hj / b_local]) * hq;
    float* dst = dx32 + static_cast<size_t>(hc - rank * b_local) * d;
    const __nv_bfloat16* yrow = y_all + static_cast<size_t>(hj) * y_ld;
    for (int c0 = lane * 2; c0 < d; c0 += 64) {
      float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(yrow + c0));
      if (y_lo >= 0) {
        const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(yrow + y_lo + c0));
        v.x += v2.x;
        v.y += v2.y;
      }
      atomicAdd(dst + c0, qc * v.x);
      atomicAdd(dst + c0 + 1, qc * v.y);
    }
  }
}
template <typename T>
__global__ void cast_out_kernel(const float* __restrict__ src, T* __restrict__ dst, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    if constexpr (sizeof(T) == 2) {
      if constexpr (std::is_same<T, __nv_bfloat16>::value) dst[i] = __float2bfloat16_rn(src[i]);
      else dst[i] = __float2half_rn(src[i]);
    }
  }
}

cudaError_t launch_bwd_finish(const float* dx_partial, int chunks, int m_pad, int m_rows, int d, const void* y_all,
                              const int32_t* pos_col, const float* pos_q, int kp1, const int32_t* opp_col_all,
                              const float* opp_q_all, int n_global, int b_local, int rank, const float* gaps,
                              const float* scalars, const float* grad_out, float c, float w, float mult, int col_mode,
                              int split, float* dx32, void* dx_out, int out_dtype, cudaStream_t stream) {
  auto yb = static_cast<const __nv_bfloat16*>(y_all);
  const int y_ld = split ? 3 * d : d;  // split: y_all is the (h | l | h) column operand
  const int y_lo = split ? d : -1;
  bwd_gather_kernel<<<(m_rows + 7) / 8, 256, 0, stream>>>(dx_partial, chunks, m_pad, m_rows, d, yb, pos_col, pos_q,
                                                          kp1, rank, gaps, scalars, grad_out, c, w, mult, y_ld, y_lo,
                                                          dx32);
  if (col_mode != 0) {
    const long long entries = static_cast<long long>(n_global) * kp1;
    bwd_scatter_kernel<<<static_cast<unsigned>((entries + 255) / 256), 256, 0, stream>>>(
        opp_col_all, opp_q_all, n_global, kp1, b_local, rank, d, yb, gaps, scalars, grad_out, c, w, mult, col_mode,
        y_ld, y_lo, dx32);
  }
  const size_t n = static_cast<size_t>(m_rows) * d;
  if (out_dtype == 1)
    cast_out_kernel<__nv_bfloat16><<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
        dx32, static_cast<__nv_bfloat16*>(dx_out), n);
  else if (out_dtype == 2)
    cast_out_kernel<__half><<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(dx32, static_cast<__half*>(dx_out), n);
  return cudaGetLastError();
}


// =====================================================================================
// statistics exchange: split the all-gathered per-rank records back into contiguous arrays
// =====================================================================================
// in: [world][rec_floats] (each rank's flat record); component k of rank r lives at in[r*rec + off[k] .. + len[k])
// and goes to out[k][r*len[k] ..].  One launch instead of one strided copy per component.
struct UnpackDesc {
  float* out[8];
  int off[8];
  int len[8];
  int n;
};
__global__ void unpack_records_kernel(const float* __restrict__ in, int world, int rec_floats, UnpackDesc dsc) {
  const int k = blockIdx.y;
  if (k >= dsc.n) return;
  const long long total = static_cast<long long>(world) * dsc.len[k];
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / dsc.len[k]);
    const int o = static_cast<int>(i - static_cast<long long>(r) * dsc.len[k]);
    dsc.out[k][i] = in[static_cast<size_t>(r) * rec_floats + dsc.off[k] + o];
  }
}
cudaError_t launch_unpack_records(const float* in, int world, int rec_floats, int n_comp, float* const* outs,
                                  const int* offs, const int* lens, cudaStream_t stream) {
  if (n_comp < 1 || n_comp > 8) return cudaErrorInvalidValue;
  UnpackDesc dsc;
  dsc.n = n_comp;
  int max_len = 0;
  for (int k = 0; k < n_comp; ++k) {
    dsc.out[k] = outs[k];
    dsc.off[k] = offs[k];
    dsc.len[k] = lens[k];
    max_len = max(max_len, lens[k]);
  }
  const long long total = static_cast<long long>(world) * max_len;
  long long blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  const int bx = static_cast<int>(blocks);
  unpack_records_kernel<<<dim3(max(bx, 1), n_comp), 256, 0, stream>>>(in, world, rec_floats, dsc);
  return cudaGetLastError();
}

}  // namespace scl
