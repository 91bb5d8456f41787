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
__device__ __forceinline__ uint2 pack4_bf16(const float (&v)[4], float s) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0] * s, v[1] * s);
  const __nv_bfloat162 hi = __floats2bfloat162_rn(v[2] * s, v[3] * s);
  uint2 p;
  p.x = *reinterpret_cast<const uint32_t*>(&lo);
  p.y = *reinterpret_cast<const uint32_t*>(&hi);
  return p;
}

// Streaming cast of up to two [rows, d] matrices (blockIdx.y selects the matrix) to bf16: one thread = 8 consecutive
// elements (two 16-byte loads for fp32, one 16-byte store), grid-stride.  Block (0, 0) also writes the scalars when
// logit_scale != nullptr (scl_prepare: cap + both casts in one launch).
template <typename T>
__global__ void __launch_bounds__(256) cast_bf16_kernel(const T* __restrict__ x0, __nv_bfloat16* __restrict__ y0,
                                                        const T* __restrict__ x1, __nv_bfloat16* __restrict__ y1,
                                                        size_t n_oct, const float* __restrict__ logit_scale, float cap,
                                                        float* __restrict__ scalars) {
  if (logit_scale != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    const float s = logit_scale[0];
    const float s_eff = cap > 0.f ? fminf(s, cap) : s;
    scalars[0] = s_eff;
    scalars[1] = s_eff * kLog2e;
    scalars[2] = s;
  }
  const T* x = blockIdx.y == 0 ? x0 : x1;
  __nv_bfloat16* y = blockIdx.y == 0 ? y0 : y1;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_oct;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float a[4], b[4];
    load4(x + i * 8, a);
    load4(x + i * 8 + 4, b);
    const uint2 pa = pack4_bf16(a, 1.f), pb = pack4_bf16(b, 1.f);
    *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(pa.x, pa.y, pb.x, pb.y);
  }
}
// F.normalize(x, dim=-1) + cast (producer contract, open_clip model.py:326-345): one warp per row, the row is read
// twice (the second read hits L1/L2)
template <typename T>
__global__ void __launch_bounds__(256) normalize_cast_bf16_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                                  int rows, int d) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + static_cast<size_t>(row) * d;
  float ss = 0.f;
  for (int c = lane * 4; c < d; c += 128) {
    float v[4];
    load4(xr + c, v);
    ss = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], fmaf(v[3], v[3], ss))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
  for (int c = lane * 4; c < d; c += 128) {
    float v[4];
    load4(xr + c, v);
    *reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * d + c) = pack4_bf16(v, inv);
  }
}

template <typename T>
static cudaError_t launch_cast_t(const void* x0, void* y0, const void* x1, void* y1, int rows, int d, int normalize,
                                 const float* logit_scale, float cap, float* scalars, cudaStream_t stream) {
  auto a0 = static_cast<const T*>(x0), a1 = static_cast<const T*>(x1);
  auto b0 = static_cast<__nv_bfloat16*>(y0), b1 = static_cast<__nv_bfloat16*>(y1);
  if (normalize) {
    normalize_cast_bf16_kernel<T><<<(rows + 7) / 8, 256, 0, stream>>>(a0, b0, rows, d);
    return cudaGetLastError();
  }
  const size_t n_oct = static_cast<size_t>(rows) * d / 8;
  size_t blocks = (n_oct + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  cast_bf16_kernel<T><<<dim3(static_cast<unsigned>(blocks), x1 != nullptr ? 2 : 1), 256, 0, stream>>>(
      a0, b0, a1, b1, n_oct, logit_scale, cap, scalars);
  return cudaGetLastError();
}

static cudaError_t launch_cast_any(const void* x0, void* y0, const void* x1, void* y1, int src_dtype, int rows, int d,
                                   int normalize, const float* logit_scale, float cap, float* scalars,
                                   cudaStream_t stream) {
  if (src_dtype == 0)
    return launch_cast_t<float>(x0, y0, x1, y1, rows, d, normalize, logit_scale, cap, scalars, stream);
  if (src_dtype == 1)
    return launch_cast_t<__nv_bfloat16>(x0, y0, x1, y1, rows, d, normalize, logit_scale, cap, scalars, stream);
  return launch_cast_t<__half>(x0, y0, x1, y1, rows, d, normalize, logit_scale, cap, scalars, stream);
}

cudaError_t launch_cast_bf16(const void* x, int src_dtype, void* y, int rows, int d, int normalize,
                             cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  return launch_cast_any(x, y, nullptr, nullptr, src_dtype, rows, d, normalize, nullptr, 0.f, nullptr, stream);
}
// cap + casts of both modalities: one launch
cudaError_t launch_prepare(const void* image, const void* text, int src_dtype, int rows, int d, void* image_bf16,
                           void* text_bf16, const float* logit_scale, float cap, float* scalars, cudaStream_t stream) {
  return launch_cast_any(image, image_bf16, text, text_bf16, src_dtype, rows, d, 0, logit_scale, cap, scalars, stream);
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
__global__ void hash_insert_kernel(const long long* __restrict__ ids, int n, long long* keys, int* vals, uint32_t cap,
                                   const int* __restrict__ run_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (run_flag != nullptr && *run_flag == 0)) return;
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
// G lanes per local row (G = 8 / 16 / 32 >= k), lane s = neighbour slot s.  Slot 0 of the output is the row's own
// column rank*B_l + i with weight 1 (losses.py:94-98); neighbour slots are skipped when alpha*scale <= 0 or the id is
// not in the global batch, and merged (fp32 +=, in k order) into the slot already holding the same column
// (losses.py:100-108).  The K hash probes of a row run in parallel; the order-sensitive parts (which slot is a
// column's first touch, the fp32 accumulation order, the L1 norm) are evaluated with shuffles in exactly the
// sequential order of the reference, so weights and probabilities are bit-identical to the one-thread-per-row form.
template <int G>
__global__ void __launch_bounds__(256) build_ell_kernel(const long long* __restrict__ nbr_ids,
                                                        const float* __restrict__ nbr_alpha, int b_local, int k,
                                                        float alpha_scale, int rank, const long long* __restrict__ keys,
                                                        const int* __restrict__ vals, uint32_t cap,
                                                        int* __restrict__ pos_col, float* __restrict__ pos_w,
                                                        float* __restrict__ pos_q, const int* __restrict__ copy_flag,
                                                        const int* __restrict__ src_col, const float* __restrict__ src_w,
                                                        const float* __restrict__ src_q) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / G;          // local row
  const int s = tid % G;          // neighbour slot of this lane
  const int lane = threadIdx.x & 31;
  const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (lane & ~(G - 1));
  const int g0 = lane & ~(G - 1);  // first lane of this row's group
  const int kp1 = k + 1;
  const bool row_ok = i < b_local;
  if (copy_flag != nullptr && *copy_flag == 0) {
    // the two id vectors are identical: this direction's lists equal the other direction's (already built)
    if (row_ok)
      for (int t = s; t < kp1; t += G) {
        const size_t o = static_cast<size_t>(i) * kp1 + t;
        pos_col[o] = src_col[o];
        pos_w[o] = src_w[o];
        pos_q[o] = src_q[o];
      }
    return;
  }
  const int own = rank * b_local + i;
  // ---- this lane's neighbour: scaled alpha and the column its id maps to (-1: not in the global batch)
  float a = 0.f;
  int c = -1;
  if (row_ok && s < k) {
    a = fmaxf(__fmul_rn(nbr_alpha[static_cast<size_t>(i) * k + s], alpha_scale), 0.f);
    if (a > 0.f) c = hash_lookup(nbr_ids[static_cast<size_t>(i) * k + s], keys, vals, cap);
  }
  const bool valid = c >= 0;
  // ---- first touch of a column: no earlier valid slot (and not the own column) names it
  bool first = valid && c != own;
  for (int e = 0; e < G; ++e) {
    const int ce = __shfl_sync(gmask, c, g0 + e);
    if (e < s && ce == c) first = false;
  }
  // ---- weight of a first-touch slot: its alpha, then every later duplicate in slot order (fp32 adds, k order)
  float w = a;
  float w_own = 1.0f;  // slot 0: 1.0, then every neighbour that maps to the own column (self loops), in slot order
  for (int e = 0; e < G; ++e) {
    const int ce = __shfl_sync(gmask, c, g0 + e);
    const float ae = __shfl_sync(gmask, a, g0 + e);
    if (first && e > s && ce == c) w = __fadd_rn(w, ae);
    if (ce >= 0 && ce == own) w_own = __fadd_rn(w_own, ae);
  }
  const unsigned first_mask = (__ballot_sync(gmask, first) >> g0) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
  const int cnt = 1 + __popc(first_mask);
  const int out = 1 + __popc(first_mask & ((1u << s) - 1u));  // output slot of a first-touch lane
  // ---- L1 norm in output-slot order (sequential fp32 adds, F.normalize(p=1), losses.py:110-111)
  float tot = w_own;
  for (unsigned m = first_mask; m; m &= m - 1) tot = __fadd_rn(tot, __shfl_sync(gmask, w, g0 + __ffs(m) - 1));
  const float inv = 1.f / fmaxf(tot, 1e-12f);
  if (!row_ok) return;
  int* col_o = pos_col + static_cast<size_t>(i) * kp1;
  float* w_o = pos_w + static_cast<size_t>(i) * kp1;
  float* q_o = pos_q + static_cast<size_t>(i) * kp1;
  if (s == 0) {
    col_o[0] = own;
    w_o[0] = w_own;
    q_o[0] = w_own * inv;
  }
  if (first) {
    col_o[out] = c;
    w_o[out] = w;
    q_o[out] = w * inv;
  }
  for (int t = cnt + s; t < kp1; t += G) {  // unused slots
    col_o[t] = -1;
    w_o[t] = 0.f;
    q_o[t] = 0.f;
  }
}

// flag = 1 when the two gathered id vectors differ anywhere (flag is zeroed by the caller)
__global__ void ids_differ_kernel(const long long* __restrict__ a, const long long* __restrict__ b, int n,
                                  int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && a[i] != b[i]) *flag = 1;
}
// the hash kernels of the second direction do nothing when the id vectors are identical
__global__ void hash_clear_if_kernel(long long* keys, int* vals, uint32_t cap, const int* __restrict__ flag) {
  if (*flag == 0) return;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= cap) {
    keys[i] = kEmptyKey;
    vals[i] = -1;
  }
}

template <int G>
static void launch_build_ell(const int64_t* nbr_ids, const float* nbr_alpha, int b_local, int k, float alpha_scale,
                             int rank, const long long* keys, const int* vals, uint32_t cap, int32_t* pos_col,
                             float* pos_w, float* pos_q, const int* copy_flag, const int32_t* src_col,
                             const float* src_w, const float* src_q, cudaStream_t stream) {
  const long long threads = static_cast<long long>(b_local) * G;
  build_ell_kernel<G><<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const long long*>(nbr_ids), nbr_alpha, b_local, k, alpha_scale, rank, keys, vals, cap, pos_col,
      pos_w, pos_q, copy_flag, src_col, src_w, src_q);
}

// differ_flag / src_*: (second direction only) device flag that is 0 when this direction's id vector equals the other
// direction's -- the hash kernels then return at once and the lists are copied from src_* -- or nullptr.
cudaError_t launch_build_positives(const int64_t* all_ids, int n_global, const int64_t* nbr_ids,
                                   const float* nbr_alpha, int b_local, int k, float alpha_scale, int rank,
                                   void* hash_ws, size_t hash_ws_bytes, int32_t* pos_col, float* pos_w, float* pos_q,
                                   const int* differ_flag, const int32_t* src_col, const float* src_w,
                                   const float* src_q, cudaStream_t stream) {
  if (k > 32) return cudaErrorInvalidValue;
  const uint32_t cap = hash_capacity(n_global);
  long long* keys = static_cast<long long*>(hash_ws);
  int* vals = reinterpret_cast<int*>(keys + cap + 1);
  if (k > 0) {
    if (hash_ws == nullptr || hash_ws_bytes < positives_hash_bytes(n_global)) return cudaErrorInvalidValue;
    if (differ_flag != nullptr)
      hash_clear_if_kernel<<<(cap + 1 + 255) / 256, 256, 0, stream>>>(keys, vals, cap, differ_flag);
    else
      hash_clear_kernel<<<(cap + 1 + 255) / 256, 256, 0, stream>>>(keys, vals, cap);
    hash_insert_kernel<<<(n_global + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const long long*>(all_ids),
                                                                   n_global, keys, vals, cap, differ_flag);
  }
  if (k <= 8)
    launch_build_ell<8>(nbr_ids, nbr_alpha, b_local, k, alpha_scale, rank, keys, vals, cap, pos_col, pos_w, pos_q,
                        differ_flag, src_col, src_w, src_q, stream);
  else if (k <= 16)
    launch_build_ell<16>(nbr_ids, nbr_alpha, b_local, k, alpha_scale, rank, keys, vals, cap, pos_col, pos_w, pos_q,
                         differ_flag, src_col, src_w, src_q, stream);
  else
    launch_build_ell<32>(nbr_ids, nbr_alpha, b_local, k, alpha_scale, rank, keys, vals, cap, pos_col, pos_w, pos_q,
                         differ_flag, src_col, src_w, src_q, stream);
  return cudaGetLastError();
}
cudaError_t launch_ids_differ(const int64_t* a, const int64_t* b, int n, int* flag, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  ids_differ_kernel<<<(n + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const long long*>(a),
                                                         reinterpret_cast<const long long*>(b), n, flag);
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
// warp per local row: stats = {L2 = m + log2(S0)  (LSE in log2 units), mu = S1/S0, var = S2/S0 - mu^2,
//                              zq = sum_k q_k <x_i, y_col_k>}
// 16-byte loads, all loads of a dot product in flight together, the slot list of the row read once (lane t = slot t).
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
__global__ void __launch_bounds__(256) row_finalize_kernel(const float4* __restrict__ partial, int n_slots,
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
cudaError_t launch_row_finalize(const float4* partial, int n_slots, int m_pad, int m_rows, int d, const void* x_rows,
                                const void* y_all, const int32_t* pos_col, const float* pos_q, int kp1,
                                float4* row_stats, cudaStream_t stream) {
  if (kp1 > 32) return cudaErrorInvalidValue;
  row_finalize_kernel<<<(m_rows + 7) / 8, 256, 0, stream>>>(partial, n_slots, m_pad, m_rows, d,
                                                            static_cast<const __nv_bfloat16*>(x_rows),
                                                            static_cast<const __nv_bfloat16*>(y_all), pos_col, pos_q,
                                                            kp1, row_stats);
  return cudaGetLastError();
}

// sums6 = { sum_A (LSE - s_eff zq), sum_B (...), sum_A (mu - zq), sum_B (...), sum_A var, sum_B var }
// Single CTA, fixed-order fp64 tree (deterministic); 1024 threads with four independent row loads in flight each.
// When out4 != nullptr the loss scalars (scl_loss_scalars) are evaluated by the same launch.
__device__ __forceinline__ void loss_scalars_from(const float* sums6, float c, float w, float* out4) {
  const float gsum = c * (sums6[2] + sums6[3]);
  const float gap = w > 0.f ? gsum : 0.f;
  out4[0] = c * (sums6[0] + sums6[1]) + w * gap * gap;
  out4[1] = gap;
  out4[2] = gsum + 2.f * w * gap * c * (sums6[4] + sums6[5]);
  out4[3] = 2.f * w * gap;
}
__global__ void __launch_bounds__(1024) reduce_rows_kernel(const float4* __restrict__ stats_a,
                                                            const float4* __restrict__ stats_b, int m_rows,
                                                            const float* __restrict__ scalars,
                                                            float* __restrict__ sums6, float c, float w,
                                                            float* __restrict__ out4) {
  __shared__ double sh[6][1024];
  const double s_eff = scalars[0];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i0 = threadIdx.x; i0 < m_rows; i0 += 4096) {
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 1024;
      a[u] = i < m_rows ? stats_a[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      b[u] = i < m_rows ? stats_b[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[0] += static_cast<double>(a[u].x) * kLn2 - s_eff * a[u].w;
      acc[1] += static_cast<double>(b[u].x) * kLn2 - s_eff * b[u].w;
      acc[2] += static_cast<double>(a[u].y) - a[u].w;
      acc[3] += static_cast<double>(b[u].y) - b[u].w;
      acc[4] += a[u].z;
      acc[5] += b[u].z;
    }
  }
  for (int k = 0; k < 6; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < 6; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < 6) sums6[threadIdx.x] = static_cast<float>(sh[threadIdx.x][0]);
  __syncthreads();
  if (threadIdx.x == 0 && out4 != nullptr) loss_scalars_from(sums6, c, w, out4);
}
cudaError_t launch_reduce_rows(const float4* stats_a, const float4* stats_b, int m_rows, const float* scalars,
                               float* sums6, float c, float w, float* out4, cudaStream_t stream) {
  reduce_rows_kernel<<<1, 1024, 0, stream>>>(stats_a, stats_b, m_rows, scalars, sums6, c, w, out4);
  return cudaGetLastError();
}

// out4 = { loss, gap, d loss / d logit_scale, 2*w*gap }   (losses.py:113-122; SURVEY §8a closed forms)
__global__ void loss_scalars_kernel(const float* __restrict__ sums6, float c, float w, float* __restrict__ out4) {
  loss_scalars_from(sums6, c, w, out4);
}
cudaError_t launch_loss_scalars(const float* sums6, const float* scalars, float c, float w, float* out4,
                                cudaStream_t stream) {
  (void)scalars;
  loss_scalars_kernel<<<1, 1, 0, stream>>>(sums6, c, w, out4);
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
// backward finish: chunk partial sums + sparse soft-target terms + cast  (deterministic: no atomics on dX)
// =====================================================================================
// Reverse lists: the gathered opposite-direction soft-target lists name, per entry e = (row j, slot t >= 1), a column
// cc; the entries whose column is one of OUR rows (rank*B_l <= cc < (rank+1)*B_l) contribute
//   dX[cc - rank*B_l, :] -= g*c*(s_eff + k2_owner(j)) * q_jt * Y[j, :].
// They are bucketed per local row (count -> bucket allocation -> fill; where a bucket lies and the fill order inside it
// are arbitrary) and the consumer visits every bucket in ascending e, so the fp32 summation order is fixed from run
// to run.
__device__ __forceinline__ bool rev_entry_hits(long long e, int kp1, int b_local, int rank, int col_mode, int cc) {
  const int j = static_cast<int>(e / kp1);
  const int t = static_cast<int>(e - static_cast<long long>(j) * kp1);
  const bool mode_ok = col_mode == 2 || (col_mode == 1 && j / b_local == rank);
  return t != 0 && mode_ok && cc >= rank * b_local && cc < (rank + 1) * b_local;  // slot 0: inside the tensor-core kernel
}
__global__ void __launch_bounds__(256) rev_count_kernel(const int* __restrict__ opp_col_all, long long total, int kp1,
                                                        int b_local, int rank, int col_mode, int* __restrict__ cnt) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int cc = opp_col_all[e];
  if (rev_entry_hits(e, kp1, b_local, rank, col_mode, cc)) atomicAdd(&cnt[cc - rank * b_local], 1);
}
// bucket placement: off[i] = a private range of cnt[i] entries of `list`, handed out by one atomic per non-empty row.
// WHERE a bucket lies does not matter (the consumer orders every bucket itself), so no prefix scan is needed.
__global__ void __launch_bounds__(256) rev_alloc_kernel(const int* __restrict__ cnt, int n, int* __restrict__ total,
                                                        int* __restrict__ off, int* __restrict__ cursor) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int c = i < n ? cnt[i] : 0;
  // one atomic per warp: inclusive prefix of the lanes' counts, the last lane reserves the warp's range
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  int base = 0;
  if (lane == 31 && incl > 0) base = atomicAdd(total, incl);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (i < n) {
    off[i] = base + incl - c;
    cursor[i] = base + incl - c;
  }
}
__global__ void __launch_bounds__(256) rev_fill_kernel(const int* __restrict__ opp_col_all, long long total, int kp1,
                                                       int b_local, int rank, int col_mode, int* __restrict__ cursor,
                                                       int* __restrict__ list) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const int cc = opp_col_all[e];
  if (rev_entry_hits(e, kp1, b_local, rank, col_mode, cc))
    list[atomicAdd(&cursor[cc - rank * b_local], 1)] = static_cast<int>(e);
}
size_t bwd_finish_workspace_bytes(int n_global, int b_local, int kp1) {
  // cnt[b_local] + total | off[b_local] | cursor[b_local] | list[n_global * kp1]
  return (static_cast<size_t>(3) * b_local + 1 + static_cast<size_t>(n_global) * kp1) * sizeof(int) + 256;
}

__device__ __forceinline__ void fma_row4(float4& acc, float qc, const __nv_bfloat16* __restrict__ yrow, int c0, int y_lo) {
  const uint2 raw = *reinterpret_cast<const uint2*>(yrow + c0);
  float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
  float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
  if (y_lo >= 0) {  // fp32-accurate mode: add the low-order bf16 part of the row
    const uint2 raw2 = *reinterpret_cast<const uint2*>(yrow + y_lo + c0);
    const float2 lo2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw2.x));
    const float2 hi2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw2.y));
    lo.x += lo2.x; lo.y += lo2.y; hi.x += hi2.x; hi.y += hi2.y;
  }
  acc.x = fmaf(qc, lo.x, acc.x); acc.y = fmaf(qc, lo.y, acc.y);
  acc.z = fmaf(qc, hi.x, acc.z); acc.w = fmaf(qc, hi.y, acc.w);
}
// warp per local row:
//   dX[i,:] = sum_chunks partial                                            (chunk order)
//           - g*c*(s_eff + k2_rank) * sum_{k>=1} q_ik * Y[col_ik,:]           (slot order)
//           - sum over the row's reverse bucket, ascending e                  (see above)
// written as fp32 (OutT = float) or cast on the way out.  y_ld: row pitch of y_all in elements; y_lo >= 0
// (fp32-accurate mode): offset of the low-order bf16 part of every row.
template <typename OutT>
__device__ __forceinline__ void store4(OutT* dst, const float4& v) {
  if constexpr (std::is_same<OutT, float>::value) {
    *reinterpret_cast<float4*>(dst) = v;
  } else if constexpr (std::is_same<OutT, __nv_bfloat16>::value) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst) = pk;
  } else {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&lo);
    pk.y = *reinterpret_cast<const uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dst) = pk;
  }
}
template <typename OutT>
__global__ void __launch_bounds__(256) bwd_gather_kernel(const float* __restrict__ dx_partial, int chunks, int m_pad,
                                                         int m_rows, int d, const __nv_bfloat16* __restrict__ y_all,
                                                         const int* __restrict__ pos_col,
                                                         const float* __restrict__ pos_q, int kp1, int b_local, int rank,
                                                         const float* __restrict__ gaps,
                                                         const float* __restrict__ scalars,
                                                         const float* __restrict__ grad_out, float c, float w,
                                                         float mult, int y_ld, int y_lo,
                                                         const int* __restrict__ rev_off, const int* __restrict__ rev_cnt,
                                                         const int* __restrict__ rev_list,
                                                         const float* __restrict__ opp_q_all,
                                                         OutT* __restrict__ dx_out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= m_rows) return;
  // ---- everything the lists need is requested first (independent loads), the long row streams follow
  int my_col = -1;
  float my_q = 0.f;
  if (lane >= 1 && lane < kp1) {  // own list: lane t holds slot t (slot 0 is handled inside the tensor-core kernel)
    my_col = pos_col[static_cast<size_t>(row) * kp1 + lane];
    my_q = pos_q[static_cast<size_t>(row) * kp1 + lane];
  }
  const int r_lo = rev_off != nullptr ? rev_off[row] : 0;
  const int r_n = rev_off != nullptr ? rev_cnt[row] : 0;
  const float base = grad_out[0] * mult * c;
  const float s_eff = scalars[0];
  const float coef = base * (s_eff + 2.f * w * gaps[rank]);
  const float my_qc = -coef * my_q;
  const unsigned own_mask = __ballot_sync(0xffffffffu, my_col >= 0);
  // the common case (bucket of <= 32 entries): order it once, lane p keeps the p-th smallest entry
  int s_j = 0;
  float s_qc = 0.f;
  if (r_n > 0 && r_n <= 32) {
    unsigned mine = lane < r_n ? static_cast<unsigned>(rev_list[r_lo + lane]) : 0xffffffffu;
    for (int p = 0; p < r_n; ++p) {
      const unsigned e = __reduce_min_sync(0xffffffffu, mine);
      if (mine == e) mine = 0xffffffffu;  // entries are unique
      if (lane == p) {
        s_j = static_cast<int>(e) / kp1;
        s_qc = -base * (s_eff + 2.f * w * gaps[s_j / b_local]) * opp_q_all[e];
      }
    }
  }
  // ---- D in segments of 128 columns (one float4 per lane); lanes beyond D only take part in the shuffles
  for (int c0 = lane * 4; c0 - lane * 4 < d; c0 += 128) {
    const bool ok = c0 < d;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = dx_partial + static_cast<size_t>(row) * d + c0;
    const size_t slab = static_cast<size_t>(m_pad) * d;
    if (ok) {
      for (int ch = 0; ch < chunks; ++ch) {
        const float4 p0 = *reinterpret_cast<const float4*>(src + ch * slab);
        acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w;
      }
    }
    for (unsigned m = own_mask; m; m &= m - 1) {
      const int t = __ffs(m) - 1;
      const int col = __shfl_sync(0xffffffffu, my_col, t);
      const float qc = __shfl_sync(0xffffffffu, my_qc, t);
      if (ok) fma_row4(acc, qc, y_all + static_cast<size_t>(col) * y_ld, c0, y_lo);
    }
    if (r_n <= 32) {
      for (int p = 0; p < r_n; ++p) {
        const int j = __shfl_sync(0xffffffffu, s_j, p);
        const float qc = __shfl_sync(0xffffffffu, s_qc, p);
        if (ok) fma_row4(acc, qc, y_all + static_cast<size_t>(j) * y_ld, c0, y_lo);
      }
    }
    // larger buckets (hub rows): ascending entry order by repeated warp-wide selection of the smallest entry > last
    int last = -1;
    for (int done = 0; r_n > 32 && done < r_n; ++done) {
      unsigned mine = 0xffffffffu;
      for (int k = lane; k < r_n; k += 32) {
        const int e = rev_list[r_lo + k];
        if (e > last) mine = min(mine, static_cast<unsigned>(e));
      }
      const int e = static_cast<int>(__reduce_min_sync(0xffffffffu, mine));
      last = e;
      const int j = e / kp1;
      const float qc = -base * (s_eff + 2.f * w * gaps[j / b_local]) * opp_q_all[e];
      if (ok) fma_row4(acc, qc, y_all + static_cast<size_t>(j) * y_ld, c0, y_lo);
    }
    if (ok) store4<OutT>(dx_out + static_cast<size_t>(row) * d + c0, acc);
  }
}

cudaError_t launch_bwd_finish(const float* dx_partial, int chunks, int m_pad, int m_rows, int d, const void* y_all,
                              const int32_t* pos_col, const float* pos_q, int kp1, const int32_t* opp_col_all,
                              const float* opp_q_all, int n_global, int b_local, int rank, const float* gaps,
                              const float* scalars, const float* grad_out, float c, float w, float mult, int col_mode,
                              int split, void* workspace, size_t workspace_bytes, void* dx_out, int out_dtype,
                              cudaStream_t stream) {
  if (kp1 > 32) return cudaErrorInvalidValue;
  auto yb = static_cast<const __nv_bfloat16*>(y_all);
  const int y_ld = split ? 3 * d : d;  // split: y_all is the (h | l | h) column operand
  const int y_lo = split ? d : -1;
  const int* rev_off = nullptr;
  const int* rev_cnt = nullptr;
  const int* rev_list = nullptr;
  if (col_mode != 0) {
    if (workspace == nullptr || workspace_bytes < bwd_finish_workspace_bytes(n_global, b_local, kp1))
      return cudaErrorInvalidValue;
    int* cnt = static_cast<int*>(workspace);  // cnt[b_local], then the running total
    int* off = cnt + b_local + 1;
    int* cursor = off + b_local;
    int* list = cursor + b_local;
    const long long total = static_cast<long long>(n_global) * kp1;
    const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
    cudaError_t e = cudaMemsetAsync(cnt, 0, static_cast<size_t>(b_local + 1) * sizeof(int), stream);
    if (e != cudaSuccess) return e;
    rev_count_kernel<<<blocks, 256, 0, stream>>>(opp_col_all, total, kp1, b_local, rank, col_mode, cnt);
    rev_alloc_kernel<<<(b_local + 255) / 256, 256, 0, stream>>>(cnt, b_local, cnt + b_local, off, cursor);
    rev_fill_kernel<<<blocks, 256, 0, stream>>>(opp_col_all, total, kp1, b_local, rank, col_mode, cursor, list);
    rev_off = off;
    rev_cnt = cnt;
    rev_list = list;
  }
  const unsigned grid = (m_rows + 7) / 8;
#define SCL_GATHER(T)                                                                                                 \
  bwd_gather_kernel<T><<<grid, 256, 0, stream>>>(dx_partial, chunks, m_pad, m_rows, d, yb, pos_col, pos_q, kp1,       \
                                                 b_local, rank, gaps, scalars, grad_out, c, w, mult, y_ld, y_lo,      \
                                                 rev_off, rev_cnt, rev_list, opp_q_all, static_cast<T*>(dx_out))
  if (out_dtype == 0) SCL_GATHER(float);
  else if (out_dtype == 1) SCL_GATHER(__nv_bfloat16);
  else SCL_GATHER(__half);
#undef SCL_GATHER
  return cudaGetLastError();
}

// =====================================================================================
// caller-resolved soft targets (SpatialLossFromColumns): validate and sanitise
// =====================================================================================
// The kernels index y_all + col * d for every col >= 0 and assume slot 0 is the row's own column.  Lists that come
// from outside the library are therefore copied through this pass: out-of-range columns become unused slots
// (flag bit 0), unused slots get q = 0 (bit 2 when they carried weight), a slot 0 other than rank*B_l + i sets bit 1.
__global__ void __launch_bounds__(256) check_positives_kernel(const int* __restrict__ col_in,
                                                              const float* __restrict__ q_in, int b_local, int kp1,
                                                              int n_global, int rank, int* __restrict__ col_out,
                                                              float* __restrict__ q_out, int* __restrict__ flag) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= b_local * kp1) return;
  const int i = idx / kp1, t = idx - i * kp1;
  int c = col_in[idx];
  float q = q_in[idx];
  int bits = 0;
  if (c < -1 || c >= n_global) {
    bits |= 1;
    c = -1;
  }
  if (t == 0 && c != rank * b_local + i) bits |= 2;
  if (c < 0 && q != 0.f) {
    if (col_in[idx] == -1) bits |= 4;
    q = 0.f;
  }
  col_out[idx] = c;
  q_out[idx] = q;
  if (bits) atomicOr(flag, bits);
}
cudaError_t launch_check_positives(const int32_t* col_in, const float* q_in, int b_local, int kp1, int n_global,
                                   int rank, int32_t* col_out, float* q_out, int* flag, cudaStream_t stream) {
  const int n = b_local * kp1;
  check_positives_kernel<<<(n + 255) / 256, 256, 0, stream>>>(col_in, q_in, b_local, kp1, n_global, rank, col_out,
                                                              q_out, flag);
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
