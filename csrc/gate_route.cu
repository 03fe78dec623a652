// Fused gating + routing for sm_100a: TWO launches take the gate logits to everything the dispatch needs.
//
// The reference spends ~15 PyTorch kernels on softmax / top-k / one-hot masks / GShard loss / gate normalisation
// (tutel/impls/moe_layer.py:283-305, tutel/impls/losses.py:12-19) and k cumsum passes + k compares on the locations
// (tutel/impls/fast_dispatch.py:143-204, `tutel_ops.cumsum` tutel/custom/custom_kernel.cpp:822-872).  Here:
//
//   gate_route_kernel    one warp per token: softmax in registers, iterative arg-max top-k, gate normalisation, the
//                        per-block histogram of every choice (the routing scan's input), per-block importance sums
//                        for the loss; the grid also pre-fills the slot map with -1.
//   route_finish_kernel  one thread per token: every block derives its own queue offsets from the block histograms
//                        (no separate scan launch), ranks its tokens (match.any), writes locations and the inverse
//                        slot -> (token, choice) map; block 0 also emits the per-expert counts and the auxiliary loss.
//
// Backward of the whole gate is ONE kernel (closed form through normalisation, top-k selection and softmax).
// Also here: grouped column sums (bias gradients at copy bandwidth) and the public `fast_cumsum_sub_one` scan.
#include "moe_kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace tb {
namespace {

constexpr int kTileTokens = 256;      // tokens per routing tile (histogram granularity)
constexpr int kGateThreads = 1024;    // gate kernel: 32 warps, 8 tokens per warp
constexpr int kInvalidLoc = 0x3fffffff;

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------
// launch 1: softmax + top-k + normalised gates + per-tile histograms / importance sums (+ slot map pre-fill)
// ------------------------------------------------------------------------------------------------
template <typename T, int VPT>
__global__ void __launch_bounds__(kGateThreads)
gate_route_kernel(const T* __restrict__ logits, float* __restrict__ scores, int* __restrict__ idx,
                  float* __restrict__ top, float* __restrict__ gates, float* __restrict__ me_partial,
                  int* __restrict__ hist, int* __restrict__ slot_src, long long slot_n, int S, int E, int k,
                  int normalize, float eps) {
  extern __shared__ int sm_dyn[];                 // [k * E] histogram, then [E] floats of importance sums
  int* sm_hist = sm_dyn;
  float* sm_me = reinterpret_cast<float*>(sm_dyn + k * E);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < k * E; i += kGateThreads) sm_hist[i] = 0;
  for (int i = threadIdx.x; i < E; i += kGateThreads) sm_me[i] = 0.0f;
  for (long long i = static_cast<long long>(blockIdx.x) * kGateThreads + threadIdx.x; i < slot_n;
       i += static_cast<long long>(gridDim.x) * kGateThreads)
    slot_src[i] = -1;
  __syncthreads();

  float me[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) me[i] = 0.0f;
  const long long s_begin = static_cast<long long>(blockIdx.x) * kTileTokens;
  const long long s_end = min(s_begin + kTileTokens, static_cast<long long>(S));
  for (long long s = s_begin + warp; s < s_end; s += kGateThreads / 32) {
    float v[VPT];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int e = lane + 32 * i;
      v[i] = e < E ? ldf<T>(logits + s * E + e) : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      v[i] = (lane + 32 * i < E) ? expf(v[i] - mx) : 0.0f;
      sum += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int e = lane + 32 * i;
      v[i] *= inv;
      if (e < E) {
        scores[s * E + e] = v[i];
        me[i] += v[i];
      }
    }
    // iterative arg-max (ties -> lower expert id); lane j keeps the j-th choice
    unsigned taken = 0;
    float mine = 0.0f;
    int mine_e = -1;
    for (int j = 0; j < k; ++j) {
      float best = -1.0f;
      int best_e = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < VPT; ++i) {
        const int e = lane + 32 * i;
        if (e < E && !((taken >> i) & 1u) && (v[i] > best)) { best = v[i]; best_e = e; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oe = __shfl_xor_sync(0xffffffffu, best_e, o);
        if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; }
      }
      if ((best_e & 31) == lane && best_e < E) {
        taken |= 1u << (best_e >> 5);
        atomicAdd(&sm_hist[j * E + best_e], 1);
      }
      if (lane == j) { mine = best; mine_e = best_e; }
    }
    float denom = (lane < k) ? mine : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) denom += __shfl_xor_sync(0xffffffffu, denom, o);
    if (lane < k) {
      const float g = (normalize != 0 && k > 1) ? mine / fmaxf(denom, eps) : mine;
      idx[static_cast<long long>(lane) * S + s] = mine_e;
      top[static_cast<long long>(lane) * S + s] = mine;
      gates[static_cast<long long>(lane) * S + s] = g;
    }
  }
#pragma unroll
  for (int i = 0; i < VPT; ++i)
    if (lane + 32 * i < E && me[i] != 0.0f) atomicAdd(&sm_me[lane + 32 * i], me[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < k * E; i += kGateThreads) hist[static_cast<long long>(blockIdx.x) * k * E + i] = sm_hist[i];
  for (int i = threadIdx.x; i < E; i += kGateThreads) me_partial[static_cast<long long>(blockIdx.x) * E + i] = sm_me[i];
}

// ------------------------------------------------------------------------------------------------
// launch 2: queue offsets from the tile histograms, in-tile ranking, locations, slot map, counts, loss
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kTileTokens)
route_finish_kernel(const int* __restrict__ idx, const int* __restrict__ hist, const float* __restrict__ me_partial,
                    int* __restrict__ loc, int* __restrict__ counts, int* __restrict__ slot_src,
                    float* __restrict__ ce_out, T* __restrict__ l_aux, int S, int E, int k, int C, int ntiles) {
  extern __shared__ int sm_dyn[];        // total[k*E] | base[k*E] | cnt[E]
  int* sm_total = sm_dyn;
  int* sm_base = sm_dyn + k * E;
  int* sm_cnt = sm_dyn + 2 * k * E;
  __shared__ float sm_red[kTileTokens / 32];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // (1) per (choice, expert): tokens in all tiles / in the tiles before this one
  for (int p = threadIdx.x; p < k * E; p += kTileTokens) {
    int total = 0, before = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int h = hist[static_cast<long long>(t) * k * E + p];
      total += h;
      before += (t < b) ? h : 0;
    }
    sm_total[p] = total;
    sm_base[p] = before;
  }
  __syncthreads();
  // (2) choice j queues behind ALL (j-1)-th choices (tutel/impls/fast_dispatch.py:160-166)
  for (int p = threadIdx.x; p < k * E; p += kTileTokens) {
    const int j = p / E, e = p - j * E;
    int prior = 0;
    for (int jj = 0; jj < j; ++jj) prior += sm_total[jj * E + e];
    sm_base[p] += prior;       // own slot only: no other thread reads sm_base[p] before the barrier below
  }
  __syncthreads();
  if (b == 0) {
    // per-expert token counts, first-choice counts (fp32, for the backward pass) and the GShard loss
    float part = 0.0f;
    for (int e = threadIdx.x; e < E; e += kTileTokens) {
      int c = 0;
      for (int j = 0; j < k; ++j) c += sm_total[j * E + e];
      counts[e] = c;
      const float ce = static_cast<float>(sm_total[e]);
      if (ce_out != nullptr) ce_out[e] = ce;
      float me = 0.0f;
      for (int t = 0; t < ntiles; ++t) me += me_partial[static_cast<long long>(t) * E + e];
      part += me * ce;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) sm_red[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0 && l_aux != nullptr) {
      float tot = 0.0f;
      for (int w = 0; w < kTileTokens / 32; ++w) tot += sm_red[w];
      stf<T>(l_aux, tot * static_cast<float>(E) / (static_cast<float>(S) * static_cast<float>(S)));
    }
  }
  // (3) stable rank of every token inside its tile, choice by choice
  const int s = b * kTileTokens + threadIdx.x;
  for (int j = 0; j < k; ++j) {
    for (int e = threadIdx.x; e < E; e += kTileTokens) sm_cnt[e] = 0;
    __syncthreads();
    int e = -1;
    if (s < S) {
      e = idx[static_cast<long long>(j) * S + s];
      if (e >= E) e = -1;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, e);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    int base = 0;
    for (int w = 0; w < kTileTokens / 32; ++w) {
      if (warp == w && e >= 0 && lane == leader) {
        base = sm_cnt[e];
        sm_cnt[e] = base + __popc(peers);
      }
      __syncthreads();
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (s < S) {
      int l = kInvalidLoc;
      if (e >= 0) {
        l = sm_base[j * E + e] + base + rank_in_warp;
        if (slot_src != nullptr && l < C) slot_src[static_cast<long long>(e) * C + l] = s * k + j;
      }
      loc[static_cast<long long>(j) * S + s] = l;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// gate backward (one warp per token), dgates fp32 [k,S], dl / dlogits in the logits' dtype:
//   r_j  = p[idx_j], D = sum_j r_j, Dc = max(D, eps);  g_j = r_j / Dc (normalize && k>1)
//   dr_j = dg_j / Dc - [D > eps] * (sum_i dg_i r_i) / Dc^2
//   dp_e = dl * ce_e * E / S^2 + sum_j [idx_j == e] dr_j        (l_aux = E/S^2 * sum_e me_e ce_e; ce is constant)
//   dlogit_e = p_e * (dp_e - sum_e' dp_e' p_e')
// ------------------------------------------------------------------------------------------------
template <typename T, int VPT>
__global__ void __launch_bounds__(256)
gate_route_bwd_kernel(const float* __restrict__ scores, const int* __restrict__ idx, const float* __restrict__ top,
                      const float* __restrict__ dgates, const float* __restrict__ ce, const T* __restrict__ dl,
                      T* __restrict__ dlogits, int S, int E, int k, int normalize, float eps) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const float aux_scale = (dl != nullptr ? ldf<T>(dl) : 0.0f) * static_cast<float>(E) /
                          (static_cast<float>(S) * static_cast<float>(S));
  for (long long s = static_cast<long long>(blockIdx.x) * 8 + warp; s < S; s += static_cast<long long>(gridDim.x) * 8) {
    float p[VPT], dp[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int e = lane + 32 * i;
      p[i] = e < E ? scores[s * E + e] : 0.0f;
      dp[i] = (e < E && ce != nullptr) ? aux_scale * ce[e] : 0.0f;
    }
    // lane j owns choice j
    const float r = lane < k ? top[static_cast<long long>(lane) * S + s] : 0.0f;
    const float dg = (lane < k && dgates != nullptr) ? dgates[static_cast<long long>(lane) * S + s] : 0.0f;
    const int my_e = lane < k ? idx[static_cast<long long>(lane) * S + s] : -1;
    float D = r, dot = dg * r;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      D += __shfl_xor_sync(0xffffffffu, D, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    float dr = dg;
    if (normalize != 0 && k > 1) {
      const float Dc = fmaxf(D, eps);
      dr = dg / Dc - (D > eps ? dot / (Dc * Dc) : 0.0f);
    }
    for (int j = 0; j < k; ++j) {
      const int e = __shfl_sync(0xffffffffu, my_e, j);
      const float d = __shfl_sync(0xffffffffu, dr, j);
      if (e >= 0 && (e & 31) == lane) {
#pragma unroll
        for (int i = 0; i < VPT; ++i)
          if (i == (e >> 5)) dp[i] += d;
      }
    }
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) acc += dp[i] * p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int e = lane + 32 * i;
      if (e < E) stf<T>(dlogits + s * E + e, p[i] * (dp[i] - acc));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// grouped column sums: out[g, n] = sum_t x[g, t, n]   (bias gradients), 16-byte loads, fp32 accumulate
// ------------------------------------------------------------------------------------------------
template <typename T> struct V16;
template <> struct V16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void add(const uint4& u, float* a) {
    a[0] += __uint_as_float(u.x); a[1] += __uint_as_float(u.y); a[2] += __uint_as_float(u.z); a[3] += __uint_as_float(u.w);
  }
};
template <> struct V16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void add(const uint4& u, float* a) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      a[2 * i] += t.x; a[2 * i + 1] += t.y;
    }
  }
};
template <> struct V16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void add(const uint4& u, float* a) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[2 * i] += __uint_as_float(w[i] << 16);
      a[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
};

// block = 256 threads = 16 row-lanes x 16 column-lanes; a block owns a strip of 16 * V columns of one group and the rows
// [split * rows_per_split, ...).  Each thread keeps 4 independent 16-byte loads in flight.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long ld, long long group_stride, T* __restrict__ out, float* __restrict__ acc_out,
              int rows, int N, int rows_per_split) {
  constexpr int V = V16<T>::N;
  __shared__ float sm[16][16 * V + 1];
  const int cl = threadIdx.x & 15;
  const int rl = threadIdx.x >> 4;
  const int col = (blockIdx.x * 16 + cl) * V;
  const int g = blockIdx.z;
  const int r_begin = blockIdx.y * rows_per_split;
  const int r_end = min(rows, r_begin + rows_per_split);
  float a[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = 0.0f;
  if (col < N) {
    const T* base = x + static_cast<long long>(g) * group_stride + col;
    int r = r_begin + rl;
    for (; r + 48 < r_end; r += 64) {
      uint4 u[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) u[q] = ptx::ld_nc_v4(base + static_cast<long long>(r + 16 * q) * ld);
#pragma unroll
      for (int q = 0; q < 4; ++q) V16<T>::add(u[q], a);
    }
    for (; r < r_end; r += 16) V16<T>::add(ptx::ld_nc_v4(base + static_cast<long long>(r) * ld), a);
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sm[rl][cl * V + i] = a[i];
  __syncthreads();
  for (int c = threadIdx.x; c < 16 * V; c += 256) {
    float t = 0.0f;
#pragma unroll
    for (int q = 0; q < 16; ++q) t += sm[q][c];
    const int n = blockIdx.x * 16 * V + c;
    if (n < N) {
      if (acc_out != nullptr) atomicAdd(acc_out + static_cast<long long>(g) * N + n, t);
      else stf<T>(out + static_cast<long long>(g) * N + n, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// public column scan (`tutel.moe.fast_cumsum_sub_one`): out[s, e] = sum_{s' <= s} in[s', e] - 1
// three linear passes over row tiles: tile sums, exclusive scan of the tile sums, in-tile scan
// ------------------------------------------------------------------------------------------------
constexpr int kScanRows = 32;

__global__ void __launch_bounds__(128)
cumsum_tile_kernel(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ tile_sums, int S, int E,
                   int tiles_per_block, int apply) {
  const int ecols = E < 128 ? E : 128;
  const int tl = threadIdx.x / ecols;
  const int e = blockIdx.x * ecols + (threadIdx.x - tl * ecols);
  const long long tile = static_cast<long long>(blockIdx.y) * tiles_per_block + tl;
  if (tl >= tiles_per_block || e >= E || tile * kScanRows >= S) return;
  const int r0 = static_cast<int>(tile) * kScanRows;
  const int r1 = min(S, r0 + kScanRows);
  int run = apply ? tile_sums[tile * E + e] - 1 : 0;
  for (int r = r0; r < r1; ++r) {
    run += in[static_cast<long long>(r) * E + e];
    if (apply) out[static_cast<long long>(r) * E + e] = run;
  }
  if (!apply) tile_sums[tile * E + e] = run;
}

__global__ void __launch_bounds__(128)
cumsum_scan_kernel(int* __restrict__ tile_sums, int ntiles, int E) {
  const int e = blockIdx.x * 128 + threadIdx.x;
  if (e >= E) return;
  int run = 0;
  for (int t = 0; t < ntiles; ++t) {
    const int v = tile_sums[static_cast<long long>(t) * E + e];
    tile_sums[static_cast<long long>(t) * E + e] = run;
    run += v;
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace

int gate_route_tiles(int S) { return (S + kTileTokens - 1) / kTileTokens; }

template <typename T>
static cudaError_t gate_route_forward_t(const void* logits, float* scores, int* idx, float* top, float* gates,
                                        float* me_partial, int* hist, int* loc, int* counts, int* slot_src,
                                        float* ce_out, void* l_aux, int S, int E, int k, int C, bool normalize, float eps,
                                        cudaStream_t stream) {
  const int ntiles = gate_route_tiles(S);
  const long long slot_n = slot_src != nullptr ? static_cast<long long>(E) * C : 0;
  const size_t smem1 = sizeof(int) * static_cast<size_t>(k) * E + sizeof(float) * E;
  const size_t smem2 = sizeof(int) * (2 * static_cast<size_t>(k) * E + E);
  if (smem1 > 48 * 1024 || smem2 > 48 * 1024) return cudaErrorInvalidValue;
#define TB_GR(VPTv)                                                                                                   \
  gate_route_kernel<T, VPTv><<<ntiles, kGateThreads, smem1, stream>>>(static_cast<const T*>(logits), scores, idx, top, \
                                                                      gates, me_partial, hist, slot_src, slot_n, S, E, \
                                                                      k, normalize ? 1 : 0, eps)
  if (E <= 32) TB_GR(1);
  else if (E <= 64) TB_GR(2);
  else if (E <= 128) TB_GR(4);
  else if (E <= 256) TB_GR(8);
  else if (E <= 512) TB_GR(16);
  else return cudaErrorInvalidValue;
#undef TB_GR
  route_finish_kernel<T><<<ntiles, kTileTokens, smem2, stream>>>(idx, hist, me_partial, loc, counts, slot_src, ce_out,
                                                                 static_cast<T*>(l_aux), S, E, k, C, ntiles);
  return cudaGetLastError();
}

cudaError_t gate_route_forward(const void* logits, float* scores, int* idx, float* top, float* gates, float* me_partial,
                               int* hist, int* loc, int* counts, int* slot_src, float* ce_out, void* l_aux, int S, int E,
                               int k, int C, bool normalize, float eps, int elem_type, cudaStream_t stream) {
  if (S <= 0 || k > 32 || k <= 0) return cudaErrorInvalidValue;
  switch (elem_type) {
    case ET_F32: return gate_route_forward_t<float>(logits, scores, idx, top, gates, me_partial, hist, loc, counts, slot_src, ce_out, l_aux, S, E, k, C, normalize, eps, stream);
    case ET_F16: return gate_route_forward_t<__half>(logits, scores, idx, top, gates, me_partial, hist, loc, counts, slot_src, ce_out, l_aux, S, E, k, C, normalize, eps, stream);
    case ET_BF16: return gate_route_forward_t<__nv_bfloat16>(logits, scores, idx, top, gates, me_partial, hist, loc, counts, slot_src, ce_out, l_aux, S, E, k, C, normalize, eps, stream);
  }
  return cudaErrorInvalidValue;
}

template <typename T>
static cudaError_t gate_route_backward_t(const float* scores, const int* idx, const float* top, const float* dgates,
                                         const float* ce, const void* dl, void* dlogits, int S, int E, int k,
                                         bool normalize, float eps, cudaStream_t stream) {
  const long long want = (static_cast<long long>(S) + 7) / 8;
  const int grid = static_cast<int>(want < 4LL * sm_count() ? want : 4LL * sm_count());
#define TB_GRB(VPTv)                                                                                          \
  gate_route_bwd_kernel<T, VPTv><<<grid, 256, 0, stream>>>(scores, idx, top, dgates, ce, static_cast<const T*>(dl), \
                                                           static_cast<T*>(dlogits), S, E, k, normalize ? 1 : 0, eps)
  if (E <= 32) TB_GRB(1);
  else if (E <= 64) TB_GRB(2);
  else if (E <= 128) TB_GRB(4);
  else if (E <= 256) TB_GRB(8);
  else if (E <= 512) TB_GRB(16);
  else return cudaErrorInvalidValue;
#undef TB_GRB
  return cudaGetLastError();
}

cudaError_t gate_route_backward(const float* scores, const int* idx, const float* top, const float* dgates,
                                const float* ce, const void* dl, void* dlogits, int S, int E, int k, bool normalize,
                                float eps, int elem_type, cudaStream_t stream) {
  if (S <= 0) return cudaSuccess;
  if (k > 32) return cudaErrorInvalidValue;
  switch (elem_type) {
    case ET_F32: return gate_route_backward_t<float>(scores, idx, top, dgates, ce, dl, dlogits, S, E, k, normalize, eps, stream);
    case ET_F16: return gate_route_backward_t<__half>(scores, idx, top, dgates, ce, dl, dlogits, S, E, k, normalize, eps, stream);
    case ET_BF16: return gate_route_backward_t<__nv_bfloat16>(scores, idx, top, dgates, ce, dl, dlogits, S, E, k, normalize, eps, stream);
  }
  return cudaErrorInvalidValue;
}

int colsum_row_splits(int G, int rows, int N, int elem_bytes) {
  const int strips = (N + 16 * (16 / elem_bytes) - 1) / (16 * (16 / elem_bytes));
  const long long blocks = static_cast<long long>(strips) * G;
  int splits = 1;
  while (blocks * splits < 2LL * sm_count() && rows / (splits * 2) >= 64) splits *= 2;
  return splits;
}

template <typename T>
static cudaError_t colsum_t(const void* x, long long ld, long long group_stride, void* out, float* acc, int G, int rows,
                            int N, int splits, cudaStream_t stream) {
  constexpr int V = V16<T>::N;
  const int strips = (N + 16 * V - 1) / (16 * V);
  const int rps = (rows + splits - 1) / splits;
  dim3 grid(strips, splits, G);
  colsum_kernel<T><<<grid, 256, 0, stream>>>(static_cast<const T*>(x), ld, group_stride, static_cast<T*>(out), acc, rows, N, rps);
  return cudaGetLastError();
}

cudaError_t grouped_colsum(const void* x, long long ld, long long group_stride, void* out, float* acc, int G, int rows,
                           int N, int splits, int elem_type, cudaStream_t stream) {
  if (G <= 0 || rows <= 0 || N <= 0) return cudaSuccess;
  const int eb = elem_type == ET_F32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || ((ld * eb) & 15) || ((group_stride * eb) & 15) || (N * eb) % 16)
    return cudaErrorInvalidValue;
  if (splits > 1 && acc == nullptr) return cudaErrorInvalidValue;
  if (splits <= 1) acc = nullptr;
  switch (elem_type) {
    case ET_F32: return colsum_t<float>(x, ld, group_stride, out, acc, G, rows, N, splits, stream);
    case ET_F16: return colsum_t<__half>(x, ld, group_stride, out, acc, G, rows, N, splits, stream);
    case ET_BF16: return colsum_t<__nv_bfloat16>(x, ld, group_stride, out, acc, G, rows, N, splits, stream);
  }
  return cudaErrorInvalidValue;
}

size_t cumsum_workspace_ints(int S, int E) { return static_cast<size_t>((S + kScanRows - 1) / kScanRows) * E; }

cudaError_t cumsum_sub_one(const int* in, int* out, int* workspace, int S, int E, cudaStream_t stream) {
  if (S <= 0 || E <= 0) return cudaSuccess;
  const int ntiles = (S + kScanRows - 1) / kScanRows;
  const int ecols = E < 128 ? E : 128;
  const int tpb = 128 / ecols;
  dim3 grid((E + ecols - 1) / ecols, (ntiles + tpb - 1) / tpb);
  cumsum_tile_kernel<<<grid, 128, 0, stream>>>(in, out, workspace, S, E, tpb, 0);
  cumsum_scan_kernel<<<(E + 127) / 128, 128, 0, stream>>>(workspace, ntiles, E);
  cumsum_tile_kernel<<<grid, 128, 0, stream>>>(in, out, workspace, S, E, tpb, 1);
  return cudaGetLastError();
}

}  // namespace tb
