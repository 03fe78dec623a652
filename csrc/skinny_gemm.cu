// Skinny grouped GEMM for dropless / decoder-inference routing: a handful of tokens per expert, many experts.
//
// The reference's "Megablocks" path (tutel/custom/custom_kernel.cpp:874-889) copies the per-expert counts to the host,
// synchronises, and loops over experts with one cuBLAS call each.  With <= a few rows per expert the problem is purely
// bound by streaming the weights of the ACTIVE experts once; this kernel does exactly that, driven by the device-side
// counts (experts with zero tokens cost nothing), in fp32 / fp16 / bf16 with fp32 accumulation:
//
//      y[g, r, :] = act( x[g, r, :] @ W[g] (+ bias[g]) )     for r < counts[g]
//
// W is [G, N, K] ("nk", one warp per output column, lanes stride K) or [G, K, N] ("kn", lanes stride N).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "moe_kernels.h"

namespace tb {
namespace {

constexpr int kRows = 8;       // rows (tokens) handled per pass
constexpr int kKChunk = 1024;  // K elements of x staged in smem per pass

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// grid: (ceil(N / cols_per_block), G); block: 256 threads
template <typename T, bool KN>
__global__ void __launch_bounds__(256)
skinny_kernel(const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ bias, T* __restrict__ y,
              const int* __restrict__ counts, int rows_cap, int N, int K, int relu) {
  __shared__ float xs[kRows][kKChunk];
  const int g = blockIdx.y;
  int count = counts != nullptr ? min(counts[g], rows_cap) : rows_cap;
  if (count <= 0) return;
  const T* xg = x + static_cast<long long>(g) * rows_cap * K;
  const T* wg = w + static_cast<long long>(g) * N * K;
  T* yg = y + static_cast<long long>(g) * rows_cap * N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int r0 = 0; r0 < count; r0 += kRows) {
    const int nr = min(kRows, count - r0);
    if constexpr (KN) {
      // one output column per thread; W rows (fixed k) are read coalesced across the block
      const int n = blockIdx.x * 256 + threadIdx.x;
      float acc[kRows];
#pragma unroll
      for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
      for (int k0 = 0; k0 < K; k0 += kKChunk) {
        const int kc = min(kKChunk, K - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < nr * kc; i += 256) xs[i / kc][i % kc] = ldf<T>(xg + static_cast<long long>(r0 + i / kc) * K + k0 + i % kc);
        __syncthreads();
        if (n < N) {
          for (int k = 0; k < kc; ++k) {
            const float wv = ldf<T>(wg + static_cast<long long>(k0 + k) * N + n);
#pragma unroll
            for (int r = 0; r < kRows; ++r) acc[r] = fmaf(xs[r][k], wv, acc[r]);
          }
        }
      }
      if (n < N) {
        const float b = bias != nullptr ? ldf<T>(bias + static_cast<long long>(g) * N + n) : 0.0f;
        for (int r = 0; r < nr; ++r) {
          float v = acc[r] + b;
          if (relu) v = fmaxf(v, 0.0f);
          stf<T>(yg + static_cast<long long>(r0 + r) * N + n, v);
        }
      }
    } else {
      // one output column per warp (8 per block-iteration); lanes stride K, warp-reduce at the end
      for (int nb = blockIdx.x * 64; nb < min(N, blockIdx.x * 64 + 64); nb += 8) {
        const int n = nb + warp;
        float acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0.0f;
        for (int k0 = 0; k0 < K; k0 += kKChunk) {
          const int kc = min(kKChunk, K - k0);
          __syncthreads();
          for (int i = threadIdx.x; i < nr * kc; i += 256) xs[i / kc][i % kc] = ldf<T>(xg + static_cast<long long>(r0 + i / kc) * K + k0 + i % kc);
          __syncthreads();
          if (n < N) {
            const T* wrow = wg + static_cast<long long>(n) * K + k0;
            for (int k = lane; k < kc; k += 32) {
              const float wv = ldf<T>(wrow + k);
#pragma unroll
              for (int r = 0; r < kRows; ++r) acc[r] = fmaf(xs[r][k], wv, acc[r]);
            }
          }
        }
        if (n < N) {
#pragma unroll
          for (int r = 0; r < kRows; ++r)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
          if (lane == 0) {
            const float b = bias != nullptr ? ldf<T>(bias + static_cast<long long>(g) * N + n) : 0.0f;
            for (int r = 0; r < nr; ++r) {
              float v = acc[r] + b;
              if (relu) v = fmaxf(v, 0.0f);
              stf<T>(yg + static_cast<long long>(r0 + r) * N + n, v);
            }
          }
        }
      }
    }
  }
}

template <typename T>
cudaError_t launch(const void* x, const void* w, const void* bias, void* y, const int* counts, int G, int rows_cap, int N,
                   int K, bool kn, bool relu, cudaStream_t stream) {
  if (G <= 0 || rows_cap <= 0 || N <= 0 || K <= 0) return cudaSuccess;
  if (kn) {
    dim3 grid((N + 255) / 256, G);
    skinny_kernel<T, true><<<grid, 256, 0, stream>>>(static_cast<const T*>(x), static_cast<const T*>(w),
                                                     static_cast<const T*>(bias), static_cast<T*>(y), counts, rows_cap, N,
                                                     K, relu ? 1 : 0);
  } else {
    dim3 grid((N + 63) / 64, G);
    skinny_kernel<T, false><<<grid, 256, 0, stream>>>(static_cast<const T*>(x), static_cast<const T*>(w),
                                                      static_cast<const T*>(bias), static_cast<T*>(y), counts, rows_cap, N,
                                                      K, relu ? 1 : 0);
  }
  return cudaGetLastError();
}

}  // namespace

cudaError_t skinny_grouped_gemm(const void* x, const void* w, const void* bias, void* y, const int* counts, int G,
                                int rows_cap, int N, int K, bool w_is_kn, bool relu, int elem_type, cudaStream_t stream) {
  switch (elem_type) {
    case ET_F32: return launch<float>(x, w, bias, y, counts, G, rows_cap, N, K, w_is_kn, relu, stream);
    case ET_F16: return launch<__half>(x, w, bias, y, counts, G, rows_cap, N, K, w_is_kn, relu, stream);
    case ET_BF16: return launch<__nv_bfloat16>(x, w, bias, y, counts, G, rows_cap, N, K, w_is_kn, relu, stream);
  }
  return cudaErrorInvalidValue;
}

}  // namespace tb
