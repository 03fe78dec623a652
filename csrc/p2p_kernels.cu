// Peer-to-peer collectives written as plain CUDA kernels over NVLink-mapped memory: the B200-native replacement
// for the reference's grouped ncclSend/ncclRecv all-to-alls (tutel/custom/custom_kernel.cpp:463-518, 520-654)
// and c10d all_to_all_single (tutel/impls/communicate.py:181-192).  One launch = handshake + payload + completion:
//
//   1. every rank posts, in each peer's mailbox, WHERE in its arena this call's data must land (epoch | offset/256,
//      one st.release.sys.u64) - receive buffers are therefore chosen per call and per rank (zero-copy results)
//   2. CTAs push their share of the payload with 16-byte stores into the destination GPU's heap
//   3. each CTA publishes completion with fence.acq_rel.sys + red.release.sys on the peer's done[] counter
//   4. the kernel does not exit before all of its own inbound pushes are complete (ld.acquire.sys polling with a
//      bounded spin), so plain stream order makes the received data visible to the next kernel.
#include "p2p_kernels.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "moe_kernels.h"
#include "ptx.cuh"

namespace tb {
namespace {

constexpr int kPushThreads = 512;

__global__ void __launch_bounds__(kPushThreads)
p2p_push_kernel(const uint8_t* __restrict__ src, const PushPlan plan, const unsigned long long* __restrict__ peer_table,
                long long recv_heap_off, long long mail_off, long long done_off, long long scratch_off, int rank,
                int world, uint32_t epoch, int blocks_per_peer) {
  const int pi = blockIdx.x / blocks_per_peer;           // which peer (rotated so that traffic is spread)
  const int sub = blockIdx.x - pi * blocks_per_peer;     // which slice of that peer's payload
  const int peer = (rank + pi) % world;
  uint8_t* peer_base = reinterpret_cast<uint8_t*>(peer_table[peer]);
  uint8_t* my_base = reinterpret_cast<uint8_t*>(peer_table[rank]);

  // (1) mailboxes: tell every peer where this rank receives this call's data (also the "buffer is free" credit)
  if (blockIdx.x == 0 && threadIdx.x < world) {
    uint8_t* pb = reinterpret_cast<uint8_t*>(peer_table[threadIdx.x]);
    const unsigned long long word = (static_cast<unsigned long long>(epoch) << 32) |
                                    static_cast<unsigned long long>(recv_heap_off >> 8);
    ptx::st_release_sys_u64(reinterpret_cast<unsigned long long*>(pb + mail_off) + rank, word);
  }
  // (2) wait for the destination's mailbox entry, then push
  __shared__ unsigned long long dst_heap_off_s;
  if (threadIdx.x == 0)
    dst_heap_off_s = static_cast<unsigned long long>(ptx::wait_mailbox_sys(
                         reinterpret_cast<const unsigned long long*>(my_base + mail_off) + peer, epoch)) << 8;
  __syncthreads();
  const long long dst_heap_off = static_cast<long long>(dst_heap_off_s);

  const long long total = plan.bytes[peer];
  const uint8_t* s = src + plan.src_off[peer];
  uint8_t* d = peer_base + dst_heap_off + plan.dst_off[peer];
  const bool aligned = (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0);
  if (aligned) {
    const long long nvec = total >> 4;
    const long long per = (nvec + blocks_per_peer - 1) / blocks_per_peer;
    const long long v0 = per * sub;
    const long long v1 = min(nvec, v0 + per);
    const uint4* sv = reinterpret_cast<const uint4*>(s);
    uint4* dv = reinterpret_cast<uint4*>(d);
    long long v = v0 + threadIdx.x;
    for (; v + 3LL * kPushThreads < v1; v += 4LL * kPushThreads) {
      uint4 a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = ptx::ld_nc_v4(sv + v + u * kPushThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u) ptx::st_na_v4(dv + v + u * kPushThreads, a[u]);
    }
    for (; v < v1; v += kPushThreads) ptx::st_na_v4(dv + v, ptx::ld_nc_v4(sv + v));
    if (sub == blocks_per_peer - 1)
      for (long long b = (nvec << 4) + threadIdx.x; b < total; b += kPushThreads) d[b] = s[b];
  } else {
    const long long per = (total + blocks_per_peer - 1) / blocks_per_peer;
    const long long b0 = per * sub, b1 = min(total, b0 + per);
    for (long long b = b0 + threadIdx.x; b < b1; b += kPushThreads) d[b] = s[b];
  }
  // (3) completion: the LAST block working for this peer publishes one release.sys increment, so the receiver's
  //     target is simply the call number (it does not need to know how many blocks the sender used)
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t* local_cnt = reinterpret_cast<uint32_t*>(my_base + scratch_off) + peer;
    uint32_t prev;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(local_cnt) : "memory");
    if (prev == static_cast<uint32_t>(blocks_per_peer) - 1u) {
      *local_cnt = 0u;  // every block of this call has arrived; re-arm for the next call (stream-ordered)
      ptx::fence_acq_rel_sys();
      ptx::red_add_release_sys(reinterpret_cast<uint32_t*>(peer_base + done_off) + rank, 1u);
    }
  }
  // (4) inbound completion: block i (< world) watches source i
  if (blockIdx.x < world && threadIdx.x == 0)
    ptx::wait_flag_ge_sys(reinterpret_cast<const uint32_t*>(my_base + done_off) + blockIdx.x, epoch);
}

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256)
p2p_reduce_slice_kernel(T* __restrict__ out, const unsigned long long* __restrict__ peer_table, long long stage_off,
                        long long slice_off_bytes, long long n, int rank, int world, bool is_max) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float acc = is_max ? -INFINITY : 0.0f;
    for (int q = 0; q < world; ++q) {
      const int p = (rank + q) % world;
      const T* pp = reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(peer_table[p]) + stage_off + slice_off_bytes);
      const float v = ldf<T>(pp + i);
      acc = is_max ? fmaxf(acc, v) : acc + v;
    }
    stf<T>(out + i, acc);
  }
}

__global__ void p2p_barrier_kernel(const unsigned long long* __restrict__ peer_table, long long bar_off, int rank,
                                   int world, uint32_t epoch) {
  const int p = threadIdx.x;
  if (p < world) {
    uint8_t* pb = reinterpret_cast<uint8_t*>(peer_table[p]);
    ptx::fence_acq_rel_sys();
    ptx::red_add_release_sys(reinterpret_cast<uint32_t*>(pb + bar_off) + rank, 1u);
    const uint8_t* mb = reinterpret_cast<const uint8_t*>(peer_table[rank]);
    ptx::wait_flag_ge_sys(reinterpret_cast<const uint32_t*>(mb + bar_off) + p, epoch);
  }
}

}  // namespace

cudaError_t p2p_push(const void* src, const PushPlan& plan, const unsigned long long* peer_table,
                     long long recv_heap_off, long long mail_off, long long done_off, long long scratch_off, int rank,
                     int world, uint32_t epoch, int blocks_per_peer, cudaStream_t stream) {
  if (world > kMaxPeers) return cudaErrorInvalidValue;
  if (blocks_per_peer < 1) blocks_per_peer = 1;
  const int grid = world * blocks_per_peer;
  p2p_push_kernel<<<grid, kPushThreads, 0, stream>>>(static_cast<const uint8_t*>(src), plan, peer_table, recv_heap_off,
                                                     mail_off, done_off, scratch_off, rank, world, epoch, blocks_per_peer);
  return cudaGetLastError();
}

cudaError_t p2p_reduce_slice(void* out, const unsigned long long* peer_table, long long stage_off,
                             long long slice_off_bytes, long long n_elems, int elem_type, int rank, int world,
                             bool is_max, cudaStream_t stream) {
  if (n_elems <= 0) return cudaSuccess;
  long long want = (n_elems + 255) / 256;
  const int grid = static_cast<int>(want < 592 ? want : 592);
  switch (elem_type) {
    case ET_F32:
      p2p_reduce_slice_kernel<float><<<grid, 256, 0, stream>>>(static_cast<float*>(out), peer_table, stage_off,
                                                               slice_off_bytes, n_elems, rank, world, is_max);
      break;
    case ET_F16:
      p2p_reduce_slice_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<__half*>(out), peer_table, stage_off,
                                                                slice_off_bytes, n_elems, rank, world, is_max);
      break;
    case ET_BF16:
      p2p_reduce_slice_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<__nv_bfloat16*>(out), peer_table,
                                                                       stage_off, slice_off_bytes, n_elems, rank,
                                                                       world, is_max);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t p2p_barrier(const unsigned long long* peer_table, long long bar_off, int rank, int world, uint32_t epoch,
                        cudaStream_t stream) {
  p2p_barrier_kernel<<<1, 32, 0, stream>>>(peer_table, bar_off, rank, world, epoch);
  return cudaGetLastError();
}

}  // namespace tb
