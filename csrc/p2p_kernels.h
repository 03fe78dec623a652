// In-kernel peer-to-peer collectives over NVLink (p2p_kernels.cu).  All buffers live in the symmetric heap
// (symm_heap.h): every rank holds the same offsets, `peer_table[r]` is rank r's heap base mapped in this process.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tb {

constexpr int kMaxPeers = 16;

struct PushPlan {
  // For every destination peer p: copy `bytes[p]` bytes from local `src + src_off[p]` to
  // peer_table[p] + recv_off(p) + dst_off[p], where recv_off(p) is the receive-buffer offset peer p announced for this
  // call through the mailbox (multiple of 256 bytes).   All offsets/sizes are multiples of 16 bytes... or not:
  // unaligned tails are handled with byte copies.
  long long src_off[kMaxPeers];
  long long dst_off[kMaxPeers];
  long long bytes[kMaxPeers];
};

// Counters (uint32, inside the heap, zero-initialised, monotonically increasing):
//   mail[W] (uint64) at heap offset mail_off: mail[src] on rank d = (epoch << 32 | recv_offset / 256) posted by src:
//             "for call `epoch`, push my data to this offset of my arena".   done[W] (uint32) at done_off: done[src]
//             on rank d counts finished pushes of src into d.
//   scratch[W] at scratch_off: LOCAL block-arrival counters used to elect the last block per destination.
// `epoch` is the 1-based call number on this (ready, done) counter pair.
cudaError_t p2p_push(const void* src, const PushPlan& plan, const unsigned long long* peer_table,
                     long long recv_heap_off, long long mail_off, long long done_off, long long scratch_off, int rank,
                     int world, uint32_t epoch, int blocks_per_peer, bool small_blocks, cudaStream_t stream);

// out[i] = sum_p peer_p[stage_off + slice_off + i]  for i in [0, n)   (one-shot pull-reduce; fp32 accumulate).
// Callers bracket it with p2p_barrier so that all stages are written / may be overwritten.
cudaError_t p2p_reduce_slice(void* out, const unsigned long long* peer_table, long long stage_off,
                             long long slice_off_bytes, long long n_elems, int elem_type, int rank, int world,
                             bool is_max, cudaStream_t stream);

// One-shot all-reduce in ONE launch (small tensors: gate gradients, capacity maxima, scalars).  Every rank stores its
// input into slot [parity][rank] of EVERY peer's inbox (16-byte NVLink stores), publishes one flag per (source, block)
// with st.release.sys, waits for the W flags of its own block and reduces the W copies in rank order (so all ranks
// obtain bit-identical results).  Inbox: heap offset inbox_off, 2 x W x slot_bytes;  flags: uint32 at flag_off,
// 2 x W x kOneShotMaxBlocks.  `epoch` = 1-based call number on this inbox (parity = epoch & 1); successive calls must
// be issued on one stream.  in/out may alias.  Supported: fp32/fp16/bf16 (fp32 accumulate), int32, int64; SUM and MAX.
constexpr int kOneShotMaxBlocks = 64;
cudaError_t p2p_allreduce_oneshot(const void* in, void* out, const unsigned long long* peer_table, long long inbox_off,
                                  long long slot_bytes, long long flag_off, long long n_elems, int elem_type, int rank,
                                  int world, uint32_t epoch, bool is_max, cudaStream_t stream);

// dst[(i % rows) * cols + i / rows, :] = src[i, :] for i in [0, rows*cols): the [rows, cols] -> [cols, rows] block
// transpose of `width`-byte records that re-orders all-to-all payloads between the intra-node and the inter-node phase
// of the 2-D hierarchical exchange (reference: the stride-copy kernel of tutel/custom/custom_kernel.cpp:408-429).
cudaError_t p2p_stride_copy(const void* src, void* dst, int rows, int cols, long long width_bytes, cudaStream_t stream);

// All ranks arrive and wait (system-scope release/acquire); `epoch` 1-based per counter array at bar_off (uint32[W]).
cudaError_t p2p_barrier(const unsigned long long* peer_table, long long bar_off, int rank, int world, uint32_t epoch,
                        cudaStream_t stream);

}  // namespace tb
