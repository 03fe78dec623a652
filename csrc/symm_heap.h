// Symmetric heap: one cudaMalloc'd arena per rank, exported with CUDA IPC and mapped into every peer process on the
// node, so that kernels can address any rank's buffers directly over NVLink.  B200-native replacement for the
// reference's private NCCL communicators (tutel/custom/custom_kernel.cpp:327-431).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace tb {

class SymmHeap {
 public:
  SymmHeap(size_t bytes, int device);
  ~SymmHeap();
  SymmHeap(const SymmHeap&) = delete;
  SymmHeap& operator=(const SymmHeap&) = delete;

  std::string ipc_handle() const;                                   // 64 opaque bytes
  void open_peers(int rank, const std::vector<std::string>& handles);  // handles[r] from rank r (own entry ignored)
  void close();

  size_t bytes() const { return bytes_; }
  int rank() const { return rank_; }
  int world() const { return static_cast<int>(peer_base_.size()); }
  void* base(int r) const { return peer_base_.at(r); }
  const unsigned long long* device_peer_table() const { return d_peer_table_; }

 private:
  size_t bytes_ = 0;
  int device_ = 0;
  int rank_ = 0;
  void* local_ = nullptr;
  std::vector<void*> peer_base_;
  unsigned long long* d_peer_table_ = nullptr;
  bool closed_ = false;
};

}  // namespace tb
