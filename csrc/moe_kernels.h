// Host API of the routing / sparse dispatch kernels (moe_kernels.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tb {

enum ElemType : int { ET_F32 = 0, ET_F16 = 1, ET_BF16 = 2 };

// Stable slot assignment.  idx: [k, S] int32 expert id of the j-th choice of token s (or <0 = none).
// Produces loc[k, S] (position of the token inside its expert's queue: all 1st choices in token order, then all
// 2nd choices, ... - tutel/impls/fast_dispatch.py:155-171) and counts[E] (tokens routed to e, all choices).
// workspace: int32[ (nblocks + 1) * k * E ], nblocks = ceil(S / 1024).
cudaError_t route_locations(const int* idx, int* loc, int* counts, int* workspace, int S, int E, int k,
                            cudaStream_t stream);
size_t route_workspace_ints(int S, int E, int k);

// slot_src[E*C]: token*k + j occupying that slot, or -1.  Must be called after route_locations.
cudaError_t build_slot_map(const int* idx, const int* loc, int* slot_src, int S, int E, int k, int C,
                           cudaStream_t stream);

// out[slot, :] = scale * x[token(slot), :]   (zeros for empty slots);  scale = gates[j*S + token] or 1.
// dst_ptr_table (optional, uint64[E]): base pointer of expert e's [C, M] block - may be PEER memory (dispatch
// fusion); when null the block of expert e is out + e*C*M.
// signal_ptr_table/signal_rows (optional): after rows [r0, r0+signal_rows) of expert e are stored, the uint32
// counter at signal_ptr_table[e] + (r0/signal_rows) is set to `signal_value` with st.release.sys (each chunk has a
// single publisher - the block that finishes the chunk's last 16-row unit, elected through `chunk_counters`
// (uint32[E * ceil(C/signal_rows)], zero-initialised, self-resetting) - so the epoch number itself is published;
// 0 = increment instead).
cudaError_t encode_rows(const void* x, const void* gates, const int* slot_src, void* out,
                        const unsigned long long* dst_ptr_table, const unsigned long long* signal_ptr_table,
                        unsigned int* chunk_counters, int signal_rows, int S, int E, int k, int C, int M, int elem_type,
                        int rot_chunks, int signal_value, cudaStream_t stream);

// out[s, :] = sum_j w_j * buf[idx_j[s]*C + loc_j[s], :]  (choices with loc >= C or idx < 0 contribute 0).
// wait_flags (optional): uint32[E] counters that must reach wait_target (acquire.sys) before expert e's rows
// are read (combine fusion).
cudaError_t decode_rows(const void* buf, const void* gates, const int* idx, const int* loc, void* out,
                        const uint32_t* wait_flags, uint32_t wait_target, int S, int E, int k, int C, int M,
                        int elem_type, cudaStream_t stream);

// dgate[j*S + s] = dot(a[s, :], buf[slot_j(s), :])   (fp32 accumulate, 0 for dropped choices).
cudaError_t gate_grad(const void* a, const void* buf, const int* idx, const int* loc, void* dgate, int S, int E,
                      int k, int C, int M, int elem_type, cudaStream_t stream);

// Fused gating forward: logits[S,E] (fp32) -> softmax scores, top-k ids, raw top-k scores, and the per-expert
// partial sums needed by the GShard auxiliary loss, one warp per token.
cudaError_t gate_topk_forward(const float* logits, float* scores, int* idx, float* topk_scores, float* me_partial,
                              int* ce_partial, int S, int E, int k, cudaStream_t stream);

// Fused gating backward (see the kernel for the formulas): gradients of the normalised top-k gates [k,S] and of the
// GShard loss (device scalar `dl`, may be null) -> d logits [S,E].  `ce` = first-choice counts per expert (fp32 [E]).
cudaError_t gate_topk_backward(const float* scores, const int* idx, const float* topk_scores, const float* dgates,
                               const float* ce, const float* dl, float* dlogits, int S, int E, int k, bool normalize,
                               float eps, cudaStream_t stream);

// q[r, :] = e4m3(x[r, :] / scale[r]),  scale[r] = max|x[r, :]| / 448   (one scale per row; rows are K-major GEMM
// operands, so the scale factors out of the dot product and is applied in the GEMM epilogue).
cudaError_t quantize_rows_e4m3(const void* x, void* q, float* scale, long long R, int K, int elem_type,
                               cudaStream_t stream);

// Dropless / decoder inference: y[g, r, :] = act(x[g, r, :] @ W[g] + bias[g]) for r < counts[g] (device counts, no
// host sync); x [G, rows_cap, K], y [G, rows_cap, N], W [G, N, K] or (w_is_kn) [G, K, N].  Rows past the count are
// left untouched.  (csrc/skinny_gemm.cu)
cudaError_t skinny_grouped_gemm(const void* x, const void* w, const void* bias, void* y, const int* counts, int G,
                                int rows_cap, int N, int K, bool w_is_kn, bool relu, int elem_type, cudaStream_t stream);

}  // namespace tb
