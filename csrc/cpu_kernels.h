// CPU reference kernels (fp32 / fp64) for routing and sparse dispatch - counterpart of
// tutel/custom/custom_kernel.cpp:280-323 (`invoke_cpu<dtype>`), slot-centric like the CUDA path.
#pragma once
#include <cstdint>

namespace tb {

// idx [k,S] -> loc [k,S], counts [E]; stable token order, j-th choices queue after all (j-1)-th choices.
void cpu_route_locations(const int32_t* idx, int32_t* loc, int32_t* counts, int S, int E, int k);

template <typename T>
void cpu_encode(const T* x, const T* gates /*[k,S] or null*/, const int32_t* idx, const int32_t* loc, T* out /*[E*C,M], zeroed here*/,
                int S, int E, int k, int C, int M);
template <typename T>
void cpu_decode(const T* buf, const T* gates, const int32_t* idx, const int32_t* loc, T* out /*[S,M]*/, int S, int E, int k,
                int C, int M);
template <typename T>
void cpu_gate_grad(const T* a, const T* buf, const int32_t* idx, const int32_t* loc, T* dgate /*[k,S]*/, int S, int E, int k,
                   int C, int M);

}  // namespace tb
