// CPU kernels + their bindings.  See cpu_kernels.h.
#include "cpu_kernels.h"

#include <torch/extension.h>

#include <cstring>
#include <vector>

namespace tb {

void cpu_route_locations(const int32_t* idx, int32_t* loc, int32_t* counts, int S, int E, int k) {
  std::vector<int32_t> running(E, 0);
  for (int j = 0; j < k; ++j) {
    for (int s = 0; s < S; ++s) {
      const int e = idx[static_cast<int64_t>(j) * S + s];
      if (e < 0 || e >= E) { loc[static_cast<int64_t>(j) * S + s] = 0x3fffffff; continue; }
      loc[static_cast<int64_t>(j) * S + s] = running[e]++;
    }
  }
  std::memcpy(counts, running.data(), sizeof(int32_t) * E);
}

template <typename T>
void cpu_encode(const T* x, const T* gates, const int32_t* idx, const int32_t* loc, T* out, int S, int E, int k, int C,
                int M) {
  std::memset(out, 0, sizeof(T) * static_cast<size_t>(E) * C * M);
  for (int j = 0; j < k; ++j) {
    for (int s = 0; s < S; ++s) {
      const int e = idx[static_cast<int64_t>(j) * S + s], l = loc[static_cast<int64_t>(j) * S + s];
      if (e < 0 || e >= E || l < 0 || l >= C) continue;
      const T g = gates ? gates[static_cast<int64_t>(j) * S + s] : T(1);
      const T* src = x + static_cast<int64_t>(s) * M;
      T* dst = out + (static_cast<int64_t>(e) * C + l) * M;
      for (int m = 0; m < M; ++m) dst[m] = g * src[m];
    }
  }
}

template <typename T>
void cpu_decode(const T* buf, const T* gates, const int32_t* idx, const int32_t* loc, T* out, int S, int E, int k, int C,
                int M) {
  for (int s = 0; s < S; ++s) {
    T* dst = out + static_cast<int64_t>(s) * M;
    for (int m = 0; m < M; ++m) dst[m] = T(0);
    for (int j = 0; j < k; ++j) {
      const int e = idx[static_cast<int64_t>(j) * S + s], l = loc[static_cast<int64_t>(j) * S + s];
      if (e < 0 || e >= E || l < 0 || l >= C) continue;
      const T g = gates ? gates[static_cast<int64_t>(j) * S + s] : T(1);
      const T* src = buf + (static_cast<int64_t>(e) * C + l) * M;
      for (int m = 0; m < M; ++m) dst[m] += g * src[m];
    }
  }
}

template <typename T>
void cpu_gate_grad(const T* a, const T* buf, const int32_t* idx, const int32_t* loc, T* dgate, int S, int E, int k, int C,
                   int M) {
  for (int j = 0; j < k; ++j) {
    for (int s = 0; s < S; ++s) {
      const int e = idx[static_cast<int64_t>(j) * S + s], l = loc[static_cast<int64_t>(j) * S + s];
      T acc = T(0);
      if (e >= 0 && e < E && l >= 0 && l < C) {
        const T* pa = a + static_cast<int64_t>(s) * M;
        const T* pb = buf + (static_cast<int64_t>(e) * C + l) * M;
        for (int m = 0; m < M; ++m) acc += pa[m] * pb[m];
      }
      dgate[static_cast<int64_t>(j) * S + s] = acc;
    }
  }
}

}  // namespace tb

namespace {

void check_cpu(const at::Tensor& t) { TORCH_CHECK(!t.is_cuda() && t.is_contiguous(), "expected contiguous CPU tensor"); }

std::vector<at::Tensor> cpu_route(const at::Tensor& idx, int64_t E) {
  check_cpu(idx);
  TORCH_CHECK(idx.scalar_type() == at::kInt && idx.dim() == 2);
  at::Tensor loc = at::empty_like(idx);
  at::Tensor counts = at::empty({E}, idx.options());
  tb::cpu_route_locations(idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(), counts.data_ptr<int32_t>(),
                          static_cast<int>(idx.size(1)), static_cast<int>(E), static_cast<int>(idx.size(0)));
  return {loc, counts};
}

at::Tensor cpu_encode(const at::Tensor& x, const c10::optional<at::Tensor>& gates, const at::Tensor& idx,
                      const at::Tensor& loc, int64_t E, int64_t C) {
  check_cpu(x); check_cpu(idx); check_cpu(loc);
  const int S = static_cast<int>(x.size(0)), M = static_cast<int>(x.size(1)), k = static_cast<int>(idx.size(0));
  at::Tensor out = at::empty({E * C, M}, x.options());
  const bool hg = gates.has_value() && gates->defined();
  if (hg) { check_cpu(*gates); TORCH_CHECK(gates->scalar_type() == x.scalar_type()); }
  AT_DISPATCH_FLOATING_TYPES(x.scalar_type(), "cpu_encode", [&] {
    tb::cpu_encode<scalar_t>(x.data_ptr<scalar_t>(), hg ? gates->data_ptr<scalar_t>() : nullptr, idx.data_ptr<int32_t>(),
                             loc.data_ptr<int32_t>(), out.data_ptr<scalar_t>(), S, static_cast<int>(E), k,
                             static_cast<int>(C), M);
  });
  return out;
}

at::Tensor cpu_decode(const at::Tensor& buf, const c10::optional<at::Tensor>& gates, const at::Tensor& idx,
                      const at::Tensor& loc, int64_t E, int64_t C) {
  check_cpu(buf); check_cpu(idx); check_cpu(loc);
  const int S = static_cast<int>(idx.size(1)), k = static_cast<int>(idx.size(0));
  const int M = static_cast<int>(buf.numel() / (E * C));
  at::Tensor out = at::empty({S, M}, buf.options());
  const bool hg = gates.has_value() && gates->defined();
  if (hg) { check_cpu(*gates); TORCH_CHECK(gates->scalar_type() == buf.scalar_type()); }
  AT_DISPATCH_FLOATING_TYPES(buf.scalar_type(), "cpu_decode", [&] {
    tb::cpu_decode<scalar_t>(buf.data_ptr<scalar_t>(), hg ? gates->data_ptr<scalar_t>() : nullptr,
                             idx.data_ptr<int32_t>(), loc.data_ptr<int32_t>(), out.data_ptr<scalar_t>(), S,
                             static_cast<int>(E), k, static_cast<int>(C), M);
  });
  return out;
}

at::Tensor cpu_gate_grad(const at::Tensor& a, const at::Tensor& buf, const at::Tensor& idx, const at::Tensor& loc,
                         int64_t E, int64_t C) {
  check_cpu(a); check_cpu(buf); check_cpu(idx); check_cpu(loc);
  const int S = static_cast<int>(idx.size(1)), k = static_cast<int>(idx.size(0)), M = static_cast<int>(a.size(1));
  at::Tensor out = at::empty({k, S}, a.options());
  AT_DISPATCH_FLOATING_TYPES(a.scalar_type(), "cpu_gate_grad", [&] {
    tb::cpu_gate_grad<scalar_t>(a.data_ptr<scalar_t>(), buf.data_ptr<scalar_t>(), idx.data_ptr<int32_t>(),
                                loc.data_ptr<int32_t>(), out.data_ptr<scalar_t>(), S, static_cast<int>(E), k,
                                static_cast<int>(C), M);
  });
  return out;
}

}  // namespace

void register_cpu_bindings(pybind11::module& m) {
  m.def("cpu_route_locations", &cpu_route);
  m.def("cpu_encode", &cpu_encode);
  m.def("cpu_decode", &cpu_decode);
  m.def("cpu_gate_grad", &cpu_gate_grad);
}
