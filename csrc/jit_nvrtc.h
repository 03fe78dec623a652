// Runtime CUDA-C compilation for the public `tutel_b200.jit.create_cuda_kernel` API
// (parity with tutel/jit.py:4 / tutel/custom/custom_kernel.cpp:94-275), implemented with NVRTC (dlopen'ed) and the
// CUDA runtime's library-management API - no nvcc fork/exec, no driver-API link dependency; kernels launch on the
// CURRENT stream (the reference uses the legacy default stream, custom_kernel.cpp:268-274).
#pragma once
