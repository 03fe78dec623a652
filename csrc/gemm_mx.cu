// MX block-scaled fp8 GEMM for sm_100a:  D[g] = A[g] * B[g]^T  with e4m3 operands that carry one UE8M0 scale per
// 32 consecutive K elements (OCP MX), multiplied INSIDE the tensor core:
//     tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale  [d_tmem], a_desc, b_desc, idesc, [sfa_tmem], [sfb_tmem], p
// The reference has no reduced-precision expert path at all (tutel/experts/ffn.py runs torch.matmul in the model
// dtype); the row-scaled e4m3 path of gemm_sm100.cu is what the fused engine uses, this kernel is the finer-grained
// alternative (outliers only cost the 32 elements next to them their precision, not the whole row).
//
// Layout of one CTA (192 threads, one 128 x BN output tile, K walked in 128-element = 128-byte steps):
//   warp 0      TMA producer: A tile [128 x 128 B] and B tile [BN x 128 B] (SWIZZLE_128B) plus the two scale atoms
//               (512 B per 128 rows, plain bulk copies) per stage, all completing on the stage's "full" mbarrier
//   warp 1      one elected lane: tcgen05.cp (scales smem -> TMEM; 32 lanes x 4 columns per 128 rows, replicated over
//               the four lane quarters) followed by four K=32 MMAs whose descriptors select byte 0..3 of those columns;
//               tcgen05.commit releases the stage.  tcgen05.cp and tcgen05.mma of one thread execute in issue order,
//               so the scale columns are single-buffered.
//   warps 2-5   epilogue: tcgen05.ld (thread = accumulator row), optional ReLU, bf16, 16-byte global stores
// BN = 256: 4 stages (198 KB), one CTA per SM.  BN = 128: 3 stages (100 KB) so that TWO CTAs share an SM (TMEM 2 x 256
// columns) and one's epilogue hides behind the other's main loop.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <mutex>

#include "gemm_mx.h"
#include "moe_kernels.h"
#include "ptx.cuh"

namespace tb {
namespace {

constexpr int kBM = 128;
constexpr int kBK = 128;                 // e4m3 elements = bytes per K step (one 128-byte swizzle row)
constexpr int kSfAtomBytes = 512;        // scales of 128 rows x 128 K elements
constexpr int kMxThreads = 192;

template <int BN>
struct MxCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 3;
  static constexpr uint32_t A_BYTES = kBM * kBK;
  static constexpr uint32_t B_BYTES = BN * kBK;
  static constexpr uint32_t OP_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t SFB_BYTES = kSfAtomBytes * (BN / 128);
  static constexpr uint32_t SF_BYTES = kSfAtomBytes + SFB_BYTES;
  static constexpr uint32_t BAR_BYTES = 192;
  static constexpr uint32_t SMEM_BYTES = 1024 + STAGES * (OP_BYTES + SF_BYTES) + BAR_BYTES;
  static constexpr uint32_t SFA_COL = BN;            // TMEM columns: [0, BN) accumulator, then 4 of A scales, then B's
  static constexpr uint32_t SFB_COL = BN + 4;
  static constexpr uint32_t TMEM_COLS = (BN == 256) ? 512 : 256;
};

struct MxArgs {
  const uint8_t* sfa;
  const uint8_t* sfb;
  __nv_bfloat16* d;
  long long ldd, d_group_stride;
  int M, N, K, G;
  int tiles_m, tiles_n;
  int sfa_row_tiles, sfb_row_tiles;   // 128-row tiles of the scale arrays
  int relu;
  int sf_addr_plain;   // debug: do not mirror the scale byte index into bits [30,32) of the scale TMEM addresses
};

__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}

// 32 rows x 16 bytes of shared memory -> TMEM lanes 0..31 (copied to all four lane quarters), 4 columns.
__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__device__ __forceinline__ void umma_mxf8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t sfa,
                                          uint32_t sfb, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
      : "memory");
}

// Output tiles are walked in bands of 8 row tiles so that co-resident CTAs share A and B tiles in L2.
__device__ __forceinline__ void decode_tile(long long t, int tiles_m, int tiles_n, int& g, int& m_blk, int& n_blk) {
  const long long per_group = static_cast<long long>(tiles_m) * tiles_n;
  g = static_cast<int>(t / per_group);
  const int r = static_cast<int>(t % per_group);
  constexpr int kBand = 8;
  const int band = r / (kBand * tiles_n);
  const int in_band = r % (kBand * tiles_n);
  const int rows = min(kBand, tiles_m - band * kBand);
  m_blk = band * kBand + in_band % rows;
  n_blk = in_band / rows;
}

template <int BN>
__global__ void __launch_bounds__(kMxThreads, (BN == 256) ? 1 : 2)
mx_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const MxArgs args) {
  using C = MxCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sf_base = smem_base + C::STAGES * C::OP_BYTES;
  const uint32_t bar_base = sf_base + C::STAGES * C::SF_BYTES;
  auto smem_a = [&](int s) { return smem_base + s * C::OP_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * C::OP_BYTES + C::A_BYTES; };
  auto smem_sfa = [&](int s) { return sf_base + s * C::SF_BYTES; };
  auto smem_sfb = [&](int s) { return sf_base + s * C::SF_BYTES + kSfAtomBytes; };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 64u + 8u * s; };
  const uint32_t tfull_bar = bar_base + 128u;
  const uint32_t tmem_slot = bar_base + 136u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
  }
  if (warp == 2 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tfull_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<1>(tmem_slot, C::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  int g, m_blk, n_blk;
  decode_tile(blockIdx.x, args.tiles_m, args.tiles_n, g, m_blk, n_blk);
  const int m0 = m_blk * kBM;
  const int n0 = n_blk * BN;
  const int num_kb = args.K / kBK;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    int s = 0;
    uint32_t ph = 0;
    const uint8_t* sfa_g = args.sfa + static_cast<long long>(g) * num_kb * args.sfa_row_tiles * kSfAtomBytes;
    const uint8_t* sfb_g = args.sfb + static_cast<long long>(g) * num_kb * args.sfb_row_tiles * kSfAtomBytes;
    for (int kb = 0; kb < num_kb; ++kb) {
      ptx::mbar_wait(empty_bar(s), ph ^ 1u);
      if (ptx::elect_one()) {
        const uint32_t fb = full_bar(s);
        ptx::mbar_expect_tx(fb, C::OP_BYTES + C::SF_BYTES);
        ptx::tma_load_3d(smem_a(s), &tmA, fb, kb * kBK, m0, g);
        ptx::tma_load_3d(smem_b(s), &tmB, fb, kb * kBK, n0, g);
        bulk_load(smem_sfa(s), sfa_g + (static_cast<long long>(kb) * args.sfa_row_tiles + m_blk) * kSfAtomBytes,
                  kSfAtomBytes, fb);
        bulk_load(smem_sfb(s),
                  sfb_g + (static_cast<long long>(kb) * args.sfb_row_tiles + n_blk * (BN / 128)) * kSfAtomBytes,
                  C::SFB_BYTES, fb);
      }
      __syncwarp();
      if (++s == C::STAGES) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    int s = 0;
    uint32_t ph = 0;
    // Operand descriptors: K-major, SWIZZLE_128B, 8-row groups 1024 B apart; a K=32 step advances the start by 32 B.
    constexpr uint32_t op_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    // Scale descriptors: K-major, no swizzle: 8-row x 16-byte core matrices of 128 contiguous bytes, 128 B apart.
    constexpr uint32_t sf_hi = (128u >> 4) | (1u << 14);
    // Instruction descriptor (block-scaled form): A/B format e4m3 (0), both K-major, N >> 3 at [17,23),
    // scale format UE8M0 at [23], M >> 4 at [24,29); the scale byte of a K=32 step goes to [4,6) (B) and [29,31) (A).
    constexpr uint32_t idesc0 = (static_cast<uint32_t>(BN >> 3) << 17) | (1u << 23) | (static_cast<uint32_t>(kBM >> 4) << 24);
    const uint32_t d_tmem = tmem_base;
    const uint32_t sfa_tmem = tmem_base + C::SFA_COL;
    const uint32_t sfb_tmem = tmem_base + C::SFB_COL;
    for (int kb = 0; kb < num_kb; ++kb) {
      ptx::mbar_wait(full_bar(s), ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t a_lo = ((smem_a(s) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = ((smem_b(s) >> 4) & 0x3FFFu) | (1u << 16);
        tmem_cp_32x128b_warpx4(sfa_tmem, (static_cast<uint64_t>(sf_hi) << 32) | ((smem_sfa(s) >> 4) & 0x3FFFu) | (1u << 16));
#pragma unroll
        for (int j = 0; j < BN / 128; ++j)
          tmem_cp_32x128b_warpx4(sfb_tmem + 4u * j, (static_cast<uint64_t>(sf_hi) << 32) |
                                                        (((smem_sfb(s) + j * kSfAtomBytes) >> 4) & 0x3FFFu) | (1u << 16));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = (static_cast<uint64_t>(op_hi) << 32) | (a_lo + 2u * k);
          const uint64_t bd = (static_cast<uint64_t>(op_hi) << 32) | (b_lo + 2u * k);
          const uint32_t idesc = idesc0 | (static_cast<uint32_t>(k) << 4) | (static_cast<uint32_t>(k) << 29);
          const uint32_t sub = args.sf_addr_plain ? 0u : (static_cast<uint32_t>(k) << 30);
          umma_mxf8(d_tmem, ad, bd, idesc, sfa_tmem + sub, sfb_tmem + sub, (kb | k) != 0);
        }
        ptx::umma_commit<1>(empty_bar(s));
        if (kb == num_kb - 1) ptx::umma_commit<1>(tfull_bar);
      }
      __syncwarp();
      if (++s == C::STAGES) { s = 0; ph ^= 1u; }
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = m0 + q * 32 + lane;
    ptx::mbar_wait(tfull_bar, 0);
    ptx::tc_fence_after();
    __nv_bfloat16* drow = args.d + static_cast<long long>(g) * args.d_group_stride + static_cast<long long>(row) * args.ldd + n0;
    const bool relu = args.relu != 0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
      ptx::tmem_ld_wait();
      if (row < args.M) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo = __uint_as_float(r[v * 8 + 2 * j]);
            float hi = __uint_as_float(r[v * 8 + 2 * j + 1]);
            if (relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
            const __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
            w[j] = *reinterpret_cast<const uint32_t*>(&p);
          }
          *reinterpret_cast<uint4*>(drow + c * 32 + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<1>(tmem_base, C::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// quantiser: 16-bit rows -> e4m3 + UE8M0 block scales in the atom layout of gemm_mx.h
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// Four consecutive threads own one 32-element block (8 elements = one 16-byte load each).  The shared exponent is the
// smallest power of two that brings the block's largest magnitude inside e4m3's finite range (448):
//     e = ceil(log2(amax / 448)),   q = rn_satfinite(x * 2^-e),   scale byte = e + 127.
template <typename T>
__global__ void __launch_bounds__(256)
mx_quantize_kernel(const T* __restrict__ x, uint8_t* __restrict__ q, uint8_t* __restrict__ sf, long long total, int R, int K,
                   int row_tiles) {
  const int k8 = K / 8;
  const int num_kb = K / kBK;
  for (long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x; i0 < total; i0 += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = i0 + threadIdx.x;
    const bool valid = i < total;
    float f[8];
    float amax = 0.f;
    if (valid) {
      const uint4 raw = ptx::ld_nc_v4(x + i * 8);
      const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = to_f32<T>(e[j]);
        amax = fmaxf(amax, fabsf(f[j]));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = 0.f;
    }
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    const uint32_t bits = __float_as_uint(amax * (1.0f / 448.0f));
    int e = static_cast<int>((bits >> 23) & 0xFFu) - 127 + ((bits & 0x7FFFFFu) != 0u ? 1 : 0);
    e = max(-127, min(126, e));
    const float inv = __uint_as_float(static_cast<uint32_t>(127 - e) << 23);   // 2^-e
    if (valid) {
      const __nv_fp8x4_e4m3 lo(make_float4(f[0] * inv, f[1] * inv, f[2] * inv, f[3] * inv));
      const __nv_fp8x4_e4m3 hi(make_float4(f[4] * inv, f[5] * inv, f[6] * inv, f[7] * inv));
      uint2 w;
      w.x = *reinterpret_cast<const uint32_t*>(&lo);
      w.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(q + i * 8) = w;
      if ((threadIdx.x & 3) == 0) {
        const long long grow = i / k8;
        const int k = static_cast<int>(i % k8) * 8;
        const int g = static_cast<int>(grow / R);
        const int r = static_cast<int>(grow % R);
        const long long atom = (static_cast<long long>(g) * num_kb + k / kBK) * row_tiles + r / 128;
        sf[atom * kSfAtomBytes + (r % 32) * 16 + ((r % 128) / 32) * 4 + (k % kBK) / 32] = static_cast<uint8_t>(e + 127);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn mx_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// e4m3 [groups, rows, k] row-major -> boxes of `box_rows` rows x 128 bytes, 128-byte swizzle
bool mx_operand_map(CUtensorMap* map, const void* base, long long rows, long long k, int groups, int box_rows) {
  EncodeTiledFn enc = mx_encode_fn();
  if (enc == nullptr) return false;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(groups)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows) * static_cast<cuuint64_t>(k)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
cudaError_t mx_launch(const MxGemmProblem& p, cudaStream_t stream, const char** why) {
  using C = MxCfg<BN>;
  CUtensorMap ta, tb_;
  if (!mx_operand_map(&ta, p.a, p.M, p.K, p.G, kBM) || !mx_operand_map(&tb_, p.b, p.N, p.K, p.G, BN)) {
    if (why) *why = "cuTensorMapEncodeTiled failed for an MX operand";
    return cudaErrorInvalidValue;
  }
  MxArgs a;
  a.sfa = static_cast<const uint8_t*>(p.sfa);
  a.sfb = static_cast<const uint8_t*>(p.sfb);
  a.d = static_cast<__nv_bfloat16*>(p.d);
  a.ldd = p.ldd;
  a.d_group_stride = p.d_group_stride;
  a.M = p.M; a.N = p.N; a.K = p.K; a.G = p.G;
  a.tiles_m = (p.M + kBM - 1) / kBM;
  a.tiles_n = p.N / BN;
  a.sfa_row_tiles = (p.M + 127) / 128;
  a.sfb_row_tiles = (p.N + 127) / 128;
  a.relu = p.relu;
  a.sf_addr_plain = p.sf_addr_plain;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(mx_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  });
  if (attr_err != cudaSuccess) return attr_err;
  const long long tiles = static_cast<long long>(a.tiles_m) * a.tiles_n * p.G;
  mx_gemm_kernel<BN><<<static_cast<unsigned>(tiles), kMxThreads, C::SMEM_BYTES, stream>>>(ta, tb_, a);
  return cudaGetLastError();
}

}  // namespace

cudaError_t mx_gemm_launch(const MxGemmProblem& p, cudaStream_t stream, const char** why) {
  auto fail = [&](const char* msg) { if (why) *why = msg; return cudaErrorInvalidValue; };
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || p.G <= 0) return fail("empty MX GEMM");
  if (p.K % kBK != 0) return fail("MX GEMM: K must be a multiple of 128");
  if (p.N % 128 != 0) return fail("MX GEMM: N must be a multiple of 128");
  if ((reinterpret_cast<uintptr_t>(p.a) | reinterpret_cast<uintptr_t>(p.b) | reinterpret_cast<uintptr_t>(p.sfa) |
       reinterpret_cast<uintptr_t>(p.sfb) | reinterpret_cast<uintptr_t>(p.d)) & 15)
    return fail("MX GEMM: operands must be 16-byte aligned");
  if (p.ldd % 8 != 0 || p.d_group_stride % 8 != 0) return fail("MX GEMM: output strides must be multiples of 8 elements");
  int bn = p.block_n;
  if (bn == 0) bn = (p.N % 256 == 0) ? 256 : 128;
  if (bn == 256 && p.N % 256 != 0) return fail("MX GEMM: block_n 256 needs N % 256 == 0");
  if (static_cast<long long>((p.M + kBM - 1) / kBM) * (p.N / bn) * p.G > 0x7fffffffLL) return fail("MX GEMM: too many tiles");
  if (bn == 256) return mx_launch<256>(p, stream, why);
  if (bn == 128) return mx_launch<128>(p, stream, why);
  return fail("MX GEMM: block_n must be 128 or 256");
}

cudaError_t mx_quantize(const void* x, void* q, void* sf, int groups, int rows, int k, int elem_type, cudaStream_t stream) {
  if (k % kBK != 0 || (elem_type != ET_F16 && elem_type != ET_BF16)) return cudaErrorInvalidValue;
  const long long total = static_cast<long long>(groups) * rows * (k / 8);
  if (total == 0) return cudaSuccess;
  const int row_tiles = (rows + 127) / 128;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
  if (elem_type == ET_BF16)
    mx_quantize_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(q),
                                                                  static_cast<uint8_t*>(sf), total, rows, k, row_tiles);
  else
    mx_quantize_kernel<__half><<<blocks, 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<uint8_t*>(q),
                                                           static_cast<uint8_t*>(sf), total, rows, k, row_tiles);
  return cudaGetLastError();
}

cudaError_t set_spin_timeout_mx(unsigned long long ns) {
  return cudaMemcpyToSymbol(tb_spin_timeout_ns, &ns, sizeof(ns));
}

}  // namespace tb
