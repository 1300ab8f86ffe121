// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM,
// operands staged by TMA into 128B-swizzled shared memory, mbarrier producer/consumer ring).
//
//     D[M,N] = act( A[M,K] * B[N,K]^T + bias[N] (+ D_old) )
//
// This one kernel serves every dense contraction of the decoder path (SURVEY.md §2.1 op map):
//   forward : att_embed / ctx2att / fc_embed prologue, the per-step gate GEMMs (K-concatenated
//             [xt|h] x [W_i2h|W_h2h]), a2c, h2att, the LSTMCell GEMMs and the vocab logit GEMM
//             (reference: nn.Linear / nn.LSTMCell calls in models/AttModel.py:79-92,426-441,574-592);
//   backward: dgrad (B operand MN-major = the untransposed weight) and wgrad (both operands
//             MN-major = the untransposed activations), so no transposed copies are ever made.
//
// Tile: BLOCK_M = 128 (one tcgen05.mma M=128, cta_group::1), BLOCK_N in {64,128}, BLOCK_K = 64
// (= one 128-byte swizzle row of bf16).  8 warps: warp 0 TMA producer, warp 1 MMA issuer (one
// elected thread), warp 2 TMEM allocator, warps 4-7 epilogue (TMEM -> registers -> global).
#include <map>
#include <mutex>
#include <string>
#include <tuple>

#include <cuda_fp16.h>

#include "uic_internal.h"
#include "uic_ptx.cuh"

namespace uic {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 256;

struct GemmEpilogue {
  float* c_f32;
  long long ldc;
  __nv_bfloat16* c_bf16;
  long long ldcb;
  const float* bias;
  int relu;
  int accumulate;
  int out_f16;  // the 16-bit output is IEEE fp16 instead of bf16
  int exp_col0;     // columns >= exp_col0 become exp_scale * exp(2 x) when exp_scale != 0
  float exp_scale;
};

__device__ __forceinline__ uint32_t pack16(float a, float b, int f16) {
  if (f16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return f2_to_bf16x2(a, b);
}
__device__ __forceinline__ __nv_bfloat16 store16(float a, int f16) {
  if (f16) {
    __half h = __float2half_rn(a);
    return *reinterpret_cast<__nv_bfloat16*>(&h);
  }
  return __float2bfloat16_rn(a);
}

template <int BN, int STAGES>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
};

template <int BN, int STAGES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         GemmEpilogue ep, int M, int N, int K) {
  using L = GemmSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        if (!A_MN) {
          tma_load_2d(sa, &tmap_a, &full_bar[s], kb * BK, m0);
        } else {  // A stored [K, M]: two boxes of 64 m-columns x 64 k-rows
#pragma unroll
          for (int i = 0; i < BM / 64; ++i) tma_load_2d(sa + i * (BK * 128), &tmap_a, &full_bar[s], m0 + 64 * i, kb * BK);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmap_b, &full_bar[s], kb * BK, n0);
        } else {
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) tma_load_2d(sb + i * (BK * 128), &tmap_b, &full_bar[s], n0 + 64 * i, kb * BK);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t a_base = smem_u32(smem + s * L::STAGE_BYTES);
        const uint32_t b_base = a_base + L::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major : 8-row groups 1024 B apart (SBO); a 16-element K step is 32 B inside the swizzle row.
          // MN-major: 64-element MN chunks BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO);
          //           a 16-row K step is 2048 B.
          const uint64_t da = A_MN ? make_smem_desc_sw128(a_base + k * 2048, BK * 128, 1024)
                                   : make_smem_desc_sw128(a_base + k * 32, 16, 1024);
          const uint64_t db = B_MN ? make_smem_desc_sw128(b_base + k * 2048, BK * 128, 1024)
                                   : make_smem_desc_sw128(b_base + k * 32, 16, 1024);
          umma_bf16_ss(tmem_base, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may access
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int row = m0 + ew * 32 + lane;
    const bool row_ok = row < M;
    const bool vec_f32 = ep.c_f32 != nullptr && (ep.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.c_f32) & 15) == 0);
    const bool vec_bf16 = ep.c_bf16 != nullptr && (ep.ldcb % 8 == 0) && ((reinterpret_cast<uintptr_t>(ep.c_bf16) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-row predicated stores
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      const int col0 = n0 + c * 32;
      if (row_ok && col0 < N) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      const bool full = col0 + 32 <= N;
      if (ep.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || col0 + j < N) f[j] += __ldg(ep.bias + col0 + j);
      }
      if (ep.accumulate && ep.c_f32 != nullptr) {
        const float* src = ep.c_f32 + static_cast<long long>(row) * ep.ldc + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || col0 + j < N) f[j] += src[j];
      }
      if (ep.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
      }
      if (ep.exp_scale != 0.0f && col0 + 32 > ep.exp_col0) {  // attention operands: scale * exp(2x), see attention.cu
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j >= ep.exp_col0) f[j] = fminf(ep.exp_scale * __expf(2.0f * f[j]), ep.out_f16 ? 65504.0f : 1.0e30f);
      }
      if (ep.c_f32 != nullptr) {
        float* dst = ep.c_f32 + static_cast<long long>(row) * ep.ldc + col0;
        if (full && vec_f32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N) dst[j] = f[j];
        }
      }
      if (ep.c_bf16 != nullptr) {
        __nv_bfloat16* dst = ep.c_bf16 + static_cast<long long>(row) * ep.ldcb + col0;
        if (full && vec_bf16) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 q;
            q.x = pack16(f[j], f[j + 1], ep.out_f16);
            q.y = pack16(f[j + 2], f[j + 3], ep.out_f16);
            q.z = pack16(f[j + 4], f[j + 5], ep.out_f16);
            q.w = pack16(f[j + 6], f[j + 7], ep.out_f16);
            *reinterpret_cast<uint4*>(dst + j) = q;
          }
        } else {
          for (int j = 0; j < 32; ++j)
            if (col0 + j < N) dst[j] = store16(f[j], ep.out_f16);
        }
      }
      }  // row_ok
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BN);
}

// ------------------------------------------------------------------------------------------------
// Verification kernel: plain CUDA-core tiled GEMM with the same contract.  Used by the test-suite
// to cross-check the tensor-core kernel on the device and selectable with UIC_GEMM=simt for
// debugging; never the default.
// ------------------------------------------------------------------------------------------------
__global__ void gemm_bf16_simt_kernel(const __nv_bfloat16* __restrict__ A, long long lda, int a_mn,
                                      const __nv_bfloat16* __restrict__ B, long long ldb, int b_mn, GemmEpilogue ep,
                                      int M, int N, int K) {
  __shared__ float sa[16][17];
  __shared__ float sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty;
  const int col = blockIdx.x * 16 + tx;
  float acc = 0.0f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    {  // A tile: element (row, k0+tx)
      const int k = k0 + tx;
      float v = 0.0f;
      if (row < M && k < K) v = __bfloat162float(a_mn ? A[static_cast<long long>(k) * lda + row] : A[static_cast<long long>(row) * lda + k]);
      sa[ty][tx] = v;
    }
    {  // B tile: element (n = blockIdx.x*16+ty, k0+tx)
      const int n = blockIdx.x * 16 + ty;
      const int k = k0 + tx;
      float v = 0.0f;
      if (n < N && k < K) v = __bfloat162float(b_mn ? B[static_cast<long long>(k) * ldb + n] : B[static_cast<long long>(n) * ldb + k]);
      sb[ty][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(sa[ty][k], sb[tx][k], acc);
    __syncthreads();
  }
  if (row < M && col < N) {
    if (ep.bias) acc += ep.bias[col];
    if (ep.accumulate && ep.c_f32) acc += ep.c_f32[static_cast<long long>(row) * ep.ldc + col];
    if (ep.relu) acc = fmaxf(acc, 0.0f);
    if (ep.exp_scale != 0.0f && col >= ep.exp_col0) acc = fminf(ep.exp_scale * __expf(2.0f * acc), ep.out_f16 ? 65504.0f : 1.0e30f);
    if (ep.c_f32) ep.c_f32[static_cast<long long>(row) * ep.ldc + col] = acc;
    if (ep.c_bf16) ep.c_bf16[static_cast<long long>(row) * ep.ldcb + col] = store16(acc, ep.out_f16);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Interned "gemm_MxNxK" labels for the live profiler (pointers must outlive the records).
static const char* gemm_label(int M, int N, int K) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int>, std::string> names;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_tuple(M, N, K);
  auto it = names.find(key);
  if (it == names.end()) {
    char buf[64];
    snprintf(buf, sizeof(buf), "gemm_%dx%dx%d", M, N, K);
    it = names.emplace(key, buf).first;
  }
  return it->second.c_str();
}

template <int BN, int STAGES, bool A_MN, bool B_MN>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                     cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES, A_MN, B_MN>;
  constexpr int smem = GemmSmem<BN, STAGES>::TOTAL;
  static bool configured = false;  // benign race: attribute set is idempotent
  if (!configured) {
    UIC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  launch_begin(gemm_label(M, N, K), stream);
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, ep, M, N, K);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

template <int BN, int STAGES>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep,
                          int M, int N, int K, cudaStream_t stream) {
  if (!a_mn && !b_mn) return launch_tc<BN, STAGES, false, false>(ta, tb, ep, M, N, K, stream);
  if (!a_mn && b_mn) return launch_tc<BN, STAGES, false, true>(ta, tb, ep, M, N, K, stream);
  if (a_mn && !b_mn) return launch_tc<BN, STAGES, true, false>(ta, tb, ep, M, N, K, stream);
  return launch_tc<BN, STAGES, true, true>(ta, tb, ep, M, N, K, stream);
}

int gemm_bf16(const void* A, long long lda, const void* B, long long ldb, float* c_f32, long long ldc, void* c_bf16,
              long long ldcb, const float* bias, int M, int N, int K, int flags, cudaStream_t stream, int exp_col0,
              float exp_scale) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(UIC_ERR_SHAPE, "gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  if (c_f32 == nullptr && c_bf16 == nullptr) return set_error(UIC_ERR_ARG, "gemm_bf16: no output buffer");
  const bool a_mn = flags & UIC_GEMM_A_MN_MAJOR;
  const bool b_mn = flags & UIC_GEMM_B_MN_MAJOR;
  GemmEpilogue ep{c_f32, ldc, static_cast<__nv_bfloat16*>(c_bf16), ldcb, bias, (flags & UIC_GEMM_RELU) ? 1 : 0,
                  (flags & UIC_GEMM_ACCUMULATE) ? 1 : 0, (flags & UIC_GEMM_OUT_F16) ? 1 : 0, exp_col0, exp_scale};
  if (gemm_impl() == GEMM_IMPL_SIMT) {
    dim3 grid((N + 15) / 16, (M + 15) / 16), block(16, 16);
    launch_begin("gemm_bf16_simt", stream);
    gemm_bf16_simt_kernel<<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(A), lda, a_mn,
                                                      static_cast<const __nv_bfloat16*>(B), ldb, b_mn, ep, M, N, K);
    UIC_CUDA_OK(cudaGetLastError());
    launch_end(stream);
    return 0;
  }
  // TMA needs 16-byte aligned bases and row pitches.
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda % 8) || (ldb % 8))
    return set_error(UIC_ERR_ALIGN, "gemm_bf16: operands must be 16-byte aligned with pitches that are multiples of 8 "
                     "elements (lda=%lld ldb=%lld)", lda, ldb);
  // Narrow tiles when the wide grid would leave most SMs idle (per-step GEMMs with few rows).
  const long long tiles128 = static_cast<long long>((M + BM - 1) / BM) * ((N + 127) / 128);
  const bool narrow = tiles128 < 120 || N <= 64;
  const int bn = narrow ? 64 : 128;
  CUtensorMap ta, tb;
  int rc;
  // operand stored [rows = M or N, cols = K] (K-major) or [rows = K, cols = M or N] (MN-major)
  rc = a_mn ? get_tensor_map_bf16(&ta, A, K, M, lda, 64, 64) : get_tensor_map_bf16(&ta, A, M, K, lda, BM, 64);
  if (rc) return rc;
  rc = b_mn ? get_tensor_map_bf16(&tb, B, K, N, ldb, 64, 64) : get_tensor_map_bf16(&tb, B, N, K, ldb, bn, 64);
  if (rc) return rc;
  if (bn == 64) return dispatch_major<64, 4>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
  return dispatch_major<128, 3>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
}

}  // namespace uic
