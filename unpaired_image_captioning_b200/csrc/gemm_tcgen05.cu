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
// Structure (second iteration; the first one's pipeline trace showed 3 stages leaving the k-loop
// TMA-latency bound at ~560 cycles per k-block and a row-per-thread epilogue taking 3x the k-loop):
//   * persistent: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ...
//     (m fastest, so concurrently running CTAs share the weight tile in L2);
//   * tile 128 x BN (BN in {64,128,256}), BLOCK_K = 64 (one 128-byte swizzle row), 3-7 stage TMA ring;
//   * warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, tcgen05.mma cta_group::1, M = 128),
//     warp 2 = TMEM allocator, warps 4-7 = epilogue;
//   * two TMEM accumulator buffers: the epilogue of tile i overlaps the k-loop of tile i+1;
//   * epilogue: tcgen05.ld (row per lane) -> per-warp shared-memory transpose -> 16-byte stores that
//     cover 4 rows x 128 contiguous bytes per warp instruction; the bias is loaded as one float4 per lane.
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <type_traits>

#include <cuda_fp16.h>

#include "uic_internal.h"
#include "uic_ptx.cuh"
#include "uic_vocab.cuh"

namespace uic {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr float EXP_CAP = 1152921504606846976.0f;  // 2^60: cap of the exponential-form attention operands (attention.cu)
constexpr int EPI_PITCH = 36;  // floats per staged row: 32 + 4 keeps every 16-byte access conflict-free

struct GemmEpilogue {
  float* c_f32;
  long long ldc;
  __nv_bfloat16* c_bf16;
  long long ldcb;
  const float* bias;
  int relu;
  int accumulate;
  int out_f16;   // the 16-bit output is IEEE fp16 instead of bf16
  int exp_col0;  // columns >= exp_col0 become exp_scale * exp(2 x) when exp_scale != 0
  float exp_scale;
  long long* trace;  // optional device buffer: CTA 0 records clock64() at pipeline events of its first tile
  int debug;         // UIC_GEMM_DEBUG bits (experiments only): 1 = no global stores, 2 = no TMEM load, 4 = no smem transpose
  // fused vocabulary statistics (STATS > 0 kernels): per (row, column part) the running (max, sum exp) of
  // x = acc + bias and the STATS best keys (UNK column shifted by -1000, banned token -inf) with their columns
  float* stats;
  const long long* banned;
  long long banned_stride;
  int parts;
  int unk_col;  // column that gets -1000 (beam search only), -1 = none
  int sample;   // multinomial sampling: candidate keys become x / T + Gumbel noise (uic_vocab.cuh)
  float inv_temperature;
  const unsigned long long* seed;  // device memory: a captured CUDA graph can be replayed with a new seed
  int step;
  // Persistent tile walk: consecutive tiles share the B tile (m fastest, default) or the A tile (n fastest).  When A is
  // the big operand (feature prologue: 50176 x 2048 against 512 x 2048) the m-fastest walk re-reads every A tile from HBM
  // once per column of tiles; n-fastest reads it once and finds it in L2 for the others.
  int n_fastest;
  // L2 eviction hints of the operand loads (uic_ptx.cuh): weights re-read by every decode step are kept (evict-last),
  // operands streamed once (raw features of the prologue) are marked evict-first
  unsigned long long policy_a, policy_b;
  // optional per-column affine AFTER the activation: y = act(x) * post_scale[n] + post_shift[n] (an eval-mode BatchNorm1d
  // behind Linear + ReLU, use_bn = 2 of models/AttModel.py:79-84); takes the element-wise store path
  const float* post_scale;
  const float* post_shift;
};

#define UIC_TRACE(slot)                                                              \
  do {                                                                               \
    if (ep.trace != nullptr && blockIdx.x == 0 && it == 0) ep.trace[slot] = clock64(); \
  } while (0)

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
__device__ __forceinline__ float epi_act(float v, int col, const GemmEpilogue& ep) {
  if (ep.relu) v = fmaxf(v, 0.0f);
  if (ep.post_scale != nullptr) v = fmaf(v, __ldg(ep.post_scale + col), __ldg(ep.post_shift + col));
  if (ep.exp_scale != 0.0f && col >= ep.exp_col0) v = fminf(ep.exp_scale * __expf(2.0f * v), ep.out_f16 ? 65504.0f : EXP_CAP);
  return v;
}

// Branch-free insertion of (x, c) into a list sorted by value (descending).  Strict ">" keeps the earlier entry on
// ties, so feeding columns in increasing order reproduces a stable descending sort.  -inf keys are never inserted.
template <int K>
__device__ __forceinline__ void topk_insert(float (&v)[K], int (&i)[K], float x, int c) {
  bool p[K];
#pragma unroll
  for (int q = 0; q < K; ++q) p[q] = x > v[q];
#pragma unroll
  for (int q = K - 1; q > 0; --q) {
    v[q] = p[q - 1] ? v[q - 1] : (p[q] ? x : v[q]);
    i[q] = p[q - 1] ? i[q - 1] : (p[q] ? c : i[q]);
  }
  v[0] = p[0] ? x : v[0];
  i[0] = p[0] ? c : i[0];
}
// Same, for candidates that arrive out of column order: equal values rank by the smaller column.
template <int K>
__device__ __forceinline__ void topk_insert_tie(float (&v)[K], int (&i)[K], float x, int c) {
  bool p[K];
#pragma unroll
  for (int q = 0; q < K; ++q) p[q] = x > v[q] || (x == v[q] && c < i[q]);
#pragma unroll
  for (int q = K - 1; q > 0; --q) {
    v[q] = p[q - 1] ? v[q - 1] : (p[q] ? x : v[q]);
    i[q] = p[q - 1] ? i[q - 1] : (p[q] ? c : i[q]);
  }
  v[0] = p[0] ? x : v[0];
  i[0] = p[0] ? c : i[0];
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int BN, int STAGES, int STATS = 0>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
  // storing kernels: one 32 x 32 fp32 staging tile per epilogue warp; statistics kernels: the double-buffered bias copy
  static constexpr int EPI_BYTES = STATS > 0 ? 2 * BN * 4 : EPI_WARPS * 32 * EPI_PITCH * 4;
  static constexpr int BAR_OFFSET = (EPI_OFFSET + EPI_BYTES + 15) / 16 * 16;
  static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // barriers + alignment slack
  // TMEM columns per accumulator buffer (allocations are powers of two; 224-wide tiles get 256-column buffers)
  static constexpr int ACC_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
};

template <int BN, int STAGES, bool A_MN, bool B_MN, int STATS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         GemmEpilogue ep, int M, int N, int K) {
  using L = GemmSmem<BN, STAGES, STATS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (ep.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 448) {  // (1024-slot debug buffer) CTA entry time
    long long tnow;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tnow));
    ep.trace[128 + 2 * blockIdx.x] = tnow;
    if (blockIdx.x == 0) ep.trace[120] = clock64();  // CTA 0 entry on the cycle counter the pipeline events use
  }
  const int num_kb = (K + BK - 1) / BK;
  const int tiles_m = (M + BM - 1) / BM;
  const int tiles_n = (N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], EPI_WARPS);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * L::ACC_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (ep.trace != nullptr && threadIdx.x == 0 && blockIdx.x == 0) ep.trace[121] = clock64();  // set-up done (barriers, TMEM)
  // The dependency wait sits in front of each role's FIRST global-memory access instead of here: the roles' loop prologues
  // (tile coordinates, descriptor constants, and above all the first fetch of their code -- the CTA timeline showed ~830
  // cycles between a common wait and the first TMA issue, cold instruction cache of a 13 k-instruction kernel) then overlap
  // the predecessor's tail.  The MMA issuer touches shared memory and TMEM only and does not wait at all.

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int g = 0;  // global k-block counter: ring position carries over from tile to tile
      for (int tile = blockIdx.x, it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
        const int m0 = (ep.n_fastest ? tile / tiles_n : tile % tiles_m) * BM, n0 = (ep.n_fastest ? tile % tiles_n : tile / tiles_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (g == 0) {
            pdl_wait();  // operands written by the predecessor are read from here on
            if (ep.trace != nullptr && blockIdx.x == 0) ep.trace[122] = clock64();  // dependency wait passed
          }
          if (kb < 32) UIC_TRACE(50 + kb);
          mbar_arrive_expect_tx(&full_bar[s], L::STAGE_BYTES);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          uint8_t* sb = sa + L::A_BYTES;
          if (!A_MN) {
            tma_load_2d_hint(sa, &tmap_a, &full_bar[s], kb * BK, m0, ep.policy_a);
          } else {  // A stored [K, M]: two boxes of 64 m-columns x 64 k-rows
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d_hint(sa + i * (BK * 128), &tmap_a, &full_bar[s], m0 + 64 * i, kb * BK, ep.policy_a);
          }
          if (!B_MN) {
            tma_load_2d_hint(sb, &tmap_b, &full_bar[s], kb * BK, n0, ep.policy_b);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d_hint(sb + i * (BK * 128), &tmap_b, &full_bar[s], n0 + 64 * i, kb * BK, ep.policy_b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // Shared-memory descriptors: everything but the start address is a compile-time constant, and a 16-element K step
      // only adds to the address field.  The four descriptor pairs of a k-block are ready before its first MMA, so the
      // four tcgen05.mma issue back to back (a first version rebuilt both 64-bit descriptors between two MMAs: 565 cycles
      // per 128 x 128 x 64 block; now 520 = 4 x 130, the rate of the 1-CTA SS instruction at N <= 128 on this part --
      // wider tiles amortise it: 128 x 256 x 64 takes 750).
      //   K-major : 8-row groups 1024 B apart (SBO), LBO 16 B; a 16-element K step is 32 B inside the swizzle row.
      //   MN-major: 64-element MN chunks BK*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO); a K step is 2048 B.
      constexpr uint64_t DESC_A = (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61) |
                                  (static_cast<uint64_t>(((A_MN ? BK * 128 : 16) >> 4) & 0x3FFF) << 16) |
                                  (static_cast<uint64_t>((1024 >> 4) & 0x3FFF) << 32);
      constexpr uint64_t DESC_B = (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61) |
                                  (static_cast<uint64_t>(((B_MN ? BK * 128 : 16) >> 4) & 0x3FFF) << 16) |
                                  (static_cast<uint64_t>((1024 >> 4) & 0x3FFF) << 32);
      constexpr uint32_t KSTEP_A = (A_MN ? 2048 : 32) >> 4, KSTEP_B = (B_MN ? 2048 : 32) >> 4;  // in 16-byte units
      const uint32_t ring0 = smem_u32(smem) >> 4;
      uint32_t s = 0, ph = 0;  // ring position: carries over from tile to tile
      for (int tile = blockIdx.x, it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * L::ACC_COLS;
        UIC_TRACE(0);
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t a_lo = ring0 + s * (L::STAGE_BYTES >> 4), b_lo = a_lo + (L::A_BYTES >> 4);
          uint64_t da[BK / 16], db[BK / 16];
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            da[k] = DESC_A | static_cast<uint64_t>(a_lo + k * KSTEP_A);
            db[k] = DESC_B | static_cast<uint64_t>(b_lo + k * KSTEP_B);
          }
          mbar_wait(&full_bar[s], ph);
          if (kb < 32) UIC_TRACE(1 + kb);
          tcgen05_fence_after();
          umma_bf16_ss(tmem_d, da[0], db[0], idesc, kb != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < BK / 16; ++k) umma_bf16_acc(tmem_d, da[k], db[k], idesc);
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        UIC_TRACE(40);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> smem transpose -> coalesced global stores =====
    const int ew = warp & 3;            // the TMEM lane quarter this warp may access (hardware rule: warp % 4)
    const int ehalf = (warp - 4) >> 2;  // warps 4-7 take the first half of the tile's column chunks, warps 8-11 the second
    constexpr int NCH = BN / 32;          // 32-column chunks of a tile (7 for the 224-wide statistics tiles)
    constexpr int CHUNKS = (NCH + 1) / 2; // chunks of the first half; the second half has NCH - CHUNKS
    const uint32_t stage_s = smem_u32(smem + L::EPI_OFFSET) + (warp - 4) * 32 * EPI_PITCH * 4;  // this warp's 32 x 32 fp32 staging tile
    const int q = lane & 7, rsub = lane >> 3;  // this lane stores columns 4q..4q+3 of rows rsub, rsub+4, ...
    const bool vec_f32 = ep.c_f32 != nullptr && (ep.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.c_f32) & 15) == 0);
    const bool vec_16 = ep.c_bf16 != nullptr && (ep.ldcb % 4 == 0) && ((reinterpret_cast<uintptr_t>(ep.c_bf16) & 7) == 0);
    const bool vec_bias = ep.bias != nullptr && ((reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0);
    pdl_wait();  // bias / banned-token loads, accumulate reads and all output stores come after this
    for (int tile = blockIdx.x, it = 0; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (ep.n_fastest ? tile / tiles_n : tile % tiles_m) * BM, n0 = (ep.n_fastest ? tile % tiles_n : tile / tiles_m) * BN;
      const int acc = it & 1;
      if constexpr (STATS > 0) {
        // ---- fused vocabulary statistics: the (rows, V) logits are never written to memory ----------------
        constexpr int EMPTY = 0x7fffffff;
        float* s_bias = reinterpret_cast<float*>(smem + L::EPI_OFFSET) + acc * BN;  // double-buffered by accumulator
        const int et = threadIdx.x - 128;
        if (et < BN) s_bias[et] = (ep.bias != nullptr && n0 + et < N) ? __ldg(ep.bias + n0 + et) : 0.0f;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
        mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
        tcgen05_fence_after();
        const int row = m0 + ew * 32 + lane;
        int banned = -1;
        if (ep.banned != nullptr && row < M) banned = static_cast<int>(ep.banned[static_cast<long long>(row) * ep.banned_stride]);
        constexpr float LOG2E = 1.4426950408889634f;
        constexpr int ES = (2 + 2 * STATS + 3) / 4 * 4;  // floats per (row, part) entry, see logit_stats_entry_floats
        float run_m = -INFINITY, run_s = 0.0f;
        // Two candidate lists (even / odd columns) so that consecutive insertions are independent instruction chains;
        // ids are column offsets inside this warp's half tile (compile-time constants in the unrolled loops).
        // (From kslots = 3 on the kernel is bound by these insertions -- scripts/stats_times.py: configs[4] 124 us with
        // kslots 1, 223 us with kslots 5.  FOUR lists were measured too: slower, 245 us -- the two extra list merges cost
        // more than the added instruction-level parallelism returns, i.e. the epilogue is issue-bound, not latency-bound.)
        float va[STATS], vb[STATS];
        int ia[STATS], ib[STATS];
#pragma unroll
        for (int qq = 0; qq < STATS; ++qq) {
          va[qq] = vb[qq] = -INFINITY;
          ia[qq] = ib[qq] = EMPTY;
        }
#pragma unroll
        for (int cc = 0; cc < CHUNKS; ++cc) {
          const int c = ehalf * CHUNKS + cc;
          if (c >= NCH) continue;  // (odd chunk count: the second half is one chunk shorter; warp-uniform)
          const int col0 = n0 + c * 32;
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32(tmem_base + acc * L::ACC_COLS + (static_cast<uint32_t>(ew * 32) << 16) + c * 32, v);
          tmem_ld_wait();
          if (col0 >= N) continue;  // warp-uniform
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + s_bias[c * 32 + j];
          if (col0 + 32 > N) {  // ragged last tile (warp-uniform)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j >= N) x[j] = -INFINITY;
          }
          float bm = x[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) bm = fmaxf(bm, x[j]);
          {
            const float off = -bm * LOG2E;
            float bs0 = 0.0f, bs1 = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              bs0 += ex2_approx(fmaf(x[j], LOG2E, off));
              bs1 += ex2_approx(fmaf(x[j + 1], LOG2E, off));
            }
            const float nm = fmaxf(run_m, bm);
            run_s = run_s * ex2_approx((run_m - nm) * LOG2E) + (bs0 + bs1) * ex2_approx((bm - nm) * LOG2E);
            run_m = nm;
          }
          if (ep.sample) {  // multinomial sampling by Gumbel-max (AttModel.py:231-239); -inf stays -inf
            const uint32_t rk = rng_row_key(rng_step_key(*ep.seed, ep.step), row);
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaf(x[j], ep.inv_temperature, rng_gumbel(rk, col0 + j));
          }
          // beam-search edits of the candidate keys; the statistics above use the unedited logits
          if (ep.unk_col >= col0 && ep.unk_col < col0 + 32) {  // UNK suppression (CaptionModel.py:133), warp-uniform
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j == ep.unk_col) x[j] -= 1000.0f;
          }
          const int bj = banned - col0;
          if (__any_sync(0xffffffffu, static_cast<unsigned>(bj) < 32u)) {  // decoding constraint (:130-131, AttModel.py:220-223)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j == bj) x[j] = -INFINITY;
          }
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            topk_insert<STATS>(va, ia, x[j], cc * 32 + j);
            topk_insert<STATS>(vb, ib, x[j + 1], cc * 32 + j + 1);
          }
        }
#pragma unroll
        for (int qq = 0; qq < STATS; ++qq) topk_insert_tie<STATS>(va, ia, vb[qq], ib[qq]);
        if (row < M) {
          float out[ES];
          out[0] = run_m;
          out[1] = run_s;
          const int cbase = n0 + ehalf * CHUNKS * 32;
#pragma unroll
          for (int qq = 0; qq < STATS; ++qq) {
            out[2 + qq] = va[qq];
            out[2 + STATS + qq] = __int_as_float(ia[qq] == EMPTY ? EMPTY : cbase + ia[qq]);
          }
#pragma unroll
          for (int qq = 2 + 2 * STATS; qq < ES; ++qq) out[qq] = 0.0f;
          float4* dst = reinterpret_cast<float4*>(ep.stats + (static_cast<long long>(row) * ep.parts + (n0 / BN) * 2 + ehalf) * ES);
#pragma unroll
          for (int qq = 0; qq < ES / 4; ++qq) dst[qq] = make_float4(out[4 * qq], out[4 * qq + 1], out[4 * qq + 2], out[4 * qq + 3]);
        }
      } else {
      // bias of this lane's four columns in each of the warp's chunks: loaded while the k-loop still runs
      float4 bias4[CHUNKS];
#pragma unroll
      for (int cc = 0; cc < CHUNKS; ++cc) {
        bias4[cc] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        const int cq = n0 + (ehalf * CHUNKS + cc) * 32 + 4 * q;
        if (ep.bias != nullptr) {
          if (vec_bias && cq + 4 <= N) {
            bias4[cc] = __ldg(reinterpret_cast<const float4*>(ep.bias + cq));
          } else {
            if (cq < N) bias4[cc].x = __ldg(ep.bias + cq);
            if (cq + 1 < N) bias4[cc].y = __ldg(ep.bias + cq + 1);
            if (cq + 2 < N) bias4[cc].z = __ldg(ep.bias + cq + 2);
            if (cq + 3 < N) bias4[cc].w = __ldg(ep.bias + cq + 3);
          }
        }
      }
      mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
      if (threadIdx.x == 128) UIC_TRACE(41);
      tcgen05_fence_after();
      const int row_base = m0 + ew * 32;
#pragma unroll
      for (int cc = 0; cc < CHUNKS; ++cc) {
        const int c = ehalf * CHUNKS + cc;
        const int col0 = n0 + c * 32;
        uint32_t v[32];
        __syncwarp();
        if (!(ep.debug & 2)) tmem_ld_32x32(tmem_base + acc * L::ACC_COLS + (static_cast<uint32_t>(ew * 32) << 16) + c * 32, v);
        const float4 b4 = bias4[cc];
        const int cq = col0 + 4 * q;
        tmem_ld_wait();
        if (threadIdx.x == 128 && cc < 2) UIC_TRACE(43 + 3 * cc);
        if (col0 >= N || row_base >= M) continue;  // warp-uniform
        if (ep.debug & 4) continue;
        // lane = row: stage the 32 accumulators of this row (explicit shared-space stores: the staging
        // pointer is derived through an integer alignment cast, so generic ST/LD would be emitted otherwise)
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(stage_s + (lane * EPI_PITCH + j) * 4), "r"(v[j]), "r"(v[j + 1]),
                       "r"(v[j + 2]), "r"(v[j + 3])
                       : "memory");
        __syncwarp();
        if (threadIdx.x == 128 && cc < 2) UIC_TRACE(44 + 3 * cc);
        // The activation mode is uniform per 32-column chunk: decide it ONCE and run a specialised store
        // loop (the first version evaluated the epilogue flags per element: ~40 instructions per value,
        // which made the epilogue 3x longer than the k-loop).
        const bool chunk_vec = (col0 + 32 <= N);
        const bool exp_all = ep.exp_scale != 0.0f && col0 >= ep.exp_col0;
        const bool exp_mixed = ep.exp_scale != 0.0f && !exp_all && col0 + 32 > ep.exp_col0;
        // Vector path: the eight 16-byte row pieces of this lane are loaded from the staging tile first, then stored through
        // ONE base pointer per output advanced by a constant pitch; output set, activation and accumulate are template
        // constants.  (The first version evaluated the epilogue flags per element; the second re-derived the 64-bit row
        // address, re-read the epilogue struct from the constant bank and branched on the output pointers in every one of
        // the eight row iterations -- ~70 dependent instructions each: the pipeline trace showed 3 600 cycles per 32 x 32
        // chunk, i.e. an epilogue longer than the k-loop of the 768-row GEMMs whatever the output width.)
        const int rows_valid = M - row_base;  // > 0 (checked above)
        const bool has32 = ep.c_f32 != nullptr, has16 = ep.c_bf16 != nullptr, accum = has32 && ep.accumulate != 0;
        auto store_fast = [&](auto mode_c) {
          constexpr int MODE = decltype(mode_c)::value;  // 0 plain, 1 relu, 2 exp on the whole chunk
          float* d32 = ep.c_f32 + static_cast<long long>(row_base + rsub) * ep.ldc + cq;             // (only dereferenced if has32)
          __nv_bfloat16* d16 = ep.c_bf16 + static_cast<long long>(row_base + rsub) * ep.ldcb + cq;  // (only if has16)
          const long long p32 = 4 * ep.ldc, p16 = 4 * ep.ldcb;
          const int f16 = ep.out_f16;
          const float cap = f16 ? 65504.0f : EXP_CAP, es = ep.exp_scale;
#pragma unroll
          for (int half = 0; half < 2; ++half) {  // two groups of four rows: 16 + 16 live registers instead of 64
            float4 f[4], o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r4 = half * 4 + i;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                           : "=f"(f[i].x), "=f"(f[i].y), "=f"(f[i].z), "=f"(f[i].w)
                           : "r"(stage_s + ((r4 * 4 + rsub) * EPI_PITCH + 4 * q) * 4));
              o[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
              if (accum && r4 * 4 + rsub < rows_valid) o[i] = *reinterpret_cast<const float4*>(d32 + r4 * p32);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r4 = half * 4 + i;
              float4 v4 = f[i];
              v4.x += b4.x + o[i].x; v4.y += b4.y + o[i].y; v4.z += b4.z + o[i].z; v4.w += b4.w + o[i].w;
              if constexpr (MODE == 1) { v4.x = fmaxf(v4.x, 0.0f); v4.y = fmaxf(v4.y, 0.0f); v4.z = fmaxf(v4.z, 0.0f); v4.w = fmaxf(v4.w, 0.0f); }
              if constexpr (MODE == 2) {
                v4.x = fminf(es * __expf(2.0f * v4.x), cap); v4.y = fminf(es * __expf(2.0f * v4.y), cap);
                v4.z = fminf(es * __expf(2.0f * v4.z), cap); v4.w = fminf(es * __expf(2.0f * v4.w), cap);
              }
              if (r4 * 4 + rsub < rows_valid) {
                if (has32) *reinterpret_cast<float4*>(d32 + r4 * p32) = v4;
                if (has16) *reinterpret_cast<uint2*>(d16 + r4 * p16) = make_uint2(pack16(v4.x, v4.y, f16), pack16(v4.z, v4.w, f16));
              }
            }
          }
        };
        auto store_rows = [&](auto mode_c) {
          constexpr int MODE = decltype(mode_c)::value;  // 0 plain, 1 relu, 2 exp on the whole chunk, 3 generic
          if constexpr (MODE != 3) {
            store_fast(mode_c);
          } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + rsub;
            const int row = row_base + r;
            float4 f;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(stage_s + (r * EPI_PITCH + 4 * q) * 4));
            if (row >= M) continue;
            f.x += b4.x; f.y += b4.y; f.z += b4.z; f.w += b4.w;
            {  // tails, unaligned outputs, chunk straddling exp_col0: element-wise with all the flags
              float e4[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (cq + k >= N) continue;
                if (ep.c_f32 != nullptr) {
                  float* dst = ep.c_f32 + static_cast<long long>(row) * ep.ldc + cq + k;
                  if (ep.accumulate) e4[k] += *dst;
                  e4[k] = epi_act(e4[k], cq + k, ep);
                  *dst = e4[k];
                } else {
                  e4[k] = epi_act(e4[k], cq + k, ep);
                }
                if (ep.c_bf16 != nullptr) ep.c_bf16[static_cast<long long>(row) * ep.ldcb + cq + k] = store16(e4[k], ep.out_f16);
              }
            }
          }
          }
        };
        const bool fast = chunk_vec && !exp_mixed && (ep.c_f32 == nullptr || vec_f32) && (ep.c_bf16 == nullptr || vec_16) && !(ep.debug & 1) &&
                          ep.post_scale == nullptr;
        if (ep.debug & 1) {
        } else if (!fast) {
          store_rows(std::integral_constant<int, 3>{});
        } else if (exp_all) {
          store_rows(std::integral_constant<int, 2>{});
        } else if (ep.relu) {
          store_rows(std::integral_constant<int, 1>{});
        } else {
          store_rows(std::integral_constant<int, 0>{});
        }
        if (threadIdx.x == 128 && cc < 2) UIC_TRACE(45 + 3 * cc);
      }
      }  // STATS == 0
      // all four epilogue warps have read this accumulator: hand it back to the MMA issuer
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (threadIdx.x == 128) UIC_TRACE(42);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * L::ACC_COLS);
  if (ep.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 448) {  // CTA exit time
    long long tnow;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tnow));
    ep.trace[129 + 2 * blockIdx.x] = tnow;
  }
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
    acc = epi_act(acc, col, ep);
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

static int gemm_debug_flags() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UIC_GEMM_DEBUG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// Operand L2 hints.  B is a weight matrix in every forward contraction of the decoder (K-major): up to 32 MB it is kept
// evict-last, the decode loop re-reads it every step while 100 MB of feature tiles stream through L2 in between.
// UIC_GEMM_A_STREAM / UIC_GEMM_B_STREAM mark an operand that is read once (the raw feature matrix of the prologue).
static void gemm_l2_policies(GemmEpilogue& ep, int M, int N, int K, int flags) {
  static int mode = -1;  // UIC_GEMM_L2=0 switches the hints off (A/B experiments)
  if (mode < 0) {
    const char* e = getenv("UIC_GEMM_L2");
    mode = e ? atoi(e) : 1;
  }
  ep.policy_a = ep.policy_b = L2_EVICT_NORMAL;
  if (mode == 0) return;
  const bool b_weight = !(flags & UIC_GEMM_B_MN_MAJOR) && !(flags & UIC_GEMM_A_MN_MAJOR);
  if (flags & UIC_GEMM_A_STREAM) ep.policy_a = L2_EVICT_FIRST;
  if (flags & UIC_GEMM_B_STREAM)
    ep.policy_b = L2_EVICT_FIRST;
  else if (b_weight && static_cast<long long>(N) * K * 2 <= (32LL << 20))
    ep.policy_b = L2_EVICT_LAST;
  (void)M;
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int STATS = 0>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                     cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05_kernel<BN, STAGES, A_MN, B_MN, STATS>;
  constexpr int smem = GemmSmem<BN, STAGES, STATS>::TOTAL;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool configured = false;  // benign race: attribute set is idempotent
  if (!configured) {
    UIC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const long long tiles = static_cast<long long>((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  launch_begin(gemm_label(M, N, K), stream);
  UIC_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), smem, stream, ta, tb, ep, M, N, K));
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
              float exp_scale, const float* post_scale, const float* post_shift) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(UIC_ERR_SHAPE, "gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  if (c_f32 == nullptr && c_bf16 == nullptr) return set_error(UIC_ERR_ARG, "gemm_bf16: no output buffer");
  const bool a_mn = flags & UIC_GEMM_A_MN_MAJOR;
  const bool b_mn = flags & UIC_GEMM_B_MN_MAJOR;
  GemmEpilogue ep{c_f32, ldc, static_cast<__nv_bfloat16*>(c_bf16), ldcb, bias, (flags & UIC_GEMM_RELU) ? 1 : 0,
                  (flags & UIC_GEMM_ACCUMULATE) ? 1 : 0, (flags & UIC_GEMM_OUT_F16) ? 1 : 0, exp_col0, exp_scale, gemm_trace_buffer(), gemm_debug_flags(), nullptr, nullptr, 0, 0, -1};
  // tile walk: keep the bigger operand's tile hot (see GemmEpilogue::n_fastest); the smaller one must fit L2 comfortably
  ep.n_fastest = (M > N && static_cast<long long>(N) * K * 2 <= (32LL << 20)) ? 1 : 0;
  gemm_l2_policies(ep, M, N, K, flags);
  if ((post_scale == nullptr) != (post_shift == nullptr)) return set_error(UIC_ERR_ARG, "gemm_bf16: post_scale and post_shift come together");
  ep.post_scale = post_scale;
  ep.post_shift = post_shift;
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
  // Tile width: the k-loop is bound by what one SM can pull in per k-block (A tile + B tile, ~57 B/clk measured:
  // 430 / 571 / 750 cycles for 128 x {64, 128, 256}), so wide tiles do more math per byte but leave fewer tiles to
  // spread over the SMs.  Pick the width with the smallest estimated makespan rounds(tiles / SMs) * cycles per k-block.
  const long long tiles_m = (M + BM - 1) / BM;
  const int sms = sm_count();
  int bn = 64;
  long long best = -1;
  const int widths[3] = {64, 128, 256}, cycles[3] = {520, 521, 750};
  for (int i = 0; i < 3; ++i) {
    if (widths[i] > 64 && N <= widths[i] / 2) continue;  // mostly padding
    const long long tiles = tiles_m * ((N + widths[i] - 1) / widths[i]);
    const long long est = ((tiles + sms - 1) / sms) * cycles[i];
    if (best < 0 || est < best) {
      best = est;
      bn = widths[i];
    }
  }
  CUtensorMap ta, tb;
  int rc;
  // operand stored [rows = M or N, cols = K] (K-major) or [rows = K, cols = M or N] (MN-major)
  rc = a_mn ? get_tensor_map_bf16(&ta, A, K, M, lda, 64, 64) : get_tensor_map_bf16(&ta, A, M, K, lda, BM, 64);
  if (rc) return rc;
  rc = b_mn ? get_tensor_map_bf16(&tb, B, K, N, ldb, 64, 64) : get_tensor_map_bf16(&tb, B, N, K, ldb, bn, 64);
  if (rc) return rc;
  if (gemm_debug_flags() & 8) {  // experiment: shallow ring, most of the 228 KB stays L1
    if (bn == 64) return dispatch_major<64, 3>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
    return dispatch_major<128, 2>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
  }
  if (bn == 64) return dispatch_major<64, 7>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
  if (bn == 256) return dispatch_major<256, 3>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
  return dispatch_major<128, 5>(a_mn, b_mn, ta, tb, ep, M, N, K, stream);
}

// Tile width of the statistics GEMM: 224 columns where that shortens the makespan (rounds of tiles over the SMs x measured
// cycles per k-block: 520 for 128 x 128 x 64, ~700 for 128 x 224 x 64) -- 768 x 10000 is 270 tiles = 2 rounds instead of
// 474 = 4, and the step tail merges 90 parts per row instead of 158.  A function of the problem shape only: the stats
// layout (two parts per tile) follows it.
static int stats_tile_width(int M, int N) {
  static int forced = -1;  // UIC_STATS_BN=128|224 pins it (A/B experiments)
  if (forced < 0) {
    const char* e = getenv("UIC_STATS_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 128 || forced == 224) return forced;
  const long long tiles_m = (M + BM - 1) / BM, sms = sm_count();
  const long long r128 = (tiles_m * ((N + 127) / 128) + sms - 1) / sms * 520;
  const long long r224 = (tiles_m * ((N + 223) / 224) + sms - 1) / sms * 700;
  return r224 < r128 ? 224 : 128;
}
int logit_stats_parts(int M, int N) {
  const int bn = stats_tile_width(M, N);
  return 2 * ((N + bn - 1) / bn);
}
int logit_stats_entry_floats(int kslots) { return (2 + 2 * kslots + 3) / 4 * 4; }

// Logit projection with the fused statistics epilogue: stats[row][part][logit_stats_entry_floats(kslots)].
int logit_stats(const void* A, long long lda, const void* B, long long ldb, const float* bias, const long long* banned,
                long long banned_stride, float* stats, int M, int N, int K, int kslots, int unk_suppress, float temperature,
                const unsigned long long* seed, int step, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(UIC_ERR_SHAPE, "logit_stats: empty problem M=%d N=%d K=%d", M, N, K);
  if (kslots != 1 && kslots != 3 && kslots != 5 && kslots != 8)
    return set_error(UIC_ERR_ARG, "logit_stats: kslots must be 1, 3, 5 or 8 (got %d)", kslots);
  if (temperature > 0.0f && seed == nullptr) return set_error(UIC_ERR_ARG, "logit_stats: sampling needs a seed (device pointer)");
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda % 8) || (ldb % 8) ||
      (reinterpret_cast<uintptr_t>(stats) & 15))
    return set_error(UIC_ERR_ALIGN, "logit_stats: operands and stats must be 16-byte aligned with pitches that are multiples of 8 elements");
  GemmEpilogue ep{nullptr, 0, nullptr, 0, bias, 0, 0, 0, 0, 0.0f, nullptr, 0, stats, banned, banned_stride, logit_stats_parts(M, N), unk_suppress ? N - 1 : -1,
                  temperature > 0.0f ? 1 : 0, temperature > 0.0f ? 1.0f / temperature : 1.0f, seed, step};
  gemm_l2_policies(ep, M, N, K, 0);
  CUtensorMap ta, tb;
  int rc = get_tensor_map_bf16(&ta, A, M, K, lda, BM, 64);
  if (rc) return rc;
  const int bn = stats_tile_width(M, N);
  rc = get_tensor_map_bf16(&tb, B, N, K, ldb, bn, 64);
  if (rc) return rc;
  ep.debug = gemm_debug_flags();
  ep.trace = gemm_trace_buffer();
  if (bn == 224) {
    if (kslots == 1) return launch_tc<224, 4, false, false, 1>(ta, tb, ep, M, N, K, stream);
    if (kslots == 3) return launch_tc<224, 4, false, false, 3>(ta, tb, ep, M, N, K, stream);
    if (kslots == 5) return launch_tc<224, 4, false, false, 5>(ta, tb, ep, M, N, K, stream);
    return launch_tc<224, 4, false, false, 8>(ta, tb, ep, M, N, K, stream);
  }
  if (kslots == 1) return launch_tc<128, 5, false, false, 1>(ta, tb, ep, M, N, K, stream);
  if (kslots == 3) return launch_tc<128, 5, false, false, 3>(ta, tb, ep, M, N, K, stream);
  if (kslots == 5) return launch_tc<128, 5, false, false, 5>(ta, tb, ep, M, N, K, stream);
  return launch_tc<128, 5, false, false, 8>(ta, tb, ep, M, N, K, stream);
}

}  // namespace uic
