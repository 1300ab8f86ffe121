// Device helpers shared by the vocabulary kernels (vocab.cu) and the fused beam / greedy advance kernels
// (beam.cu): running (max, sum-exp) pairs, ordered candidates and the merge of the statistics written by the
// logit GEMM's STATS epilogue (gemm_tcgen05.cu).
#pragma once
#include "uic_internal.h"

namespace uic {

// ---- counter-based noise for multinomial sampling ------------------------------------------------------------------
// torch.multinomial(exp(logprobs / T), 1) (models/AttModel.py:231-239) draws token v with probability
// softmax(x / T)[v]; so does argmax_v (x_v / T + g_v) with independent standard Gumbel noise g (Gumbel-max).
// g is a pure function of (seed, step, row, column): the GEMM epilogue perturbs the keys with it, the merge
// recomputes it for the winner to recover the unperturbed logit, and the oracle restates it in numpy.
__host__ __device__ __forceinline__ uint32_t rng_mix(uint32_t h) {  // "lowbias32" finaliser
  h ^= h >> 16;
  h *= 0x7feb352dU;
  h ^= h >> 15;
  h *= 0x846ca68bU;
  h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t rng_step_key(unsigned long long seed, int step) {
  return static_cast<uint32_t>(seed ^ (seed >> 32)) ^ (static_cast<uint32_t>(step) * 0x9E3779B1U);
}
__host__ __device__ __forceinline__ uint32_t rng_row_key(uint32_t step_key, int row) {
  return rng_mix(step_key + static_cast<uint32_t>(row) * 0x85EBCA77U);
}
__host__ __device__ __forceinline__ float rng_uniform(uint32_t row_key, int col) {
  const uint32_t h = rng_mix(row_key ^ (static_cast<uint32_t>(col) * 0xC2B2AE3DU));
  return (static_cast<float>(h >> 9) + 0.5f) * (1.0f / 8388608.0f);  // 23 bits + 1/2: exact in fp32, strictly inside (0, 1)
}
// g = -log(-log(u)).  Close to u = 1 the inner value -log(u) ~ 1 - u is as small as 6e-8, below the ABSOLUTE error of
// __logf (2^-21.4): it could come out 0 or negative and turn g into +inf / NaN for ~5e-7 of the draws.  There the series
// -log(1 - t) = t + t^2/2 + t^3/3 + t^4/4 (t = 1 - u is exact in fp32, relative error < 1e-7 for t < 1/32) replaces it.
__device__ __forceinline__ float rng_gumbel(uint32_t row_key, int col) {
  const float u = rng_uniform(row_key, col);
  const float t = 1.0f - u;
  const float series = t * fmaf(t, fmaf(t, fmaf(t, 0.25f, 1.0f / 3.0f), 0.5f), 1.0f);
  const float inner = t < 0.03125f ? series : -__logf(u);
  return -__logf(inner);
}
constexpr int RNG_ROW_DRAW_COL = 0x7fffffff;  // "column" of a per-row uniform draw (scheduled-sampling coin), outside any vocabulary

struct MaxSum {
  float m, s;
};
__device__ __forceinline__ MaxSum ms_merge(MaxSum a, MaxSum b) {
  MaxSum r;
  r.m = fmaxf(a.m, b.m);
  if (r.m == -INFINITY) {
    r.s = 0.0f;
    return r;
  }
  r.s = a.s * __expf(a.m - r.m) + b.s * __expf(b.m - r.m);
  return r;
}
struct Best {
  float v;
  int i;
};
__device__ __forceinline__ bool better(float v, int i, const Best& b) { return v > b.v || (v == b.v && i < b.i); }

// ---- merges of the fused logit statistics (gemm_tcgen05.cu, STATS epilogue) ----------------------------------
// stats[row][part][ES] with ES = logit_stats_entry_floats(KS): (max, sum exp, KS keys, KS columns, padding).
// One warp per row; lanes read consecutive parts with 16-byte loads.
constexpr int MERGE_ROWS_PER_CTA = 4;  // rows (warps) per CTA of the stand-alone merge kernels

struct RowStats {
  float M, log_s;
};

// Merges the (max, sum exp) pairs of a row and leaves each lane with the best KS candidates of its share of
// the parts (sorted by key descending, smaller column first on ties).
template <int KS>
__device__ __forceinline__ RowStats merge_row_stats(const float* __restrict__ st, int parts, float (&kv)[KS], int (&ki)[KS]) {
  constexpr int EMPTY = 0x7fffffff;
  constexpr int ES = (2 + 2 * KS + 3) / 4 * 4;
  const int lane = threadIdx.x & 31;
  MaxSum a{-INFINITY, 0.0f};
#pragma unroll
  for (int q = 0; q < KS; ++q) {
    kv[q] = -INFINITY;
    ki[q] = EMPTY;
  }
  // Lanes take parts lane, lane + 32, ...; the entries of four such rounds are requested back to back (the parts were
  // written by the GEMM epilogue a moment ago: L2 hits whose latency would otherwise be paid once per round).
  const float4* base = reinterpret_cast<const float4*>(st);
  constexpr int ROUNDS = 4;
  for (int p0 = lane; p0 < parts; p0 += 32 * ROUNDS) {
    float4 ent[ROUNDS][ES / 4];
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      const int p = p0 + 32 * rd;
      if (p < parts) {
#pragma unroll
        for (int w = 0; w < ES / 4; ++w) ent[rd][w] = base[static_cast<long long>(p) * (ES / 4) + w];
      }
    }
#pragma unroll
    for (int rd = 0; rd < ROUNDS; ++rd) {
      if (p0 + 32 * rd >= parts) break;
      float e[ES];
#pragma unroll
      for (int w = 0; w < ES / 4; ++w) {
        e[4 * w] = ent[rd][w].x;
        e[4 * w + 1] = ent[rd][w].y;
        e[4 * w + 2] = ent[rd][w].z;
        e[4 * w + 3] = ent[rd][w].w;
      }
      a = ms_merge(a, MaxSum{e[0], e[1]});
#pragma unroll
      for (int c = 0; c < KS; ++c) {
        const float x = e[2 + c];
        const int col = __float_as_int(e[2 + KS + c]);
        bool pr[KS];
#pragma unroll
        for (int q = 0; q < KS; ++q) pr[q] = col != EMPTY && (x > kv[q] || (x == kv[q] && col < ki[q]));
#pragma unroll
        for (int q = KS - 1; q > 0; --q) {
          kv[q] = pr[q - 1] ? kv[q - 1] : (pr[q] ? x : kv[q]);
          ki[q] = pr[q - 1] ? ki[q - 1] : (pr[q] ? col : ki[q]);
        }
        kv[0] = pr[0] ? x : kv[0];
        ki[0] = pr[0] ? col : ki[0];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum other;
    other.m = __shfl_xor_sync(0xffffffffu, a.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, a.s, o);
    a = ms_merge(a, other);
  }
  return RowStats{a.m, logf(a.s)};
}

// Pops the globally best remaining candidate of the warp (every lane holds a sorted list).
template <int KS>
__device__ __forceinline__ Best warp_pop_best(float (&kv)[KS], int (&ki)[KS]) {
  constexpr int EMPTY = 0x7fffffff;
  Best b{kv[0], ki[0]};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (oi != EMPTY && (b.i == EMPTY || better(ov, oi, b))) {
      b.v = ov;
      b.i = oi;
    }
  }
  if (ki[0] == b.i && b.i != EMPTY) {  // the owner shifts its list up
#pragma unroll
    for (int q = 0; q + 1 < KS; ++q) {
      kv[q] = kv[q + 1];
      ki[q] = ki[q + 1];
    }
    kv[KS - 1] = -INFINITY;
    ki[KS - 1] = EMPTY;
  }
  return b;
}

}  // namespace uic
