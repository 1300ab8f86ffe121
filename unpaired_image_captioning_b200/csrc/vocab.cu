// Row kernels over the vocabulary axis: log-softmax, fused masked cross-entropy, greedy argmax
// step and per-row top-k for beam search.  One CTA per decoder row.  The logits were just written
// by the logit GEMM, so the reads are L2 hits and the kernels are latency/issue bound: each thread
// pulls batches of 16 values (four 16-byte loads issued back to back), reduces a batch to its
// (max, sum-exp) with branch-free code, and only enters the arg-max / top-k update when the batch
// maximum can beat the thread's current threshold (rare after the first batches).
#include "uic_internal.h"
#include "uic_ptx.cuh"
#include "uic_vocab.cuh"

namespace uic {

constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;
constexpr int ROW_U = 4;               // 16-byte loads per batch
constexpr int ROW_BATCH = 4 * ROW_U;   // values per batch

__device__ __forceinline__ MaxSum ms_block_reduce(MaxSum v, MaxSum* s_part) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum other;
    other.m = __shfl_xor_sync(0xffffffffu, v.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, v.s, o);
    v = ms_merge(v, other);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  MaxSum r = s_part[0];
#pragma unroll
  for (int q = 1; q < ROW_WARPS; ++q) r = ms_merge(r, s_part[q]);
  return r;
}

// Calls f(x, idx) once per batch of this thread: x[16] values (-inf padding), idx[16] their column
// indices, in increasing index order.  16-byte loads when the row start is 16-byte aligned.
template <typename F>
__device__ __forceinline__ void scan_row_batches(const float* __restrict__ row, int V, F&& f) {
  const bool vec = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
  const int v4 = (V + 3) >> 2;  // float4 slots, the last one may be partial
  for (int i = threadIdx.x; i < v4; i += ROW_U * ROW_THREADS) {
    float x[ROW_BATCH];
    int idx[ROW_BATCH];
#pragma unroll
    for (int u = 0; u < ROW_U; ++u) {
      const int slot = i + u * ROW_THREADS;
      const int b = 4 * slot;
      float4 q = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      if (slot < v4) {
        if (vec && b + 4 <= V) {
          q = *reinterpret_cast<const float4*>(row + b);
        } else {
          if (b < V) q.x = row[b];
          if (b + 1 < V) q.y = row[b + 1];
          if (b + 2 < V) q.z = row[b + 2];
          if (b + 3 < V) q.w = row[b + 3];
        }
      }
      x[4 * u] = q.x; x[4 * u + 1] = q.y; x[4 * u + 2] = q.z; x[4 * u + 3] = q.w;
      idx[4 * u] = b; idx[4 * u + 1] = b + 1; idx[4 * u + 2] = b + 2; idx[4 * u + 3] = b + 3;
    }
    f(x, idx);
  }
}

// (max, sum exp) of one batch, merged into the running pair; returns the batch maximum.
__device__ __forceinline__ float ms_push_batch(MaxSum& a, const float (&x)[ROW_BATCH]) {
  float bm = x[0];
#pragma unroll
  for (int k = 1; k < ROW_BATCH; ++k) bm = fmaxf(bm, x[k]);
  if (bm == -INFINITY) return bm;
  float bs = 0.0f;
#pragma unroll
  for (int k = 0; k < ROW_BATCH; ++k) bs += __expf(x[k] - bm);  // exp(-inf) = 0 for the padding
  a = ms_merge(a, MaxSum{bm, bs});
  return bm;
}

// ---- log_softmax (models/AttModel.py:163) ---------------------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) log_softmax_rows_kernel(const float* __restrict__ logits, long long ld,
                                                                       float* __restrict__ out, long long ld_out, int V) {
  __shared__ MaxSum s_part[ROW_WARPS];
  const float* row = logits + static_cast<long long>(blockIdx.x) * ld;
  float* orow = out + static_cast<long long>(blockIdx.x) * ld_out;
  MaxSum a{-INFINITY, 0.0f};
  scan_row_batches(row, V, [&](const float(&x)[ROW_BATCH], const int(&)[ROW_BATCH]) { ms_push_batch(a, x); });
  const MaxSum r = ms_block_reduce(a, s_part);
  const float log_s = logf(r.s);
  const bool vec_out = (reinterpret_cast<uintptr_t>(orow) & 15) == 0;
  scan_row_batches(row, V, [&](const float(&x)[ROW_BATCH], const int(&idx)[ROW_BATCH]) {
#pragma unroll
    for (int u = 0; u < ROW_U; ++u) {
      const int b = idx[4 * u];
      if (vec_out && b + 4 <= V) {
        *reinterpret_cast<float4*>(orow + b) = make_float4((x[4 * u] - r.m) - log_s, (x[4 * u + 1] - r.m) - log_s,
                                                           (x[4 * u + 2] - r.m) - log_s, (x[4 * u + 3] - r.m) - log_s);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (b + k < V) orow[b + k] = (x[4 * u + k] - r.m) - log_s;
      }
    }
  });
}

int log_softmax_rows(const float* logits, long long ld, float* out, long long ld_out, int rows, int V, cudaStream_t stream) {
  launch_begin("log_softmax_rows", stream);
  log_softmax_rows_kernel<<<rows, ROW_THREADS, 0, stream>>>(logits, ld, out, ld_out, V);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- fused masked cross-entropy forward (misc/criterion.py:143-150) --------------------------------
__global__ void __launch_bounds__(ROW_THREADS) lse_xent_fwd_kernel(const float* __restrict__ logits, long long ld,
                                                                   const int64_t* __restrict__ target,
                                                                   const float* __restrict__ mask, float* __restrict__ lse,
                                                                   float* __restrict__ nll, int V) {
  __shared__ MaxSum s_part[ROW_WARPS];
  const int r = blockIdx.x;
  const float* row = logits + static_cast<long long>(r) * ld;
  MaxSum a{-INFINITY, 0.0f};
  scan_row_batches(row, V, [&](const float(&x)[ROW_BATCH], const int(&)[ROW_BATCH]) { ms_push_batch(a, x); });
  const MaxSum red = ms_block_reduce(a, s_part);
  if (threadIdx.x == 0) {
    const float log_s = logf(red.s);
    lse[r] = red.m + log_s;
    long long t = target[r];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    nll[r] = -((row[t] - red.m) - log_s) * mask[r];
  }
}

int lse_xent_fwd(const float* logits, long long ld, const int64_t* target, const float* mask, float* lse, float* nll, int rows,
                 int V, cudaStream_t stream) {
  launch_begin("lse_xent_fwd", stream);
  lse_xent_fwd_kernel<<<rows, ROW_THREADS, 0, stream>>>(logits, ld, target, mask, lse, nll, V);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- greedy step (models/AttModel.py:218-251, sample_max=1) ------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) greedy_step_kernel(const float* __restrict__ logits, long long ld,
                                                                  int64_t* __restrict__ seq, float* __restrict__ seq_lp,
                                                                  uint8_t* __restrict__ unfinished,
                                                                  int64_t* __restrict__ next_tok,
                                                                  int32_t* __restrict__ n_unfinished, int t, int T, int V,
                                                                  int flags) {
  // The reference leaves the loop once every row has finished (AttModel.py:250-251): later
  // columns of seq / seqLogprobs stay zero.
  if (t > 0 && n_unfinished[t - 1] == 0) return;
  __shared__ MaxSum s_part[ROW_WARPS];
  __shared__ Best s_best[ROW_WARPS];
  const int r = blockIdx.x;
  const float* row = logits + static_cast<long long>(r) * ld;
  const int banned = ((flags & UIC_SAMPLE_DECODING_CONSTRAINT) && t > 0) ? static_cast<int>(seq[static_cast<long long>(r) * T + t - 1]) : -1;
  MaxSum a{-INFINITY, 0.0f};
  Best b{-INFINITY, 0x7fffffff};
  scan_row_batches(row, V, [&](const float(&x)[ROW_BATCH], const int(&idx)[ROW_BATCH]) {
    const float bm = ms_push_batch(a, x);
    if (bm > b.v || b.i == 0x7fffffff) {  // indices increase within and across batches: ">" keeps the first maximum
#pragma unroll
      for (int k = 0; k < ROW_BATCH; ++k) {
        const float xs = (idx[k] == banned) ? -INFINITY : x[k];
        if (idx[k] < V && (xs > b.v || b.i == 0x7fffffff)) {
          b.v = xs;
          b.i = idx[k];
        }
      }
    }
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (better(ov, oi, b)) {
      b.v = ov;
      b.i = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = b;
  const MaxSum red = ms_block_reduce(a, s_part);  // contains the __syncthreads that publishes s_best
  if (threadIdx.x == 0) {
    Best w = s_best[0];
    for (int q = 1; q < ROW_WARPS; ++q)
      if (better(s_best[q].v, s_best[q].i, w)) w = s_best[q];
    const float lp = (w.v - red.m) - logf(red.s);
    long long it = w.i;
    const bool u = (t == 0 ? true : unfinished[r] != 0) && it > 0;  // :242-245
    it = u ? it : 0;                                                 // :246
    seq[static_cast<long long>(r) * T + t] = it;
    seq_lp[static_cast<long long>(r) * T + t] = lp;                  // :248, not masked
    unfinished[r] = u ? 1 : 0;
    next_tok[r] = it;
    if (u) atomicAdd(&n_unfinished[t], 1);
  }
}

int greedy_step(const float* logits, long long ld, int64_t* seq, float* seq_lp, uint8_t* unfinished, int64_t* next_tok,
                int32_t* n_unfinished, int t, int seq_length, int rows, int V, int flags, cudaStream_t stream) {
  launch_begin("greedy_step", stream);
  greedy_step_kernel<<<rows, ROW_THREADS, 0, stream>>>(logits, ld, seq, seq_lp, unfinished, next_tok, n_unfinished, t, seq_length,
                                                       V, flags);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- per-row top-k after the beam-search edits (models/CaptionModel.py:128-133,61) -------------------
template <int KMAX>
__global__ void __launch_bounds__(ROW_THREADS) row_topk_kernel(const float* __restrict__ logits, long long ld,
                                                               const int64_t* __restrict__ prev_tok, float* __restrict__ topk_val,
                                                               int32_t* __restrict__ topk_idx, int V, int k, int flags) {
  __shared__ MaxSum s_part[ROW_WARPS];
  __shared__ float s_val[ROW_THREADS * KMAX];
  __shared__ int s_idx[ROW_THREADS * KMAX];
  __shared__ Best s_best[ROW_WARPS];
  __shared__ int s_winner;
  const int r = blockIdx.x;
  const float* row = logits + static_cast<long long>(r) * ld;
  const int banned = ((flags & UIC_SAMPLE_DECODING_CONSTRAINT) && prev_tok) ? static_cast<int>(prev_tok[r]) : -1;

  float val[KMAX];
  int idx_k[KMAX];
#pragma unroll
  for (int q = 0; q < KMAX; ++q) {
    val[q] = -INFINITY;
    idx_k[q] = 0x7fffffff;
  }
  MaxSum a{-INFINITY, 0.0f};
  scan_row_batches(row, V, [&](const float(&x)[ROW_BATCH], const int(&idx)[ROW_BATCH]) {
    const float bm = ms_push_batch(a, x);
    // the edits only lower values, so a batch whose raw maximum cannot beat the current k-th value is skipped
    if (bm > val[KMAX - 1] || idx_k[KMAX - 1] == 0x7fffffff) {
#pragma unroll
      for (int e = 0; e < ROW_BATCH; ++e) {
        if (idx[e] >= V) continue;
        float xs = (idx[e] == V - 1) ? x[e] - 1000.0f : x[e];  // UNK suppression (:133)
        if (idx[e] == banned) xs = -INFINITY;                   // decoding constraint (:130-131)
        // indices arrive in increasing order per thread, so ">" keeps the smaller index on ties
        if (xs > val[KMAX - 1] || idx_k[KMAX - 1] == 0x7fffffff) {
          val[KMAX - 1] = xs;
          idx_k[KMAX - 1] = idx[e];
#pragma unroll
          for (int q = KMAX - 1; q > 0; --q) {
            if (val[q] > val[q - 1] || idx_k[q - 1] == 0x7fffffff) {
              const float tv = val[q];
              val[q] = val[q - 1];
              val[q - 1] = tv;
              const int ti = idx_k[q];
              idx_k[q] = idx_k[q - 1];
              idx_k[q - 1] = ti;
            }
          }
        }
      }
    }
  });
#pragma unroll
  for (int q = 0; q < KMAX; ++q) {
    s_val[threadIdx.x * KMAX + q] = val[q];
    s_idx[threadIdx.x * KMAX + q] = idx_k[q];
  }
  const MaxSum red = ms_block_reduce(a, s_part);
  const float log_s = logf(red.s);

  int head = 0;
  for (int round = 0; round < k; ++round) {
    Best b{-INFINITY, 0x7fffffff};
    if (head < KMAX && s_idx[threadIdx.x * KMAX + head] != 0x7fffffff) {
      b.v = s_val[threadIdx.x * KMAX + head];
      b.i = s_idx[threadIdx.x * KMAX + head];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, b.v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, b.i, o);
      if (oi != 0x7fffffff && (b.i == 0x7fffffff || better(ov, oi, b))) {
        b.v = ov;
        b.i = oi;
      }
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
      Best w = s_best[0];
      for (int q = 1; q < ROW_WARPS; ++q)
        if (s_best[q].i != 0x7fffffff && (w.i == 0x7fffffff || better(s_best[q].v, s_best[q].i, w))) w = s_best[q];
      s_winner = w.i;
      // ys[q, c] of the reference: the edited log-probability (the UNK shift is applied to the
      // log-prob in fp32, like `logprobsf[:, V-1] - 1000`)
      // (no candidate left -- fewer than k finite entries in the row: -inf and column 0, never an out-of-range read)
      float lp = w.i == 0x7fffffff ? -INFINITY : ((row[w.i] - red.m) - log_s);
      if (w.i == 0x7fffffff) w.i = 0;
      if (w.i == V - 1) lp -= 1000.0f;
      if (w.i == banned) lp = -INFINITY;
      topk_val[static_cast<long long>(r) * k + round] = lp;
      topk_idx[static_cast<long long>(r) * k + round] = w.i;
    }
    __syncthreads();
    if (head < KMAX && s_idx[threadIdx.x * KMAX + head] == s_winner) ++head;
  }
}

int row_topk(const float* logits, long long ld, const int64_t* prev_tok, float* topk_val, int32_t* topk_idx, int rows, int V,
             int k, int flags, cudaStream_t stream) {
  // every thread keeps its KMAX best entries: with k > 16 a thread that holds more than 16 of the row's top k would
  // silently drop some (the instantiations stop at 16; 32 would not fit the static shared-memory budget)
  if (k > 16) return set_error(UIC_ERR_SHAPE, "row_topk: k=%d candidates per row (max 16)", k);
  launch_begin("row_topk", stream);
  if (k <= 4)
    row_topk_kernel<4><<<rows, ROW_THREADS, 0, stream>>>(logits, ld, prev_tok, topk_val, topk_idx, V, k, flags);
  else if (k <= 8)
    row_topk_kernel<8><<<rows, ROW_THREADS, 0, stream>>>(logits, ld, prev_tok, topk_val, topk_idx, V, k, flags);
  else
    row_topk_kernel<16><<<rows, ROW_THREADS, 0, stream>>>(logits, ld, prev_tok, topk_val, topk_idx, V, k, flags);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- merges of the fused logit statistics (helpers in uic_vocab.cuh) ------------------------------------------
template <int KS>
__global__ void __launch_bounds__(32 * MERGE_ROWS_PER_CTA) beam_topk_merge_kernel(const float* __restrict__ stats, int parts,
                                                                                   float* __restrict__ topk_val,
                                                                                   int32_t* __restrict__ topk_idx, int rows, int k) {
  const int r = blockIdx.x * MERGE_ROWS_PER_CTA + (threadIdx.x >> 5);
  if (r >= rows) return;
  float kv[KS];
  int ki[KS];
  const RowStats rs = merge_row_stats<KS>(stats + static_cast<long long>(r) * parts * ((2 + 2 * KS + 3) / 4 * 4), parts, kv, ki);
  for (int round = 0; round < k; ++round) {
    const Best b = warp_pop_best<KS>(kv, ki);
    if ((threadIdx.x & 31) == 0) {
      // the key already carries the beam-search edits (UNK - 1000, banned -inf): ys[q, c] of the reference
      topk_val[static_cast<long long>(r) * k + round] = (b.v - rs.M) - rs.log_s;
      topk_idx[static_cast<long long>(r) * k + round] = b.i;
    }
  }
}

int beam_topk_merge(const float* stats, int parts, int kslots, float* topk_val, int32_t* topk_idx, int rows, int k,
                    cudaStream_t stream) {
  if (k > kslots) return set_error(UIC_ERR_ARG, "beam_topk_merge: k=%d > kslots=%d", k, kslots);
  if (kslots != 1 && kslots != 3 && kslots != 5 && kslots != 8)
    return set_error(UIC_ERR_ARG, "beam_topk_merge: kslots must be 1, 3, 5 or 8");
  const int grid = (rows + MERGE_ROWS_PER_CTA - 1) / MERGE_ROWS_PER_CTA;
  launch_begin("beam_topk_merge", stream);
  if (kslots == 1)
    beam_topk_merge_kernel<1><<<grid, 32 * MERGE_ROWS_PER_CTA, 0, stream>>>(stats, parts, topk_val, topk_idx, rows, k);
  else if (kslots == 3)
    beam_topk_merge_kernel<3><<<grid, 32 * MERGE_ROWS_PER_CTA, 0, stream>>>(stats, parts, topk_val, topk_idx, rows, k);
  else if (kslots == 5)
    beam_topk_merge_kernel<5><<<grid, 32 * MERGE_ROWS_PER_CTA, 0, stream>>>(stats, parts, topk_val, topk_idx, rows, k);
  else if (kslots == 8)
    beam_topk_merge_kernel<8><<<grid, 32 * MERGE_ROWS_PER_CTA, 0, stream>>>(stats, parts, topk_val, topk_idx, rows, k);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// Greedy step from the fused statistics (kslots = 1): same bookkeeping as greedy_step_kernel.
__global__ void __launch_bounds__(32 * MERGE_ROWS_PER_CTA) greedy_merge_kernel(const float* __restrict__ stats, int parts,
                                                                                int64_t* __restrict__ seq, float* __restrict__ seq_lp,
                                                                                uint8_t* __restrict__ unfinished,
                                                                                int64_t* __restrict__ next_tok,
                                                                                int32_t* __restrict__ n_unfinished, int t, int T,
                                                                                int rows) {
  if (t > 0 && n_unfinished[t - 1] == 0) return;  // the reference has left its loop (AttModel.py:250-251)
  const int r = blockIdx.x * MERGE_ROWS_PER_CTA + (threadIdx.x >> 5);
  if (r >= rows) return;
  float kv[1];
  int ki[1];
  const RowStats rs = merge_row_stats<1>(stats + static_cast<long long>(r) * parts * 4, parts, kv, ki);
  const Best b = warp_pop_best<1>(kv, ki);
  if ((threadIdx.x & 31) == 0) {
    long long it = b.i;
    const bool u = (t == 0 ? true : unfinished[r] != 0) && it > 0;
    it = u ? it : 0;
    seq[static_cast<long long>(r) * T + t] = it;
    seq_lp[static_cast<long long>(r) * T + t] = (b.v - rs.M) - rs.log_s;
    unfinished[r] = u ? 1 : 0;
    next_tok[r] = it;
    if (u) atomicAdd(&n_unfinished[t], 1);
  }
}

int greedy_merge(const float* stats, int parts, int64_t* seq, float* seq_lp, uint8_t* unfinished, int64_t* next_tok,
                 int32_t* n_unfinished, int t, int seq_length, int rows, cudaStream_t stream) {
  const int grid = (rows + MERGE_ROWS_PER_CTA - 1) / MERGE_ROWS_PER_CTA;
  launch_begin("greedy_merge", stream);
  greedy_merge_kernel<<<grid, 32 * MERGE_ROWS_PER_CTA, 0, stream>>>(stats, parts, seq, seq_lp, unfinished, next_tok, n_unfinished, t,
                                                                    seq_length, rows);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

}  // namespace uic
