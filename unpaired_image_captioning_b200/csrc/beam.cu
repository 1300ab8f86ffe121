// Beam-search bookkeeping for ALL images of the batch in one launch (one warp per image).
//
// The reference runs this per image in Python with O(beams^2) `.item()` host syncs per step
// (models/CaptionModel.py:48-97,155-172).  The rules reproduced here, per image:
//   * candidates (c, q) = c-th best column of beam row q, score p = sum[q] + ys[q, c] in fp32;
//     at t == 0 only row q = 0 is used (:64-66);
//   * candidates are ordered c-major / q-minor and stably sorted by -p (:67-74): ties go to the
//     smaller c, then the smaller q; the first `beams` become the new beams vix = 0..beams-1;
//   * the new beam copies its parent's prefix (seq and per-token log-probs), appends the token and
//     its (edited, "unaugmented") log-prob, and takes p as its new running sum (:82-95);
//   * a beam whose token is 0, or any beam at the last step, is recorded as finished with
//     p (divided by t+1 under max_ppl), then its running sum is set to -1000; it stays in the
//     beam and keeps being fed token 0 (:155-172);
//   * finished hypotheses are kept sorted by p descending, earlier-recorded first among equals,
//     truncated to `beams` entries (:175): running insertion is equivalent to the final stable sort.
#include "uic_internal.h"
#include "uic_ptx.cuh"
#include "uic_vocab.cuh"

namespace uic {

constexpr int BEAM_MAX = 16;
constexpr int BEAM_T_MAX = 64;
constexpr int CAND_PER_LANE = BEAM_MAX * BEAM_MAX / 32;

// One image, one warp.  tkv / tki: the image's (beams x beams) candidate tables [q * b + c] (global or shared).
__device__ __forceinline__ void beam_step_image(const float* tkv, const int32_t* tki, const float* tku, int32_t* __restrict__ beam_seq,
                                                float* __restrict__ beam_lp, float* __restrict__ beam_sum,
                                                int32_t* __restrict__ done_seq, float* __restrict__ done_lp,
                                                double* __restrict__ done_p, float* __restrict__ done_unaug,
                                                int32_t* __restrict__ done_cnt, int32_t* __restrict__ parent_row,
                                                int64_t* __restrict__ next_tok, int img, int t, int T, int b, int flags) {
  __shared__ int32_t s_seq[BEAM_MAX * BEAM_T_MAX];
  __shared__ float s_lp[BEAM_MAX * BEAM_T_MAX];
  __shared__ int s_q[BEAM_MAX], s_tok[BEAM_MAX];
  __shared__ float s_p[BEAM_MAX], s_r[BEAM_MAX];
  const int lane = threadIdx.x & 31;
  const long long row0 = static_cast<long long>(img) * b;
  int32_t* seq_img = beam_seq + row0 * T;
  float* lp_img = beam_lp + row0 * T;
  float* sum_img = beam_sum + row0;

  // ---- candidate scores, c-major / q-minor -------------------------------------------------
  const int rows = (t == 0) ? 1 : b;
  const int n_cand = b * rows;
  float cp[CAND_PER_LANE];
  bool used[CAND_PER_LANE];
#pragma unroll
  for (int i = 0; i < CAND_PER_LANE; ++i) {
    const int n = lane + 32 * i;
    used[i] = n >= n_cand;
    cp[i] = -INFINITY;
    if (!used[i]) {
      const int c = n / rows, q = n - c * rows;
      cp[i] = sum_img[q] + tkv[q * b + c];
    }
  }
  // ---- stable top-`b` selection: b rounds of warp arg-max, ties -> smaller candidate number ----
  for (int vix = 0; vix < b; ++vix) {
    float bv = -INFINITY;
    int bn = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < CAND_PER_LANE; ++i) {
      const int n = lane + 32 * i;
      if (!used[i] && (bn == 0x7fffffff || cp[i] > bv)) {  // per lane n increases with i: ">" keeps the smaller n
        bv = cp[i];
        bn = n;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int on = __shfl_xor_sync(0xffffffffu, bn, o);
      if (on != 0x7fffffff && (bn == 0x7fffffff || ov > bv || (ov == bv && on < bn))) {
        bv = ov;
        bn = on;
      }
    }
#pragma unroll
    for (int i = 0; i < CAND_PER_LANE; ++i)
      if (lane + 32 * i == bn) used[i] = true;
    if (lane == 0) {
      const int c = bn / rows, q = bn - c * rows;
      s_q[vix] = q;
      s_p[vix] = bv;
      s_tok[vix] = tki[q * b + c];
      s_r[vix] = tku != nullptr ? tku[q * b + c] : tkv[q * b + c];  // stored log-prob: without the diversity penalty (:38,86)
    }
  }
  // ---- fork the tables: stage the old prefixes, then write parents' prefixes + the new token ----
  for (int e = lane; e < b * t; e += 32) {
    const int q = e / t, s = e - q * t;
    s_seq[q * BEAM_T_MAX + s] = seq_img[q * T + s];
    s_lp[q * BEAM_T_MAX + s] = lp_img[q * T + s];
  }
  __syncwarp();
  for (int e = lane; e < b * t; e += 32) {
    const int vix = e / t, s = e - vix * t;
    const int q = s_q[vix];
    seq_img[vix * T + s] = s_seq[q * BEAM_T_MAX + s];
    lp_img[vix * T + s] = s_lp[q * BEAM_T_MAX + s];
  }
  if (lane < b) {
    seq_img[lane * T + t] = s_tok[lane];
    lp_img[lane * T + t] = s_r[lane];
    sum_img[lane] = s_p[lane];
    parent_row[row0 + lane] = static_cast<int32_t>(row0) + s_q[lane];
    next_tok[row0 + lane] = s_tok[lane];
  }
  __syncwarp();
  __threadfence_block();
  // ---- record finished hypotheses ---------------------------------------------------------------
  int32_t* dseq = done_seq + row0 * T;
  float* dlp = done_lp + row0 * T;
  double* dp = done_p + row0;
  float* dun = done_unaug + row0;
  int cnt = done_cnt[img];
  for (int vix = 0; vix < b; ++vix) {
    if (!(s_tok[vix] == 0 || t == T - 1)) continue;  // warp-uniform
    double p = static_cast<double>(s_p[vix]);
    if (flags & UIC_BEAM_MAX_PPL) p = p / static_cast<double>(t + 1);
    int pos = 0;
    while (pos < cnt && dp[pos] >= p) ++pos;  // stable: after every entry with an equal or better score
    if (pos < b) {
      const int last = (cnt < b) ? cnt : b - 1;  // index that the tail entry moves into / is dropped from
      for (int j = last; j > pos; --j) {
        for (int s = lane; s < T; s += 32) {
          dseq[j * T + s] = dseq[(j - 1) * T + s];
          dlp[j * T + s] = dlp[(j - 1) * T + s];
        }
        if (lane == 0) {
          dp[j] = dp[j - 1];
          dun[j] = dun[j - 1];
        }
        __syncwarp();
      }
      float unaug = 0.0f;
      for (int s = 0; s <= t; ++s) unaug += lp_img[vix * T + s];
      for (int s = lane; s < T; s += 32) {
        dseq[pos * T + s] = (s <= t) ? seq_img[vix * T + s] : 0;
        dlp[pos * T + s] = (s <= t) ? lp_img[vix * T + s] : 0.0f;
      }
      if (lane == 0) {
        dp[pos] = p;
        dun[pos] = unaug;
      }
      if (cnt < b) ++cnt;
      __syncwarp();
    }
    if (lane == 0) sum_img[vix] = -1000.0f;  // :167
  }
  if (lane == 0) done_cnt[img] = cnt;
}

__global__ void __launch_bounds__(32) beam_step_kernel(const float* __restrict__ topk_val, const int32_t* __restrict__ topk_idx,
                                                       const float* __restrict__ topk_unaug,
                                                       int32_t* __restrict__ beam_seq, float* __restrict__ beam_lp,
                                                       float* __restrict__ beam_sum, int32_t* __restrict__ done_seq,
                                                       float* __restrict__ done_lp, double* __restrict__ done_p,
                                                       float* __restrict__ done_unaug, int32_t* __restrict__ done_cnt,
                                                       int32_t* __restrict__ parent_row, int64_t* __restrict__ next_tok, int t,
                                                       int T, int b, int flags) {
  const long long off = static_cast<long long>(blockIdx.x) * b * b;
  beam_step_image(topk_val + off, topk_idx + off, topk_unaug != nullptr ? topk_unaug + off : nullptr, beam_seq, beam_lp, beam_sum, done_seq, done_lp, done_p, done_unaug, done_cnt,
                  parent_row, next_tok, blockIdx.x, t, T, b, flags);
}

int beam_step(const float* topk_val, const int32_t* topk_idx, const float* topk_unaug, int32_t* beam_seq, float* beam_lp, float* beam_sum,
              int32_t* done_seq, float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt, int32_t* parent_row,
              int64_t* next_tok, int t, int seq_length, int n_img, int beams, int flags, cudaStream_t stream) {
  if (beams > BEAM_MAX || seq_length > BEAM_T_MAX)
    return set_error(UIC_ERR_SHAPE, "beam_step: beams=%d (max %d), seq_length=%d (max %d)", beams, BEAM_MAX, seq_length, BEAM_T_MAX);
  launch_begin("beam_step", stream);
  beam_step_kernel<<<n_img, 32, 0, stream>>>(topk_val, topk_idx, topk_unaug, beam_seq, beam_lp, beam_sum, done_seq, done_lp, done_p,
                                             done_unaug, done_cnt, parent_row, next_tok, t, seq_length, beams, flags);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- diverse beam search: candidates of one group after the diversity penalty (models/CaptionModel.py:36-45) ----------
// Input: per row the K' = k + n_prev best (edited) log-probs and their columns (uic_row_topk with k = K'), enough to
// contain the penalised top-k (a penalty only lowers values).  Every candidate loses `lambda` once per occurrence of its
// token among the tokens that the groups before this one hold at local step `lt` (tables beam_seq [group][image][beam][T]);
// the k best by (penalised value, smaller column) come out with both the penalised and the unpenalised value.
__global__ void __launch_bounds__(128) diverse_select_kernel(const float* __restrict__ cand_val, const int32_t* __restrict__ cand_idx,
                                                              int kp, const int32_t* __restrict__ beam_seq, int group, int n_img,
                                                              int bdash, int T, int lt, float lambda, float* __restrict__ topk_val,
                                                              float* __restrict__ topk_unaug, int32_t* __restrict__ topk_idx,
                                                              int rows) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int img = r / bdash;
  float un = -INFINITY, aug = -INFINITY;
  int col = 0x7fffffff;
  if (lane < kp) {
    un = cand_val[static_cast<long long>(r) * kp + lane];
    col = cand_idx[static_cast<long long>(r) * kp + lane];
    int hits = 0;
    for (int g = 0; g < group; ++g)
      for (int j = 0; j < bdash; ++j)
        hits += beam_seq[((static_cast<long long>(g) * n_img + img) * bdash + j) * T + lt] == col;
    aug = un - lambda * static_cast<float>(hits);
  }
  bool taken = lane >= kp;
  for (int round = 0; round < bdash; ++round) {
    float bv = taken ? -INFINITY : aug;
    int bi = taken ? 0x7fffffff : col, bl = taken ? 32 : lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o), ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ol < 32 && (bl == 32 || ov > bv || (ov == bv && oi < bi))) {
        bv = ov;
        bi = oi;
        bl = ol;
      }
    }
    const float bu = __shfl_sync(0xffffffffu, un, bl & 31);
    if (lane == 0) {
      topk_val[static_cast<long long>(r) * bdash + round] = bv;
      topk_unaug[static_cast<long long>(r) * bdash + round] = bu;
      topk_idx[static_cast<long long>(r) * bdash + round] = bi;
    }
    if (lane == bl) taken = true;
  }
}

int diverse_select(const float* cand_val, const int32_t* cand_idx, int kp, const int32_t* beam_seq, int group, int n_img, int bdash,
                   int T, int lt, float lambda, float* topk_val, float* topk_unaug, int32_t* topk_idx, cudaStream_t stream) {
  if (kp > 32 || kp < bdash) return set_error(UIC_ERR_SHAPE, "diverse_select: %d candidates for %d beams (max 32)", kp, bdash);
  const int rows = n_img * bdash;
  launch_begin("diverse_select", stream);
  diverse_select_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(cand_val, cand_idx, kp, beam_seq, group, n_img, bdash, T, lt, lambda, topk_val,
                                                           topk_unaug, topk_idx, rows);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- fused step tail of the sampling loops ------------------------------------------------------------------------
// Everything between the statistics GEMM of step t and the gate GEMM of step t + 1 in ONE launch per step (it was
// four: uic_beam_topk_merge, uic_beam_step, uic_beam_gather, uic_embed_rows -- each a few microseconds of launch
// latency and dependent-kernel gap on a 130 us step).  One CTA of four warps per image:
//   1. warps merge the statistics parts of the image's beam rows into their top-`beams` candidates (shared memory);
//   2. warp 0 runs the beam bookkeeping above on them;
//   3. all warps move the recurrent state of the chosen parents into the other state buffer and write the next
//      step's word embeddings next to it.
struct AdvanceIO {
  // state re-ordering (uic_beam_gather): two column ranges of the bf16 activation matrix + n_state fp32 matrices
  const __nv_bfloat16* x_src;
  __nv_bfloat16* x_dst;
  long long ld_x;
  int col0_a, ncol_a, col0_b, ncol_b;
  const float* c_src;
  float* c_dst;
  int n_state, rows, H;
  // beams per image in the SOURCE buffers (stats, x_src, c_src).  Equal to the beam width, except at the first step when
  // the caller ran the step on one row per image: only beam 0 is read at t = 0 (rows = 1, models/CaptionModel.py:56).
  int src_beams;
  // next input (uic_embed_rows): x_dst[r, xt_col0 : xt_col0 + E] = table[tok[r]]
  const __nv_bfloat16* table;
  long long ld_table;
  int xt_col0, E, V;
  long long* trace;  // debug (uic_gemm_set_trace buffer): CTA 0 records %globaltimer after each phase in slots 110..113
};
#define ADV_TRACE(slot)                                                         \
  do {                                                                         \
    if (io.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {          \
      long long tnow;                                                          \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tnow));                  \
      io.trace[slot] = tnow;                                                   \
    }                                                                          \
  } while (0)

__device__ __forceinline__ void copy_row_bf16(__nv_bfloat16* dst, const __nv_bfloat16* src, int n, int tid, int nthreads) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0 && (n & 7) == 0) {
    for (int c = tid; c < n / 8; c += nthreads) reinterpret_cast<uint4*>(dst)[c] = reinterpret_cast<const uint4*>(src)[c];
  } else {
    for (int c = tid; c < n; c += nthreads) dst[c] = src[c];
  }
}
__device__ __forceinline__ void copy_row_f32(float* dst, const float* src, int n, int tid, int nthreads) {
  if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0 && (n & 3) == 0) {
    for (int c = tid; c < n / 4; c += nthreads) reinterpret_cast<float4*>(dst)[c] = reinterpret_cast<const float4*>(src)[c];
  } else {
    for (int c = tid; c < n; c += nthreads) dst[c] = src[c];
  }
}

constexpr int ADV_THREADS = 128;

// (8 warps: the candidate merge of phase 1 and the state copies of phase 3 run one warp per beam row, so every row of up to
//  8 beams is in flight at once -- with 4 warps the fifth beam of configs[4] waited for a second round)
constexpr int BEAM_ADV_THREADS = 256;

template <int KS>
__global__ void __launch_bounds__(BEAM_ADV_THREADS) beam_advance_kernel(const float* __restrict__ stats, int parts,
                                                                    int32_t* __restrict__ beam_seq, float* __restrict__ beam_lp,
                                                                    float* __restrict__ beam_sum, int32_t* __restrict__ done_seq,
                                                                    float* __restrict__ done_lp, double* __restrict__ done_p,
                                                                    float* __restrict__ done_unaug, int32_t* __restrict__ done_cnt,
                                                                    int32_t* __restrict__ parent_row, int64_t* __restrict__ next_tok,
                                                                    int t, int T, int b, int flags, int move_state, AdvanceIO io) {
  __shared__ float s_tkv[KS * KS];
  __shared__ int32_t s_tki[KS * KS];
  pdl_launch_dependents();
  pdl_wait();
  ADV_TRACE(110);
  constexpr int ES = (2 + 2 * KS + 3) / 4 * 4;
  const int img = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = static_cast<long long>(img) * b;
  // 1. candidates of every beam row: ys[q, c] = log-prob of the c-th best column after the beam-search edits
  const long long src_row0 = static_cast<long long>(img) * io.src_beams;
  for (int q = warp; q < io.src_beams; q += BEAM_ADV_THREADS / 32) {
    float kv[KS];
    int ki[KS];
    const RowStats rs = merge_row_stats<KS>(stats + (src_row0 + q) * parts * ES, parts, kv, ki);
    for (int c = 0; c < b; ++c) {
      const Best best = warp_pop_best<KS>(kv, ki);
      if (lane == 0) {
        s_tkv[q * b + c] = (best.v - rs.M) - rs.log_s;
        s_tki[q * b + c] = best.i;
      }
    }
  }
  __syncthreads();
  ADV_TRACE(111);
  // 2. beam bookkeeping
  if (warp == 0)
    beam_step_image(s_tkv, s_tki, nullptr, beam_seq, beam_lp, beam_sum, done_seq, done_lp, done_p, done_unaug, done_cnt, parent_row,
                    next_tok, img, t, T, b, flags);
  if (!move_state) return;
  __syncthreads();
  ADV_TRACE(112);
  // 3. state of the parents + next embeddings into the other buffer: one warp per beam row, so the rows' loads are
  //    all in flight together
  const int lane3 = threadIdx.x & 31;
  for (int v = warp; v < b; v += BEAM_ADV_THREADS / 32) {
    const long long r = row0 + v;
    const long long q = src_row0 + (parent_row[r] - row0);   // the parent's row in the source buffers
    long long tk = next_tok[r];
    tk = tk < 0 ? 0 : (tk >= io.V ? io.V - 1 : tk);
    copy_row_bf16(io.x_dst + r * io.ld_x + io.col0_a, io.x_src + q * io.ld_x + io.col0_a, io.ncol_a, lane3, 32);
    copy_row_bf16(io.x_dst + r * io.ld_x + io.col0_b, io.x_src + q * io.ld_x + io.col0_b, io.ncol_b, lane3, 32);
    copy_row_bf16(io.x_dst + r * io.ld_x + io.xt_col0, io.table + tk * io.ld_table, io.E, lane3, 32);
    for (int s = 0; s < io.n_state; ++s)
      copy_row_f32(io.c_dst + (static_cast<long long>(s) * io.rows + r) * io.H,
                   io.c_src + (static_cast<long long>(s) * gridDim.x * io.src_beams + q) * io.H, io.H, lane3, 32);
  }
  __syncthreads();
  ADV_TRACE(113);
}

int beam_advance(const float* stats, int parts, int kslots, int32_t* beam_seq, float* beam_lp, float* beam_sum, int32_t* done_seq,
                 float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt, int32_t* parent_row, int64_t* next_tok, int t,
                 int seq_length, int n_img, int beams, int flags, int move_state, const void* x_src, void* x_dst, long long ld_x,
                 int col0_a, int ncol_a, int col0_b, int ncol_b, const float* c_src, float* c_dst, int n_state, int H,
                 const void* table, long long ld_table, int xt_col0, int E, int V, int src_beams, cudaStream_t stream) {
  if (src_beams != beams && !(src_beams == 1 && t == 0))
    return set_error(UIC_ERR_ARG, "beam_advance: src_beams=%d must be beams=%d, or 1 at the first step (t=%d)", src_beams, beams, t);
  if (beams > kslots || seq_length > BEAM_T_MAX)
    return set_error(UIC_ERR_SHAPE, "beam_advance: beams=%d (kslots %d), seq_length=%d (max %d)", beams, kslots, seq_length, BEAM_T_MAX);
  AdvanceIO io{static_cast<const __nv_bfloat16*>(x_src), static_cast<__nv_bfloat16*>(x_dst), ld_x, col0_a, ncol_a, col0_b, ncol_b,
               c_src, c_dst, n_state, n_img * beams, H, src_beams, static_cast<const __nv_bfloat16*>(table), ld_table, xt_col0, E, V,
               gemm_trace_buffer()};
  launch_begin("beam_advance", stream);
#define UIC_ADV(KS_)                                                                                                         \
  UIC_CUDA_OK(launch_pdl(beam_advance_kernel<KS_>, dim3(n_img), dim3(BEAM_ADV_THREADS), 0, stream, stats, parts, beam_seq, beam_lp, beam_sum, \
                         done_seq, done_lp, done_p, done_unaug, done_cnt, parent_row, next_tok, t, seq_length, beams, flags,       \
                         move_state, io))
  if (kslots == 1)
    UIC_ADV(1);
  else if (kslots == 3)
    UIC_ADV(3);
  else if (kslots == 5)
    UIC_ADV(5);
  else if (kslots == 8)
    UIC_ADV(8);
  else
    return set_error(UIC_ERR_ARG, "beam_advance: kslots must be 1, 3, 5 or 8");
#undef UIC_ADV
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// Greedy analogue: uic_greedy_merge + uic_embed_rows of the next step in one launch (one warp per row).
__global__ void __launch_bounds__(ADV_THREADS) greedy_advance_kernel(const float* __restrict__ stats, int parts,
                                                                      int64_t* __restrict__ seq, float* __restrict__ seq_lp,
                                                                      uint8_t* __restrict__ unfinished, int64_t* __restrict__ next_tok,
                                                                      int32_t* __restrict__ n_unfinished, int t, int T, int rows,
                                                                      const __nv_bfloat16* __restrict__ table, long long ld_table,
                                                                      __nv_bfloat16* __restrict__ x, long long ld_x, int E, int V,
                                                                      float temperature, const unsigned long long* __restrict__ seed) {
  pdl_launch_dependents();
  pdl_wait();
  if (t > 0 && n_unfinished[t - 1] == 0) return;  // the reference has left its loop (AttModel.py:250-251)
  const int r = blockIdx.x * (ADV_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float kv[1];
  int ki[1];
  const RowStats rs = merge_row_stats<1>(stats + static_cast<long long>(r) * parts * 4, parts, kv, ki);
  const Best best = warp_pop_best<1>(kv, ki);
  long long it = best.i;
  const bool u = (t == 0 ? true : unfinished[r] != 0) && it > 0;
  it = u ? it : 0;
  __syncwarp();
  // sampling: the key is x / T + g(row, column); the reported log-prob is that of the unperturbed logit (AttModel.py:238)
  const float xbest = temperature > 0.0f ? (best.v - rng_gumbel(rng_row_key(rng_step_key(*seed, t), r), best.i)) * temperature : best.v;
  if (lane == 0) {
    seq[static_cast<long long>(r) * T + t] = it;
    seq_lp[static_cast<long long>(r) * T + t] = (xbest - rs.M) - rs.log_s;
    unfinished[r] = u ? 1 : 0;
    next_tok[r] = it;
    if (u) atomicAdd(&n_unfinished[t], 1);
  }
  if (x != nullptr) {
    const long long tk = it >= V ? V - 1 : it;
    copy_row_bf16(x + static_cast<long long>(r) * ld_x, table + tk * ld_table, E, lane, 32);
  }
}

// Scheduled sampling (models/AttModel.py:130-143): the input token of teacher-forced step t is, with probability ss_prob
// per row, a draw from the model's own distribution of step t - 1 (Gumbel-max over the statistics of that step's logits,
// temperature 1) instead of the ground-truth token; the choice is written to tokens_out and its embedding row to x_xt.
__global__ void __launch_bounds__(ADV_THREADS) ss_advance_kernel(const float* __restrict__ stats, int parts,
                                                                  const int64_t* __restrict__ gt_tok, long long gt_stride,
                                                                  float ss_prob, const unsigned long long* __restrict__ seed, int t,
                                                                  int64_t* __restrict__ tokens_out, int rows,
                                                                  const __nv_bfloat16* __restrict__ table, long long ld_table,
                                                                  __nv_bfloat16* __restrict__ x, long long ld_x, int E, int V) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x * (ADV_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float kv[1];
  int ki[1];
  merge_row_stats<1>(stats + static_cast<long long>(r) * parts * 4, parts, kv, ki);
  const Best best = warp_pop_best<1>(kv, ki);
  const uint32_t rk = rng_row_key(rng_step_key(*seed, t), r);
  long long tk = rng_uniform(rk, RNG_ROW_DRAW_COL) < ss_prob ? static_cast<long long>(best.i) : gt_tok[r * gt_stride];
  tk = tk < 0 ? 0 : (tk >= V ? V - 1 : tk);
  if (lane == 0) tokens_out[r] = tk;
  copy_row_bf16(x + static_cast<long long>(r) * ld_x, table + tk * ld_table, E, lane, 32);
}

int ss_advance(const float* stats, int parts, const int64_t* gt_tok, long long gt_stride, float ss_prob,
               const unsigned long long* seed, int t, int64_t* tokens_out, int rows, const void* table, long long ld_table,
               void* x_xt, long long ld_x, int E, int V, cudaStream_t stream) {
  const int per = ADV_THREADS / 32;
  launch_begin("ss_advance", stream);
  UIC_CUDA_OK(launch_pdl(ss_advance_kernel, dim3((rows + per - 1) / per), dim3(ADV_THREADS), 0, stream, stats, parts, gt_tok, gt_stride,
                         ss_prob, seed, t, tokens_out, rows, static_cast<const __nv_bfloat16*>(table), ld_table,
                         static_cast<__nv_bfloat16*>(x_xt), ld_x, E, V));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

int greedy_advance(const float* stats, int parts, int64_t* seq, float* seq_lp, uint8_t* unfinished, int64_t* next_tok,
                   int32_t* n_unfinished, int t, int seq_length, int rows, const void* table, long long ld_table, void* x_xt,
                   long long ld_x, int E, int V, float temperature, const unsigned long long* seed, cudaStream_t stream) {
  if (temperature > 0.0f && seed == nullptr) return set_error(UIC_ERR_ARG, "greedy_advance: sampling needs a seed (device pointer)");
  const int per = ADV_THREADS / 32;
  launch_begin("greedy_advance", stream);
  UIC_CUDA_OK(launch_pdl(greedy_advance_kernel, dim3((rows + per - 1) / per), dim3(ADV_THREADS), 0, stream, stats, parts, seq, seq_lp, unfinished,
                         next_tok, n_unfinished, t, seq_length, rows, static_cast<const __nv_bfloat16*>(table), ld_table,
                         static_cast<__nv_bfloat16*>(x_xt), ld_x, E, V, temperature, seed));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

}  // namespace uic
