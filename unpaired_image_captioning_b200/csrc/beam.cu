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

namespace uic {

constexpr int BEAM_MAX = 16;
constexpr int BEAM_T_MAX = 64;
constexpr int CAND_PER_LANE = BEAM_MAX * BEAM_MAX / 32;

__global__ void __launch_bounds__(32) beam_step_kernel(const float* __restrict__ topk_val, const int32_t* __restrict__ topk_idx,
                                                       int32_t* __restrict__ beam_seq, float* __restrict__ beam_lp,
                                                       float* __restrict__ beam_sum, int32_t* __restrict__ done_seq,
                                                       float* __restrict__ done_lp, double* __restrict__ done_p,
                                                       float* __restrict__ done_unaug, int32_t* __restrict__ done_cnt,
                                                       int32_t* __restrict__ parent_row, int64_t* __restrict__ next_tok, int t,
                                                       int T, int b, int flags) {
  __shared__ int32_t s_seq[BEAM_MAX * BEAM_T_MAX];
  __shared__ float s_lp[BEAM_MAX * BEAM_T_MAX];
  __shared__ int s_q[BEAM_MAX], s_tok[BEAM_MAX];
  __shared__ float s_p[BEAM_MAX], s_r[BEAM_MAX];
  const int img = blockIdx.x;
  const int lane = threadIdx.x;
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
      cp[i] = sum_img[q] + topk_val[(row0 + q) * b + c];
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
      s_tok[vix] = topk_idx[(row0 + q) * b + c];
      s_r[vix] = topk_val[(row0 + q) * b + c];
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

int beam_step(const float* topk_val, const int32_t* topk_idx, int32_t* beam_seq, float* beam_lp, float* beam_sum,
              int32_t* done_seq, float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt, int32_t* parent_row,
              int64_t* next_tok, int t, int seq_length, int n_img, int beams, int flags, cudaStream_t stream) {
  if (beams > BEAM_MAX || seq_length > BEAM_T_MAX)
    return set_error(UIC_ERR_SHAPE, "beam_step: beams=%d (max %d), seq_length=%d (max %d)", beams, BEAM_MAX, seq_length, BEAM_T_MAX);
  launch_begin("beam_step", stream);
  beam_step_kernel<<<n_img, 32, 0, stream>>>(topk_val, topk_idx, beam_seq, beam_lp, beam_sum, done_seq, done_lp, done_p,
                                             done_unaug, done_cnt, parent_row, next_tok, t, seq_length, beams, flags);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

}  // namespace uic
