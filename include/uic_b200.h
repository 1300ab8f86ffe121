/*
 * uic_b200.h -- C ABI of the B200-native attention-LSTM caption decoder kernels.
 *
 * The reference (gujiuxiang/unpaired_image_captioning, pivot_based_eccv2018/) has no FFI or operator
 * registry on this path: every operation below is a chain of torch ops inside models/AttModel.py,
 * models/CaptionModel.py and misc/criterion.py.  Each entry point names the reference lines it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated or freed here;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered and asynchronous;
 *   - return value 0 = ok, negative = UIC_ERR_*; uic_last_error() gives the message (thread-local);
 *   - bf16 buffers are raw uint16 storage (torch.bfloat16); "ld*" are row pitches in ELEMENTS;
 *   - token ids are int64 (torch.long), like the reference; token 0 = BOS = EOS = pad, the last
 *     vocabulary index V-1 is UNK (scripts/prepro_labels.py:86-88).
 */
#ifndef UIC_B200_H_
#define UIC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UIC_OK 0
#define UIC_ERR_ARG (-1)     /* null / inconsistent arguments */
#define UIC_ERR_SHAPE (-2)   /* unsupported or empty shape */
#define UIC_ERR_ALIGN (-3)   /* pointer or pitch alignment */
#define UIC_ERR_CUDA (-4)    /* CUDA runtime / driver error (message has the detail) */
#define UIC_ERR_DEVICE (-5)  /* not an sm_100 device */

/* uic_gemm_bf16 flags */
#define UIC_GEMM_RELU 1
#define UIC_GEMM_ACCUMULATE 2   /* c_f32 += result (used for gradient accumulation) */
#define UIC_GEMM_A_MN_MAJOR 4   /* A is stored [K, M] row-major instead of [M, K] */
#define UIC_GEMM_B_MN_MAJOR 8   /* B is stored [K, N] row-major instead of [N, K] */
#define UIC_GEMM_OUT_F16 16     /* the 16-bit output buffer receives IEEE fp16 instead of bf16 */
#define UIC_GEMM_A_STREAM 32    /* A is read once (raw features): load it with the L2 evict-first hint */
#define UIC_GEMM_B_STREAM 64    /* B is read once: evict-first (default: K-major B up to 32 MB is a weight, kept evict-last) */

/* sampling flags (uic_greedy_step / uic_row_topk) */
#define UIC_SAMPLE_DECODING_CONSTRAINT 1 /* -inf on the previous token (AttModel.py:220-223, CaptionModel.py:130-131) */
#define UIC_BEAM_MAX_PPL 2               /* divide finished scores by length (CaptionModel.py:163-164) */

/* ---- library ------------------------------------------------------------------------------- */
const char* uic_last_error(void);
int uic_version(void);
/* Kernel launches issued through this library since process start (bench.py's gpu_launches). */
int64_t uic_launch_count(void);
/* 0 = tcgen05 tensor-core GEMM (default), 1 = CUDA-core verification GEMM (tests/debug only). */
int uic_set_gemm_impl(int impl);
/* Debug aid: with a non-NULL device buffer of 1024 int64, CTA 0 of every following GEMM launch records
 * clock64() at its pipeline events ([0] setup done, [1+kb] operands of k-block kb landed, [50+kb] TMA for kb
 * issued, [40] last MMA issued, [41] accumulator ready, [42] epilogue done) and CTA 0 of every attention
 * launch the %globaltimer of its batch pipeline ([0] start, [1+4i..] batch i: data ready / scored / previous
 * batch finished, [100] end); slots [128+2c], [129+2c] receive the %globaltimer at entry / exit of CTA c < 448
 * of both kernels.  NULL turns it off. */
int uic_gemm_set_trace(void* device_buffer_1024_i64);
/* Live per-kernel device timing: while enabled, every launch is bracketed by a CUDA event pair on
 * its stream (do not enable during CUDA-graph capture).  uic_profile_dump synchronises, writes one
 * "name launches total_ms" line per kernel label into `out` (host buffer) and returns the number
 * of labels; uic_profile_enable(0/1) clears the records. */
int uic_profile_enable(int on);
int64_t uic_profile_dump(char* out, int64_t cap);
/* Fails with UIC_ERR_DEVICE unless the current device is compute capability 10.x. */
int uic_check_device(void);

/* ---- dense contractions -------------------------------------------------------------------- */
/* D[M,N] = act(A[M,K] * B[N,K]^T + bias[N] (+ D)), bf16 operands, fp32 accumulate on tcgen05.
 * Replaces every nn.Linear / nn.LSTMCell matmul on the path: models/AttModel.py:76-92 (fc_embed,
 * att_embed, logit, ctx2att), :426-427 (LSTMCell), :535 (h2att), :574-576 (a2c, i2h, h2h) and their
 * autograd backward (dgrad: B MN-major, wgrad: A and B MN-major).
 * Either output may be NULL (but not both): c_f32 (fp32, pitch ldc) and c_bf16 (bf16, pitch ldcb). */
int uic_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, float* c_f32, int64_t ldc, void* c_bf16,
                  int64_t ldcb, const float* bias, int M, int N, int K, int flags, void* stream);

/* Same, with the attention-operand epilogue: when exp_scale != 0, output columns >= exp_col0 become
 * exp_scale * exp(2 x).  The additive attention needs tanh(p_att + att_h) for every (region, unit,
 * beam, step); with E = exp(2 p_att) stored once per image (ctx2att, a bf16 tile: fp32's exponent
 * range, so no p_att saturates it) and F = exp(2 att_h) produced by the h2att projection each step,
 * tanh(p + a) = 1 - 2 / (E F + 1) costs one FMA and (a share of) one reciprocal instead of one
 * MUFU.TANH (a quarter-rate op).  The exponentials are capped at 2^60 (|x| <= 20.8). */
int uic_gemm_bf16_ex(const void* A, int64_t lda, const void* B, int64_t ldb, float* c_f32, int64_t ldc, void* c_16, int64_t ldc16,
                     const float* bias, int M, int N, int K, int flags, int exp_col0, float exp_scale, void* stream);

/* D = act(A B^T + bias) * post_scale[n] + post_shift[n]: a per-column affine behind the activation -- the eval-mode
 * nn.BatchNorm1d(rnn_size) that use_bn = 2 appends to att_embed (models/AttModel.py:79-84), fused into the Linear + ReLU GEMM.
 * post_scale / post_shift: fp32 vectors of length N. */
int uic_gemm_bf16_affine(const void* A, int64_t lda, const void* B, int64_t ldb, float* c_f32, int64_t ldc, void* c_16, int64_t ldc16,
                         const float* bias, const float* post_scale, const float* post_shift, int M, int N, int K, int flags,
                         void* stream);

/* ---- elementwise / layout ------------------------------------------------------------------ */
/* dst_bf16[r, c] = bf16(relu?(src[r, c])) for an (rows x cols) fp32 matrix. Used to stage fp32
 * features and weights as tensor-core operands. */
int uic_cast_f32_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int64_t cols, int relu,
                      void* stream);
/* out[r, 0:E] = table[tok[r], 0:E] (bf16 rows; ReLU is pre-applied to the table copy).
 * Replaces self.embed = Embedding+ReLU (models/AttModel.py:73-75,160). */
int uic_embed_rows(const void* table_bf16, int64_t ld_table, const int64_t* tok, void* out_bf16, int64_t ld_out, int rows,
                   int E, int V, void* stream);

/* x[i, l, :] = 0 for l >= sum(att_masks[i, :]): the zero padding pad_packed_sequence leaves behind
 * in pack_wrapper (models/AttModel.py:44-53). x is (n_img*L, H) bf16 contiguous. */
int uic_zero_padded_rows(void* x_bf16, const float* att_masks, int n_img, int L, int H, void* stream);

/* ---- fused additive attention step (models/AttModel.py:538-558) ------------------------------
 * For image i (n_img of them) and beam j < beams (row r = i*beams + j):
 *   e[l]  = sum_a w_alpha[a] * tanh(p_att[i,l,a] + att_h[r,a])        (alpha_net bias cancels in softmax)
 *   alpha = softmax_l(e);  if att_masks: alpha = alpha*m / sum(alpha*m)
 *   ctx[r,:] = sum_l alpha[l] * att[i,l,:]
 * Operands are passed in the exponential form produced by uic_gemm_bf16_ex (exp epilogue, scale 1, capped at 2^60):
 *   att_h  -> F[r,a]   = exp(2 * (h2att(h)[r,a] + bias))   fp32, pitch ld_att_h
 *   p_att  -> E[i,l,a] = exp(2 * p_att[i,l,a])             bf16 (n_img, L, A)   (the parameter keeps its round-1 name)
 * so that tanh(p_att + att_h) = 1 - 2 / (E F + 1).  att is bf16 (n_img,L,H).  Both tiles are read
 * once per image and shared by the image's beams (up to 3 per pass; up to 5 where A = 512 < H <= 1024).
 * Outputs (each optional): ctx_bf16, ctx_f32, alpha (rows x L fp32, saved for backward).
 * att_h must be 16-byte aligned with a pitch that is a multiple of 4 floats (its rows are staged by bulk copies);
 * the context weights are rounded to bf16 for the tensor-core product (fp32 accumulation, fp32 normalisation).
 * `workspace`: uic_att_step_workspace_bytes(...) bytes, 16-byte aligned, zeroed ONCE by the caller (the kernels leave
 * their arrival counters / publication flags at zero).  Two launch plans share it: with at least half as many
 * (image, beam group) jobs as CTA slots the flat list of 16-region batches is cut into equal contiguous ranges, one per
 * CTA, and a job cut by a range boundary is finished by the CTA that owns its first batch from the partial records the
 * later CTAs publish there (all CTAs of the grid are co-resident; a wait that cannot be satisfied traps after ~1 s
 * instead of hanging); smaller batches cut every job into equal segments merged by the last CTA to arrive. */
int uic_att_step_fwd(const float* att_h, int64_t ld_att_h, const void* p_att_f16, const void* att_bf16,
                     const float* w_alpha, const float* att_masks, void* ctx_bf16, int64_t ld_ctx_bf16, float* ctx_f32,
                     int64_t ld_ctx_f32, float* alpha, void* workspace, int64_t workspace_bytes, int n_img, int beams, int L,
                     int A, int H, void* stream);
int64_t uic_att_step_workspace_bytes(int n_img, int beams, int L, int A, int H);

/* ---- LSTM pointwise ------------------------------------------------------------------------- */
/* Att2in2 maxout cell (models/AttModel.py:584-601): sums = i2h(xt)+h2h(h) (rows x 5H, pitch ld_sums),
 * a2c = a2c(ctx) (rows x 2H, pitch ld_a2c); i,f,o = sigmoid(sums[:, :3H]);
 * g = max(sums[:,3H:4H]+a2c[:, :H], sums[:,4H:]+a2c[:,H:]); c = f*c_prev + i*g; h = o*tanh(c).
 * h is written as fp32 (optional) and as bf16 to up to two destinations (next step's GEMM operand
 * slots).  c_prev == NULL means zero state.  a2c == NULL: no separate context term -- Att2all2Core
 * (models/AttModel.py:618-654) adds a2h(ctx) to all five gate sums, which its GEMM accumulates into `sums`. */
int uic_lstm_maxout_fwd(const float* sums, int64_t ld_sums, const float* a2c, int64_t ld_a2c, const float* c_prev,
                        float* c_out, float* h_f32, void* h_bf16_a, int64_t ld_ha, void* h_bf16_b, int64_t ld_hb, int rows,
                        int H, void* stream);
/* torch.nn.LSTMCell pointwise, gate order i,f,g,o (models/AttModel.py:434,441): gates (rows x 4H). */
int uic_lstm_cell_fwd(const float* gates, int64_t ld_gates, const float* c_prev, float* c_out, float* h_f32,
                      void* h_bf16_a, int64_t ld_ha, void* h_bf16_b, int64_t ld_hb, int rows, int H, void* stream);
/* The same two cells with per-row ADDENDS to the gate pre-activations, for the decode loops:
 *   add_tok (V x n_gates*H fp32, pitch ld_add_tok) is gathered by tok[row] -- the input-word term
 *     i2h(relu(embed(it))) (models/AttModel.py:160,584) resp. the xt columns of att_lstm.weight_ih (:432-434) is a
 *     function of the token alone, so it is tabulated once per weight version and the per-step gate GEMM contracts
 *     over the recurrent columns only;
 *   add_grp (rows/group x n_gates*H fp32) is indexed by row / group -- a term that is constant per image (the fc
 *     columns of att_lstm.weight_ih) and shared by the `group` beams of the image.
 * Either table may be NULL.  Out-of-range tokens are clamped to [0, V). */
int uic_lstm_maxout_fwd_add(const float* sums, int64_t ld_sums, const float* a2c, int64_t ld_a2c, const float* c_prev,
                            float* c_out, float* h_f32, void* h_bf16_a, int64_t ld_ha, void* h_bf16_b, int64_t ld_hb, int rows,
                            int H, const float* add_tok, int64_t ld_add_tok, const int64_t* tok, int V, const float* add_grp,
                            int64_t ld_add_grp, int group, void* stream);
int uic_lstm_cell_fwd_add(const float* gates, int64_t ld_gates, const float* c_prev, float* c_out, float* h_f32,
                          void* h_bf16_a, int64_t ld_ha, void* h_bf16_b, int64_t ld_hb, int rows, int H, const float* add_tok,
                          int64_t ld_add_tok, const int64_t* tok, int V, const float* add_grp, int64_t ld_add_grp, int group,
                          void* stream);

/* ---- vocabulary softmax family --------------------------------------------------------------- */
/* out[r, :] = log_softmax(logits[r, :]) (F.log_softmax, models/AttModel.py:163). */
int uic_log_softmax_rows(const float* logits, int64_t ld_logits, float* out, int64_t ld_out, int rows, int V, void* stream);
/* Fused masked cross-entropy forward (misc/criterion.py:143-150 without materialising log-probs):
 * lse[r] = logsumexp(logits[r,:]); nll[r] = -(logits[r,target[r]] - lse[r]) * mask[r]. */
int uic_lse_xent_fwd(const float* logits, int64_t ld_logits, const int64_t* target, const float* mask, float* lse,
                     float* nll, int rows, int V, void* stream);

/* One greedy decoding step (models/AttModel.py:218-251, sample_max=1) for `rows` sequences:
 * log-softmax normaliser + argmax over V fused; writes seq[r,t] / seq_logprobs[r,t], updates the
 * unfinished flags, emits the next input token (0 for finished rows) and counts the rows still
 * unfinished into n_unfinished[t] so that later steps reproduce the reference's early `break`. */
int uic_greedy_step(const float* logits, int64_t ld_logits, int64_t* seq, float* seq_logprobs, uint8_t* unfinished,
                    int64_t* next_tok, int32_t* n_unfinished, int t, int seq_length, int rows, int V, int flags, void* stream);

/* Per-row top-k of the log-probabilities after the beam-search edits (models/CaptionModel.py:
 * 128-133,61): optional -inf on the previous token, -1000 on the UNK column; ties -> smaller id. */
int uic_row_topk(const float* logits, int64_t ld_logits, const int64_t* prev_tok, float* topk_val, int32_t* topk_idx,
                 int rows, int V, int k, int flags, void* stream);

/* Fused vocabulary projection + statistics for sampling: logits = h W^T + b are reduced inside the GEMM
 * epilogue and NEVER written to memory (replaces self.logit + F.log_softmax + torch.max / torch.sort of
 * models/AttModel.py:163,229 and models/CaptionModel.py:61,128-133 for inference).  For every row and
 * each of uic_logit_stats_parts(rows, V) column parts (two per GEMM tile; the tile width, 128 or 224 columns, is chosen from
 * the problem shape) it stores one entry of uic_logit_stats_entry_floats(kslots)
 * floats (2 + 2*kslots rounded up to a multiple of 4): max, sum exp(x - max), the kslots best keys (logit,
 * with -1000 on the UNK column V-1 when unk_suppress != 0 and -inf on the banned token) and their columns
 * (int bits, 0x7fffffff = empty slot), padding.  `stats` is (rows, parts, entry) fp32, 16-byte aligned.
 * banned_tok may be NULL; row r reads banned_tok[r * banned_stride].  kslots: 1, 3, 5 or 8.
 * temperature > 0 turns the best-key search into multinomial sampling (models/AttModel.py:231-239,
 * torch.multinomial(exp(logprobs / temperature), 1)) by Gumbel-max: keys become x / temperature + g with
 * g = -log(-log(u)), u a counter-based hash of (seed, step, row, column) (csrc/uic_vocab.cuh; restated in
 * oracle/decoder_oracle.py).  `seed` points to DEVICE memory (so a captured CUDA graph can be replayed with a new
 * seed).  temperature = 0: deterministic (seed may be NULL, step is ignored). */
int uic_logit_stats_parts(int rows, int V);
int uic_logit_stats_entry_floats(int kslots);
int uic_logit_stats(const void* h_bf16, int64_t ld_h, const void* w_logit_bf16, int64_t ld_w, const float* bias,
                    const int64_t* banned_tok, int64_t banned_stride, float* stats, int rows, int V, int H, int kslots, int unk_suppress,
                    float temperature, const uint64_t* seed, int step, void* stream);
/* Merges the parts of uic_logit_stats: same outputs as uic_row_topk (k <= kslots). */
int uic_beam_topk_merge(const float* stats, int parts, int kslots, float* topk_val, int32_t* topk_idx, int rows, int k,
                        void* stream);
/* Merges the parts of uic_logit_stats (kslots = 1): same bookkeeping as uic_greedy_step. */
int uic_greedy_merge(const float* stats, int parts, int64_t* seq, float* seq_logprobs, uint8_t* unfinished, int64_t* next_tok,
                     int32_t* n_unfinished, int t, int seq_length, int rows, void* stream);

/* The whole tail of a beam-search step in one launch: uic_beam_topk_merge + uic_beam_step and, when move_state != 0,
 * uic_beam_gather + uic_embed_rows for the next step (x_dst[r, xt_col0 : xt_col0 + E] = emb_table[next_tok[r]]).
 * Same results as the four separate calls; the candidate tables stay in shared memory.  beams <= kslots.
 * src_beams: beams per image in the SOURCE buffers (stats, x_src, c_src).  Normally == beams.  At the first step only
 * beam 0 of every image is read (rows = 1, models/CaptionModel.py:56), so the caller may run that step on ONE row per
 * image and pass src_beams = 1 with t = 0: stats / x_src / c_src then have n_img rows and every beam forks from it. */
int uic_beam_advance(const float* stats, int parts, int kslots, int32_t* beam_seq, float* beam_lp, float* beam_sum,
                     int32_t* done_seq, float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt,
                     int32_t* parent_row, int64_t* next_tok, int t, int seq_length, int n_img, int beams, int flags,
                     int move_state, const void* x_src, void* x_dst, int64_t ld_x, int col0_a, int ncol_a, int col0_b,
                     int ncol_b, const float* c_src, float* c_dst, int n_state, int H, const void* emb_table_bf16,
                     int64_t ld_table, int xt_col0, int E, int V, int src_beams, void* stream);
/* Greedy analogue: uic_greedy_merge and, when x_xt_bf16 != NULL, the next step's embedding rows
 * x_xt_bf16[r, 0:E] = emb_table[token r] (pitch ld_x), in one launch.  With temperature > 0 (same temperature and
 * seed as the uic_logit_stats call of step t) the winner's key is converted back to its unperturbed log-prob. */
int uic_greedy_advance(const float* stats, int parts, int64_t* seq, float* seq_logprobs, uint8_t* unfinished, int64_t* next_tok,
                       int32_t* n_unfinished, int t, int seq_length, int rows, const void* emb_table_bf16, int64_t ld_table,
                       void* x_xt_bf16, int64_t ld_x, int E, int V, float temperature, const uint64_t* seed, void* stream);

/* In-place nn.Dropout(p) in training mode (models/AttModel.py:73-84 embed / fc_embed / att_embed, :431,599 core output):
 * x[r, c] = keep ? x[r, c] / (1 - p) : 0 with keep = u(seed, site, row0 + r * row_stride, c) >= p, u the library's
 * counter-based uniform.  Calling it again on the gradient of the same tensor applies the same mask (the backward).
 * is_bf16: 1 = bf16 storage, 0 = fp32.  `seed`: device pointer. */
int uic_dropout(void* x, int is_bf16, int64_t ld, int64_t rows, int cols, float p, const uint64_t* seed, int site, int64_t row0,
                int64_t row_stride, void* stream);

/* Scheduled sampling (models/AttModel.py:130-143): input token of teacher-forced step t = with probability ss_prob per
 * row a draw from softmax(logits of step t-1) (statistics from uic_logit_stats(kslots = 1, temperature = 1, seed, step = t)),
 * else gt_tok[r * gt_stride]; written to tokens_out[r], its embedding row to x_xt_bf16[r, 0:E].  The per-row coin is
 * the same counter-based hash (column 0x7fffffff).  `seed`: device pointer. */
int uic_ss_advance(const float* stats, int parts, const int64_t* gt_tok, int64_t gt_stride, float ss_prob, const uint64_t* seed,
                   int t, int64_t* tokens_out, int rows, const void* emb_table_bf16, int64_t ld_table, void* x_xt_bf16, int64_t ld_x,
                   int E, int V, void* stream);

/* One beam-search bookkeeping step for all images at once (models/CaptionModel.py:48-97,155-172):
 * merges the beams x k candidates of each image (c-major, q-minor stable order), forks the
 * sequence tables, records finished hypotheses (token 0 or last step) into the sorted done lists,
 * and emits parent row + next token for every beam.  topk_unaug (may be NULL = topk_val): the candidates'
 * log-probs before the diversity penalty; they are what the log-prob tables store (:38,86) while topk_val ranks. */
int uic_beam_step(const float* topk_val, const int32_t* topk_idx, const float* topk_unaug, int32_t* beam_seq, float* beam_lp, float* beam_sum,
                  int32_t* done_seq, float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt,
                  int32_t* parent_row, int64_t* next_tok, int t, int seq_length, int n_img, int beams, int flags,
                  void* stream);
/* Diverse beam search (group_size > 1, models/CaptionModel.py:36-45 add_diversity): candidates of group `group` at
 * its local step t.  cand_val / cand_idx: the n_cand >= beams + group * beams best edited log-probs per row
 * (uic_row_topk with k = n_cand; a penalty only lowers values, so they contain the penalised top-`beams`).  Each
 * candidate loses diversity_lambda once per occurrence of its token among beam_seq[g][image][*][t] of the groups
 * g < group (tables laid out [group][image][beam][seq_length]).  Out: the `beams` best by (penalised value,
 * smaller column) with the penalised (ranking) and unpenalised (stored) values, ready for uic_beam_step. */
int uic_diverse_select(const float* cand_val, const int32_t* cand_idx, int n_cand, const int32_t* beam_seq, int group, int n_img,
                       int beams, int seq_length, int t, float diversity_lambda, float* topk_val, float* topk_unaug,
                       int32_t* topk_idx, void* stream);
/* Re-order recurrent state by parent beam (CaptionModel.py:89-91): for two column ranges of the
 * bf16 activation matrix and `n_state` fp32 state matrices (rows x H each, contiguous). */
int uic_beam_gather(const int32_t* parent_row, const void* x_src, void* x_dst, int64_t ld_x, int col0_a, int ncol_a,
                    int col0_b, int ncol_b, const float* c_src, float* c_dst, int n_state, int rows, int H, void* stream);

/* ---- backward (the reference gets these from torch autograd, trainer.py:173) --------------------
 * Gradients that feed a tcgen05 dgrad/wgrad GEMM are emitted directly as bf16 operands. */

/* nn.LSTMCell backward.  dh = dh0 + dh1 + dh2 (each optional, own pitch); gates are the saved
 * pre-activations (rows x 4H); writes d gates (bf16) and dc_prev. */
int uic_lstm_cell_bwd(const float* gates, int64_t ld_gates, const float* c_prev, const float* c, const float* dh0, int64_t ld0,
                      const float* dh1, int64_t ld1, const float* dh2, int64_t ld2, const float* dc_next, void* dgates_bf16,
                      int64_t ld_dg, float* dc_prev, int rows, int H, void* stream);
/* Att2in2 maxout cell backward (models/AttModel.py:585-597): writes d sums (rows x 5H, bf16),
 * d a2c (rows x 2H, bf16; a2c and da2c_bf16 both NULL for the att2all2 cell) and dc_prev. */
int uic_lstm_maxout_bwd(const float* sums, int64_t ld_sums, const float* a2c, int64_t ld_a2c, const float* c_prev, const float* c,
                        const float* dh0, int64_t ld0, const float* dh1, int64_t ld1, const float* dc_next, void* dsums_bf16,
                        int64_t ld_ds, void* da2c_bf16, int64_t ld_da, float* dc_prev, int rows, int H, void* stream);
/* Attention step backward for rows == images (teacher forcing): from d ctx and the saved alpha,
 * de[r,l] = alpha (d alpha - sum alpha d alpha) and d att_h[r,a] = w_a sum_l de (1 - tanh^2). */
int uic_att_step_bwd(const float* dctx, int64_t ld_dctx, const float* alpha, const void* p_att_bf16, const void* att_bf16,
                     const float* att_h, int64_t ld_att_h, const float* w_alpha, float* de, void* datt_h_bf16, int64_t ld_dah,
                     int rows, int L, int A, int H, void* stream);
/* Deferred gradients of the feature tiles over all T steps at once:
 * d att (B,L,H) fp32, d p_att (B,L,A) bf16, and accumulated into dw_alpha (2A floats, zeroed by the
 * caller): [0,A) d w_alpha, [A,2A) the fp32 column sum of d p_att (= d bias of ctx2att).  The
 * per-step vectors are addressed as base + t*stride_t + b*ld. */
int uic_att_tiles_bwd(const float* de_all, const float* alpha_all, const float* dctx_all, int64_t dctx_stride_t, int64_t ld_dctx,
                      const float* att_h_all, int64_t ah_stride_t, int64_t ld_ah, const void* p_att_bf16, const float* w_alpha,
                      float* datt, void* dp_att_bf16, float* dw_alpha, int T, int B, int L, int A, int H, void* stream);
/* d logits = (softmax(logits) - onehot(target)) * mask * inv_norm[0] * grad_scale, bf16, pitch ld_d
 * (>= V, padding columns are zeroed).  Backward of uic_lse_xent_fwd / misc/criterion.py:143-150. */
int uic_lse_xent_bwd(const float* logits, int64_t ld, const float* lse, const int64_t* target, const float* mask,
                     const float* inv_norm, float grad_scale, void* dlogits_bf16, int64_t ld_d, int rows, int V, void* stream);
/* d logits = d lp - exp(lp) * rowsum(d lp): backward of uic_log_softmax_rows for the dense API path. */
int uic_log_softmax_bwd(const float* dlp, int64_t ld_dlp, const float* lp, int64_t ld_lp, void* dlogits_bf16, int64_t ld_d,
                        int rows, int V, void* stream);
/* Batch statistics of nn.BatchNorm1d(att_feat_size) in att_embed (use_bn, models/AttModel.py:79-84) over the PACKED
 * regions that pack_wrapper (:44-53) feeds it: sum[c] += sum x[r,c], sumsq[c] += sum x[r,c]^2 over the rows
 * (image i, region l < lens[i]) of x (n_img * L rows, pitch ld; bf16 or fp32).  lens == NULL: every row.  fp64 sums. */
int uic_col_moments(const void* x, int is_bf16, int64_t ld, const int32_t* lens, int n_img, int L, int cols, double* sum,
                    double* sumsq, void* stream);
/* out[c] += sum_r x[r,c] (bias gradients); x is bf16 (is_bf16 != 0) or fp32. */
int uic_col_sum(const void* x, int is_bf16, int64_t ld, float* out, int rows, int cols, void* stream);
/* dEmb[tok[r], :] += dxt[r, :] where ReLU(Emb) was active (backward of uic_embed_rows). */
int uic_embed_bwd(const float* dxt, int64_t ld, const int64_t* tok, const void* table_relu_bf16, float* demb, int64_t rows, int E,
                  int V, void* stream);
/* x[i] = y_bf16[i] > 0 ? x[i] : 0 in place, and out_bf16[i] = bf16(x[i]): ReLU backward fused with
 * the operand cast (the fp32 copy is what the bias gradient is summed from). */
int uic_relu_bwd_cast(float* x, const void* y_bf16, void* out_bf16, int64_t n, void* stream);
/* dst[b, j] = sum_t src[t*stride_t + b*ld + col0 + j]. */
int uic_reduce_time(const float* src, int64_t stride_t, int64_t ld, int col0, float* dst, int T, int rows, int n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UIC_B200_H_ */
