// Internal declarations shared by the kernel translation units and the C-ABI layer (api.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/uic_b200.h"

namespace uic {

enum { GEMM_IMPL_TCGEN05 = 0, GEMM_IMPL_SIMT = 1 };

int set_error(int code, const char* fmt, ...);
// Launch accounting: every kernel launch is bracketed by launch_begin / launch_end.  They count
// launches (uic_launch_count) and, when profiling is enabled (uic_profile_enable), record a CUDA
// event pair on the launching stream so bench.py can report per-kernel device time live.
void launch_begin(const char* name, cudaStream_t stream);
void launch_end(cudaStream_t stream);
int gemm_impl();
long long* gemm_trace_buffer();

// Cached cuTensorMapEncodeTiled for a row-major bf16 matrix [rows, cols] with pitch ld (elements),
// box = box_rows x box_cols, 128-byte swizzle, zero fill out of bounds.
int get_tensor_map_bf16(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_rows,
                        int box_cols);
// [rows, cols] (cols % 64 == 0) as [cols/64 slabs][rows][64]: one box = box_rows rows of all slabs (see api.cu).
int get_tensor_map_bf16_slabs(CUtensorMap* out, const void* base, long long rows, long long cols, int box_rows);

#define UIC_CUDA_OK(expr)                                                                                   \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess)                                                                                  \
      return ::uic::set_error(UIC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Programmatic dependent launch (PDL): a kernel launched through launch_pdl may be scheduled while its stream
// predecessor is still draining, so its launch latency and set-up (barrier init, TMEM allocation, descriptor
// prefetch) overlap the predecessor's tail.  Contract for such a kernel: call pdl_launch_dependents() first and
// pdl_wait() before the first access to any global memory a predecessor writes, or that it writes itself.
// pdl_wait() returns when every predecessor grid has completed and flushed.  UIC_PDL=0 turns the attribute off.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// kernels.  Each returns 0 or a negative UIC_ERR_* (after set_error).
int gemm_bf16(const void* A, long long lda, const void* B, long long ldb, float* c_f32, long long ldc, void* c_bf16,
              long long ldcb, const float* bias, int M, int N, int K, int flags, cudaStream_t stream, int exp_col0 = 0,
              float exp_scale = 0.0f, const float* post_scale = nullptr, const float* post_shift = nullptr);
int logit_stats_parts(int M, int N);
int logit_stats_entry_floats(int kslots);
int logit_stats(const void* A, long long lda, const void* B, long long ldb, const float* bias, const long long* banned,
                long long banned_stride, float* stats, int M, int N, int K, int kslots, int unk_suppress, float temperature,
                const unsigned long long* seed, int step, cudaStream_t stream);
int beam_topk_merge(const float* stats, int parts, int kslots, float* topk_val, int32_t* topk_idx, int rows, int k,
                    cudaStream_t stream);
int greedy_merge(const float* stats, int parts, int64_t* seq, float* seq_lp, uint8_t* unfinished, int64_t* next_tok,
                 int32_t* n_unfinished, int t, int seq_length, int rows, cudaStream_t stream);
int cast_f32_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, long long cols, int relu,
                  cudaStream_t stream);
int zero_padded_rows(void* x, const float* masks, int n_img, int L, int H, cudaStream_t stream);
int embed_rows(const void* table, long long ld_table, const int64_t* tok, void* out, long long ld_out, int rows, int E, int V,
               cudaStream_t stream);
int att_step_fwd(const float* att_h, long long ld_att_h, const void* p_att, const void* att, const float* w_alpha,
                 const float* masks, void* ctx_bf16, long long ld_ctx_bf16, float* ctx_f32, long long ld_ctx_f32, float* alpha,
                 void* workspace, long long workspace_bytes, int n_img, int beams, int L, int A, int H, cudaStream_t stream);
long long att_step_workspace_bytes(int n_img, int beams, int L, int A, int H);
int lstm_maxout_fwd(const float* sums, long long ld_sums, const float* a2c, long long ld_a2c, const float* c_prev, float* c_out,
                    float* h_f32, void* h_a, long long ld_ha, void* h_b, long long ld_hb, int rows, int H, cudaStream_t stream,
                    const float* add_tok = nullptr, long long ld_add_tok = 0, const long long* tok = nullptr, int V = 0,
                    const float* add_grp = nullptr, long long ld_add_grp = 0, int group = 1);
int lstm_cell_fwd(const float* gates, long long ld_gates, const float* c_prev, float* c_out, float* h_f32, void* h_a,
                  long long ld_ha, void* h_b, long long ld_hb, int rows, int H, cudaStream_t stream, const float* add_tok = nullptr,
                  long long ld_add_tok = 0, const long long* tok = nullptr, int V = 0, const float* add_grp = nullptr,
                  long long ld_add_grp = 0, int group = 1);
int log_softmax_rows(const float* logits, long long ld, float* out, long long ld_out, int rows, int V, cudaStream_t stream);
int lse_xent_fwd(const float* logits, long long ld, const int64_t* target, const float* mask, float* lse, float* nll, int rows,
                 int V, cudaStream_t stream);
int greedy_step(const float* logits, long long ld, int64_t* seq, float* seq_lp, uint8_t* unfinished, int64_t* next_tok,
                int32_t* n_unfinished, int t, int seq_length, int rows, int V, int flags, cudaStream_t stream);
int row_topk(const float* logits, long long ld, const int64_t* prev_tok, float* topk_val, int32_t* topk_idx, int rows, int V,
             int k, int flags, cudaStream_t stream);
int diverse_select(const float* cand_val, const int32_t* cand_idx, int kp, const int32_t* beam_seq, int group, int n_img, int bdash,
                   int T, int lt, float lambda, float* topk_val, float* topk_unaug, int32_t* topk_idx, cudaStream_t stream);
int beam_step(const float* topk_val, const int32_t* topk_idx, const float* topk_unaug, int32_t* beam_seq, float* beam_lp, float* beam_sum,
              int32_t* done_seq, float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt, int32_t* parent_row,
              int64_t* next_tok, int t, int seq_length, int n_img, int beams, int flags, cudaStream_t stream);
int beam_advance(const float* stats, int parts, int kslots, int32_t* beam_seq, float* beam_lp, float* beam_sum, int32_t* done_seq,
                 float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt, int32_t* parent_row, int64_t* next_tok, int t,
                 int seq_length, int n_img, int beams, int flags, int move_state, const void* x_src, void* x_dst, long long ld_x,
                 int col0_a, int ncol_a, int col0_b, int ncol_b, const float* c_src, float* c_dst, int n_state, int H,
                 const void* table, long long ld_table, int xt_col0, int E, int V, int src_beams, cudaStream_t stream);
int greedy_advance(const float* stats, int parts, int64_t* seq, float* seq_lp, uint8_t* unfinished, int64_t* next_tok,
                   int32_t* n_unfinished, int t, int seq_length, int rows, const void* table, long long ld_table, void* x_xt,
                   long long ld_x, int E, int V, float temperature, const unsigned long long* seed, cudaStream_t stream);
int ss_advance(const float* stats, int parts, const int64_t* gt_tok, long long gt_stride, float ss_prob,
               const unsigned long long* seed, int t, int64_t* tokens_out, int rows, const void* table, long long ld_table,
               void* x_xt, long long ld_x, int E, int V, cudaStream_t stream);
int dropout(void* x, int is_bf16, long long ld, long long rows, int cols, float p, const unsigned long long* seed, int site,
            long long row0, long long row_stride, cudaStream_t stream);
int beam_gather(const int32_t* parent_row, const void* x_src, void* x_dst, long long ld_x, int col0_a, int ncol_a, int col0_b,
                int ncol_b, const float* c_src, float* c_dst, int n_state, int rows, int H, cudaStream_t stream);

}  // namespace uic

namespace uic {
// backward kernels (backward.cu)
int lstm_cell_bwd(const float* gates, long long ld_gates, const float* c_prev, const float* c, const float* dh0, long long ld0,
                  const float* dh1, long long ld1, const float* dh2, long long ld2, const float* dc_next, void* dgates,
                  long long ld_dg, float* dc_prev, int rows, int H, cudaStream_t stream);
int lstm_maxout_bwd(const float* sums, long long ld_sums, const float* a2c, long long ld_a2c, const float* c_prev, const float* c,
                    const float* dh0, long long ld0, const float* dh1, long long ld1, const float* dc_next, void* dsums,
                    long long ld_ds, void* da2c, long long ld_da, float* dc_prev, int rows, int H, cudaStream_t stream);
int att_step_bwd(const float* dctx, long long ld_dctx, const float* alpha, const void* p_att, const void* att, const float* att_h,
                 long long ld_att_h, const float* w_alpha, float* de, void* datt_h, long long ld_dah, int rows, int L, int A,
                 int H, cudaStream_t stream);
int att_tiles_bwd(const float* de_all, const float* alpha_all, const float* dctx_all, long long dctx_stride_t, long long ld_dctx,
                  const float* att_h_all, long long ah_stride_t, long long ld_ah, const void* p_att, const float* w_alpha,
                  float* datt, void* dp_att, float* dw_alpha, int T, int B, int L, int A, int H, cudaStream_t stream);
int lse_xent_bwd(const float* logits, long long ld, const float* lse, const int64_t* target, const float* mask,
                 const float* inv_norm, float grad_scale, void* dlogits, long long ld_d, int rows, int V, cudaStream_t stream);
int log_softmax_bwd(const float* dlp, long long ld_dlp, const float* lp, long long ld_lp, void* dlogits, long long ld_d, int rows,
                    int V, cudaStream_t stream);
int col_moments(const void* x, int is_bf16, long long ld, const int32_t* lens, int n_img, int L, int cols, double* sum, double* sumsq,
                cudaStream_t stream);
int col_sum(const void* x, int is_bf16, long long ld, float* out, int rows, int cols, cudaStream_t stream);
int embed_bwd(const float* dxt, long long ld, const int64_t* tok, const void* table_relu, float* demb, long long rows, int E, int V,
              cudaStream_t stream);
int relu_bwd_cast(float* x, const void* y, void* out, long long n, cudaStream_t stream);
int reduce_time(const float* src, long long stride_t, long long ld, int col0, float* dst, int T, int rows, int n,
                cudaStream_t stream);
}  // namespace uic
