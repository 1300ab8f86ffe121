// C-ABI layer of libuic_b200.so: argument validation, error strings, launch accounting and the
// TMA descriptor cache.  Signatures and contracts are documented in include/uic_b200.h.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "uic_internal.h"

namespace uic {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_gemm_impl{-1};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---- launch accounting / live per-kernel timing ----------------------------------------------
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static thread_local int g_prof_open = -1;

void launch_begin(const char* name, cudaStream_t stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{name, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, stream);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back(r);
  g_prof_open = static_cast<int>(g_prof.size()) - 1;
}

void launch_end(cudaStream_t stream) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_prof_open < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (g_prof_open < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[g_prof_open].b, stream);
  g_prof_open = -1;
}

static std::atomic<long long*> g_gemm_trace{nullptr};
long long* gemm_trace_buffer() { return g_gemm_trace.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UIC_PDL");
    v = (e != nullptr && strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v != 0;
}

int gemm_impl() {
  int v = g_gemm_impl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("UIC_GEMM");
    v = (e != nullptr && strcmp(e, "simt") == 0) ? GEMM_IMPL_SIMT : GEMM_IMPL_TCGEN05;
    g_gemm_impl.store(v);
  }
  return v;
}

// ---- TMA descriptor cache ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* base;
  long long rows, cols, ld;
  int box_rows, box_cols;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && box_cols == o.box_cols;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(static_cast<size_t>(k.rows));
    mix(static_cast<size_t>(k.cols));
    mix(static_cast<size_t>(k.ld));
    mix(static_cast<size_t>(k.box_rows) * 1315423911u + static_cast<size_t>(k.box_cols));
    return h;
  }
};

int get_tensor_map_bf16(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_rows,
                        int box_cols) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{base, rows, cols, ld, box_rows, box_cols};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return set_error(UIC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(UIC_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for base=%p rows=%lld cols=%lld ld=%lld box=%dx%d",
                     static_cast<int>(r), base, rows, cols, ld, box_rows, box_cols);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

// A row-major bf16 matrix [rows, cols] (cols % 64 == 0, contiguous rows) seen as [cols / 64 slabs][rows][64]:
// one box = box_rows rows of ALL slabs, stored slab by slab, each slab a 128-byte-swizzled tile like the 2-D
// boxes above.  One TMA instruction then stages box_rows full rows.
int get_tensor_map_bf16_slabs(CUtensorMap* out, const void* base, long long rows, long long cols, int box_rows) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{base, rows, cols, -1, box_rows, 64};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return set_error(UIC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if (cols % 64 != 0 || cols / 64 > 256) return set_error(UIC_ERR_SHAPE, "slab tensor map: cols=%lld must be a multiple of 64", cols);
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols / 64)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(cols) * 2, 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(cols / 64)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(UIC_ERR_CUDA, "cuTensorMapEncodeTiled (slabs) failed (%d) for base=%p rows=%lld cols=%lld box_rows=%d",
                     static_cast<int>(r), base, rows, cols, box_rows);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 65536) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

}  // namespace uic

// ---- extern "C" ---------------------------------------------------------------------------------
using namespace uic;
#define ST(s) static_cast<cudaStream_t>(s)
#define REQUIRE(cond, code, ...) \
  do {                           \
    if (!(cond)) return set_error(code, __VA_ARGS__); \
  } while (0)

extern "C" {

const char* uic_last_error(void) { return g_err; }
int uic_version(void) { return 100; }
int64_t uic_launch_count(void) { return g_launches.load(); }
int uic_set_gemm_impl(int impl) {
  REQUIRE(impl == 0 || impl == 1, UIC_ERR_ARG, "uic_set_gemm_impl: impl must be 0 or 1");
  g_gemm_impl.store(impl);
  return 0;
}
int uic_gemm_set_trace(void* device_buffer_1024_i64) {
  g_gemm_trace.store(static_cast<long long*>(device_buffer_1024_i64));
  return 0;
}

int uic_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  for (auto& r : g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  g_prof_on.store(on ? 1 : 0);
  return 0;
}

int64_t uic_profile_dump(char* out, int64_t cap) {
  REQUIRE(out != nullptr && cap > 0, UIC_ERR_ARG, "uic_profile_dump: no buffer");
  std::lock_guard<std::mutex> lock(g_prof_mu);
  struct Agg {
    long long n = 0;
    double ms = 0;
  };
  std::vector<std::pair<const char*, Agg>> agg;
  for (auto& r : g_prof) {
    float ms = 0.0f;
    if (cudaEventSynchronize(r.b) != cudaSuccess || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
    size_t i = 0;
    for (; i < agg.size(); ++i)
      if (strcmp(agg[i].first, r.name) == 0) break;
    if (i == agg.size()) agg.push_back({r.name, Agg()});
    agg[i].second.n += 1;
    agg[i].second.ms += ms;
  }
  int64_t used = 0;
  for (auto& a : agg) {
    int w = snprintf(out + used, static_cast<size_t>(cap - used), "%s %lld %.6f\n", a.first, a.second.n, a.second.ms);
    if (w < 0 || used + w >= cap) break;
    used += w;
  }
  out[used < cap ? used : cap - 1] = '\0';
  return static_cast<int64_t>(agg.size());
}

int uic_check_device(void) {
  int dev = 0;
  UIC_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  UIC_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  REQUIRE(p.major == 10, UIC_ERR_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, p.major, p.minor);
  return 0;
}

int uic_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, float* c_f32, int64_t ldc, void* c_bf16,
                  int64_t ldcb, const float* bias, int M, int N, int K, int flags, void* stream) {
  REQUIRE(A && B, UIC_ERR_ARG, "uic_gemm_bf16: null operand");
  return gemm_bf16(A, lda, B, ldb, c_f32, ldc, c_bf16, ldcb, bias, M, N, K, flags, ST(stream));
}

int uic_gemm_bf16_ex(const void* A, int64_t lda, const void* B, int64_t ldb, float* c_f32, int64_t ldc, void* c_16, int64_t ldc16,
                     const float* bias, int M, int N, int K, int flags, int exp_col0, float exp_scale, void* stream) {
  REQUIRE(A && B, UIC_ERR_ARG, "uic_gemm_bf16_ex: null operand");
  REQUIRE(exp_col0 >= 0, UIC_ERR_ARG, "uic_gemm_bf16_ex: exp_col0=%d", exp_col0);
  return gemm_bf16(A, lda, B, ldb, c_f32, ldc, c_16, ldc16, bias, M, N, K, flags, ST(stream), exp_col0, exp_scale);
}

int uic_gemm_bf16_affine(const void* A, int64_t lda, const void* B, int64_t ldb, float* c_f32, int64_t ldc, void* c_16, int64_t ldc16,
                         const float* bias, const float* post_scale, const float* post_shift, int M, int N, int K, int flags, void* stream) {
  REQUIRE(A && B && post_scale && post_shift, UIC_ERR_ARG, "uic_gemm_bf16_affine: null pointer");
  return gemm_bf16(A, lda, B, ldb, c_f32, ldc, c_16, ldc16, bias, M, N, K, flags, ST(stream), 0, 0.0f, post_scale, post_shift);
}

int uic_logit_stats_parts(int rows, int V) { return V > 0 ? logit_stats_parts(rows, V) : 0; }
int uic_logit_stats_entry_floats(int kslots) { return kslots > 0 ? logit_stats_entry_floats(kslots) : 0; }

int uic_logit_stats(const void* h_bf16, int64_t ld_h, const void* w_logit_bf16, int64_t ld_w, const float* bias,
                    const int64_t* banned_tok, int64_t banned_stride, float* stats, int rows, int V, int H, int kslots, int unk_suppress,
                    float temperature, const uint64_t* seed, int step, void* stream) {
  REQUIRE(h_bf16 && w_logit_bf16 && stats, UIC_ERR_ARG, "uic_logit_stats: null pointer");
  if (rows == 0) return 0;
  return logit_stats(h_bf16, ld_h, w_logit_bf16, ld_w, bias, reinterpret_cast<const long long*>(banned_tok), banned_stride, stats,
                     rows, V, H, kslots, unk_suppress, temperature,
                     reinterpret_cast<const unsigned long long*>(seed), step, ST(stream));
}

int uic_beam_topk_merge(const float* stats, int parts, int kslots, float* topk_val, int32_t* topk_idx, int rows, int k,
                        void* stream) {
  REQUIRE(stats && topk_val && topk_idx, UIC_ERR_ARG, "uic_beam_topk_merge: null pointer");
  REQUIRE(k >= 1 && parts >= 1, UIC_ERR_ARG, "uic_beam_topk_merge: k=%d parts=%d", k, parts);
  if (rows == 0) return 0;
  return beam_topk_merge(stats, parts, kslots, topk_val, topk_idx, rows, k, ST(stream));
}

int uic_greedy_merge(const float* stats, int parts, int64_t* seq, float* seq_logprobs, uint8_t* unfinished, int64_t* next_tok,
                     int32_t* n_unfinished, int t, int seq_length, int rows, void* stream) {
  REQUIRE(stats && seq && seq_logprobs && unfinished && next_tok && n_unfinished, UIC_ERR_ARG, "uic_greedy_merge: null pointer");
  REQUIRE(t >= 0 && t < seq_length && parts >= 1, UIC_ERR_ARG, "uic_greedy_merge: t=%d seq_length=%d parts=%d", t, seq_length, parts);
  if (rows == 0) return 0;
  return greedy_merge(stats, parts, seq, seq_logprobs, unfinished, next_tok, n_unfinished, t, seq_length, rows, ST(stream));
}

int uic_cast_f32_bf16(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int64_t cols, int relu,
                      void* stream) {
  REQUIRE(src && dst, UIC_ERR_ARG, "uic_cast_f32_bf16: null pointer");
  if (rows == 0 || cols == 0) return 0;
  return cast_f32_bf16(src, ld_src, dst, ld_dst, rows, cols, relu, ST(stream));
}

int uic_embed_rows(const void* table, int64_t ld_table, const int64_t* tok, void* out, int64_t ld_out, int rows, int E, int V,
                   void* stream) {
  REQUIRE(table && tok && out, UIC_ERR_ARG, "uic_embed_rows: null pointer");
  if (rows == 0) return 0;
  return embed_rows(table, ld_table, tok, out, ld_out, rows, E, V, ST(stream));
}

int uic_zero_padded_rows(void* x_bf16, const float* att_masks, int n_img, int L, int H, void* stream) {
  REQUIRE(x_bf16 && att_masks, UIC_ERR_ARG, "uic_zero_padded_rows: null pointer");
  if (n_img == 0 || L == 0) return 0;
  return zero_padded_rows(x_bf16, att_masks, n_img, L, H, ST(stream));
}

int uic_att_step_fwd(const float* att_h, int64_t ld_att_h, const void* p_att, const void* att, const float* w_alpha,
                     const float* masks, void* ctx_bf16, int64_t ld_ctx_bf16, float* ctx_f32, int64_t ld_ctx_f32, float* alpha,
                     void* workspace, int64_t workspace_bytes, int n_img, int beams, int L, int A, int H, void* stream) {
  REQUIRE(att_h && p_att && att && w_alpha, UIC_ERR_ARG, "uic_att_step_fwd: null input");
  REQUIRE(ctx_bf16 || ctx_f32, UIC_ERR_ARG, "uic_att_step_fwd: no output buffer");
  if (n_img == 0) return 0;
  return att_step_fwd(att_h, ld_att_h, p_att, att, w_alpha, masks, ctx_bf16, ld_ctx_bf16, ctx_f32, ld_ctx_f32, alpha, workspace,
                      workspace_bytes, n_img, beams, L, A, H, ST(stream));
}

int64_t uic_att_step_workspace_bytes(int n_img, int beams, int L, int A, int H) {
  if (n_img <= 0 || beams <= 0 || L <= 0 || H <= 0) return 0;
  return att_step_workspace_bytes(n_img, beams, L, A, H);
}

int uic_lstm_maxout_fwd(const float* sums, int64_t ld_sums, const float* a2c, int64_t ld_a2c, const float* c_prev, float* c_out,
                        float* h_f32, void* h_a, int64_t ld_ha, void* h_b, int64_t ld_hb, int rows, int H, void* stream) {
  REQUIRE(sums && c_out, UIC_ERR_ARG, "uic_lstm_maxout_fwd: null pointer");
  if (rows == 0) return 0;
  return lstm_maxout_fwd(sums, ld_sums, a2c, ld_a2c, c_prev, c_out, h_f32, h_a, ld_ha, h_b, ld_hb, rows, H, ST(stream));
}

int uic_lstm_cell_fwd(const float* gates, int64_t ld_gates, const float* c_prev, float* c_out, float* h_f32, void* h_a,
                      int64_t ld_ha, void* h_b, int64_t ld_hb, int rows, int H, void* stream) {
  REQUIRE(gates && c_out, UIC_ERR_ARG, "uic_lstm_cell_fwd: null pointer");
  if (rows == 0) return 0;
  return lstm_cell_fwd(gates, ld_gates, c_prev, c_out, h_f32, h_a, ld_ha, h_b, ld_hb, rows, H, ST(stream));
}

int uic_lstm_maxout_fwd_add(const float* sums, int64_t ld_sums, const float* a2c, int64_t ld_a2c, const float* c_prev, float* c_out,
                            float* h_f32, void* h_a, int64_t ld_ha, void* h_b, int64_t ld_hb, int rows, int H, const float* add_tok,
                            int64_t ld_add_tok, const int64_t* tok, int V, const float* add_grp, int64_t ld_add_grp, int group,
                            void* stream) {
  REQUIRE(sums && c_out, UIC_ERR_ARG, "uic_lstm_maxout_fwd_add: null pointer");
  if (rows == 0) return 0;
  return lstm_maxout_fwd(sums, ld_sums, a2c, ld_a2c, c_prev, c_out, h_f32, h_a, ld_ha, h_b, ld_hb, rows, H, ST(stream), add_tok, ld_add_tok,
                         reinterpret_cast<const long long*>(tok), V, add_grp, ld_add_grp, group);
}

int uic_lstm_cell_fwd_add(const float* gates, int64_t ld_gates, const float* c_prev, float* c_out, float* h_f32, void* h_a,
                          int64_t ld_ha, void* h_b, int64_t ld_hb, int rows, int H, const float* add_tok, int64_t ld_add_tok,
                          const int64_t* tok, int V, const float* add_grp, int64_t ld_add_grp, int group, void* stream) {
  REQUIRE(gates && c_out, UIC_ERR_ARG, "uic_lstm_cell_fwd_add: null pointer");
  if (rows == 0) return 0;
  return lstm_cell_fwd(gates, ld_gates, c_prev, c_out, h_f32, h_a, ld_ha, h_b, ld_hb, rows, H, ST(stream), add_tok, ld_add_tok,
                       reinterpret_cast<const long long*>(tok), V, add_grp, ld_add_grp, group);
}

int uic_log_softmax_rows(const float* logits, int64_t ld, float* out, int64_t ld_out, int rows, int V, void* stream) {
  REQUIRE(logits && out, UIC_ERR_ARG, "uic_log_softmax_rows: null pointer");
  if (rows == 0) return 0;
  REQUIRE(V > 0, UIC_ERR_SHAPE, "uic_log_softmax_rows: V=%d", V);
  return log_softmax_rows(logits, ld, out, ld_out, rows, V, ST(stream));
}

int uic_lse_xent_fwd(const float* logits, int64_t ld, const int64_t* target, const float* mask, float* lse, float* nll, int rows,
                     int V, void* stream) {
  REQUIRE(logits && target && mask && lse && nll, UIC_ERR_ARG, "uic_lse_xent_fwd: null pointer");
  if (rows == 0) return 0;
  REQUIRE(V > 0, UIC_ERR_SHAPE, "uic_lse_xent_fwd: V=%d", V);
  return lse_xent_fwd(logits, ld, target, mask, lse, nll, rows, V, ST(stream));
}

int uic_greedy_step(const float* logits, int64_t ld, int64_t* seq, float* seq_lp, uint8_t* unfinished, int64_t* next_tok,
                    int32_t* n_unfinished, int t, int seq_length, int rows, int V, int flags, void* stream) {
  REQUIRE(logits && seq && seq_lp && unfinished && next_tok && n_unfinished, UIC_ERR_ARG, "uic_greedy_step: null pointer");
  REQUIRE(t >= 0 && t < seq_length, UIC_ERR_ARG, "uic_greedy_step: t=%d outside [0,%d)", t, seq_length);
  if (rows == 0) return 0;
  return greedy_step(logits, ld, seq, seq_lp, unfinished, next_tok, n_unfinished, t, seq_length, rows, V, flags, ST(stream));
}

int uic_row_topk(const float* logits, int64_t ld, const int64_t* prev_tok, float* topk_val, int32_t* topk_idx, int rows, int V,
                 int k, int flags, void* stream) {
  REQUIRE(logits && topk_val && topk_idx, UIC_ERR_ARG, "uic_row_topk: null pointer");
  REQUIRE(k >= 1 && k <= 16 && k <= V, UIC_ERR_SHAPE, "uic_row_topk: k=%d must be in [1,16] and <= V=%d", k, V);
  REQUIRE(!(flags & UIC_SAMPLE_DECODING_CONSTRAINT) || prev_tok, UIC_ERR_ARG, "uic_row_topk: constraint needs prev_tok");
  if (rows == 0) return 0;
  return row_topk(logits, ld, prev_tok, topk_val, topk_idx, rows, V, k, flags, ST(stream));
}

int uic_beam_advance(const float* stats, int parts, int kslots, int32_t* beam_seq, float* beam_lp, float* beam_sum,
                     int32_t* done_seq, float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt,
                     int32_t* parent_row, int64_t* next_tok, int t, int seq_length, int n_img, int beams, int flags,
                     int move_state, const void* x_src, void* x_dst, int64_t ld_x, int col0_a, int ncol_a, int col0_b,
                     int ncol_b, const float* c_src, float* c_dst, int n_state, int H, const void* emb_table_bf16,
                     int64_t ld_table, int xt_col0, int E, int V, int src_beams, void* stream) {
  REQUIRE(stats && beam_seq && beam_lp && beam_sum && done_seq && done_lp && done_p && done_unaug && done_cnt && parent_row && next_tok,
          UIC_ERR_ARG, "uic_beam_advance: null pointer");
  REQUIRE(parts > 0 && beams > 0 && seq_length > 0 && t >= 0 && t < seq_length, UIC_ERR_SHAPE,
          "uic_beam_advance: parts=%d beams=%d t=%d seq_length=%d", parts, beams, t, seq_length);
  REQUIRE((reinterpret_cast<uintptr_t>(stats) & 15) == 0, UIC_ERR_ALIGN, "uic_beam_advance: stats must be 16-byte aligned");
  if (move_state) {
    REQUIRE(x_src && x_dst && emb_table_bf16 && (n_state == 0 || (c_src && c_dst)), UIC_ERR_ARG, "uic_beam_advance: null state buffer");
    REQUIRE(x_src != x_dst && c_src != c_dst, UIC_ERR_ARG, "uic_beam_advance: the state must move between two different buffers");
    // (E == 0: no embedding rows are written -- the caller gathers the input-word term from a gate table by next_tok)
    REQUIRE(E >= 0 && V > 0 && H > 0 && ncol_a >= 0 && ncol_b >= 0, UIC_ERR_SHAPE, "uic_beam_advance: bad state shape");
  }
  if (n_img == 0) return 0;
  return beam_advance(stats, parts, kslots, beam_seq, beam_lp, beam_sum, done_seq, done_lp, done_p, done_unaug, done_cnt, parent_row,
                      next_tok, t, seq_length, n_img, beams, flags, move_state, x_src, x_dst, ld_x, col0_a, ncol_a, col0_b, ncol_b,
                      c_src, c_dst, n_state, H, emb_table_bf16, ld_table, xt_col0, E, V, src_beams, ST(stream));
}

int uic_greedy_advance(const float* stats, int parts, int64_t* seq, float* seq_logprobs, uint8_t* unfinished, int64_t* next_tok,
                       int32_t* n_unfinished, int t, int seq_length, int rows, const void* emb_table_bf16, int64_t ld_table,
                       void* x_xt_bf16, int64_t ld_x, int E, int V, float temperature, const uint64_t* seed, void* stream) {
  REQUIRE(stats && seq && seq_logprobs && unfinished && next_tok && n_unfinished, UIC_ERR_ARG, "uic_greedy_advance: null pointer");
  REQUIRE(parts > 0 && seq_length > 0 && t >= 0 && t < seq_length, UIC_ERR_SHAPE, "uic_greedy_advance: parts=%d t=%d seq_length=%d",
          parts, t, seq_length);
  REQUIRE((reinterpret_cast<uintptr_t>(stats) & 15) == 0, UIC_ERR_ALIGN, "uic_greedy_advance: stats must be 16-byte aligned");
  REQUIRE(x_xt_bf16 == nullptr || (emb_table_bf16 && E > 0 && V > 0), UIC_ERR_ARG, "uic_greedy_advance: embedding table missing");
  if (rows == 0) return 0;
  return greedy_advance(stats, parts, seq, seq_logprobs, unfinished, next_tok, n_unfinished, t, seq_length, rows, emb_table_bf16,
                        ld_table, x_xt_bf16, ld_x, E, V, temperature, reinterpret_cast<const unsigned long long*>(seed), ST(stream));
}

int uic_ss_advance(const float* stats, int parts, const int64_t* gt_tok, int64_t gt_stride, float ss_prob, const uint64_t* seed,
                   int t, int64_t* tokens_out, int rows, const void* emb_table_bf16, int64_t ld_table, void* x_xt_bf16, int64_t ld_x,
                   int E, int V, void* stream) {
  REQUIRE(stats && gt_tok && seed && tokens_out && emb_table_bf16 && x_xt_bf16, UIC_ERR_ARG, "uic_ss_advance: null pointer");
  REQUIRE(parts > 0 && E > 0 && V > 0 && t >= 0, UIC_ERR_SHAPE, "uic_ss_advance: parts=%d E=%d V=%d t=%d", parts, E, V, t);
  REQUIRE((reinterpret_cast<uintptr_t>(stats) & 15) == 0, UIC_ERR_ALIGN, "uic_ss_advance: stats must be 16-byte aligned");
  if (rows == 0) return 0;
  return ss_advance(stats, parts, gt_tok, gt_stride, ss_prob, reinterpret_cast<const unsigned long long*>(seed), t, tokens_out, rows,
                    emb_table_bf16, ld_table, x_xt_bf16, ld_x, E, V, ST(stream));
}

int uic_dropout(void* x, int is_bf16, int64_t ld, int64_t rows, int cols, float p, const uint64_t* seed, int site, int64_t row0,
                int64_t row_stride, void* stream) {
  REQUIRE(x && seed, UIC_ERR_ARG, "uic_dropout: null pointer");
  REQUIRE(p >= 0.0f && p < 1.0f, UIC_ERR_ARG, "uic_dropout: p=%f must be in [0, 1)", p);
  REQUIRE(cols > 0 && ld >= cols && rows >= 0, UIC_ERR_SHAPE, "uic_dropout: rows=%lld cols=%d ld=%lld", static_cast<long long>(rows), cols,
          static_cast<long long>(ld));
  if (rows == 0) return 0;
  return dropout(x, is_bf16, ld, rows, cols, p, reinterpret_cast<const unsigned long long*>(seed), site, row0, row_stride, ST(stream));
}

int uic_diverse_select(const float* cand_val, const int32_t* cand_idx, int n_cand, const int32_t* beam_seq, int group, int n_img,
                       int beams, int seq_length, int t, float diversity_lambda, float* topk_val, float* topk_unaug,
                       int32_t* topk_idx, void* stream) {
  REQUIRE(cand_val && cand_idx && topk_val && topk_unaug && topk_idx, UIC_ERR_ARG, "uic_diverse_select: null pointer");
  REQUIRE(group == 0 || beam_seq, UIC_ERR_ARG, "uic_diverse_select: group %d needs the sequence tables of the earlier groups", group);
  REQUIRE(beams >= 1 && beams <= 16 && n_cand >= beams && n_cand <= 16, UIC_ERR_SHAPE,
          "uic_diverse_select: beams=%d candidates=%d (1 <= beams <= candidates <= 16)", beams, n_cand);
  REQUIRE(group >= 0 && t >= 0 && t < seq_length, UIC_ERR_SHAPE, "uic_diverse_select: group=%d t=%d seq_length=%d", group, t, seq_length);
  if (n_img == 0) return 0;
  return diverse_select(cand_val, cand_idx, n_cand, beam_seq, group, n_img, beams, seq_length, t, diversity_lambda, topk_val, topk_unaug,
                        topk_idx, ST(stream));
}

int uic_beam_step(const float* topk_val, const int32_t* topk_idx, const float* topk_unaug, int32_t* beam_seq, float* beam_lp, float* beam_sum,
                  int32_t* done_seq, float* done_lp, double* done_p, float* done_unaug, int32_t* done_cnt, int32_t* parent_row,
                  int64_t* next_tok, int t, int seq_length, int n_img, int beams, int flags, void* stream) {
  REQUIRE(topk_val && topk_idx && beam_seq && beam_lp && beam_sum && done_seq && done_lp && done_p && done_unaug && done_cnt &&
              parent_row && next_tok,
          UIC_ERR_ARG, "uic_beam_step: null pointer");
  REQUIRE(beams >= 1 && beams <= 16, UIC_ERR_SHAPE, "uic_beam_step: beams=%d must be in [1,16]", beams);
  REQUIRE(t >= 0 && t < seq_length && seq_length <= 64, UIC_ERR_SHAPE, "uic_beam_step: t=%d seq_length=%d (max 64)", t, seq_length);
  if (n_img == 0) return 0;
  return beam_step(topk_val, topk_idx, topk_unaug, beam_seq, beam_lp, beam_sum, done_seq, done_lp, done_p, done_unaug, done_cnt, parent_row,
                   next_tok, t, seq_length, n_img, beams, flags, ST(stream));
}

int uic_beam_gather(const int32_t* parent_row, const void* x_src, void* x_dst, int64_t ld_x, int col0_a, int ncol_a, int col0_b,
                    int ncol_b, const float* c_src, float* c_dst, int n_state, int rows, int H, void* stream) {
  REQUIRE(parent_row && x_src && x_dst, UIC_ERR_ARG, "uic_beam_gather: null pointer");
  REQUIRE(x_src != x_dst && (c_src == nullptr || c_src != c_dst), UIC_ERR_ARG, "uic_beam_gather: must not run in place");
  if (rows == 0) return 0;
  return beam_gather(parent_row, x_src, x_dst, ld_x, col0_a, ncol_a, col0_b, ncol_b, c_src, c_dst, n_state, rows, H, ST(stream));
}

}  // extern "C"

// ---- backward entry points ------------------------------------------------------------------------
extern "C" {

int uic_lstm_cell_bwd(const float* gates, int64_t ld_gates, const float* c_prev, const float* c, const float* dh0, int64_t ld0,
                      const float* dh1, int64_t ld1, const float* dh2, int64_t ld2, const float* dc_next, void* dgates_bf16,
                      int64_t ld_dg, float* dc_prev, int rows, int H, void* stream) {
  REQUIRE(gates && c && dgates_bf16 && dc_prev, UIC_ERR_ARG, "uic_lstm_cell_bwd: null pointer");
  if (rows == 0) return 0;
  return lstm_cell_bwd(gates, ld_gates, c_prev, c, dh0, ld0, dh1, ld1, dh2, ld2, dc_next, dgates_bf16, ld_dg, dc_prev, rows, H,
                       ST(stream));
}

int uic_lstm_maxout_bwd(const float* sums, int64_t ld_sums, const float* a2c, int64_t ld_a2c, const float* c_prev, const float* c,
                        const float* dh0, int64_t ld0, const float* dh1, int64_t ld1, const float* dc_next, void* dsums_bf16,
                        int64_t ld_ds, void* da2c_bf16, int64_t ld_da, float* dc_prev, int rows, int H, void* stream) {
  REQUIRE(sums && c && dsums_bf16 && dc_prev && (a2c != nullptr) == (da2c_bf16 != nullptr), UIC_ERR_ARG,
          "uic_lstm_maxout_bwd: null pointer (a2c and da2c are given together or not at all)");
  if (rows == 0) return 0;
  return lstm_maxout_bwd(sums, ld_sums, a2c, ld_a2c, c_prev, c, dh0, ld0, dh1, ld1, dc_next, dsums_bf16, ld_ds, da2c_bf16, ld_da,
                         dc_prev, rows, H, ST(stream));
}

int uic_att_step_bwd(const float* dctx, int64_t ld_dctx, const float* alpha, const void* p_att_bf16, const void* att_bf16,
                     const float* att_h, int64_t ld_att_h, const float* w_alpha, float* de, void* datt_h_bf16, int64_t ld_dah,
                     int rows, int L, int A, int H, void* stream) {
  REQUIRE(dctx && alpha && p_att_bf16 && att_bf16 && att_h && w_alpha && de && datt_h_bf16, UIC_ERR_ARG,
          "uic_att_step_bwd: null pointer");
  if (rows == 0) return 0;
  return att_step_bwd(dctx, ld_dctx, alpha, p_att_bf16, att_bf16, att_h, ld_att_h, w_alpha, de, datt_h_bf16, ld_dah, rows, L, A, H,
                      ST(stream));
}

int uic_att_tiles_bwd(const float* de_all, const float* alpha_all, const float* dctx_all, int64_t dctx_stride_t, int64_t ld_dctx,
                      const float* att_h_all, int64_t ah_stride_t, int64_t ld_ah, const void* p_att_bf16, const float* w_alpha,
                      float* datt, void* dp_att_bf16, float* dw_alpha, int T, int B, int L, int A, int H, void* stream) {
  REQUIRE(de_all && alpha_all && dctx_all && att_h_all && p_att_bf16 && w_alpha && datt && dp_att_bf16 && dw_alpha, UIC_ERR_ARG,
          "uic_att_tiles_bwd: null pointer");
  if (B == 0 || T == 0) return 0;
  return att_tiles_bwd(de_all, alpha_all, dctx_all, dctx_stride_t, ld_dctx, att_h_all, ah_stride_t, ld_ah, p_att_bf16, w_alpha, datt,
                       dp_att_bf16, dw_alpha, T, B, L, A, H, ST(stream));
}

int uic_lse_xent_bwd(const float* logits, int64_t ld, const float* lse, const int64_t* target, const float* mask,
                     const float* inv_norm, float grad_scale, void* dlogits_bf16, int64_t ld_d, int rows, int V, void* stream) {
  REQUIRE(logits && lse && target && mask && inv_norm && dlogits_bf16, UIC_ERR_ARG, "uic_lse_xent_bwd: null pointer");
  REQUIRE(ld_d >= V, UIC_ERR_SHAPE, "uic_lse_xent_bwd: ld_d=%lld < V=%d", (long long)ld_d, V);
  if (rows == 0) return 0;
  return lse_xent_bwd(logits, ld, lse, target, mask, inv_norm, grad_scale, dlogits_bf16, ld_d, rows, V, ST(stream));
}

int uic_log_softmax_bwd(const float* dlp, int64_t ld_dlp, const float* lp, int64_t ld_lp, void* dlogits_bf16, int64_t ld_d,
                        int rows, int V, void* stream) {
  REQUIRE(dlp && lp && dlogits_bf16, UIC_ERR_ARG, "uic_log_softmax_bwd: null pointer");
  REQUIRE(ld_d >= V, UIC_ERR_SHAPE, "uic_log_softmax_bwd: ld_d=%lld < V=%d", (long long)ld_d, V);
  if (rows == 0) return 0;
  return log_softmax_bwd(dlp, ld_dlp, lp, ld_lp, dlogits_bf16, ld_d, rows, V, ST(stream));
}

int uic_col_sum(const void* x, int is_bf16, int64_t ld, float* out, int rows, int cols, void* stream) {
  REQUIRE(x && out, UIC_ERR_ARG, "uic_col_sum: null pointer");
  if (rows == 0 || cols == 0) return 0;
  return col_sum(x, is_bf16, ld, out, rows, cols, ST(stream));
}

int uic_col_moments(const void* x, int is_bf16, int64_t ld, const int32_t* lens, int n_img, int L, int cols, double* sum, double* sumsq,
                    void* stream) {
  REQUIRE(x && sum && sumsq, UIC_ERR_ARG, "uic_col_moments: null pointer");
  REQUIRE(ld >= cols, UIC_ERR_SHAPE, "uic_col_moments: pitch %lld < cols %d", static_cast<long long>(ld), cols);
  if (n_img <= 0 || L <= 0 || cols <= 0) return 0;
  return col_moments(x, is_bf16, ld, lens, n_img, L, cols, sum, sumsq, ST(stream));
}

int uic_embed_bwd(const float* dxt, int64_t ld, const int64_t* tok, const void* table_relu_bf16, float* demb, int64_t rows, int E,
                  int V, void* stream) {
  REQUIRE(dxt && tok && table_relu_bf16 && demb, UIC_ERR_ARG, "uic_embed_bwd: null pointer");
  if (rows == 0) return 0;
  return embed_bwd(dxt, ld, tok, table_relu_bf16, demb, rows, E, V, ST(stream));
}

int uic_relu_bwd_cast(float* x, const void* y_bf16, void* out_bf16, int64_t n, void* stream) {
  REQUIRE(x && y_bf16 && out_bf16, UIC_ERR_ARG, "uic_relu_bwd_cast: null pointer");
  if (n == 0) return 0;
  return relu_bwd_cast(x, y_bf16, out_bf16, n, ST(stream));
}

int uic_reduce_time(const float* src, int64_t stride_t, int64_t ld, int col0, float* dst, int T, int rows, int n, void* stream) {
  REQUIRE(src && dst, UIC_ERR_ARG, "uic_reduce_time: null pointer");
  if (rows == 0 || n == 0) return 0;
  return reduce_time(src, stride_t, ld, col0, dst, T, rows, n, ST(stream));
}

}  // extern "C"
