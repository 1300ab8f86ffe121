// Element-wise and row-movement kernels of the decoder step: fp32->bf16 operand staging, embedding
// row gather, the two LSTM pointwise updates and the beam-state gather.  All are HBM/L2-bound
// streaming kernels: 16-byte vector accesses where alignment allows, grid sized from the data.
#include "uic_internal.h"
#include "uic_ptx.cuh"
#include "uic_vocab.cuh"

namespace uic {

static inline int blocks_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = 148LL * 16;  // a few waves of the 148 SMs; kernels are grid-stride
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- fp32 -> bf16 (optionally ReLU) -------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, long long ld_src, __nv_bfloat16* __restrict__ dst,
                                     long long ld_dst, long long rows, long long cols, int relu, int vec) {
  if (vec) {  // cols % 8 == 0, pitches % 8 == 0, 16/32-byte aligned bases
    const long long cg = cols / 8;
    const long long total = rows * cg;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long r = i / cg, c = (i - r * cg) * 8;
      const float4 a = *reinterpret_cast<const float4*>(src + r * ld_src + c);
      const float4 b = *reinterpret_cast<const float4*>(src + r * ld_src + c + 4);
      float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      if (relu) {
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
      }
      uint4 q;
      q.x = f2_to_bf16x2(f[0], f[1]);
      q.y = f2_to_bf16x2(f[2], f[3]);
      q.z = f2_to_bf16x2(f[4], f[5]);
      q.w = f2_to_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(dst + r * ld_dst + c) = q;
    }
  } else {
    const long long total = rows * cols;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long r = i / cols, c = i - r * cols;
      float v = src[r * ld_src + c];
      if (relu) v = fmaxf(v, 0.0f);
      dst[r * ld_dst + c] = __float2bfloat16_rn(v);
    }
  }
}

int cast_f32_bf16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, long long cols, int relu,
                  cudaStream_t stream) {
  const int vec = (cols % 8 == 0) && (ld_src % 4 == 0) && (ld_dst % 8 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
  const long long work = vec ? rows * (cols / 8) : rows * cols;
  launch_begin("cast_f32_bf16", stream);
  cast_f32_bf16_kernel<<<blocks_for(work, 256), 256, 0, stream>>>(src, ld_src, static_cast<__nv_bfloat16*>(dst), ld_dst, rows,
                                                                   cols, relu, vec);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- embedding rows -------------------------------------------------------------------------------
__global__ void embed_rows_kernel(const __nv_bfloat16* __restrict__ table, long long ld_table, const int64_t* __restrict__ tok,
                                  __nv_bfloat16* __restrict__ out, long long ld_out, int rows, int E, int V, int vec) {
  pdl_launch_dependents();
  pdl_wait();
  const int per_row = vec ? E / 8 : E;
  const long long total = static_cast<long long>(rows) * per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / per_row);
    const int c = static_cast<int>(i - static_cast<long long>(r) * per_row);
    long long t = tok[r];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    if (vec)
      *reinterpret_cast<uint4*>(out + r * ld_out + c * 8) = *reinterpret_cast<const uint4*>(table + t * ld_table + c * 8);
    else
      out[r * ld_out + c] = table[t * ld_table + c];
  }
}

int embed_rows(const void* table, long long ld_table, const int64_t* tok, void* out, long long ld_out, int rows, int E, int V,
               cudaStream_t stream) {
  const int vec = (E % 8 == 0) && (ld_table % 8 == 0) && (ld_out % 8 == 0) && ((reinterpret_cast<uintptr_t>(table) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const long long work = static_cast<long long>(rows) * (vec ? E / 8 : E);
  launch_begin("embed_rows", stream);
  UIC_CUDA_OK(launch_pdl(embed_rows_kernel, dim3(blocks_for(work, 256)), dim3(256), 0, stream, static_cast<const __nv_bfloat16*>(table),
                         ld_table, tok, static_cast<__nv_bfloat16*>(out), ld_out, rows, E, V, vec));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- LSTM pointwise ---------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
struct HOut {
  float* h_f32;
  __nv_bfloat16* h_a;
  long long ld_ha;
  __nv_bfloat16* h_b;
  long long ld_hb;
};

// VEC consecutive hidden units per thread: 16-byte loads / 8-byte bf16 stores when VEC == 4 (the launchers check alignment),
// scalar accesses when VEC == 1.  These kernels are a few microseconds of pure launch + memory latency; a quarter of the
// threads with four independent chains each is what shortens them.
template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    v[0] = p[0];
  }
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4)
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else
    p[0] = v[0];
}
template <int VEC>
__device__ __forceinline__ void stv_bf16(__nv_bfloat16* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4)
    *reinterpret_cast<uint2*>(p) = make_uint2(f2_to_bf16x2(v[0], v[1]), f2_to_bf16x2(v[2], v[3]));
  else
    p[0] = __float2bfloat16_rn(v[0]);
}
template <int VEC>
__device__ __forceinline__ void store_h(const HOut& o, int r, int j, int H, const float (&h)[VEC]) {
  if (o.h_f32) stv<VEC>(o.h_f32 + static_cast<long long>(r) * H + j, h);
  if (o.h_a) stv_bf16<VEC>(o.h_a + static_cast<long long>(r) * o.ld_ha + j, h);
  if (o.h_b) stv_bf16<VEC>(o.h_b + static_cast<long long>(r) * o.ld_hb + j, h);
}
static bool hout_vec_ok(const HOut& o, int H) {
  return H % 4 == 0 && aligned16(o.h_f32) && aligned8(o.h_a) && aligned8(o.h_b) && o.ld_ha % 4 == 0 && o.ld_hb % 4 == 0;
}

// Optional addends to the gate pre-activations, looked up per row (decode loops): the input-word term W_x relu(Emb[tok])
// is a function of the token alone, so it is precomputed as a (V, n_gates H) fp32 table per weight version and gathered
// here instead of being re-multiplied every step (the gate GEMM then contracts over the recurrent columns only); the
// image-constant term W_fc fc of the TopDown attention LSTM is one row per image, shared by its beams.
struct GateAdd {
  const float* tok_table;  // [V][n_gates H] or nullptr
  long long ld_tok;
  const long long* tok;    // [rows] token ids
  int V;
  const float* grp_table;  // [rows / group][n_gates H] or nullptr
  long long ld_grp;
  int group;
};
template <int VEC>
__device__ __forceinline__ void gate_add(const GateAdd& ga, int r, int col, float (&v)[VEC]) {
  if (ga.tok_table != nullptr) {
    long long t = ga.tok[r];
    t = t < 0 ? 0 : (t >= ga.V ? ga.V - 1 : t);
    float a[VEC];
    ldv<VEC>(ga.tok_table + t * ga.ld_tok + col, a);
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] += a[k];
  }
  if (ga.grp_table != nullptr) {
    float a[VEC];
    ldv<VEC>(ga.grp_table + static_cast<long long>(r / ga.group) * ga.ld_grp + col, a);
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] += a[k];
  }
}
static bool gate_add_vec_ok(const GateAdd& ga) {
  return aligned16(ga.tok_table) && aligned16(ga.grp_table) && ga.ld_tok % 4 == 0 && ga.ld_grp % 4 == 0;
}

// Att2in2Core.forward pointwise part, models/AttModel.py:585-597.
template <int VEC>
__global__ void lstm_maxout_fwd_kernel(const float* __restrict__ sums, long long ld_sums, const float* __restrict__ a2c,
                                       long long ld_a2c, const float* __restrict__ c_prev, float* __restrict__ c_out, HOut o,
                                       int rows, int H, GateAdd ga) {
  pdl_launch_dependents();
  pdl_wait();
  const int per_row = H / VEC;
  const long long total = static_cast<long long>(rows) * per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / per_row), j = static_cast<int>(i - static_cast<long long>(r) * per_row) * VEC;
    const float* s = sums + r * ld_sums;
    const float* a = a2c + r * ld_a2c;
    float si[VEC], sf[VEC], so[VEC], s3[VEC], s4[VEC], a0[VEC], a1[VEC], cp[VEC], c[VEC], h[VEC];
    ldv<VEC>(s + j, si);
    ldv<VEC>(s + H + j, sf);
    ldv<VEC>(s + 2 * H + j, so);
    ldv<VEC>(s + 3 * H + j, s3);
    ldv<VEC>(s + 4 * H + j, s4);
    if (ga.tok_table != nullptr || ga.grp_table != nullptr) {
      gate_add<VEC>(ga, r, j, si);
      gate_add<VEC>(ga, r, H + j, sf);
      gate_add<VEC>(ga, r, 2 * H + j, so);
      gate_add<VEC>(ga, r, 3 * H + j, s3);
      gate_add<VEC>(ga, r, 4 * H + j, s4);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) a0[v] = a1[v] = 0.0f;
    if (a2c) {  // att2all2 (:618-654): the context term was accumulated into all five gate sums by its GEMM
      ldv<VEC>(a + j, a0);
      ldv<VEC>(a + H + j, a1);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) cp[v] = 0.0f;
    if (c_prev) ldv<VEC>(c_prev + static_cast<long long>(r) * H + j, cp);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float ig = sigmoid_acc(si[v]), fg = sigmoid_acc(sf[v]), og = sigmoid_acc(so[v]);
      const float g = fmaxf(s3[v] + a0[v], s4[v] + a1[v]);
      c[v] = fg * cp[v] + ig * g;
      h[v] = og * tanhf(c[v]);
    }
    stv<VEC>(c_out + static_cast<long long>(r) * H + j, c);
    store_h<VEC>(o, r, j, H, h);
  }
}

// torch.nn.LSTMCell pointwise part (gate order i, f, g, o), used at models/AttModel.py:434,441.
template <int VEC>
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gates, long long ld_gates, const float* __restrict__ c_prev,
                                     float* __restrict__ c_out, HOut o, int rows, int H, GateAdd ga) {
  pdl_launch_dependents();
  pdl_wait();
  const int per_row = H / VEC;
  const long long total = static_cast<long long>(rows) * per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / per_row), j = static_cast<int>(i - static_cast<long long>(r) * per_row) * VEC;
    const float* g4 = gates + r * ld_gates;
    float gi[VEC], gf[VEC], gg[VEC], go[VEC], cp[VEC], c[VEC], h[VEC];
    ldv<VEC>(g4 + j, gi);
    ldv<VEC>(g4 + H + j, gf);
    ldv<VEC>(g4 + 2 * H + j, gg);
    ldv<VEC>(g4 + 3 * H + j, go);
    if (ga.tok_table != nullptr || ga.grp_table != nullptr) {
      gate_add<VEC>(ga, r, j, gi);
      gate_add<VEC>(ga, r, H + j, gf);
      gate_add<VEC>(ga, r, 2 * H + j, gg);
      gate_add<VEC>(ga, r, 3 * H + j, go);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) cp[v] = 0.0f;
    if (c_prev) ldv<VEC>(c_prev + static_cast<long long>(r) * H + j, cp);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      c[v] = sigmoid_acc(gf[v]) * cp[v] + sigmoid_acc(gi[v]) * tanhf(gg[v]);
      h[v] = sigmoid_acc(go[v]) * tanhf(c[v]);
    }
    stv<VEC>(c_out + static_cast<long long>(r) * H + j, c);
    store_h<VEC>(o, r, j, H, h);
  }
}

static int gate_add_check(const char* who, const float* add_tok, const long long* tok, int V, const float* add_grp, int group) {
  if (add_tok != nullptr && (tok == nullptr || V <= 0)) return set_error(UIC_ERR_ARG, "%s: the token table needs token ids and V > 0", who);
  if (add_grp != nullptr && group <= 0) return set_error(UIC_ERR_ARG, "%s: rows per group must be positive (got %d)", who, group);
  return 0;
}

int lstm_maxout_fwd(const float* sums, long long ld_sums, const float* a2c, long long ld_a2c, const float* c_prev, float* c_out,
                    float* h_f32, void* h_a, long long ld_ha, void* h_b, long long ld_hb, int rows, int H, cudaStream_t stream,
                    const float* add_tok, long long ld_add_tok, const long long* tok, int V, const float* add_grp, long long ld_add_grp,
                    int group) {
  if (int rc = gate_add_check("lstm_maxout_fwd", add_tok, tok, V, add_grp, group)) return rc;
  HOut o{h_f32, static_cast<__nv_bfloat16*>(h_a), ld_ha, static_cast<__nv_bfloat16*>(h_b), ld_hb};
  GateAdd ga{add_tok, ld_add_tok, tok, V, add_grp, ld_add_grp, group > 0 ? group : 1};
  const bool vec = hout_vec_ok(o, H) && aligned16(sums) && aligned16(a2c) && aligned16(c_prev) && aligned16(c_out) && ld_sums % 4 == 0 &&
                   ld_a2c % 4 == 0 && gate_add_vec_ok(ga);
  launch_begin("lstm_maxout_fwd", stream);
  if (vec)
    UIC_CUDA_OK(launch_pdl(lstm_maxout_fwd_kernel<4>, dim3(blocks_for(static_cast<long long>(rows) * H / 4, 256)), dim3(256), 0, stream, sums,
                           ld_sums, a2c, ld_a2c, c_prev, c_out, o, rows, H, ga));
  else
    UIC_CUDA_OK(launch_pdl(lstm_maxout_fwd_kernel<1>, dim3(blocks_for(static_cast<long long>(rows) * H, 256)), dim3(256), 0, stream, sums,
                           ld_sums, a2c, ld_a2c, c_prev, c_out, o, rows, H, ga));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

int lstm_cell_fwd(const float* gates, long long ld_gates, const float* c_prev, float* c_out, float* h_f32, void* h_a,
                  long long ld_ha, void* h_b, long long ld_hb, int rows, int H, cudaStream_t stream, const float* add_tok,
                  long long ld_add_tok, const long long* tok, int V, const float* add_grp, long long ld_add_grp, int group) {
  if (int rc = gate_add_check("lstm_cell_fwd", add_tok, tok, V, add_grp, group)) return rc;
  HOut o{h_f32, static_cast<__nv_bfloat16*>(h_a), ld_ha, static_cast<__nv_bfloat16*>(h_b), ld_hb};
  GateAdd ga{add_tok, ld_add_tok, tok, V, add_grp, ld_add_grp, group > 0 ? group : 1};
  const bool vec = hout_vec_ok(o, H) && aligned16(gates) && aligned16(c_prev) && aligned16(c_out) && ld_gates % 4 == 0 && gate_add_vec_ok(ga);
  launch_begin("lstm_cell_fwd", stream);
  if (vec)
    UIC_CUDA_OK(launch_pdl(lstm_cell_fwd_kernel<4>, dim3(blocks_for(static_cast<long long>(rows) * H / 4, 256)), dim3(256), 0, stream, gates,
                           ld_gates, c_prev, c_out, o, rows, H, ga));
  else
    UIC_CUDA_OK(launch_pdl(lstm_cell_fwd_kernel<1>, dim3(blocks_for(static_cast<long long>(rows) * H, 256)), dim3(256), 0, stream, gates,
                           ld_gates, c_prev, c_out, o, rows, H, ga));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- dropout (nn.Dropout of embed / fc_embed / att_embed and the core output, models/AttModel.py:73-84,431,599) --------
// keep(row, col) = u >= p with u the counter-based uniform of (seed, site, row, col) (uic_vocab.cuh); kept values are
// scaled by 1 / (1 - p).  The SAME call on a gradient buffer applies the same mask: that is the whole backward.
template <typename T>
__global__ void dropout_kernel(T* __restrict__ x, long long ld, long long rows, int cols, float p, float scale,
                               const unsigned long long* __restrict__ seed, int site, long long row0, long long row_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t sk = rng_step_key(*seed, 0x40000000 + site);
  const long long total = rows * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    const int c = static_cast<int>(i - r * cols);
    const bool keep = rng_uniform(rng_row_key(sk, static_cast<int>(row0 + r * row_stride)), c) >= p;
    T& v = x[r * ld + c];
    v = keep ? static_cast<T>(static_cast<float>(v) * scale) : static_cast<T>(0.0f);
  }
}

int dropout(void* x, int is_bf16, long long ld, long long rows, int cols, float p, const unsigned long long* seed, int site,
            long long row0, long long row_stride, cudaStream_t stream) {
  const float scale = 1.0f / (1.0f - p);
  const int grid = blocks_for(rows * cols, 256);
  launch_begin("dropout", stream);
  if (is_bf16)
    UIC_CUDA_OK(launch_pdl(dropout_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream, static_cast<__nv_bfloat16*>(x), ld, rows, cols, p,
                           scale, seed, site, row0, row_stride));
  else
    UIC_CUDA_OK(launch_pdl(dropout_kernel<float>, dim3(grid), dim3(256), 0, stream, static_cast<float*>(x), ld, rows, cols, p, scale, seed,
                           site, row0, row_stride));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- beam state gather (CaptionModel.py:89-91) -------------------------------------------------------
__global__ void beam_gather_kernel(const int32_t* __restrict__ parent, const __nv_bfloat16* __restrict__ x_src,
                                   __nv_bfloat16* __restrict__ x_dst, long long ld_x, int col0_a, int ncol_a, int col0_b,
                                   int ncol_b, const float* __restrict__ c_src, float* __restrict__ c_dst, int n_state, int rows,
                                   int H) {
  const int r = blockIdx.x;
  const int q = parent[r];
  for (int c = threadIdx.x; c < ncol_a; c += blockDim.x) x_dst[r * ld_x + col0_a + c] = x_src[q * ld_x + col0_a + c];
  for (int c = threadIdx.x; c < ncol_b; c += blockDim.x) x_dst[r * ld_x + col0_b + c] = x_src[q * ld_x + col0_b + c];
  if (c_src != nullptr) {
    for (int s = 0; s < n_state; ++s) {
      const float* src = c_src + (static_cast<long long>(s) * rows + q) * H;
      float* dst = c_dst + (static_cast<long long>(s) * rows + r) * H;
      for (int c = threadIdx.x; c < H; c += blockDim.x) dst[c] = src[c];
    }
  }
}

int beam_gather(const int32_t* parent_row, const void* x_src, void* x_dst, long long ld_x, int col0_a, int ncol_a, int col0_b,
                int ncol_b, const float* c_src, float* c_dst, int n_state, int rows, int H, cudaStream_t stream) {
  launch_begin("beam_gather", stream);
  beam_gather_kernel<<<rows, 128, 0, stream>>>(parent_row, static_cast<const __nv_bfloat16*>(x_src),
                                               static_cast<__nv_bfloat16*>(x_dst), ld_x, col0_a, ncol_a, col0_b, ncol_b, c_src,
                                               c_dst, n_state, rows, H);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

}  // namespace uic

namespace uic {

// ---- zero the padded regions of the embedded features (pad_packed_sequence, AttModel.py:50-51) -----
__global__ void zero_padded_rows_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ masks, int L, int H) {
  const long long row = blockIdx.x;  // row = img * L + l
  const int img = static_cast<int>(row / L), l = static_cast<int>(row - static_cast<long long>(img) * L);
  // number of valid regions = sum of the mask row (masks are prefix-shaped in the reference's packed path)
  float n = 0.0f;
  for (int q = 0; q < L; ++q) n += masks[static_cast<long long>(img) * L + q];
  if (l < static_cast<int>(n)) return;
  for (int c = threadIdx.x; c < H; c += blockDim.x) x[row * H + c] = __float2bfloat16_rn(0.0f);
}

int zero_padded_rows(void* x, const float* masks, int n_img, int L, int H, cudaStream_t stream) {
  launch_begin("zero_padded_rows", stream);
  zero_padded_rows_kernel<<<n_img * L, 128, 0, stream>>>(static_cast<__nv_bfloat16*>(x), masks, L, H);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

}  // namespace uic
