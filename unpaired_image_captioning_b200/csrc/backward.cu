// Backward kernels of the teacher-forced decoder step (BPTT through Att2in2Core / TopDownCore,
// Attention, log-softmax + masked cross-entropy).  The reference gets all of this from autograd
// over ~30 ATen kernels per step (trainer.py:173); here every step's backward is a handful of
// fused kernels plus tcgen05 dgrad GEMMs, and everything that can be deferred is time-batched:
//   * weight gradients: one wgrad GEMM per weight over all T steps (K = T*B), MN-major operands;
//   * the gradients w.r.t. the feature tiles (att, p_att) are NOT read-modify-written every step:
//     each step only stores de (B, L); one pass at the end rebuilds sum_t(...) per tile element.

#include "uic_internal.h"
#include "uic_ptx.cuh"

namespace uic {

static inline int grid_for(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

struct Addends {  // up to three fp32 (rows x H) matrices with their own pitches, any may be null
  const float* p[3];
  long long ld[3];
};
__device__ __forceinline__ float add3(const Addends& a, int r, int j) {
  float v = 0.0f;
#pragma unroll
  for (int q = 0; q < 3; ++q)
    if (a.p[q]) v += a.p[q][static_cast<long long>(r) * a.ld[q] + j];
  return v;
}

// ---- torch.nn.LSTMCell backward (gate order i, f, g, o) ---------------------------------------------
// VEC consecutive hidden units per thread (16-byte loads, 8-byte bf16 stores when VEC == 4), like the forward kernels.
template <int VEC>
__device__ __forceinline__ void ldv_b(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 4) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
    v[0] = p[0];
  }
}
template <int VEC>
__device__ __forceinline__ void stv_bf16_b(__nv_bfloat16* p, const float (&v)[VEC]) {
  if constexpr (VEC == 4)
    *reinterpret_cast<uint2*>(p) = make_uint2(f2_to_bf16x2(v[0], v[1]), f2_to_bf16x2(v[2], v[3]));
  else
    p[0] = __float2bfloat16_rn(v[0]);
}

template <int VEC>
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gates, long long ld_gates, const float* __restrict__ c_prev,
                                     const float* __restrict__ c, Addends dh, const float* __restrict__ dc_next,
                                     __nv_bfloat16* __restrict__ dgates, long long ld_dg, float* __restrict__ dc_prev, int rows,
                                     int H) {
  const int per_row = H / VEC;
  const long long total = static_cast<long long>(rows) * per_row;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / per_row), j = static_cast<int>(i - static_cast<long long>(r) * per_row) * VEC;
    const long long idx = static_cast<long long>(r) * H + j;
    const float* g4 = gates + r * ld_gates;
    float gi[VEC], gf[VEC], gg[VEC], go[VEC], cv[VEC], cp[VEC], dcn[VEC], dhv[VEC], tmp[VEC];
    ldv_b<VEC>(g4 + j, gi);
    ldv_b<VEC>(g4 + H + j, gf);
    ldv_b<VEC>(g4 + 2 * H + j, gg);
    ldv_b<VEC>(g4 + 3 * H + j, go);
    ldv_b<VEC>(c + idx, cv);
#pragma unroll
    for (int v = 0; v < VEC; ++v) cp[v] = dcn[v] = dhv[v] = 0.0f;
    if (c_prev) ldv_b<VEC>(c_prev + idx, cp);
    if (dc_next) ldv_b<VEC>(dc_next + idx, dcn);
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (dh.p[q]) {
        ldv_b<VEC>(dh.p[q] + static_cast<long long>(r) * dh.ld[q] + j, tmp);
#pragma unroll
        for (int v = 0; v < VEC; ++v) dhv[v] += tmp[v];
      }
    float di[VEC], df[VEC], dg[VEC], dO[VEC], dcp[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float ig = sigmoid_acc(gi[v]), fg = sigmoid_acc(gf[v]), g = tanhf(gg[v]), og = sigmoid_acc(go[v]);
      const float tc = tanhf(cv[v]);
      const float dc = dcn[v] + dhv[v] * og * (1.0f - tc * tc);
      di[v] = dc * g * ig * (1.0f - ig);
      df[v] = dc * cp[v] * fg * (1.0f - fg);
      dg[v] = dc * ig * (1.0f - g * g);
      dO[v] = dhv[v] * tc * og * (1.0f - og);
      dcp[v] = dc * fg;
    }
    __nv_bfloat16* d = dgates + r * ld_dg;
    stv_bf16_b<VEC>(d + j, di);
    stv_bf16_b<VEC>(d + H + j, df);
    stv_bf16_b<VEC>(d + 2 * H + j, dg);
    stv_bf16_b<VEC>(d + 3 * H + j, dO);
    if constexpr (VEC == 4)
      *reinterpret_cast<float4*>(dc_prev + idx) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    else
      dc_prev[idx] = dcp[0];
  }
}

int lstm_cell_bwd(const float* gates, long long ld_gates, const float* c_prev, const float* c, const float* dh0, long long ld0,
                  const float* dh1, long long ld1, const float* dh2, long long ld2, const float* dc_next, void* dgates,
                  long long ld_dg, float* dc_prev, int rows, int H, cudaStream_t stream) {
  Addends a{{dh0, dh1, dh2}, {ld0, ld1, ld2}};
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = H % 4 == 0 && ld_gates % 4 == 0 && ld_dg % 4 == 0 && a16(gates) && a16(c_prev) && a16(c) && a16(dc_next) &&
                   a16(dc_prev) && (reinterpret_cast<uintptr_t>(dgates) & 7) == 0 && a16(dh0) && a16(dh1) && a16(dh2) &&
                   ld0 % 4 == 0 && ld1 % 4 == 0 && ld2 % 4 == 0;
  launch_begin("lstm_cell_bwd", stream);
  if (vec)
    lstm_cell_bwd_kernel<4><<<grid_for(static_cast<long long>(rows) * H / 4, 256), 256, 0, stream>>>(
        gates, ld_gates, c_prev, c, a, dc_next, static_cast<__nv_bfloat16*>(dgates), ld_dg, dc_prev, rows, H);
  else
    lstm_cell_bwd_kernel<1><<<grid_for(static_cast<long long>(rows) * H, 256), 256, 0, stream>>>(
        gates, ld_gates, c_prev, c, a, dc_next, static_cast<__nv_bfloat16*>(dgates), ld_dg, dc_prev, rows, H);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- Att2in2 maxout cell backward (models/AttModel.py:585-597) ------------------------------------------
__global__ void lstm_maxout_bwd_kernel(const float* __restrict__ sums, long long ld_sums, const float* __restrict__ a2c,
                                       long long ld_a2c, const float* __restrict__ c_prev, const float* __restrict__ c,
                                       Addends dh, const float* __restrict__ dc_next, __nv_bfloat16* __restrict__ dsums,
                                       long long ld_ds, __nv_bfloat16* __restrict__ da2c, long long ld_da,
                                       float* __restrict__ dc_prev, int rows, int H) {
  const long long total = static_cast<long long>(rows) * H;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(idx / H), j = static_cast<int>(idx - static_cast<long long>(r) * H);
    const float* s = sums + r * ld_sums;
    const float* a = a2c + r * ld_a2c;
    const float ig = sigmoid_acc(s[j]), fg = sigmoid_acc(s[H + j]), og = sigmoid_acc(s[2 * H + j]);
    const float p1 = s[3 * H + j] + (a2c ? a[j] : 0.0f), p2 = s[4 * H + j] + (a2c ? a[H + j] : 0.0f);
    const float g = fmaxf(p1, p2);
    const float tc = tanhf(c[idx]);
    const float dhv = add3(dh, r, j);
    const float dc = (dc_next ? dc_next[idx] : 0.0f) + dhv * og * (1.0f - tc * tc);
    const float cp = c_prev ? c_prev[idx] : 0.0f;
    const float dg = dc * ig;
    // torch.max(a, b) routes the gradient to the larger input and splits it on exact ties
    const float d1 = p1 > p2 ? dg : (p1 == p2 ? 0.5f * dg : 0.0f);
    const float d2 = dg - d1;
    __nv_bfloat16* d = dsums + r * ld_ds;
    d[j] = __float2bfloat16_rn(dc * g * ig * (1.0f - ig));
    d[H + j] = __float2bfloat16_rn(dc * cp * fg * (1.0f - fg));
    d[2 * H + j] = __float2bfloat16_rn(dhv * tc * og * (1.0f - og));
    d[3 * H + j] = __float2bfloat16_rn(d1);
    d[4 * H + j] = __float2bfloat16_rn(d2);
    if (da2c) {
      da2c[r * ld_da + j] = __float2bfloat16_rn(d1);
      da2c[r * ld_da + H + j] = __float2bfloat16_rn(d2);
    }
    dc_prev[idx] = dc * fg;
  }
}

int lstm_maxout_bwd(const float* sums, long long ld_sums, const float* a2c, long long ld_a2c, const float* c_prev, const float* c,
                    const float* dh0, long long ld0, const float* dh1, long long ld1, const float* dc_next, void* dsums,
                    long long ld_ds, void* da2c, long long ld_da, float* dc_prev, int rows, int H, cudaStream_t stream) {
  Addends a{{dh0, dh1, nullptr}, {ld0, ld1, 0}};
  launch_begin("lstm_maxout_bwd", stream);
  lstm_maxout_bwd_kernel<<<grid_for(static_cast<long long>(rows) * H, 256), 256, 0, stream>>>(
      sums, ld_sums, a2c, ld_a2c, c_prev, c, a, dc_next, static_cast<__nv_bfloat16*>(dsums), ld_ds,
      static_cast<__nv_bfloat16*>(da2c), ld_da, dc_prev, rows, H);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- attention step backward: de (rows, L) and d att_h (rows, A) --------------------------------------
// One CTA per row (= image in teacher-forced training).  Phase 1 streams the att tile to get
// d alpha_l = <dctx, att_l>; the softmax Jacobian gives de_l = alpha_l (d alpha_l - sum alpha d alpha);
// phase 2 streams the p_att tile for d att_h[a] = w_a sum_l de_l (1 - tanh^2(p_att[l,a] + att_h[a])).
constexpr int ATTB_THREADS = 256;
constexpr int ATTB_WARPS = ATTB_THREADS / 32;
constexpr int ATTB_GROUP = 2;  // regions whose loads a warp issues before it consumes the first (the kernel is latency bound)

// C = 256-wide chunks of max(A, H): lane owns elements [256c + 8*lane, +8) of chunk c.
template <int C>
__global__ void __launch_bounds__(ATTB_THREADS, C <= 2 ? 4 : 2) att_step_bwd_kernel(const float* __restrict__ dctx, long long ld_dctx,
                                                                                   const float* __restrict__ alpha,
                                                                                   const __nv_bfloat16* __restrict__ p_att,
                                                                                   const __nv_bfloat16* __restrict__ att,
                                                                                   const float* __restrict__ att_h, long long ld_att_h,
                                                                                   const float* __restrict__ w_alpha, float* __restrict__ de,
                                                                                   __nv_bfloat16* __restrict__ datt_h, long long ld_dah,
                                                                                   int L, int A, int H) {
  extern __shared__ float sm[];
  float* s_de = sm;           // [L]              d alpha, then de
  float* s_acc = sm + ((L + 3) & ~3);  // [ATTB_WARPS][A]  per-warp d att_h partial sums (16-byte aligned)
  __shared__ float s_red[ATTB_WARPS];
  const int r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* a_img = att + static_cast<long long>(r) * L * H;
  const __nv_bfloat16* p_img = p_att + static_cast<long long>(r) * L * A;
  const float* al = alpha + static_cast<long long>(r) * L;

  // phase 1: d alpha_l = <dctx, att_l>
  float dc[C * 8];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int h = c * 256 + lane * 8 + k;
      dc[c * 8 + k] = h < H ? dctx[static_cast<long long>(r) * ld_dctx + h] : 0.0f;
    }
  for (int l0 = warp; l0 < L; l0 += ATTB_WARPS * ATTB_GROUP) {
    uint4 q[ATTB_GROUP][C];
#pragma unroll
    for (int gi = 0; gi < ATTB_GROUP; ++gi) {
      const int l = l0 + gi * ATTB_WARPS;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int h0 = c * 256 + lane * 8;
        q[gi][c] = make_uint4(0, 0, 0, 0);
        if (l < L && h0 < H) q[gi][c] = ldg_nc_v4(a_img + static_cast<long long>(l) * H + h0);
      }
    }
#pragma unroll
    for (int gi = 0; gi < ATTB_GROUP; ++gi) {
      const int l = l0 + gi * ATTB_WARPS;
      if (l >= L) break;  // warp-uniform
      float part = 0.0f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t u[4] = {q[gi][c].x, q[gi][c].y, q[gi][c].z, q[gi][c].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = bf16x2_to_f2(u[k]);
          part = fmaf(f.x, dc[c * 8 + 2 * k], part);
          part = fmaf(f.y, dc[c * 8 + 2 * k + 1], part);
        }
      }
      part = warp_sum(part);
      if (lane == 0) s_de[l] = part;
    }
  }
  __syncthreads();
  float s = 0.0f;
  for (int l = threadIdx.x; l < L; l += ATTB_THREADS) s = fmaf(al[l], s_de[l], s);
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  float tot = 0.0f;
#pragma unroll
  for (int q2 = 0; q2 < ATTB_WARPS; ++q2) tot += s_red[q2];
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += ATTB_THREADS) {
    const float v = al[l] * (s_de[l] - tot);
    s_de[l] = v;
    de[static_cast<long long>(r) * L + l] = v;
  }
  __syncthreads();

  // phase 2: d att_h[a] = w_a sum_l de_l (1 - tanh^2(p_att[l,a] + att_h[a]))
  float ah[C * 8], acc[C * 8];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int a = c * 256 + lane * 8 + k;
      ah[c * 8 + k] = a < A ? att_h[static_cast<long long>(r) * ld_att_h + a] : 0.0f;
      acc[c * 8 + k] = 0.0f;
    }
  for (int l0 = warp; l0 < L; l0 += ATTB_WARPS * ATTB_GROUP) {
    uint4 q[ATTB_GROUP][C];
#pragma unroll
    for (int gi = 0; gi < ATTB_GROUP; ++gi) {
      const int l = l0 + gi * ATTB_WARPS;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int a0 = c * 256 + lane * 8;
        q[gi][c] = make_uint4(0, 0, 0, 0);
        if (l < L && a0 < A) q[gi][c] = ldg_nc_v4(p_img + static_cast<long long>(l) * A + a0);
      }
    }
#pragma unroll
    for (int gi = 0; gi < ATTB_GROUP; ++gi) {
      const int l = l0 + gi * ATTB_WARPS;
      if (l >= L) break;  // warp-uniform
      const float del = 4.0f * s_de[l];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t u[4] = {q[gi][c].x, q[gi][c].y, q[gi][c].z, q[gi][c].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = bf16x2_to_f2(u[k]);
          // operands are E = exp(2 p) (bf16 tile) and F = exp(2 att_h): tanh = 1 - 2r, 1 - tanh^2 = 4 r (1 - r), r = 1/(E F + 1)
          const float r0 = rcp_approx(fmaf(f.x, ah[c * 8 + 2 * k], 1.0f)), r1 = rcp_approx(fmaf(f.y, ah[c * 8 + 2 * k + 1], 1.0f));
          acc[c * 8 + 2 * k] = fmaf(del, r0 * (1.0f - r0), acc[c * 8 + 2 * k]);
          acc[c * 8 + 2 * k + 1] = fmaf(del, r1 * (1.0f - r1), acc[c * 8 + 2 * k + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int a0 = c * 256 + lane * 8;
    if (a0 < A) {
      float* dst = s_acc + warp * A + a0;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[c * 8], acc[c * 8 + 1], acc[c * 8 + 2], acc[c * 8 + 3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[c * 8 + 4], acc[c * 8 + 5], acc[c * 8 + 6], acc[c * 8 + 7]);
    }
  }
  __syncthreads();
  for (int a = threadIdx.x; a < A; a += ATTB_THREADS) {
    float v = 0.0f;
#pragma unroll
    for (int q2 = 0; q2 < ATTB_WARPS; ++q2) v += s_acc[q2 * A + a];
    datt_h[static_cast<long long>(r) * ld_dah + a] = __float2bfloat16_rn(v * w_alpha[a]);
  }
}

int att_step_bwd(const float* dctx, long long ld_dctx, const float* alpha, const void* p_att, const void* att, const float* att_h,
                 long long ld_att_h, const float* w_alpha, float* de, void* datt_h, long long ld_dah, int rows, int L, int A,
                 int H, cudaStream_t stream) {
  if (A % 8 || H % 8 || A > 1024 || H > 1024) return set_error(UIC_ERR_SHAPE, "att_step_bwd: A=%d H=%d", A, H);
  // s_de [L] (padded to 4 floats so that the float4 stores of the partial sums stay aligned) + per-warp partials
  const int Lp = (L + 3) / 4 * 4;
  const size_t smem = sizeof(float) * (static_cast<size_t>(Lp) + static_cast<size_t>(ATTB_WARPS) * A);
  if (smem > 48 * 1024) return set_error(UIC_ERR_SHAPE, "att_step_bwd: L=%d too large", L);
  const int chunks = ((A > H ? A : H) + 255) / 256;
  launch_begin("att_step_bwd", stream);
#define UIC_ATTB(C_)                                                                                                          \
  att_step_bwd_kernel<C_><<<rows, ATTB_THREADS, smem, stream>>>(dctx, ld_dctx, alpha, static_cast<const __nv_bfloat16*>(p_att),          \
                                                                static_cast<const __nv_bfloat16*>(att), att_h, ld_att_h, w_alpha, \
                                                                de, static_cast<__nv_bfloat16*>(datt_h), ld_dah, L, A, H)
  if (chunks <= 1)
    UIC_ATTB(1);
  else if (chunks <= 2)
    UIC_ATTB(2);
  else
    UIC_ATTB(4);
#undef UIC_ATTB
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- deferred tile gradients: d att (B,L,H), d p_att (B,L,A), d w_alpha (A) ----------------------------
//   d att[b,l,h]   = sum_t alpha_t[b,l] * dctx_t[b,h]
//   d p_att[b,l,a] = w_a * sum_t de_t[b,l] * (1 - tanh^2(p_att[b,l,a] + att_h_t[b,a]))
//   d w_alpha[a]   = sum_{b,l,t} de_t[b,l] * tanh(p_att[b,l,a] + att_h_t[b,a])
// One CTA per (image, chunk of regions); the per-step vectors of the image are staged in shared memory.
constexpr int TILE_THREADS = 256;

__global__ void __launch_bounds__(TILE_THREADS) att_tiles_bwd_kernel(
    const float* __restrict__ de_all, const float* __restrict__ alpha_all, const float* __restrict__ dctx_all,
    long long dctx_stride_t, long long ld_dctx, const float* __restrict__ att_h_all, long long ah_stride_t, long long ld_ah,
    const __nv_bfloat16* __restrict__ p_att, const float* __restrict__ w_alpha, float* __restrict__ datt,
    __nv_bfloat16* __restrict__ dp_att, float* __restrict__ dw_alpha, int T, int B, int L, int A, int H, int l_chunk) {
  extern __shared__ float sm[];
  float* s_ah = sm;                       // [T][A]
  float* s_dctx = s_ah + T * A;           // [T][H]
  float* s_de = s_dctx + T * H;           // [T][l_chunk]
  float* s_al = s_de + T * l_chunk;       // [T][l_chunk]
  const int b = blockIdx.x;
  const int l0 = blockIdx.y * l_chunk;
  const int nl = min(l_chunk, L - l0);
  for (int i = threadIdx.x; i < T * A; i += TILE_THREADS) {
    const int t = i / A, a = i - t * A;
    s_ah[i] = att_h_all[t * ah_stride_t + static_cast<long long>(b) * ld_ah + a];
  }
  for (int i = threadIdx.x; i < T * H; i += TILE_THREADS) {
    const int t = i / H, h = i - t * H;
    s_dctx[i] = dctx_all[t * dctx_stride_t + static_cast<long long>(b) * ld_dctx + h];
  }
  for (int i = threadIdx.x; i < T * l_chunk; i += TILE_THREADS) {
    const int t = i / l_chunk, q = i - t * l_chunk;
    const bool ok = q < nl;
    const long long off = (static_cast<long long>(t) * B + b) * L + l0 + q;
    s_de[i] = ok ? de_all[off] : 0.0f;
    s_al[i] = ok ? alpha_all[off] : 0.0f;
  }
  __syncthreads();
  // Four regions per pass: their p_att values are loaded up front and every shared-memory read of att_h / dctx serves four
  // of them (the first version re-read shared memory once per region and waited for one global load per region).
  constexpr int QU = 4;
  for (int a = threadIdx.x; a < A; a += TILE_THREADS) {
    const float wa = w_alpha[a];
    float dw = 0.0f, dbias = 0.0f;
    for (int q0 = 0; q0 < nl; q0 += QU) {
      float p[QU], acc[QU];
#pragma unroll
      for (int u = 0; u < QU; ++u) {
        const int q = min(q0 + u, nl - 1);
        p[u] = __bfloat162float(p_att[(static_cast<long long>(b) * L + l0 + q) * A + a]);
        acc[u] = 0.0f;
      }
      for (int t = 0; t < T; ++t) {
        const float f = s_ah[t * A + a];  // p = E, s_ah = F (exponential operand form)
#pragma unroll
        for (int u = 0; u < QU; ++u) {
          const float r = rcp_approx(fmaf(p[u], f, 1.0f));
          const float d = (q0 + u < nl) ? s_de[t * l_chunk + q0 + u] : 0.0f;
          acc[u] = fmaf(d, 4.0f * r * (1.0f - r), acc[u]);
          dw = fmaf(d, 1.0f - 2.0f * r, dw);
        }
      }
#pragma unroll
      for (int u = 0; u < QU; ++u) {
        if (q0 + u < nl) {
          dp_att[(static_cast<long long>(b) * L + l0 + q0 + u) * A + a] = __float2bfloat16_rn(acc[u] * wa);
          dbias += acc[u] * wa;
        }
      }
    }
    atomicAdd(dw_alpha + a, dw);
    atomicAdd(dw_alpha + A + a, dbias);  // fp32 column sum of d p_att = d bias of ctx2att
  }
  for (int h = threadIdx.x; h < H; h += TILE_THREADS) {
    for (int q0 = 0; q0 < nl; q0 += QU) {
      float acc[QU];
#pragma unroll
      for (int u = 0; u < QU; ++u) acc[u] = 0.0f;
      for (int t = 0; t < T; ++t) {
        const float dcx = s_dctx[t * H + h];
#pragma unroll
        for (int u = 0; u < QU; ++u) acc[u] = fmaf((q0 + u < nl) ? s_al[t * l_chunk + q0 + u] : 0.0f, dcx, acc[u]);
      }
#pragma unroll
      for (int u = 0; u < QU; ++u)
        if (q0 + u < nl) datt[(static_cast<long long>(b) * L + l0 + q0 + u) * H + h] = acc[u];
    }
  }
}

int att_tiles_bwd(const float* de_all, const float* alpha_all, const float* dctx_all, long long dctx_stride_t, long long ld_dctx,
                  const float* att_h_all, long long ah_stride_t, long long ld_ah, const void* p_att, const float* w_alpha,
                  float* datt, void* dp_att, float* dw_alpha, int T, int B, int L, int A, int H, cudaStream_t stream) {
  int l_chunk = L < 32 ? L : 28;
  if (L >= 32 && L <= 40) l_chunk = L;
  const size_t smem = sizeof(float) * (static_cast<size_t>(T) * (A + H) + 2 * static_cast<size_t>(T) * l_chunk);
  if (smem > 200 * 1024) return set_error(UIC_ERR_SHAPE, "att_tiles_bwd: T=%d A=%d H=%d need %zu bytes of shared memory", T, A, H, smem);
  if (smem > 48 * 1024)
    UIC_CUDA_OK(cudaFuncSetAttribute(att_tiles_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(B, (L + l_chunk - 1) / l_chunk);
  launch_begin("att_tiles_bwd", stream);
  att_tiles_bwd_kernel<<<grid, TILE_THREADS, smem, stream>>>(de_all, alpha_all, dctx_all, dctx_stride_t, ld_dctx, att_h_all,
                                                             ah_stride_t, ld_ah, static_cast<const __nv_bfloat16*>(p_att), w_alpha,
                                                             datt, static_cast<__nv_bfloat16*>(dp_att), dw_alpha, T, B, L, A, H,
                                                             l_chunk);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- masked cross-entropy backward: d logits = (softmax - onehot) * mask * scale, bf16 -------------------
__global__ void __launch_bounds__(256) lse_xent_bwd_kernel(const float* __restrict__ logits, long long ld,
                                                           const float* __restrict__ lse, const int64_t* __restrict__ target,
                                                           const float* __restrict__ mask, const float* __restrict__ inv_norm,
                                                           float grad_scale, __nv_bfloat16* __restrict__ dlogits, long long ld_d,
                                                           int V) {
  const int r = blockIdx.x;
  const float* row = logits + static_cast<long long>(r) * ld;
  __nv_bfloat16* drow = dlogits + static_cast<long long>(r) * ld_d;
  const float sc = mask[r] * inv_norm[0] * grad_scale;
  const float l = lse[r];
  long long t = target[r];
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);
  for (int v = threadIdx.x; v < ld_d; v += 256) {
    float g = 0.0f;
    if (v < V && sc != 0.0f) g = (__expf(row[v] - l) - (v == t ? 1.0f : 0.0f)) * sc;
    drow[v] = __float2bfloat16_rn(g);
  }
}

int lse_xent_bwd(const float* logits, long long ld, const float* lse, const int64_t* target, const float* mask,
                 const float* inv_norm, float grad_scale, void* dlogits, long long ld_d, int rows, int V, cudaStream_t stream) {
  launch_begin("lse_xent_bwd", stream);
  lse_xent_bwd_kernel<<<rows, 256, 0, stream>>>(logits, ld, lse, target, mask, inv_norm, grad_scale,
                                                static_cast<__nv_bfloat16*>(dlogits), ld_d, V);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- log_softmax backward for the dense (B,T,V) API path: d logits = d lp - exp(lp) * sum(d lp) ----------
__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ dlp, long long ld_dlp,
                                                              const float* __restrict__ lp, long long ld_lp,
                                                              __nv_bfloat16* __restrict__ dlogits, long long ld_d, int V) {
  __shared__ float s_red[8];
  const int r = blockIdx.x;
  const float* g = dlp + static_cast<long long>(r) * ld_dlp;
  const float* y = lp + static_cast<long long>(r) * ld_lp;
  float s = 0.0f;
  for (int v = threadIdx.x; v < V; v += 256) s += g[v];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
  __syncthreads();
  float tot = 0.0f;
#pragma unroll
  for (int q = 0; q < 8; ++q) tot += s_red[q];
  __nv_bfloat16* d = dlogits + static_cast<long long>(r) * ld_d;
  for (int v = threadIdx.x; v < ld_d; v += 256) d[v] = __float2bfloat16_rn(v < V ? g[v] - __expf(y[v]) * tot : 0.0f);
}

int log_softmax_bwd(const float* dlp, long long ld_dlp, const float* lp, long long ld_lp, void* dlogits, long long ld_d, int rows,
                    int V, cudaStream_t stream) {
  launch_begin("log_softmax_bwd", stream);
  log_softmax_bwd_kernel<<<rows, 256, 0, stream>>>(dlp, ld_dlp, lp, ld_lp, static_cast<__nv_bfloat16*>(dlogits), ld_d, V);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- column sums (bias gradients): out[c] += sum_r x[r, c] ------------------------------------------------
// Each lane owns VEC consecutive columns (one 16-byte load per row), a CTA covers 32*VEC columns and a slice
// of the rows (8 warps x 4 rows in flight each); partial sums are merged through shared memory and atomics.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) col_sum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out, int rows,
                                                      int cols, int vec_ok) {
  __shared__ float s[8][32 * VEC + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * VEC;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.0f;
  const int rows_per_cta = (rows + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
  if (c0 < cols) {
    if (vec_ok && c0 + VEC <= cols) {
      for (int r = r_begin + warp; r < r_end; r += 8) {
        const uint4 q = ldg_nc_v4(x + static_cast<long long>(r) * ld + c0);
        if (sizeof(T) == 2) {
          const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = bf16x2_to_f2(u[k]);
            acc[(2 * k) % VEC] += f.x;
            acc[(2 * k + 1) % VEC] += f.y;
          }
        } else {
          acc[0] += __uint_as_float(q.x);
          acc[1 % VEC] += __uint_as_float(q.y);
          acc[2 % VEC] += __uint_as_float(q.z);
          acc[3 % VEC] += __uint_as_float(q.w);
        }
      }
    } else {
      for (int r = r_begin + warp; r < r_end; r += 8)
#pragma unroll
        for (int k = 0; k < VEC; ++k)
          if (c0 + k < cols) acc[k] += static_cast<float>(x[static_cast<long long>(r) * ld + c0 + k]);
    }
  }
#pragma unroll
  for (int k = 0; k < VEC; ++k) s[warp][lane * VEC + k] = acc[k];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * VEC; c += 256) {
    const int col = blockIdx.x * 32 * VEC + c;
    if (col < cols) {
      float t = 0.0f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += s[q][c];
      atomicAdd(out + col, t);
    }
  }
}

int col_sum(const void* x, int is_bf16, long long ld, float* out, int rows, int cols, cudaStream_t stream) {
  const int vec = is_bf16 ? 8 : 4;
  const int vec_ok = (ld % vec == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int gx = (cols + 32 * vec - 1) / (32 * vec);
  int gy = (2 * 148 + gx - 1) / gx;  // about two CTAs per SM
  const int max_gy = (rows + 31) / 32;
  gy = gy > max_gy ? max_gy : gy;
  gy = gy < 1 ? 1 : gy;
  dim3 grid(gx, gy);
  launch_begin("col_sum", stream);
  if (is_bf16)
    col_sum_kernel<__nv_bfloat16, 8><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, out, rows, cols, vec_ok);
  else
    col_sum_kernel<float, 4><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), ld, out, rows, cols, vec_ok);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- BatchNorm1d statistics over the packed (valid) regions (models/AttModel.py:80 with pack_wrapper :44-53) ----------
// sum[c] += sum_r x[r, c], sumsq[c] += sum_r x[r, c]^2 over the rows (image i, region l < lens[i]) of x (n_img * L, cols).
// One thread per column (a warp reads 128 contiguous bytes of a row), 8 rows in flight per thread, fp64 accumulation
// so that var = E[x^2] - mean^2 keeps fp32 accuracy.
template <typename T>
__global__ void __launch_bounds__(256) col_moments_kernel(const T* __restrict__ x, long long ld, const int32_t* __restrict__ lens,
                                                          int n_img, int L, int cols, double* __restrict__ sum,
                                                          double* __restrict__ sumsq) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int rows = n_img * L;
  const int rows_per_cta = (rows + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
  if (c >= cols) return;
  double s1 = 0.0, s2 = 0.0;
  for (int r0 = r_begin; r0 < r_end; r0 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = r0 + k;
      v[k] = 0.0f;
      if (r < r_end) {
        const int img = r / L;
        if (lens == nullptr || r - img * L < lens[img]) v[k] = static_cast<float>(x[static_cast<long long>(r) * ld + c]);
      }
    }
    float p1 = 0.0f, p2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      p1 += v[k];
      p2 = fmaf(v[k], v[k], p2);
    }
    s1 += static_cast<double>(p1);
    s2 += static_cast<double>(p2);
  }
  atomicAdd(sum + c, s1);
  atomicAdd(sumsq + c, s2);
}

int col_moments(const void* x, int is_bf16, long long ld, const int32_t* lens, int n_img, int L, int cols, double* sum, double* sumsq,
                cudaStream_t stream) {
  const int gx = (cols + 255) / 256;
  const int rows = n_img * L;
  int gy = (4 * 148 + gx - 1) / gx;
  const int max_gy = (rows + 63) / 64;
  gy = gy > max_gy ? max_gy : gy;
  gy = gy < 1 ? 1 : gy;
  dim3 grid(gx, gy);
  launch_begin("col_moments", stream);
  if (is_bf16)
    col_moments_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, lens, n_img, L, cols, sum, sumsq);
  else
    col_moments_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), ld, lens, n_img, L, cols, sum, sumsq);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- embedding backward: dEmb[tok[r], e] += dxt[r, e] where relu(emb) was active -----------------------------
__global__ void embed_bwd_kernel(const float* __restrict__ dxt, long long ld, const int64_t* __restrict__ tok,
                                 const __nv_bfloat16* __restrict__ table_relu, float* __restrict__ demb, long long rows, int E,
                                 int V) {
  const long long total = rows * E;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / E;
    const int e = static_cast<int>(i - r * E);
    long long t = tok[r];
    t = t < 0 ? 0 : (t >= V ? V - 1 : t);
    if (__bfloat162float(table_relu[t * E + e]) > 0.0f) atomicAdd(demb + t * E + e, dxt[r * ld + e]);
  }
}

int embed_bwd(const float* dxt, long long ld, const int64_t* tok, const void* table_relu, float* demb, long long rows, int E, int V,
              cudaStream_t stream) {
  launch_begin("embed_bwd", stream);
  embed_bwd_kernel<<<grid_for(rows * E, 256), 256, 0, stream>>>(dxt, ld, tok, static_cast<const __nv_bfloat16*>(table_relu), demb,
                                                                rows, E, V);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- out_bf16 = x_f32 where y_bf16 > 0 else 0 (ReLU backward + operand cast) -----------------------------------
__global__ void relu_bwd_cast_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ y,
                                     __nv_bfloat16* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = __bfloat162float(y[i]) > 0.0f ? x[i] : 0.0f;
    x[i] = v;  // masked fp32 copy stays in place: the bias gradient is summed from it at full precision
    out[i] = __float2bfloat16_rn(v);
  }
}

int relu_bwd_cast(float* x, const void* y, void* out, long long n, cudaStream_t stream) {
  launch_begin("relu_bwd_cast", stream);
  relu_bwd_cast_kernel<<<grid_for(n, 256), 256, 0, stream>>>(x, static_cast<const __nv_bfloat16*>(y),
                                                             static_cast<__nv_bfloat16*>(out), n);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

// ---- dst[b, j] = sum_t src[t, b, col0 + j] (gradient of a feature that is re-fed every step) ------------------
__global__ void reduce_time_kernel(const float* __restrict__ src, long long stride_t, long long ld, int col0, float* __restrict__ dst,
                                   int T, int rows, int n) {
  const long long total = static_cast<long long>(rows) * n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / n;
    const int j = static_cast<int>(i - b * n);
    float acc = 0.0f;
    for (int t = 0; t < T; ++t) acc += src[t * stride_t + b * ld + col0 + j];
    dst[i] = acc;
  }
}

int reduce_time(const float* src, long long stride_t, long long ld, int col0, float* dst, int T, int rows, int n,
                cudaStream_t stream) {
  launch_begin("reduce_time", stream);
  reduce_time_kernel<<<grid_for(static_cast<long long>(rows) * n, 256), 256, 0, stream>>>(src, stride_t, ld, col0, dst, T, rows, n);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

}  // namespace uic
