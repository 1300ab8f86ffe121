// Fused additive-attention step: score -> softmax -> context in ONE pass over the image's
// feature tiles (reference: Attention.forward, models/AttModel.py:538-558, eight separate ATen
// kernels and a materialised (rows, L, A) tanh tensor).
//
// HBM-bound by design: per image and step the kernel reads p_att[i] (L x A bf16) and att[i]
// (L x H bf16) exactly once, with 16-byte coalesced loads, no matter how many beams share the
// image.  One CTA per (image, beam group); each warp owns the regions l = warp, warp+8, ... and
// keeps an online softmax (running max / sum) plus a partial context vector in registers; warps
// are merged once at the end through shared memory.
//
// Lane ownership: chunk c of 256 elements, lane owns elements [256c + 8*lane, +8) of both the
// A (attention hidden) and the H (feature) axis, so every global load is a full 512-byte warp
// transaction.
#include "uic_internal.h"
#include "uic_ptx.cuh"

namespace uic {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;

struct AttParams {
  const float* att_h;
  long long ld_att_h;
  const __nv_bfloat16* p_att;
  const __nv_bfloat16* att;
  const float* w_alpha;
  const float* masks;
  __nv_bfloat16* ctx_bf16;
  long long ld_ctx_bf16;
  float* ctx_f32;
  long long ld_ctx_f32;
  float* alpha;
  int beams, L, A, H;
};

template <int NB, int CA, int CH>
__global__ void __launch_bounds__(ATT_THREADS) att_step_fwd_kernel(AttParams p) {
  extern __shared__ float att_smem[];
  // layout: ctx[NB][H] | wm[NB][ATT_WARPS] | ws[NB][ATT_WARPS] | scores[NB][L] (only if alpha)
  float* s_ctx = att_smem;
  float* s_wm = s_ctx + NB * p.H;
  float* s_ws = s_wm + NB * ATT_WARPS;
  float* s_sc = s_ws + NB * ATT_WARPS;

  const int img = blockIdx.x;
  const int beam0 = blockIdx.y * NB;
  const int nb = min(NB, p.beams - beam0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L, A = p.A, H = p.H;

  for (int i = threadIdx.x; i < NB * H; i += ATT_THREADS) s_ctx[i] = 0.0f;

  // per-lane constants: alpha_net weight and the h2att projection of each beam's row
  float w[CA * 8];
  float ah[NB][CA * 8];
#pragma unroll
  for (int c = 0; c < CA; ++c) {
    const int a0 = c * 256 + lane * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool ok = a0 + k < A;
      w[c * 8 + k] = ok ? __ldg(p.w_alpha + a0 + k) : 0.0f;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const long long row = static_cast<long long>(img) * p.beams + beam0 + (j < nb ? j : 0);
        ah[j][c * 8 + k] = ok ? __ldg(p.att_h + row * p.ld_att_h + a0 + k) : 0.0f;
      }
    }
  }

  float m_run[NB], s_run[NB];
  float acc[NB][CH * 8];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    m_run[j] = -INFINITY;
    s_run[j] = 0.0f;
#pragma unroll
    for (int k = 0; k < CH * 8; ++k) acc[j][k] = 0.0f;
  }

  const __nv_bfloat16* p_img = p.p_att + static_cast<long long>(img) * L * A;
  const __nv_bfloat16* a_img = p.att + static_cast<long long>(img) * L * H;
  const float* m_img = p.masks ? p.masks + static_cast<long long>(img) * L : nullptr;

  uint4 pb[CA], ab[CH];
  auto load_row = [&](int l, uint4(&pq)[CA], uint4(&aq)[CH]) {
#pragma unroll
    for (int c = 0; c < CA; ++c) {
      const int a0 = c * 256 + lane * 8;
      pq[c] = (a0 < A) ? ldg_nc_v4(p_img + static_cast<long long>(l) * A + a0) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int h0 = c * 256 + lane * 8;
      aq[c] = (h0 < H) ? ldg_nc_v4(a_img + static_cast<long long>(l) * H + h0) : make_uint4(0, 0, 0, 0);
    }
  };

  int l = warp;
  if (l < L) load_row(l, pb, ab);
  for (; l < L; l += ATT_WARPS) {
    uint4 pn[CA], an[CH];
    const int ln = l + ATT_WARPS;
    if (ln < L) load_row(ln, pn, an);  // prefetch the next region while this one is reduced
    const float mask_l = m_img ? __ldg(m_img + l) : 1.0f;

    float pf[CA * 8];
#pragma unroll
    for (int c = 0; c < CA; ++c) {
      const uint32_t u[4] = {pb[c].x, pb[c].y, pb[c].z, pb[c].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = bf16x2_to_f2(u[q]);
        pf[c * 8 + 2 * q] = f.x;
        pf[c * 8 + 2 * q + 1] = f.y;
      }
    }
    float af[CH * 8];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const uint32_t u[4] = {ab[c].x, ab[c].y, ab[c].z, ab[c].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = bf16x2_to_f2(u[q]);
        af[c * 8 + 2 * q] = f.x;
        af[c * 8 + 2 * q + 1] = f.y;
      }
    }

#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float part = 0.0f;
#pragma unroll
      for (int k = 0; k < CA * 8; ++k) part = fmaf(w[k], tanh_approx(pf[k] + ah[j][k]), part);
      const float e = warp_sum(part);
      if (p.alpha != nullptr && lane == 0) s_sc[j * L + l] = e;
      const float m_new = fmaxf(m_run[j], e);
      const float scale = __expf(m_run[j] - m_new);  // exp(-inf) = 0 on the first region
      const float pl = __expf(e - m_new) * mask_l;
      m_run[j] = m_new;
      s_run[j] = s_run[j] * scale + pl;
#pragma unroll
      for (int k = 0; k < CH * 8; ++k) acc[j][k] = fmaf(pl, af[k], acc[j][k] * scale);
    }

    if (ln < L) {
#pragma unroll
      for (int c = 0; c < CA; ++c) pb[c] = pn[c];
#pragma unroll
      for (int c = 0; c < CH; ++c) ab[c] = an[c];
    }
  }

  // ---- merge the warps' partial softmaxes -------------------------------------------------
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      s_wm[j * ATT_WARPS + warp] = m_run[j];
      s_ws[j * ATT_WARPS + warp] = s_run[j];
    }
  }
  __syncthreads();
  float M[NB], inv_S[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < ATT_WARPS; ++q) mx = fmaxf(mx, s_wm[j * ATT_WARPS + q]);
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < ATT_WARPS; ++q) {
      const float mq = s_wm[j * ATT_WARPS + q];
      s += (mq == -INFINITY) ? 0.0f : s_ws[j * ATT_WARPS + q] * __expf(mq - mx);
    }
    M[j] = mx;
    inv_S[j] = 1.0f / s;
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (j < nb && m_run[j] != -INFINITY) {
      const float f = __expf(m_run[j] - M[j]) * inv_S[j];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int h0 = c * 256 + lane * 8;
        if (h0 < H) {
#pragma unroll
          for (int k = 0; k < 8; ++k) atomicAdd(&s_ctx[j * H + h0 + k], acc[j][c * 8 + k] * f);
        }
      }
    }
  }
  __syncthreads();

  // ---- outputs -------------------------------------------------------------------------------
  for (int j = 0; j < nb; ++j) {
    const long long row = static_cast<long long>(img) * p.beams + beam0 + j;
    for (int h = threadIdx.x * 2; h < H; h += ATT_THREADS * 2) {
      const float v0 = s_ctx[j * H + h], v1 = s_ctx[j * H + h + 1];
      if (p.ctx_bf16) *reinterpret_cast<uint32_t*>(p.ctx_bf16 + row * p.ld_ctx_bf16 + h) = f2_to_bf16x2(v0, v1);
      if (p.ctx_f32) *reinterpret_cast<float2*>(p.ctx_f32 + row * p.ld_ctx_f32 + h) = make_float2(v0, v1);
    }
    if (p.alpha != nullptr) {
      for (int q = threadIdx.x; q < L; q += ATT_THREADS) {
        const float mk = m_img ? m_img[q] : 1.0f;
        p.alpha[row * L + q] = __expf(s_sc[j * L + q] - M[j]) * mk * inv_S[j];
      }
    }
  }
}

template <int NB, int CA, int CH>
static int launch_att(const AttParams& p, int n_img, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (static_cast<size_t>(NB) * p.H + 2 * NB * ATT_WARPS + (p.alpha ? static_cast<size_t>(NB) * p.L : 0));
  auto kern = att_step_fwd_kernel<NB, CA, CH>;
  if (smem > 48 * 1024) {
    if (smem > 200 * 1024) return set_error(UIC_ERR_SHAPE, "att_step_fwd: L=%d H=%d need %zu bytes of shared memory", p.L, p.H, smem);
    UIC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  dim3 grid(n_img, (p.beams + NB - 1) / NB);
  launch_begin("att_step_fwd", stream);
  kern<<<grid, ATT_THREADS, smem, stream>>>(p);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

template <int CA, int CH>
static int dispatch_nb(const AttParams& p, int n_img, int nb_max, cudaStream_t stream) {
  // beams per pass: as many as fit the register budget (NB * (8*CA + 8*CH) accumulators per lane)
  int nb = p.beams < nb_max ? p.beams : nb_max;
  if (p.beams > nb_max) {  // balance the groups, e.g. 5 beams -> 3 + 2, 10 -> 3+3+2+2 handled as ceil
    const int groups = (p.beams + nb_max - 1) / nb_max;
    nb = (p.beams + groups - 1) / groups;
  }
  switch (nb) {
    case 1: return launch_att<1, CA, CH>(p, n_img, stream);
    case 2: return launch_att<2, CA, CH>(p, n_img, stream);
    default: return launch_att<3, CA, CH>(p, n_img, stream);
  }
}

int att_step_fwd(const float* att_h, long long ld_att_h, const void* p_att, const void* att, const float* w_alpha,
                 const float* masks, void* ctx_bf16, long long ld_ctx_bf16, float* ctx_f32, long long ld_ctx_f32, float* alpha,
                 int n_img, int beams, int L, int A, int H, cudaStream_t stream) {
  if (L <= 0 || beams <= 0) return set_error(UIC_ERR_SHAPE, "att_step_fwd: L=%d beams=%d", L, beams);
  if (A % 8 || H % 8 || A > 1024 || H > 1024)
    return set_error(UIC_ERR_SHAPE, "att_step_fwd: A=%d and H=%d must be multiples of 8 and <= 1024", A, H);
  if ((reinterpret_cast<uintptr_t>(p_att) & 15) || (reinterpret_cast<uintptr_t>(att) & 15))
    return set_error(UIC_ERR_ALIGN, "att_step_fwd: feature tiles must be 16-byte aligned");
  if ((ctx_bf16 && (ld_ctx_bf16 % 2 || (reinterpret_cast<uintptr_t>(ctx_bf16) & 3))) ||
      (ctx_f32 && (ld_ctx_f32 % 2 || (reinterpret_cast<uintptr_t>(ctx_f32) & 7))))
    return set_error(UIC_ERR_ALIGN, "att_step_fwd: ctx outputs need even pitches and 4/8-byte alignment");
  AttParams p{att_h, ld_att_h, static_cast<const __nv_bfloat16*>(p_att), static_cast<const __nv_bfloat16*>(att), w_alpha, masks,
              static_cast<__nv_bfloat16*>(ctx_bf16), ld_ctx_bf16, ctx_f32, ld_ctx_f32, alpha, beams, L, A, H};
  const int ca = (A + 255) / 256, ch = (H + 255) / 256;
  if (ca <= 1 && ch <= 1) return dispatch_nb<1, 1>(p, n_img, 3, stream);
  if (ca <= 2 && ch <= 2) return dispatch_nb<2, 2>(p, n_img, 3, stream);
  if (ca <= 2 && ch <= 4) return dispatch_nb<2, 4>(p, n_img, 2, stream);
  return dispatch_nb<4, 4>(p, n_img, 2, stream);
}

}  // namespace uic
