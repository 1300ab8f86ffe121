// Fused additive-attention step: score -> softmax -> context in ONE pass over the image's
// feature tiles (reference: Attention.forward, models/AttModel.py:538-558, eight separate ATen
// kernels and a materialised (rows, L, A) tanh tensor).
//
// HBM-bound by design: per image and step the kernel reads p_att[i] (L x A fp16) and att[i]
// (L x H bf16) exactly once, no matter how many beams share the image.
//
// Structure (round-1 pass 2, after the first ncu capture showed the register-resident version was
// latency bound at 15 % of HBM peak with 8 warps/SM):
//   * grid = (image x L-split, beam group); each CTA owns a contiguous run of regions and streams
//     them through a 3-stage shared-memory ring filled by bulk async copies (cp.async.bulk +
//     mbarrier complete_tx), so the loads in flight do not depend on registers or occupancy;
//   * per stage (<= 8 regions): phase 1, one warp per region: e[r][j] = w . tanh(p_att[r] + att_h[j])
//     with packed tanh.approx.f16x2 (half the MUFU work of the fp32 form, same 2^-11 error) and an
//     fp32 dot product; phase 2, one thread per pair of feature columns: online-softmax update of
//     ctx[j][col] over the stage's regions (6 accumulator registers instead of 48);
//   * L-splits of an image are merged by the last CTA to arrive (threadfence reduction) from a
//     small fp32 workspace; no second launch.
#include <cuda_fp16.h>

#include "uic_internal.h"
#include "uic_ptx.cuh"

namespace uic {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_STAGE_ROWS = 8;   // capacity of one ring slot (regions); the host may use 6..8
constexpr int ATT_NSTAGES = 3;
constexpr int ATT_MAX_SPLIT = 8;

struct AttParams {
  const float* att_h;
  long long ld_att_h;
  const __half* p_att;
  const __nv_bfloat16* att;
  const float* w_alpha;
  const float* masks;
  __nv_bfloat16* ctx_bf16;
  long long ld_ctx_bf16;
  float* ctx_f32;
  long long ld_ctx_f32;
  float* alpha;
  float* ws_partial;   // [img][group][split][NB][H + 2]
  int* ws_counter;     // [img][group], zero between launches
  int beams, L, A, H;
  int rows_per_stage, stages_per_cta, total_stages, nsplit;
};

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t tanh_f16x2(uint32_t x) {
  uint32_t y;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t hadd2_u32(uint32_t a, uint32_t b) {
  __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <int NB, int CA, int CHP>
__global__ void __launch_bounds__(ATT_THREADS, (CA <= 2 && CHP <= 1) ? 3 : 1) att_step_fwd_kernel(AttParams p) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  __shared__ uint64_t full_bar[ATT_NSTAGES];
  __shared__ float s_e[ATT_STAGE_ROWS * NB];
  __shared__ float s_wp[ATT_WARPS][ATT_STAGE_ROWS * NB];
  __shared__ float s_wscale[ATT_WARPS][NB];
  __shared__ int s_last;

  const int L = p.L, A = p.A, H = p.H;
  const int img = blockIdx.x / p.nsplit, split = blockIdx.x - img * p.nsplit;
  const int grp = blockIdx.y, n_grp = gridDim.y;
  const int beam0 = grp * NB;
  const int nb = min(NB, p.beams - beam0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int rs = p.rows_per_stage;
  const int stage0 = split * p.stages_per_cta;
  const int n_iters = min(p.stages_per_cta, p.total_stages - stage0);
  const uint32_t stage_bytes = ATT_STAGE_ROWS * (A + H) * 2;
  const __half* p_img = p.p_att + static_cast<long long>(img) * L * A;
  const __nv_bfloat16* a_img = p.att + static_cast<long long>(img) * L * H;
  const float* m_img = p.masks ? p.masks + static_cast<long long>(img) * L : nullptr;

  auto issue = [&](int it) {  // one thread: arm the slot's barrier and start both bulk copies
    const int s = it % ATT_NSTAGES;
    const int l0 = (stage0 + it) * rs;
    const int n = min(rs, L - l0);
    uint8_t* slot = att_smem + s * stage_bytes;
    const uint32_t bp = n * A * 2, ba = n * H * 2;
    mbar_arrive_expect_tx(&full_bar[s], bp + ba);
    bulk_g2s(slot, p_img + static_cast<long long>(l0) * A, bp, &full_bar[s]);
    bulk_g2s(slot + ATT_STAGE_ROWS * A * 2, a_img + static_cast<long long>(l0) * H, ba, &full_bar[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < ATT_NSTAGES; ++s) mbar_init(&full_bar[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0)
    for (int it = 0; it < ATT_NSTAGES && it < n_iters; ++it) issue(it);

  // per-lane constants for phase 1: alpha_net weight (fp32) and att_h of each beam as half2
  float w[CA * 8];
  uint32_t ah2[NB][CA * 4];
#pragma unroll
  for (int c = 0; c < CA; ++c) {
    const int a0 = c * 256 + lane * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) w[c * 8 + k] = (a0 + k < A) ? __ldg(p.w_alpha + a0 + k) : 0.0f;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const long long row = static_cast<long long>(img) * p.beams + beam0 + (j < nb ? j : 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int a = a0 + 2 * k;
        const float x0 = (a < A) ? __ldg(p.att_h + row * p.ld_att_h + a) : 0.0f;
        const float x1 = (a + 1 < A) ? __ldg(p.att_h + row * p.ld_att_h + a + 1) : 0.0f;
        __half2 h2 = __floats2half2_rn(x0, x1);
        ah2[j][c * 4 + k] = *reinterpret_cast<uint32_t*>(&h2);
      }
    }
  }

  // online softmax state: lane l < NB*8 tracks beam j = l % NB (replicated in every warp)
  const int my_j = lane % NB;
  float m_run = -INFINITY, s_run = 0.0f;
  float acc[NB][CHP * 2];
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int k = 0; k < CHP * 2; ++k) acc[j][k] = 0.0f;

  for (int it = 0; it < n_iters; ++it) {
    const int s = it % ATT_NSTAGES;
    const int l0 = (stage0 + it) * rs;
    const int n = min(rs, L - l0);
    const uint8_t* slot = att_smem + s * stage_bytes;
    mbar_wait(&full_bar[s], (it / ATT_NSTAGES) & 1);

    // ---- phase 1: scores, one warp per region -----------------------------------------------
    if (warp < n) {
      const uint8_t* prow = slot + static_cast<size_t>(warp) * A * 2;
      uint4 q[CA];
#pragma unroll
      for (int c = 0; c < CA; ++c) {
        const int a0 = c * 256 + lane * 8;
        q[c] = (a0 < A) ? *reinterpret_cast<const uint4*>(prow + a0 * 2) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float part = 0.0f;
#pragma unroll
        for (int c = 0; c < CA; ++c) {
          const uint32_t u[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t t2 = tanh_f16x2(hadd2_u32(u[k], ah2[j][c * 4 + k]));
            const float2 t = __half22float2(*reinterpret_cast<__half2*>(&t2));
            part = fmaf(w[c * 8 + 2 * k], t.x, part);
            part = fmaf(w[c * 8 + 2 * k + 1], t.y, part);
          }
        }
        const float e = warp_sum(part);
        if (lane == 0) {
          s_e[warp * NB + j] = e;
          if (p.alpha != nullptr && j < nb)
            p.alpha[(static_cast<long long>(img) * p.beams + beam0 + j) * L + l0 + warp] = e;  // raw score, normalised at the end
        }
      }
    }
    __syncthreads();

    // ---- softmax bookkeeping, replicated per warp (no block-wide sync needed) -------------------
    {
      const int r = lane / NB;
      const bool valid = lane < n * NB;
      float m_st = -INFINITY;
      for (int rr = 0; rr < n; ++rr) m_st = fmaxf(m_st, s_e[rr * NB + my_j]);
      const float m_new = fmaxf(m_run, m_st);
      const float scale = __expf(m_run - m_new);  // 0 on the first stage (m_run = -inf)
      const float mk = (valid && m_img) ? m_img[l0 + r] : 1.0f;
      const float pv = valid ? __expf(s_e[lane] - m_new) * mk : 0.0f;
      if (lane < ATT_STAGE_ROWS * NB) s_wp[warp][lane] = pv;
      if (lane < NB) s_wscale[warp][lane] = scale;
      __syncwarp();
      float s_st = 0.0f;
      for (int rr = 0; rr < n; ++rr) s_st += s_wp[warp][rr * NB + my_j];
      s_run = s_run * scale + s_st;
      m_run = m_new;
    }

    // ---- phase 2: context accumulation, one thread per pair of feature columns ---------------------
    {
      float sc[NB];
      bool rescale = false;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        sc[j] = s_wscale[warp][j];
        rescale |= sc[j] != 1.0f;
      }
      if (rescale) {
#pragma unroll
        for (int j = 0; j < NB; ++j)
#pragma unroll
          for (int k = 0; k < CHP * 2; ++k) acc[j][k] *= sc[j];
      }
      const uint8_t* abase = slot + ATT_STAGE_ROWS * A * 2;
#pragma unroll
      for (int i = 0; i < CHP; ++i) {
        const int col = 2 * (tid + ATT_THREADS * i);
        if (col < H) {
          for (int rr = 0; rr < n; ++rr) {
            const float2 a = bf16x2_to_f2(*reinterpret_cast<const uint32_t*>(abase + (static_cast<size_t>(rr) * H + col) * 2));
#pragma unroll
            for (int j = 0; j < NB; ++j) {
              const float pj = s_wp[warp][rr * NB + j];
              acc[j][2 * i] = fmaf(pj, a.x, acc[j][2 * i]);
              acc[j][2 * i + 1] = fmaf(pj, a.y, acc[j][2 * i + 1]);
            }
          }
        }
      }
    }
    __syncthreads();  // every warp is done with slot s (and with s_e)
    if (tid == 0 && it + ATT_NSTAGES < n_iters) issue(it + ATT_NSTAGES);
  }

  // beam j's running (max, sum) live in lane j of every warp
  float Mj[NB], Sj[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    Mj[j] = __shfl_sync(0xffffffffu, m_run, j);
    Sj[j] = __shfl_sync(0xffffffffu, s_run, j);
  }

  if (p.nsplit > 1) {
    // ---- publish this split's partial, last arriver merges ------------------------------------------
    float* part = p.ws_partial + ((static_cast<long long>(img) * n_grp + grp) * p.nsplit + split) * NB * (H + 2);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int i = 0; i < CHP; ++i) {
        const int col = 2 * (tid + ATT_THREADS * i);
        if (col < H) *reinterpret_cast<float2*>(part + j * (H + 2) + col) = make_float2(acc[j][2 * i], acc[j][2 * i + 1]);
      }
      if (tid == 0) {
        part[j * (H + 2) + H] = Mj[j];
        part[j * (H + 2) + H + 1] = Sj[j];
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int old = atomicAdd(p.ws_counter + img * n_grp + grp, 1);
      s_last = (old == p.nsplit - 1);
      if (s_last) p.ws_counter[img * n_grp + grp] = 0;  // ready for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const float* base = p.ws_partial + (static_cast<long long>(img) * n_grp + grp) * p.nsplit * NB * (H + 2);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float mx = -INFINITY;
      for (int k = 0; k < p.nsplit; ++k) mx = fmaxf(mx, __ldcg(base + (k * NB + j) * (H + 2) + H));
      float ssum = 0.0f;
      float f[ATT_MAX_SPLIT];
#pragma unroll
      for (int k = 0; k < ATT_MAX_SPLIT; ++k) {
        f[k] = 0.0f;
        if (k < p.nsplit) {
          f[k] = __expf(__ldcg(base + (k * NB + j) * (H + 2) + H) - mx);
          ssum += __ldcg(base + (k * NB + j) * (H + 2) + H + 1) * f[k];
        }
      }
#pragma unroll
      for (int i = 0; i < CHP; ++i) {
        const int col = 2 * (tid + ATT_THREADS * i);
        float2 v = make_float2(0.0f, 0.0f);
        if (col < H) {
#pragma unroll
          for (int k = 0; k < ATT_MAX_SPLIT; ++k) {
            if (k < p.nsplit) {
              const float2 a = __ldcg(reinterpret_cast<const float2*>(base + (k * NB + j) * (H + 2) + col));
              v.x = fmaf(a.x, f[k], v.x);
              v.y = fmaf(a.y, f[k], v.y);
            }
          }
        }
        acc[j][2 * i] = v.x;
        acc[j][2 * i + 1] = v.y;
      }
      Mj[j] = mx;
      Sj[j] = ssum;
    }
  }

  // ---- outputs -----------------------------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (j < nb) {
      const long long row = static_cast<long long>(img) * p.beams + beam0 + j;
      const float inv = 1.0f / Sj[j];
#pragma unroll
      for (int i = 0; i < CHP; ++i) {
        const int col = 2 * (tid + ATT_THREADS * i);
        if (col < H) {
          const float v0 = acc[j][2 * i] * inv, v1 = acc[j][2 * i + 1] * inv;
          if (p.ctx_bf16) *reinterpret_cast<uint32_t*>(p.ctx_bf16 + row * p.ld_ctx_bf16 + col) = f2_to_bf16x2(v0, v1);
          if (p.ctx_f32) *reinterpret_cast<float2*>(p.ctx_f32 + row * p.ld_ctx_f32 + col) = make_float2(v0, v1);
        }
      }
      if (p.alpha != nullptr) {  // raw scores (possibly written by the other splits) -> weights
        __syncthreads();
        for (int l = tid; l < L; l += ATT_THREADS) {
          const float mk = m_img ? m_img[l] : 1.0f;
          p.alpha[row * L + l] = __expf(__ldcg(p.alpha + row * L + l) - Mj[j]) * mk * inv;
        }
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------
struct AttPlan {
  int rows_per_stage, total_stages, nsplit, stages_per_cta;
};

static AttPlan make_plan(int L) {
  AttPlan pl;
  int best_rs = ATT_STAGE_ROWS, best_waste = 1 << 30;
  for (int rs = ATT_STAGE_ROWS; rs >= 6; --rs) {  // fewest idle score warps in the last stage
    const int waste = ((L + rs - 1) / rs) * rs - L;
    if (waste < best_waste) {
      best_waste = waste;
      best_rs = rs;
    }
  }
  if (L < 6) best_rs = L;
  pl.rows_per_stage = best_rs;
  pl.total_stages = (L + best_rs - 1) / best_rs;
  int nsplit = (pl.total_stages + 3) / 7;  // about 7 stages (~50 regions) per CTA
  nsplit = nsplit < 1 ? 1 : (nsplit > ATT_MAX_SPLIT ? ATT_MAX_SPLIT : nsplit);
  pl.stages_per_cta = (pl.total_stages + nsplit - 1) / nsplit;
  pl.nsplit = (pl.total_stages + pl.stages_per_cta - 1) / pl.stages_per_cta;
  return pl;
}

static int beams_per_group(int beams) {
  if (beams <= 3) return beams;
  const int groups = (beams + 2) / 3;
  return (beams + groups - 1) / groups;
}

long long att_step_workspace_bytes(int n_img, int beams, int L, int H) {
  const AttPlan pl = make_plan(L);
  const int nb = beams_per_group(beams);
  const int groups = (beams + nb - 1) / nb;
  const long long counters = ((static_cast<long long>(n_img) * groups * 4 + 255) / 256) * 256;
  const long long partial = pl.nsplit > 1 ? static_cast<long long>(n_img) * groups * pl.nsplit * nb * (H + 2) * 4 : 0;
  return counters + partial;
}

template <int NB, int CA, int CHP>
static int launch_att(AttParams& p, int n_img, const AttPlan& pl, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(ATT_NSTAGES) * ATT_STAGE_ROWS * (p.A + p.H) * 2;
  auto kern = att_step_fwd_kernel<NB, CA, CHP>;
  if (smem > 200 * 1024) return set_error(UIC_ERR_SHAPE, "att_step_fwd: A=%d H=%d need %zu bytes of shared memory", p.A, p.H, smem);
  if (smem > 48 * 1024) UIC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(n_img * pl.nsplit, (p.beams + NB - 1) / NB);
  launch_begin("att_step_fwd", stream);
  kern<<<grid, ATT_THREADS, smem, stream>>>(p);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

template <int CA, int CHP>
static int dispatch_nb(AttParams& p, int n_img, const AttPlan& pl, cudaStream_t stream) {
  switch (beams_per_group(p.beams)) {
    case 1: return launch_att<1, CA, CHP>(p, n_img, pl, stream);
    case 2: return launch_att<2, CA, CHP>(p, n_img, pl, stream);
    default: return launch_att<3, CA, CHP>(p, n_img, pl, stream);
  }
}

int att_step_fwd(const float* att_h, long long ld_att_h, const void* p_att, const void* att, const float* w_alpha,
                 const float* masks, void* ctx_bf16, long long ld_ctx_bf16, float* ctx_f32, long long ld_ctx_f32, float* alpha,
                 void* workspace, long long workspace_bytes, int n_img, int beams, int L, int A, int H, cudaStream_t stream) {
  if (L <= 0 || beams <= 0) return set_error(UIC_ERR_SHAPE, "att_step_fwd: L=%d beams=%d", L, beams);
  if (A % 8 || H % 8 || A > 1024 || H > 1024)
    return set_error(UIC_ERR_SHAPE, "att_step_fwd: A=%d and H=%d must be multiples of 8 and <= 1024", A, H);
  if ((reinterpret_cast<uintptr_t>(p_att) & 15) || (reinterpret_cast<uintptr_t>(att) & 15))
    return set_error(UIC_ERR_ALIGN, "att_step_fwd: feature tiles must be 16-byte aligned");
  if ((ctx_bf16 && (ld_ctx_bf16 % 2 || (reinterpret_cast<uintptr_t>(ctx_bf16) & 3))) ||
      (ctx_f32 && (ld_ctx_f32 % 2 || (reinterpret_cast<uintptr_t>(ctx_f32) & 7))))
    return set_error(UIC_ERR_ALIGN, "att_step_fwd: ctx outputs need even pitches and 4/8-byte alignment");
  const AttPlan pl = make_plan(L);
  const long long need = att_step_workspace_bytes(n_img, beams, L, H);
  if (workspace == nullptr || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return set_error(UIC_ERR_ARG, "att_step_fwd: workspace of %lld bytes (16-byte aligned, zeroed once) required, got %lld", need,
                     workspace_bytes);
  const int nbg = beams_per_group(beams);
  const int groups = (beams + nbg - 1) / nbg;
  AttParams p{};
  p.att_h = att_h;
  p.ld_att_h = ld_att_h;
  p.p_att = static_cast<const __half*>(p_att);
  p.att = static_cast<const __nv_bfloat16*>(att);
  p.w_alpha = w_alpha;
  p.masks = masks;
  p.ctx_bf16 = static_cast<__nv_bfloat16*>(ctx_bf16);
  p.ld_ctx_bf16 = ld_ctx_bf16;
  p.ctx_f32 = ctx_f32;
  p.ld_ctx_f32 = ld_ctx_f32;
  p.alpha = alpha;
  // counters first (they must stay zero between launches), partials after them
  p.ws_counter = static_cast<int*>(workspace);
  p.ws_partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + ((static_cast<long long>(n_img) * groups * 4 + 255) / 256) * 256);
  p.beams = beams;
  p.L = L;
  p.A = A;
  p.H = H;
  p.rows_per_stage = pl.rows_per_stage;
  p.stages_per_cta = pl.stages_per_cta;
  p.total_stages = pl.total_stages;
  p.nsplit = pl.nsplit;
  const int ca = (A + 255) / 256, chp = (H + 511) / 512;
  if (ca <= 1 && chp <= 1) return dispatch_nb<1, 1>(p, n_img, pl, stream);
  if (ca <= 2 && chp <= 1) return dispatch_nb<2, 1>(p, n_img, pl, stream);
  if (ca <= 2 && chp <= 2) return dispatch_nb<2, 2>(p, n_img, pl, stream);
  return dispatch_nb<4, 2>(p, n_img, pl, stream);
}

}  // namespace uic
