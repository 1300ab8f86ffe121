// Fused additive-attention step: score -> softmax -> context in ONE pass over the image's
// feature tiles (reference: Attention.forward, models/AttModel.py:538-558, eight separate ATen
// kernels and a materialised (rows, L, A) tanh tensor).
//
// HBM-bound by design: per image and step the kernel reads p_att[i] (L x A fp16) and att[i]
// (L x H bf16) exactly once, no matter how many beams share the image.
//
// Structure (third iteration, see profiles/ for the ncu captures that drove it):
//   v1 kept a region per warp in registers with one prefetch in flight: latency bound, 15 % of HBM.
//   v2 staged 8-region tiles per CTA with two block-wide phases per stage: 3x the instructions
//      (addressing, bookkeeping, barrier spinning) for no gain.
//   v3 (this file): every warp is autonomous.  Warp w of a CTA owns regions w, w+8, ... of the CTA's
//      run and streams them through its PRIVATE 4-slot shared-memory ring: lane 0 issues two bulk
//      async copies (cp.async.bulk, mbarrier complete_tx) per region, four regions ahead, so loads
//      in flight do not depend on registers or occupancy and no block-wide barrier exists in the
//      steady state.  Per region the warp computes e[j] = w . tanh(p_att + att_h[j]) for its beams
//      with packed tanh.approx.f16x2 (half the MUFU work of the fp32 form, same 2^-11 error; p_att
//      is fp16 so the add is one HADD2), an fp32 dot product, a warp-shuffle reduction, an online
//      softmax and the fp32 context accumulation.  Warps are merged once through shared memory;
//      L-splits of an image are merged by the last CTA to arrive (threadfence reduction).
#include <cuda_fp16.h>

#include "uic_internal.h"
#include "uic_ptx.cuh"

namespace uic {

constexpr int ATT_THREADS = 256;
constexpr int ATT_WARPS = ATT_THREADS / 32;
constexpr int ATT_SLOTS = 4;        // regions in flight per warp
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
  int rows_per_cta, nsplit;
};

__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc),
               "r"(bytes), "r"(bar)
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

// out[0..8) = base[a0 .. a0+8) (zero past n): two 16-byte loads when aligned, scalar otherwise.
__device__ __forceinline__ void load8(const float* __restrict__ base, int a0, int n, float* out) {
  if (a0 + 8 <= n && (reinterpret_cast<uintptr_t>(base + a0) & 15) == 0) {
    const float4 lo = __ldg(reinterpret_cast<const float4*>(base + a0));
    const float4 hi = __ldg(reinterpret_cast<const float4*>(base + a0) + 1);
    out[0] = lo.x; out[1] = lo.y; out[2] = lo.z; out[3] = lo.w;
    out[4] = hi.x; out[5] = hi.y; out[6] = hi.z; out[7] = hi.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) out[k] = (a0 + k < n) ? __ldg(base + a0 + k) : 0.0f;
  }
}

// CA = ceil(A / 256), CH = ceil(H / 256): lane owns elements [256c + 8*lane, +8) of chunk c.
template <int NB, int CA, int CH, bool EXACT>
__global__ void __launch_bounds__(ATT_THREADS, (NB * (CA + CH) <= 12) ? 2 : 1) att_step_fwd_kernel(AttParams p) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  __shared__ uint64_t s_bar[ATT_WARPS][ATT_SLOTS];
  __shared__ float s_m[NB][ATT_WARPS], s_s[NB][ATT_WARPS];
  __shared__ int s_last;

  const int L = p.L, A = p.A, H = p.H;
  const int img = blockIdx.x / p.nsplit, split = blockIdx.x - img * p.nsplit;
  const int grp = blockIdx.y, n_grp = gridDim.y;
  const int beam0 = grp * NB;
  const int nb = min(NB, p.beams - beam0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int row_begin = split * p.rows_per_cta;
  const int row_end = min(L, row_begin + p.rows_per_cta);
  const uint32_t row_bytes_p = A * 2, row_bytes_a = H * 2, slot_bytes = row_bytes_p + row_bytes_a;
  const __half* p_img = p.p_att + static_cast<long long>(img) * L * A;
  const __nv_bfloat16* a_img = p.att + static_cast<long long>(img) * L * H;
  const float* m_img = p.masks ? p.masks + static_cast<long long>(img) * L : nullptr;

  const uint32_t ring = smem_u32(att_smem) + warp * ATT_SLOTS * slot_bytes;   // this warp's private ring
  const uint32_t bar0 = smem_u32(&s_bar[warp][0]);

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < ATT_SLOTS; ++s) mbar_init(&s_bar[warp][s], 1);
    fence_barrier_init();
  }
  __syncwarp();
  auto issue = [&](int l, int s) {  // lane 0 only
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + s * 8), "r"(slot_bytes) : "memory");
    bulk_g2s(ring + s * slot_bytes, p_img + static_cast<long long>(l) * A, row_bytes_p, bar0 + s * 8);
    bulk_g2s(ring + s * slot_bytes + row_bytes_p, a_img + static_cast<long long>(l) * H, row_bytes_a, bar0 + s * 8);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < ATT_SLOTS; ++s) {
      const int l = row_begin + warp + s * ATT_WARPS;
      if (l < row_end) issue(l, s);
    }
  }

  // per-lane constant: -2 * alpha_net weight.  F = 16 exp(2 att_h) of the CTA's beams lives in shared
  // memory (NB * CA * 8 registers per lane would push the loop into local-memory spills, ncu pass 4):
  // quad (j, c, half) is stored as [lane][4] so that a warp's LDS.128 is conflict-free.
  float w[CA * 8];
  float* s_F = reinterpret_cast<float*>(att_smem + static_cast<size_t>(ATT_WARPS) * ATT_SLOTS * slot_bytes);
#pragma unroll
  for (int c = 0; c < CA; ++c) {
    load8(p.w_alpha, c * 256 + lane * 8, A, &w[c * 8]);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[c * 8 + k] *= -2.0f;
  }
  for (int jc = warp; jc < NB * CA; jc += ATT_WARPS) {
    const int j = jc / CA, c = jc - j * CA;
    const long long row = static_cast<long long>(img) * p.beams + beam0 + (j < nb ? j : 0);
    float x[8];
    load8(p.att_h + row * p.ld_att_h, c * 256 + lane * 8, A, x);
    *reinterpret_cast<float4*>(s_F + (jc * 2 + 0) * 128 + lane * 4) = make_float4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<float4*>(s_F + (jc * 2 + 1) * 128 + lane * 4) = make_float4(x[4], x[5], x[6], x[7]);
  }
  __syncthreads();

  float m_run[NB], s_run[NB];
  float acc[NB][CH * 8];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    m_run[j] = -INFINITY;
    s_run[j] = 0.0f;
#pragma unroll
    for (int k = 0; k < CH * 8; ++k) acc[j][k] = 0.0f;
  }

  int slot = 0;
  uint32_t parity = 0;
  for (int l = row_begin + warp; l < row_end; l += ATT_WARPS) {
    mbar_wait_addr(bar0 + slot * 8, parity);
    const uint32_t base = ring + slot * slot_bytes;
    uint4 q[CA], av[CH];
#pragma unroll
    for (int c = 0; c < CA; ++c) {
      const int a0 = c * 256 + lane * 8;
      q[c] = make_uint4(0, 0, 0, 0);
      if (EXACT || a0 < A) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q[c].x), "=r"(q[c].y), "=r"(q[c].z), "=r"(q[c].w) : "r"(base + a0 * 2));
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int h0 = c * 256 + lane * 8;
      av[c] = make_uint4(0, 0, 0, 0);
      if (EXACT || h0 < H)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(av[c].x), "=r"(av[c].y), "=r"(av[c].z), "=r"(av[c].w) : "r"(base + row_bytes_p + h0 * 2));
    }
    const float mask_l = m_img ? __ldg(m_img + l) : 1.0f;

    // ---- scores ------------------------------------------------------------------------------
    // tanh(p + a) = 1 - 2 / (E F + 1); the constant sum(w) drops out of the softmax, so the score is
    // e = sum_a (-2 w_a) / (E_a F_a + 1).  One reciprocal serves a PAIR of units:
    //   w1/d1 + w2/d2 = (w1 d2 + w2 d1) / (d1 d2)          (d <= 1 + 65504 * F: no fp32 overflow for |att_h| < 30)
    float Ef[CA * 8];
#pragma unroll
    for (int c = 0; c < CA; ++c) {
      const uint32_t u[4] = {q[c].x, q[c].y, q[c].z, q[c].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&u[k]));
        Ef[c * 8 + 2 * k] = t.x;
        Ef[c * 8 + 2 * k + 1] = t.y;
      }
    }
    float e[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float Fj[CA * 8];
#pragma unroll
      for (int c = 0; c < CA; ++c) {
        const float4 lo = *reinterpret_cast<const float4*>(s_F + ((j * CA + c) * 2 + 0) * 128 + lane * 4);
        const float4 hi = *reinterpret_cast<const float4*>(s_F + ((j * CA + c) * 2 + 1) * 128 + lane * 4);
        Fj[c * 8 + 0] = lo.x; Fj[c * 8 + 1] = lo.y; Fj[c * 8 + 2] = lo.z; Fj[c * 8 + 3] = lo.w;
        Fj[c * 8 + 4] = hi.x; Fj[c * 8 + 5] = hi.y; Fj[c * 8 + 6] = hi.z; Fj[c * 8 + 7] = hi.w;
      }
      float p0 = 0.0f, p1 = 0.0f;
#pragma unroll
      for (int k = 0; k < CA * 8; k += 2) {
        const float d1 = fmaf(Ef[k], Fj[k], 1.0f);
        const float d2 = fmaf(Ef[k + 1], Fj[k + 1], 1.0f);
        const float r = rcp_approx(d1 * d2);
        const float num = fmaf(w[k + 1], d1, w[k] * d2);
        if ((k & 2) == 0) p0 = fmaf(r, num, p0); else p1 = fmaf(r, num, p1);
      }
      e[j] = p0 + p1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < NB; ++j) e[j] += __shfl_xor_sync(0xffffffffu, e[j], o);
    }
    if (p.alpha != nullptr && lane < nb) {
      float ej = e[0];
#pragma unroll
      for (int j = 1; j < NB; ++j) ej = (lane == j) ? e[j] : ej;
      p.alpha[(static_cast<long long>(img) * p.beams + beam0 + lane) * L + l] = ej;  // raw score, normalised at the end
    }

    // ---- online softmax + context ---------------------------------------------------------------
    float af[CH * 8];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const uint32_t u[4] = {av[c].x, av[c].y, av[c].z, av[c].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        af[c * 8 + 2 * k] = __uint_as_float(u[k] << 16);
        af[c * 8 + 2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (e[j] > m_run[j]) {  // warp-uniform: rare after the first few regions
        const float scale = __expf(m_run[j] - e[j]);
        m_run[j] = e[j];
        s_run[j] *= scale;
#pragma unroll
        for (int k = 0; k < CH * 8; ++k) acc[j][k] *= scale;
      }
      const float pl = __expf(e[j] - m_run[j]) * mask_l;
      s_run[j] += pl;
#pragma unroll
      for (int k = 0; k < CH * 8; ++k) acc[j][k] = fmaf(pl, af[k], acc[j][k]);
    }
    // Refill the slot only now: the FMAs above consumed every lane's registers loaded from it, so no
    // shared-memory read of this slot can still be in flight when the async copy lands.
    __syncwarp();
    if (lane == 0) {
      const int ln = l + ATT_SLOTS * ATT_WARPS;
      if (ln < row_end) issue(ln, slot);
    }
    if (++slot == ATT_SLOTS) {
      slot = 0;
      parity ^= 1;
    }
  }

  // ---- merge the warps of this CTA through shared memory (the rings are idle now) ----------------------
  __syncthreads();
  float* s_acc = reinterpret_cast<float*>(att_smem);  // [ATT_WARPS][NB][H]
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      s_m[j][warp] = m_run[j];
      s_s[j][warp] = s_run[j];
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int h0 = c * 256 + lane * 8;
      if (h0 < H) {
        float* dst = s_acc + (static_cast<size_t>(warp) * NB + j) * H + h0;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[j][c * 8], acc[j][c * 8 + 1], acc[j][c * 8 + 2], acc[j][c * 8 + 3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[j][c * 8 + 4], acc[j][c * 8 + 5], acc[j][c * 8 + 6], acc[j][c * 8 + 7]);
      }
    }
  __syncthreads();
  float Mj[NB], Sj[NB], fw[NB][ATT_WARPS];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    float mx = -INFINITY;
#pragma unroll
    for (int q2 = 0; q2 < ATT_WARPS; ++q2) mx = fmaxf(mx, s_m[j][q2]);
    float ss = 0.0f;
#pragma unroll
    for (int q2 = 0; q2 < ATT_WARPS; ++q2) {
      const float mq = s_m[j][q2];
      fw[j][q2] = (mq == -INFINITY) ? 0.0f : __expf(mq - mx);
      ss += s_s[j][q2] * fw[j][q2];
    }
    Mj[j] = mx;
    Sj[j] = ss;
  }
  // thread `tid` owns the column pairs col = 2*(tid + 256*i)
  constexpr int CHP = (CH + 1) / 2;
  float out[NB][CHP * 2];
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int i = 0; i < CHP; ++i) {
      const int col = 2 * (tid + ATT_THREADS * i);
      float2 v = make_float2(0.0f, 0.0f);
      if (col < H) {
#pragma unroll
        for (int q2 = 0; q2 < ATT_WARPS; ++q2) {
          const float2 a = *reinterpret_cast<const float2*>(s_acc + (static_cast<size_t>(q2) * NB + j) * H + col);
          v.x = fmaf(a.x, fw[j][q2], v.x);
          v.y = fmaf(a.y, fw[j][q2], v.y);
        }
      }
      out[j][2 * i] = v.x;
      out[j][2 * i + 1] = v.y;
    }

  if (p.nsplit > 1) {
    // ---- publish this split's partial, last arriver merges ------------------------------------------
    float* part = p.ws_partial + ((static_cast<long long>(img) * n_grp + grp) * p.nsplit + split) * NB * (H + 2);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#pragma unroll
      for (int i = 0; i < CHP; ++i) {
        const int col = 2 * (tid + ATT_THREADS * i);
        if (col < H) *reinterpret_cast<float2*>(part + j * (H + 2) + col) = make_float2(out[j][2 * i], out[j][2 * i + 1]);
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
          const float mk = __ldcg(base + (k * NB + j) * (H + 2) + H);
          f[k] = (mk == -INFINITY) ? 0.0f : __expf(mk - mx);
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
        out[j][2 * i] = v.x;
        out[j][2 * i + 1] = v.y;
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
          const float v0 = out[j][2 * i] * inv, v1 = out[j][2 * i + 1] * inv;
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
  int rows_per_cta, nsplit;
};

static AttPlan make_plan(int L, int n_img) {
  // Splitting an image's regions over several CTAs costs a fixed prologue/epilogue per CTA (about half
  // of the kernel at 7 regions per warp, ncu pass 4), so it is only used to create parallelism for small
  // batches: aim for ~2 CTAs per SM, in multiples of 8 regions so the 8 warps of a CTA get equal shares.
  int target = (2 * 148 + n_img - 1) / n_img;
  const int by_len = (L + 24) / 49;
  target = target > by_len ? by_len : target;
  target = target < 1 ? 1 : (target > ATT_MAX_SPLIT ? ATT_MAX_SPLIT : target);
  const int groups8 = (L + 7) / 8;
  const int k = (groups8 + target - 1) / target;
  AttPlan pl;
  pl.rows_per_cta = 8 * k;
  pl.nsplit = (L + pl.rows_per_cta - 1) / pl.rows_per_cta;
  return pl;
}

static int beams_per_group(int beams) {
  if (beams <= 3) return beams;
  const int groups = (beams + 2) / 3;
  return (beams + groups - 1) / groups;
}

long long att_step_workspace_bytes(int n_img, int beams, int L, int H) {
  const AttPlan pl = make_plan(L, n_img);
  const int nb = beams_per_group(beams);
  const int groups = (beams + nb - 1) / nb;
  const long long counters = ((static_cast<long long>(n_img) * groups * 4 + 255) / 256) * 256;
  const long long partial = pl.nsplit > 1 ? static_cast<long long>(n_img) * groups * pl.nsplit * nb * (H + 2) * 4 : 0;
  return counters + partial;
}

template <int NB, int CA, int CH, bool EXACT>
static int launch_att(AttParams& p, int n_img, const AttPlan& pl, cudaStream_t stream) {
  const size_t ring = static_cast<size_t>(ATT_WARPS) * ATT_SLOTS * (p.A + p.H) * 2 + static_cast<size_t>(NB) * CA * 256 * 4;
  const size_t merge = static_cast<size_t>(ATT_WARPS) * NB * p.H * 4;
  const size_t smem = ring > merge ? ring : merge;
  auto kern = att_step_fwd_kernel<NB, CA, CH, EXACT>;
  if (smem > 200 * 1024) return set_error(UIC_ERR_SHAPE, "att_step_fwd: A=%d H=%d need %zu bytes of shared memory", p.A, p.H, smem);
  if (smem > 40 * 1024)  // dynamic + the kernel's static shared memory may exceed the 48 KB default
    UIC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  dim3 grid(n_img * pl.nsplit, (p.beams + NB - 1) / NB);
  launch_begin("att_step_fwd", stream);
  kern<<<grid, ATT_THREADS, smem, stream>>>(p);
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

template <int CA, int CH>
static int dispatch_nb(AttParams& p, int n_img, const AttPlan& pl, int nb_max, cudaStream_t stream) {
  int nb = beams_per_group(p.beams);
  nb = nb > nb_max ? nb_max : nb;
  const bool exact = (p.A == 256 * CA) && (p.H == 256 * CH);
  if (exact) {
    switch (nb) {
      case 1: return launch_att<1, CA, CH, true>(p, n_img, pl, stream);
      case 2: return launch_att<2, CA, CH, true>(p, n_img, pl, stream);
      default: return launch_att<3, CA, CH, true>(p, n_img, pl, stream);
    }
  }
  switch (nb) {
    case 1: return launch_att<1, CA, CH, false>(p, n_img, pl, stream);
    case 2: return launch_att<2, CA, CH, false>(p, n_img, pl, stream);
    default: return launch_att<3, CA, CH, false>(p, n_img, pl, stream);
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
  const AttPlan pl = make_plan(L, n_img);
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
  p.rows_per_cta = pl.rows_per_cta;
  p.nsplit = pl.nsplit;
  const int ca = (A + 255) / 256, ch = (H + 255) / 256;
  if (ca <= 1 && ch <= 1) return dispatch_nb<1, 1>(p, n_img, pl, 3, stream);
  if (ca <= 2 && ch <= 2) return dispatch_nb<2, 2>(p, n_img, pl, 3, stream);
  if (ca <= 2 && ch <= 4) return dispatch_nb<2, 4>(p, n_img, pl, 3, stream);
  return dispatch_nb<4, 4>(p, n_img, pl, 3, stream);
}

}  // namespace uic
