// Attention step, seventh structure: the v6 pipeline (attention.cu: persistent homogeneous warps, 3-stage ring of
// 16-region batches, exponential-form scoring, online softmax, mma.sync context) with
//   * a BATCH-balanced partition: CTA c owns the batches [c n / ctas, (c + 1) n / ctas) of the flat batch list, so
//     every CTA slot of the chip carries the same load (v6 gives whole jobs to CTAs: 256 jobs on 296 slots leave
//     40 SMs half empty and the other 108 set the makespan).  A job cut by a range boundary is finished by its
//     OWNER, the CTA that holds its first batch: every later CTA that starts inside the job publishes its partial
//     (max, sum, accumulators) record, per warp, behind a release flag; the owner's warps acquire the flags of their
//     own column sets at the end of their range -- by then the followers, which treat that job FIRST, are long
//     done -- merge and write the result.  All CTAs are co-resident (ctas <= slots), so the wait cannot deadlock;
//   * a lean per-batch path: the attention width is a template constant (A = 256 CA), the producer derives its
//     coordinates from the batch index when it is its turn (no cursor triple carried by every warp), rows past the
//     image are published as -inf scores by the scoring side (no guards in the softmax), masks / alpha output are
//     compiled out unless asked for (AUX), no debug timeline.
// Reference: Attention.forward, models/AttModel.py:538-558.
#include "uic_attention.cuh"

namespace uic {

// Compile-time knobs of the occupancy experiments recorded in profiles/r2_att_v7_experiments.txt (-DV7_SLOTS_N=2: 2-slot
// rings, +1.5 %; -DV7_MIN_CTAS=3: an 80-register build, +29 % at equal occupancy).  The defaults are what ships.
#ifndef V7_SLOTS_N
#define V7_SLOTS_N 3
#endif
#ifndef V7_MIN_CTAS
#define V7_MIN_CTAS 2
#endif
constexpr int V7_SLOTS = V7_SLOTS_N;  // depth of the p_att-row ring and of the att-box ring
constexpr int V7_E_SLOTS = 4;  // score buffers: the warps of a CTA drift up to two batches apart

__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_addr(bar, parity)) return;
  int n = 0;
  // (a failed try_wait comes back within a few tens of cycles whatever its suspend hint says: without the explicit sleep the
  //  warps that run ahead of their CTA issued 18 % of all instructions of the kernel polling the scored barrier)
  do {
    __nanosleep(96);
    if (++n > (1 << 23)) {  // a protocol bug becomes a launch failure instead of a hung GPU
      printf("uic: att_step_fwd_v7 mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  } while (!mbar_try_wait_addr(bar, parity));
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// CA = A / 256: lane owns units [256c + 128h + 4 lane, +4), h = 0, 1.  MT = 16-column context tiles per warp.
// AUX: region masks and/or the alpha output are present.
template <int NB, int CA, int MT, bool AUX>
__global__ void __launch_bounds__(ATT_THREADS, (CA <= 2 && MT <= 4) ? V7_MIN_CTAS : 1)
att_step_fwd_v7_kernel(const __grid_constant__ CUtensorMap tmap_att, const __grid_constant__ AttParams p) {
  extern __shared__ uint8_t att_smem_raw[];
  constexpr int A = 256 * CA;
  constexpr float LOG2E = 1.4426950408889634f;

  pdl_launch_dependents();
  const int L = p.L, H = p.H, nbpi = p.nbpi;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = static_cast<int>(static_cast<long long>(blockIdx.x) * p.n_batches / gridDim.x);
  const int nloc = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.n_batches / gridDim.x) - b0;
  const int job0 = b0 / nbpi, kb0 = b0 - job0 * nbpi;
  const int f_bufs = p.f_bufs;
  const int n_slabs = H >> 6;

  // shared memory: [3] att boxes (slot = 16 regions x H, 128-byte swizzled 64-column slabs) | [3] p_att rows (16 x A bf16) |
  // att_h buffers [f_bufs][NB][A] fp32 | scores [4][NB][16] fp32 | mbarriers full_p / full_a / scored [3] | counters.
  // The p_att rows and the att boxes of a batch live in SEPARATE rings: the rows are free again once every warp has
  // scored the batch, the boxes one iteration later (after the context step), so both are requested two iterations
  // ahead of their use (one shared 3-stage ring, as in v6, gives one iteration: 12 % of the warp time was spent waiting
  // for the tiles of the next batch).
  const uint32_t a_slot = n_slabs * ATT_SLAB_BYTES;
  constexpr uint32_t p_slot = ATT_BATCH * A * 2;
  const uint32_t s_base = (smem_u32(att_smem_raw) + 1023) & ~1023u;
  const uint32_t s_p = s_base + V7_SLOTS * a_slot;
  const uint32_t s_F = s_p + V7_SLOTS * p_slot;
  const uint32_t s_e = s_F + f_bufs * NB * A * 4;
  const uint32_t bar_full_p = s_e + V7_E_SLOTS * NB * ATT_BATCH * 4, bar_full_a = bar_full_p + V7_SLOTS * 8,
                 bar_scored = bar_full_a + V7_SLOTS * 8;
  const uint32_t cnt_scored = bar_scored + V7_SLOTS * 8, cnt_consumed = cnt_scored + V7_SLOTS * 4;  // warps done with a slot

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < V7_SLOTS; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full_p + s * 8), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full_a + s * 8), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_scored + s * 8), "r"(ATT_WARPS));
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(cnt_scored + s * 4), "r"(0) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(cnt_consumed + s * 4), "r"(0) : "memory");
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_att);
  }
  // alpha_net weights of this lane's units (a parameter: read before the dependency wait)
  float w[CA * 8];
#pragma unroll
  for (int c = 0; c < CA; ++c)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w_alpha + c * 256 + h * 128 + lane * 4));
      w[c * 8 + h * 4 + 0] = -2.0f * w4.x;
      w[c * 8 + h * 4 + 1] = -2.0f * w4.y;
      w[c * 8 + h * 4 + 2] = -2.0f * w4.z;
      w[c * 8 + h * 4 + 3] = -2.0f * w4.w;
    }
  // The first three batches of this CTA's range are pulled into L2 while the stream predecessor is still draining: an L2
  // prefetch is safe before the dependency wait whatever wrote the tiles (L2 is the point of coherence: a line written
  // after it was prefetched is simply updated), and the first tile loads below then start from L2 instead of from 296
  // simultaneous DRAM misses -- the start-up latency was ~2 us of a 31 us launch.
  if (threadIdx.x < V7_SLOTS && static_cast<int>(threadIdx.x) < nloc) {
    const int b = b0 + threadIdx.x;
    const int job = b / nbpi, kb = b - job * nbpi;
    const int img = p.n_grp > 1 ? job / p.n_grp : job;
    const int l0 = kb * ATT_BATCH;
    const int nrows = min(ATT_BATCH, L - l0);
    const long long l = static_cast<long long>(img) * L + l0;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.p_att + l * A), "r"(nrows * A * 2) : "memory");
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.att + l * H), "r"(nrows * H * 2) : "memory");
  }
  __syncthreads();  // the only block-wide barrier
  pdl_wait();

  // ---- producers (one lane): request the p_att rows (+ att_h when the job starts here) / the att boxes of local batch j
  auto coords = [&](int j, int& job, int& kb, int& img, int& grp) {
    const int b = b0 + j;
    job = b / nbpi;
    kb = b - job * nbpi;
    img = job;
    grp = 0;
    if (p.n_grp > 1) {
      img = job / p.n_grp;
      grp = job - img * p.n_grp;
    }
  };
  auto produce_p = [&](int j) {
    int job, kb, img, grp;
    coords(j, job, kb, img, grp);
    const int slot = j % V7_SLOTS;
    const int l0 = kb * ATT_BATCH;
    const int nrows = min(ATT_BATCH, L - l0);
    const bool first = (kb == 0) || (j == 0);
    const uint32_t bar = bar_full_p + slot * 8;
    const long long l = static_cast<long long>(img) * L + l0;
    mbar_expect_tx_addr(bar, nrows * A * 2 + (first ? NB * A * 4 : 0));
    bulk_g2s_hint(s_p + slot * p_slot, p.p_att + l * A, nrows * A * 2, bar, p.tile_policy);
    if (first) {
      const int fbuf = (job - job0) % f_bufs;
      const int nb = min(NB, p.beams - grp * NB);
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) {
        const long long row = static_cast<long long>(img) * p.beams + grp * NB + (jb < nb ? jb : 0);
        bulk_g2s(s_F + ((fbuf * NB + jb) * A) * 4, p.att_h + row * p.ld_att_h, A * 4, bar);
      }
    }
  };
  auto produce_a = [&](int j) {
    int job, kb, img, grp;
    coords(j, job, kb, img, grp);
    const int slot = j % V7_SLOTS;
    const uint32_t bar = bar_full_a + slot * 8;
    const long long l = static_cast<long long>(img) * L + kb * ATT_BATCH;
    mbar_expect_tx_addr(bar, n_slabs * ATT_SLAB_BYTES);
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(s_base + slot * a_slot),
                 "l"(reinterpret_cast<uint64_t>(&tmap_att)), "r"(bar), "r"(0), "r"(static_cast<int>(l)), "r"(0), "l"(p.tile_policy)
                 : "memory");
  };
  // (every CTA starts at the same time: one batch first, the next two when it has landed, lets HBM deliver 296 first
  //  batches instead of 888 before anybody can start)
  if (threadIdx.x == 0) {
    produce_p(0);
    produce_a(0);
  }

  // ---- per-lane constants -----------------------------------------------------------------------------------------
  const int g = lane >> 2, t = lane & 3;
  const int n_mtiles = H >> 4;
  const int k_in = (lane & 7) + ((lane >> 4) & 1) * 8, m_in = ((lane >> 3) & 1) * 8;
  const uint32_t a_off = s_base + (warp >> 2) * ATT_SLAB_BYTES + k_in * 128 + (((2 * (warp & 3) + (m_in >> 3)) ^ (k_in & 7)) << 4);
  const uint32_t e_rd = s_e + ((g < NB ? g : 0) * ATT_BATCH + 2 * t) * 4;  // this lane's four scores of a batch: + e-slot * NB * 64

  float m_run = -INFINITY, s_run = 0.0f;  // of beam g (lanes with g >= NB idle along)
  float acc[MT][4];
#pragma unroll
  for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0f;

  // ---- scoring: NR = 1 or 2 regions (r and r + 8) of the batch in `stage`, att_h buffer `fb` --------------------------
  // tanh(p + a) = 1 - 2 / (E F + 1), E = exp(2 p) (bf16 tile), F = exp(2 att_h); one reciprocal per PAIR of units:
  // w1/d1 + w2/d2 = (w1 d2 + w2 d1) / (d1 d2)  (see attention.cu).
  auto score_regions = [&](uint32_t st_e, uint32_t prow, uint32_t Fimg, auto nr_tag) {
    constexpr int NR = decltype(nr_tag)::value;
    // Packed fp32 arithmetic (fma.rn.f32x2 / mul.rn.f32x2, sm_100): a lane's four units (0, 1, 2, 3) of a group are paired
    // as (0, 2) and (1, 3), so that every operand is an aligned register pair as it comes out of the shared loads --
    //   (d0, d1) = (E0, E1) (F0, F1) + 1      (d2, d3) = (E2, E3) (F2, F3) + 1
    //   (d0 d2, d1 d3)                        -> two reciprocals
    //   (w0 d2 + w2 d0, w1 d3 + w3 d1)        -> numerators of the pairs
    // eight instructions for four tanh terms instead of fourteen.
    float2 pa[NR][NB];
#pragma unroll
    for (int x = 0; x < NR; ++x)
#pragma unroll
      for (int j = 0; j < NB; ++j) pa[x][j] = make_float2(0.0f, 0.0f);
    const float2 ones = make_float2(1.0f, 1.0f);
    // unit group (c, h) outermost: four units of every region are live at a time (the v6 loop keeps all of a lane's
    // 8 CA units of both regions in registers across the beams: 32 registers more, and ptxas re-derived addresses
    // and lane constants inside the batch loop to make room)
#pragma unroll
    for (int c = 0; c < CA; ++c)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float2 Ea[NR], Eb[NR];
#pragma unroll
        for (int x = 0; x < NR; ++x) {
          uint32_t u[2];
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(u[0]), "=r"(u[1]) : "r"(prow + x * ATT_WARPS * A * 2 + (c * 256 + h * 128) * 2));
          // bf16 -> fp32 is a shift / a mask
          Ea[x] = make_float2(__uint_as_float(u[0] << 16), __uint_as_float(u[0] & 0xffff0000u));
          Eb[x] = make_float2(__uint_as_float(u[1] << 16), __uint_as_float(u[1] & 0xffff0000u));
        }
        const int k = c * 8 + h * 4;
        const float2 Wa = make_float2(w[k], w[k + 1]), Wb = make_float2(w[k + 2], w[k + 3]);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          float4 f;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w)
                       : "r"(Fimg + (j * A + c * 256 + h * 128) * 4));
          const float2 Fa = make_float2(f.x, f.y), Fb = make_float2(f.z, f.w);
#pragma unroll
          for (int x = 0; x < NR; ++x) {
            const float2 da = __ffma2_rn(Ea[x], Fa, ones), db = __ffma2_rn(Eb[x], Fb, ones);
            const float2 pr = __fmul2_rn(da, db);
            const float2 r = make_float2(rcp_approx(pr.x), rcp_approx(pr.y));
            const float2 nm = __ffma2_rn(Wa, db, __fmul2_rn(Wb, da));
            pa[x][j] = __ffma2_rn(r, nm, pa[x][j]);
          }
        }
      }
    float e[NR][NB];
#pragma unroll
    for (int x = 0; x < NR; ++x)
#pragma unroll
      for (int j = 0; j < NB; ++j) e[x][j] = pa[x][j].x + pa[x][j].y;
    float v[NB];
    if constexpr (NR == 2) {
      const bool hi = lane >= 16;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float keep = hi ? e[1][j] : e[0][j], give = hi ? e[0][j] : e[1][j];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NB; ++j) v[j] = e[0][j] + __shfl_xor_sync(0xffffffffu, e[0][j], 16);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
      for (int j = 0; j < NB; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    }
    const int jl = lane & 15;
    if (jl < NB) {  // lanes 0..NB-1: region r; lanes 16..16+NB-1: region r + 8 (NR == 1: -inf, the row is past the image)
      float ej = v[0];
#pragma unroll
      for (int j = 1; j < NB; ++j) ej = (jl == j) ? v[j] : ej;
      if (NR == 1 && lane >= 16) ej = -INFINITY;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_e + (jl * ATT_BATCH + warp + (lane >> 4) * ATT_WARPS) * 4), "f"(ej) : "memory");
    }
  };

  // ---- records of cut jobs ---------------------------------------------------------------------------------------------
  constexpr int RECV = 1 + MT;  // float4 per lane
  auto rec_of = [&](int cta) { return p.v7_rec + (static_cast<size_t>(cta) * ATT_WARPS + warp) * (RECV * 32) + lane; };

  // ---- main loop: score batch i, then finish batch i - 1 --------------------------------------------------------------
  int kb = kb0, job = job0, fb = 0, stage = 0;
  uint32_t par = 0;
  int c_kb = 0, c_job = 0, c_stage = 0;  // the batch awaiting its context step
  uint32_t c_par = 0;
  for (int i = 0; i <= nloc; ++i) {
    if (i < nloc) {
      const int nrows = min(ATT_BATCH, L - kb * ATT_BATCH);
      mbar_wait_lean(bar_full_p + stage * 8, par);
      if (i == 0 && threadIdx.x == 0) {
        for (int k = 1; k < V7_SLOTS && k < nloc; ++k) produce_p(k), produce_a(k);
      }
      const uint32_t st_e = s_e + (i & (V7_E_SLOTS - 1)) * (NB * ATT_BATCH * 4);
      const uint32_t prow = s_p + stage * p_slot + (warp * A + lane * 4) * 2;
      const uint32_t Fimg = s_F + (fb * NB * A + lane * 4) * 4;
      if (warp + ATT_WARPS < nrows)
        score_regions(st_e, prow, Fimg, std::integral_constant<int, 2>{});
      else if (warp < nrows)
        score_regions(st_e, prow, Fimg, std::integral_constant<int, 1>{});
      else if (lane < NB) {  // both rows of this warp are past the image
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_e + (lane * ATT_BATCH + warp) * 4), "f"(-INFINITY) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(st_e + (lane * ATT_BATCH + warp + ATT_WARPS) * 4), "f"(-INFINITY) : "memory");
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_addr(bar_scored + stage * 8);
        if (i + V7_SLOTS < nloc) {  // the last warp to finish with the rows requests those of batch i + 3 into their slot
          uint32_t prev;
          asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(cnt_scored + stage * 4) : "memory");
          if (prev == ATT_WARPS - 1) {
            asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"(cnt_scored + stage * 4), "r"(0) : "memory");
            produce_p(i + V7_SLOTS);
          }
        }
      }
    }

    if (i > 0) {
      // ---- softmax update + context MMA of batch i - 1 (scored by everybody a whole batch ago) ----------------------
      const int j = i - 1;
      float mk[4];
      if constexpr (AUX) {
        mk[0] = mk[1] = mk[2] = mk[3] = 1.0f;
        if (p.masks != nullptr) {
          const int img = c_job / p.n_grp;
          const int l0 = c_kb * ATT_BATCH;
          const float* m_img = p.masks + static_cast<long long>(img) * L + l0;
          const int kk[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (l0 + kk[q] < L) mk[q] = __ldg(m_img + kk[q]);
        }
      }
      mbar_wait_lean(bar_scored + c_stage * 8, c_par);
      float ev[4];
      {
        const uint32_t ea = e_rd + (j & (V7_E_SLOTS - 1)) * (NB * ATT_BATCH * 4);
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(ev[0]), "=f"(ev[1]) : "r"(ea));
        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(ev[2]), "=f"(ev[3]) : "r"(ea + 32));
      }
      float mb = fmaxf(fmaxf(ev[0], ev[1]), fmaxf(ev[2], ev[3]));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
      const float m_new = fmaxf(m_run, mb);  // finite: every batch has at least one region
      const float off = -m_new * LOG2E;
      const float scale = att_ex2(fmaf(m_run, LOG2E, off));  // 0 for the first batch of a job here (m_run = -inf)
      float pl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        pl[q] = att_ex2(fmaf(ev[q], LOG2E, off));  // rows past the image: exp2(-inf) = 0
        if constexpr (AUX) pl[q] *= mk[q];
      }
      float ps = (pl[0] + pl[1]) + (pl[2] + pl[3]);
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      s_run = fmaf(s_run, scale, ps);
      m_run = m_new;
      const uint32_t bq0 = f2_to_bf16x2(pl[0], pl[1]), bq1 = f2_to_bf16x2(pl[2], pl[3]);
      if constexpr (AUX) {
        if (p.alpha != nullptr && warp == 0) {  // raw scores for the backward pass, normalised when the job is finished
          const int img = c_job / p.n_grp, beam0 = (c_job - img * p.n_grp) * NB;
          const int nb = min(NB, p.beams - beam0), l0 = c_kb * ATT_BATCH;
          for (int idx = lane; idx < NB * ATT_BATCH; idx += 32) {
            const int jb = idx / ATT_BATCH, r = idx - jb * ATT_BATCH;
            if (jb < nb && l0 + r < L) {
              float sv;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sv) : "r"(s_e + (((j & (V7_E_SLOTS - 1)) * NB + jb) * ATT_BATCH + r) * 4));
              p.alpha[(static_cast<long long>(img) * p.beams + beam0 + jb) * L + l0 + r] = sv;
            }
          }
        }
      }
      // context: D[m = column][n = beam] += A[m][k = region] * B[k][n]; accumulator columns n = 2t, 2t+1 belong to the
      // beams whose running max lives in lanes 8t and 8t+4
      const float sc0 = __shfl_sync(0xffffffffu, scale, 8 * t), sc1 = __shfl_sync(0xffffffffu, scale, 8 * t + 4);
#pragma unroll
      for (int q = 0; q < MT; ++q) {
        acc[q][0] *= sc0;
        acc[q][1] *= sc1;
        acc[q][2] *= sc0;
        acc[q][3] *= sc1;
      }
      mbar_wait_lean(bar_full_a + c_stage * 8, c_par);
      const uint32_t arow = a_off + c_stage * a_slot;
#pragma unroll
      for (int q = 0; q < MT; ++q) {
        if (warp + ATT_WARPS * q < n_mtiles) {  // warp-uniform
          uint32_t a[4];
          ldmatrix_x4_trans(arow + q * 2 * ATT_SLAB_BYTES, a);
          mma_bf16_16816(acc[q], a, bq0, bq1);
        }
      }
      // hand the boxes back: the LAST warp to get here requests those of batch j + 3 into the slot.  (v6 lets the warps take turns at
      // waiting for the other seven: the waiting warp then is the slowest of the next batch and the other seven wait
      // for it at the scored barrier -- two CTA-wide rendezvous per batch, 14 % of the instruction stream was polling)
      if (j + V7_SLOTS < nloc) {
        __syncwarp();
        if (lane == 0) {
          uint32_t prev;
          asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(prev) : "r"(cnt_consumed + c_stage * 4) : "memory");
          if (prev == ATT_WARPS - 1) {
            asm volatile("st.relaxed.cta.shared.u32 [%0], %1;" ::"r"(cnt_consumed + c_stage * 4), "r"(0) : "memory");
            produce_a(j + V7_SLOTS);
          }
        }
      }

      // ---- job finished, or the range ends inside it ----------------------------------------------------------------
      const bool job_done = (c_kb == nbpi - 1), range_end = (i == nloc);
      if (job_done || range_end) {
        float Mn[2] = {__shfl_sync(0xffffffffu, m_run, 8 * t), __shfl_sync(0xffffffffu, m_run, 8 * t + 4)};
        float Sn[2] = {__shfl_sync(0xffffffffu, s_run, 8 * t), __shfl_sync(0xffffffffu, s_run, 8 * t + 4)};
        if (c_job == job0 && kb0 != 0) {
          // a follower's share of a job it does not own: publish the record of this warp's column set
          float4* rec = rec_of(blockIdx.x);
          rec[0] = make_float4(Mn[0], Mn[1], Sn[0], Sn[1]);
#pragma unroll
          for (int q = 0; q < MT; ++q) rec[(1 + q) * 32] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
          __syncwarp();  // (orders the lanes' record stores before lane 0's release: one fence per warp, not three)
          if (lane == 0) st_release_gpu(p.v7_flag + blockIdx.x * ATT_WARPS + warp, 1);
        } else {
          if (!job_done) {
            // owner whose range ends inside the job: merge the followers' records (CTAs c + 1 .. that start before the job ends)
            const long long job_end = static_cast<long long>(c_job + 1) * nbpi;
            for (int cc = blockIdx.x + 1; cc < static_cast<int>(gridDim.x); ++cc) {
              if (static_cast<long long>(cc) * p.n_batches / gridDim.x >= job_end) break;
              int* flag = p.v7_flag + cc * ATT_WARPS + warp;
              if (lane == 0) {
                int n = 0;
                while (ld_acquire_gpu(flag) == 0) {
                  __nanosleep(64);
                  if (++n > (1 << 24)) {
                    printf("uic: att_step_fwd_v7 merge wait timed out (block %d warp %d waits for block %d)\n", blockIdx.x, warp, cc);
                    __trap();
                  }
                }
              }
              __syncwarp();  // lane 0 acquired the flag; the record is read through L2 (ld.global.cg)
              const float4* rec = rec_of(cc);
              const float4 st = __ldcg(rec);
              float4 v[MT];
#pragma unroll
              for (int q = 0; q < MT; ++q) v[q] = __ldcg(rec + (1 + q) * 32);
              __syncwarp();
              if (lane == 0) *flag = 0;  // ready for the next launch
              const float mx0 = fmaxf(Mn[0], st.x), mx1 = fmaxf(Mn[1], st.y);
              const float f0 = att_ex2((Mn[0] - mx0) * LOG2E), f1 = att_ex2((Mn[1] - mx1) * LOG2E);
              const float g0 = att_ex2((st.x - mx0) * LOG2E), g1 = att_ex2((st.y - mx1) * LOG2E);
              Sn[0] = fmaf(Sn[0], f0, st.z * g0);
              Sn[1] = fmaf(Sn[1], f1, st.w * g1);
              Mn[0] = mx0;
              Mn[1] = mx1;
#pragma unroll
              for (int q = 0; q < MT; ++q) {
                acc[q][0] = fmaf(acc[q][0], f0, v[q].x * g0);
                acc[q][1] = fmaf(acc[q][1], f1, v[q].y * g1);
                acc[q][2] = fmaf(acc[q][2], f0, v[q].z * g0);
                acc[q][3] = fmaf(acc[q][3], f1, v[q].w * g1);
              }
            }
          }
          // Element (q, e4) of the accumulators is column colb + 128 q + 8 (e4 >> 1) of beam n0 + (e4 & 1).
          int img = c_job, beam0 = 0;
          if (p.n_grp > 1) {
            img = c_job / p.n_grp;
            beam0 = (c_job - img * p.n_grp) * NB;
          }
          const int nb = min(NB, p.beams - beam0);
          const int n0 = 2 * t;
          const int colb = warp * 16 + g;
          const float inv[2] = {1.0f / Sn[0], 1.0f / Sn[1]};
          const long long row0 = static_cast<long long>(img) * p.beams + beam0;
          const long long r_n[2] = {row0 + (n0 < nb ? n0 : 0), row0 + (n0 + 1 < nb ? n0 + 1 : 0)};
          if (p.ctx_bf16 != nullptr) {
            __nv_bfloat16* o_n[2] = {p.ctx_bf16 + r_n[0] * p.ld_ctx_bf16 + colb, p.ctx_bf16 + r_n[1] * p.ld_ctx_bf16 + colb};
#pragma unroll
            for (int q = 0; q < MT; ++q)
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const int offc = 128 * q + 8 * (e4 >> 1);
                if (n0 + (e4 & 1) < nb && colb + offc < H) o_n[e4 & 1][offc] = __float2bfloat16_rn(acc[q][e4] * inv[e4 & 1]);
              }
          }
          if (p.ctx_f32 != nullptr) {
            float* o_n[2] = {p.ctx_f32 + r_n[0] * p.ld_ctx_f32 + colb, p.ctx_f32 + r_n[1] * p.ld_ctx_f32 + colb};
#pragma unroll
            for (int q = 0; q < MT; ++q)
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const int offc = 128 * q + 8 * (e4 >> 1);
                if (n0 + (e4 & 1) < nb && colb + offc < H) o_n[e4 & 1][offc] = acc[q][e4] * inv[e4 & 1];
              }
          }
          if constexpr (AUX) {
            if (p.alpha != nullptr && warp == 0) {  // raw scores (possibly written by followers) -> weights
              __syncwarp();
              const float* mfull = p.masks ? p.masks + static_cast<long long>(img) * L : nullptr;
              for (int jb = 0; jb < nb; ++jb) {
                // beam jb's statistics live in accumulator column n = jb, i.e. in the lanes with t == jb / 2
                const float Mj = __shfl_sync(0xffffffffu, (jb & 1) ? Mn[1] : Mn[0], (jb >> 1));
                const float Sj = __shfl_sync(0xffffffffu, (jb & 1) ? Sn[1] : Sn[0], (jb >> 1));
                float* arow_g = p.alpha + (row0 + jb) * L;
                for (int l = lane; l < L; l += 32) {
                  const float mkl = mfull ? mfull[l] : 1.0f;
                  arow_g[l] = __expf(__ldcg(arow_g + l) - Mj) * mkl / Sj;
                }
              }
            }
          }
        }
        m_run = -INFINITY;
        s_run = 0.0f;
#pragma unroll
        for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0f;
      }
    }

    if (i < nloc) {
      c_kb = kb;
      c_job = job;
      c_stage = stage;
      c_par = par;
      if (++kb == nbpi) {
        kb = 0;
        ++job;
        if (++fb == f_bufs) fb = 0;
      }
      if (++stage == V7_SLOTS) {
        stage = 0;
        par ^= 1;
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------
template <int NB, int CA, int MT, bool AUX>
static int launch_v7(AttParams& p, const AttPlan& pl, int n_img, int ctas, cudaStream_t stream, bool query_only, int* per_sm) {
  const size_t smem = static_cast<size_t>(V7_SLOTS) * ((p.H >> 6) * ATT_SLAB_BYTES + ATT_BATCH * p.A * 2) + pl.f_bufs * NB * p.A * 4 +
                      V7_E_SLOTS * NB * ATT_BATCH * 4 + 3 * V7_SLOTS * 8 + 2 * V7_SLOTS * 4 + 1024;
  auto kern = att_step_fwd_v7_kernel<NB, CA, MT, AUX>;
  if (smem > 226 * 1024) return 1;
  // Function attributes are per device: the opt-in shared-memory size and the occupancy it yields are cached per
  // (instantiation, device), and refreshed when a call needs more shared memory than the cached setting.
  constexpr int MAX_DEV = 64;
  static int blocks_per_sm[MAX_DEV] = {};
  static size_t smem_set[MAX_DEV] = {};
  int dev = 0;
  UIC_CUDA_OK(cudaGetDevice(&dev));
  const int slot = dev >= 0 && dev < MAX_DEV ? dev : 0;
  if (blocks_per_sm[slot] == 0 || smem > smem_set[slot] || dev >= MAX_DEV) {
    int per = 0;
    UIC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    UIC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, ATT_THREADS, smem));
    blocks_per_sm[slot] = per;
    smem_set[slot] = smem;
  }
  if (query_only) {
    *per_sm = blocks_per_sm[slot];
    return 0;
  }
  CUtensorMap tm;
  p.slab_map = 1;
  const int rc = get_tensor_map_bf16_slabs(&tm, p.att, static_cast<long long>(n_img) * p.L, p.H, ATT_BATCH);
  if (rc) return 1;  // the driver refuses the slab view: v6 handles it with 2-D boxes
  launch_begin("att_step_fwd", stream);
  UIC_CUDA_OK(launch_pdl(kern, dim3(ctas), dim3(ATT_THREADS), smem, stream, tm, p));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

template <int CA, int MT>
static int dispatch_v7(AttParams& p, const AttPlan& pl, int n_img, int ctas, cudaStream_t stream, bool q, int* per_sm) {
  const bool aux = p.masks != nullptr || p.alpha != nullptr;
  switch (pl.nb) {
    case 1: return aux ? launch_v7<1, CA, MT, true>(p, pl, n_img, ctas, stream, q, per_sm) : launch_v7<1, CA, MT, false>(p, pl, n_img, ctas, stream, q, per_sm);
    case 2: return aux ? launch_v7<2, CA, MT, true>(p, pl, n_img, ctas, stream, q, per_sm) : launch_v7<2, CA, MT, false>(p, pl, n_img, ctas, stream, q, per_sm);
    case 3: return aux ? launch_v7<3, CA, MT, true>(p, pl, n_img, ctas, stream, q, per_sm) : launch_v7<3, CA, MT, false>(p, pl, n_img, ctas, stream, q, per_sm);
    default: break;
  }
  if constexpr (CA == 2 && MT == 8) {  // the one-CTA-per-SM shape with room for 4 or 5 beams per pass (att_v7_beam_cap)
    if (pl.nb == 4) return aux ? launch_v7<4, CA, MT, true>(p, pl, n_img, ctas, stream, q, per_sm) : launch_v7<4, CA, MT, false>(p, pl, n_img, ctas, stream, q, per_sm);
    if (pl.nb == 5) return aux ? launch_v7<5, CA, MT, true>(p, pl, n_img, ctas, stream, q, per_sm) : launch_v7<5, CA, MT, false>(p, pl, n_img, ctas, stream, q, per_sm);
  }
  return 1;
}

static int dispatch_shape_v7(AttParams& p, const AttPlan& pl, int n_img, int ctas, cudaStream_t stream, bool q, int* per_sm) {
  const int ca = p.A / 256;
  if (p.H <= 512) {
    if (ca == 1) return dispatch_v7<1, 4>(p, pl, n_img, ctas, stream, q, per_sm);
    if (ca == 2) return dispatch_v7<2, 4>(p, pl, n_img, ctas, stream, q, per_sm);
    return dispatch_v7<4, 4>(p, pl, n_img, ctas, stream, q, per_sm);
  }
  if (ca <= 2) return ca == 1 ? 1 : dispatch_v7<2, 8>(p, pl, n_img, ctas, stream, q, per_sm);
  return dispatch_v7<4, 8>(p, pl, n_img, ctas, stream, q, per_sm);
}

static bool v7_shape_ok(int A, int H) { return A % 256 == 0 && (A == 256 || A == 512 || A == 1024) && H % 64 == 0 && H >= 64 && H <= 1024; }

static bool v7_enabled() {
  static int mode = -1;  // UIC_ATT_V7=0 keeps v6 everywhere
  if (mode < 0) {
    const char* e = getenv("UIC_ATT_V7");
    mode = e ? atoi(e) : 1;
  }
  return mode != 0;
}

// With more than three beams per image the 3-beam kernels read the image's tiles once per GROUP of beams.  The shape
// A = 512, 512 < H <= 1024 (configs[4]: rnn 1024, beam 5) runs one CTA per SM whatever the beam count, and its registers
// and shared memory hold five beams: one pass instead of two there (UIC_ATT_NB5=0 turns it off).
int att_v7_beam_cap(int n_img, int beams, int L, int A, int H) {
  (void)L;
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("UIC_ATT_NB5");
    on = e ? atoi(e) : 1;
  }
  if (!on || !v7_enabled() || beams <= 3 || A != 512 || H <= 512 || H > 1024 || H % 64) return 3;
  if (static_cast<long long>(n_img) * 2 < 148) return 3;  // below that v6's segmented plan runs (att_v7_ctas)
  return 5;
}

// Grid of the v7 kernel for this problem, 0 when v6 should run it: v7 covers the shapes it is instantiated for and
// the regime where jobs are plentiful (>= half of the CTA slots); below that v6 cuts every job into equal segments,
// whose partials are merged in parallel rather than by one owner.
int att_v7_ctas(int n_img, int beams, int L, int A, int H, const AttPlan& pl) {
  if (!v7_enabled() || !v7_shape_ok(A, H)) return 0;
  const long long jobs = static_cast<long long>(n_img) * pl.groups;
  const int per_sm_nominal = (A / 256 <= 2 && H <= 512) ? 2 : 1;
  if (jobs * 2 < 148LL * per_sm_nominal) return 0;
  AttParams q{};
  q.A = A;
  q.H = H;
  int per_sm = 0;
  if (dispatch_shape_v7(q, pl, n_img, 0, nullptr, true, &per_sm) != 0 || per_sm <= 0) return 0;
  if (per_sm > per_sm_nominal) per_sm = per_sm_nominal;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_batches = jobs * pl.nbpi;
  const long long slots = static_cast<long long>(sms) * per_sm;
  return static_cast<int>(n_batches < slots ? n_batches : slots);
}

long long att_v7_workspace_bytes(int ctas, int mt) {
  const long long flags = ((static_cast<long long>(ctas) * ATT_WARPS * 4 + 255) / 256) * 256;
  return flags + static_cast<long long>(ctas) * ATT_WARPS * (1 + mt) * 32 * 16;
}

int att_step_fwd_v7(AttParams& p, const AttPlan& pl, int n_img, int ctas, cudaStream_t stream) {
  return dispatch_shape_v7(p, pl, n_img, ctas, stream, false, nullptr);
}

}  // namespace uic
