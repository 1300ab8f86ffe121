// Fused additive-attention step: score -> softmax -> context with ONE read of the image's feature
// tiles (reference: Attention.forward, models/AttModel.py:538-558, eight separate ATen kernels and a
// materialised (rows, L, A) tanh tensor).
//
// HBM-bound by design: per image and step the kernel reads p_att[i] (L x A bf16) and att[i]
// (L x H bf16) exactly once, no matter how many beams share the image.
//
// Structure (sixth iteration, see profiles/ for the ncu captures and timelines that drove it):
//   v1 kept a region per warp in registers with one prefetch in flight: latency bound, 15 % of HBM.
//   v2 staged 8-region tiles per CTA with two block-wide phases per stage: 3x the instructions.
//   v3 made every warp autonomous (private rings, online softmax, fp32 context accumulators in
//      registers): 645 instructions per region, a third of them bookkeeping; 26 % of HBM.
//   v4 gave a CTA a 49-region slice, all loads up front, exact softmax, context on mma.sync: 35 % fewer
//      instructions but the same time -- a CTA lived ~10 us of which 3 were arithmetic, the rest exposed
//      load latency, three block barriers and the fence/atomic of the slice merge.
//   v5 made the CTAs persistent and warp-specialised (8 score warps, 4 context warps): the context warps'
//      short dependent chains were starved by the always-ready score warps (1.5 us per batch instead of
//      0.25) and throttled the ring.
//   v6 (this file): persistent, HOMOGENEOUS warps, no block barrier after start-up.  The work is the flat
//      list of 16-region batches of all (image, beam group) jobs; CTA c owns a contiguous range of it and
//      streams it through a 3-stage shared-memory ring (one bulk copy for the batch's p_att rows, one
//      128B-swizzled TMA box per 64 columns of its att rows, att_h when the image changes).  Per batch i
//      every one of the 8 warps
//        - scores its two regions: e[j] = sum_a w_a tanh(p_att[l,a] + att_h[j,a]) for the job's beams, in
//          the exponential form described below, and publishes them in shared memory (mbarrier);
//        - in between, finishes batch i-1: online softmax at batch granularity (redundantly per warp, it
//          is tiny) and the context product of ITS 16-column tiles on the tensor cores,
//          D[col, beam] += att^T[col, region] * alpha[region, beam] (mma.sync m16n8k16, bf16 weights, fp32
//          accumulate, operands straight from the swizzled boxes through ldmatrix.trans);
//        - the last warp to finish batch i-1 refills its stage with batch i+2.
//      With at least half as many jobs as CTA slots a CTA owns whole jobs and nothing is merged; smaller
//      batches cut every job into equal segments, one CTA each, and the last warp to arrive at the job's
//      workspace counter finishes it (threadfence reduction), column set by column set.
#include "uic_attention.cuh"

namespace uic {

// First batch of work item `it` (item = segment `it % segs` of job `it / segs`; a job's nbpi batches are cut into
// `segs` nearly equal runs).
__device__ __forceinline__ int att_item_begin(int it, int segs, int nbpi) {
  const int job = it / segs, sg = it - job * segs;
  return job * nbpi + (sg * nbpi) / segs;
}

// Position of a batch in the job list (every warp keeps identical copies and advances them in step).
struct AttCursor {
  int kb, img, grp, job, fbuf, stage;
  uint32_t parity;
};
__device__ __forceinline__ void att_advance(AttCursor& c, int nbpi, int n_grp, int f_bufs) {
  if (++c.kb == nbpi) {
    c.kb = 0;
    ++c.job;
    if (++c.fbuf == f_bufs) c.fbuf = 0;
    if (++c.grp == n_grp) {
      c.grp = 0;
      ++c.img;
    }
  }
  if (++c.stage == ATT_STAGES) {
    c.stage = 0;
    c.parity ^= 1;
  }
}

// CA = ceil(A / 256): lane owns units [256c + 128h + 4*lane, +4), h = 0, 1.  MT = 16-column context tiles per warp.
// EXA: A == 256 * CA exactly (no tail guards / register clears in the scoring loop).
template <int NB, int CA, int MT, bool EXA>
__global__ void __launch_bounds__(ATT_THREADS, (CA <= 2 && MT <= 4) ? 2 : 1)
att_step_fwd_kernel(const __grid_constant__ CUtensorMap tmap_att, AttParams p) {
  extern __shared__ uint8_t att_smem_raw[];

  pdl_launch_dependents();
  const int L = p.L, A = p.A, H = p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // (one predicate for the debug timeline: evaluated per use it cost an S2R of blockIdx and three compares at six places of
  //  the per-batch loop)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  // CTA c owns the work items [c * items / ctas, (c + 1) * items / ctas): a contiguous run of batches
  const int item0 = static_cast<int>(static_cast<long long>(blockIdx.x) * p.items / gridDim.x);
  const int item1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.items / gridDim.x);
  const int b_start = att_item_begin(item0, p.segs, p.nbpi);
  const int nloc = att_item_begin(item1, p.segs, p.nbpi) - b_start;
  if (nloc <= 0) return;

  const AttSmem sm = att_smem_layout(A, H, NB, p.f_bufs);
  uint8_t* att_smem = att_smem_raw + ((1024 - (smem_u32(att_smem_raw) & 1023)) & 1023);
  const uint32_t s_base = smem_u32(att_smem);
  const uint32_t bar_full = s_base + sm.off_bar, bar_scored = bar_full + ATT_STAGES * 8, bar_consumed = bar_scored + ATT_STAGES * 8;
  const float* s_F = reinterpret_cast<const float*>(att_smem + sm.off_F);
  float* s_e = reinterpret_cast<float*>(att_smem + sm.off_e);

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < ATT_STAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full + s * 8), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_scored + s * 8), "r"(ATT_WARPS));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_consumed + s * 8), "r"(ATT_WARPS));
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmap_att);
  }
  // alpha_net weights of this lane's units: a parameter, not written by the stream predecessor, so its loads (an L2 round
  // trip every CTA used to pay after the wait below) overlap the previous kernel's tail
  float w[CA * 8];
#pragma unroll
  for (int c = 0; c < CA; ++c)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int u0 = c * 256 + h * 128 + lane * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) w[c * 8 + h * 4 + k] = (u0 + k < A) ? -2.0f * __ldg(p.w_alpha + u0 + k) : 0.0f;
    }
  __syncthreads();  // the only block-wide barrier
  pdl_wait();       // (programmatic dependent launch: the set-up above overlapped the previous kernel's tail)
  ATT_TRACE(0);
  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 448) p.trace[128 + 2 * blockIdx.x] = att_now();  // (1024-slot debug buffer)

  // ---- the producer: whichever lane 0 calls it ----------------------------------------------------------
  // Requests the batch under the producer cursor: the p_att rows are contiguous in both memories, the att rows
  // come as 64-column TMA boxes (rows past the image belong to the next one or are zero-filled: their weights
  // are zero), and the job's att_h rows (as F) come along when the batch is the first of its image here.
  const int n_slabs = (H + 63) >> 6;
  AttCursor pr;
  {
    const int job0 = b_start / p.nbpi;
    pr.job = job0;
    pr.kb = b_start - job0 * p.nbpi;
    pr.img = job0 / p.n_grp;
    pr.grp = job0 - pr.img * p.n_grp;
    pr.fbuf = 0;
    pr.stage = 0;
    pr.parity = 0;
  }
  AttCursor sc = pr;  // the batch being scored
  int pr_index = 0;   // local index of the batch under the producer cursor
  auto produce = [&]() {
    const int l0 = pr.kb * ATT_BATCH;
    const int nrows = min(ATT_BATCH, L - l0);
    const bool first = (pr.kb == 0) || (pr_index == 0);
    const uint32_t bar = bar_full + pr.stage * 8;
    const uint32_t st = s_base + pr.stage * sm.stage_bytes;
    const long long l = static_cast<long long>(pr.img) * L + l0;
    mbar_expect_tx_addr(bar, n_slabs * ATT_SLAB_BYTES + nrows * A * 2 + (first ? NB * A * 4 : 0));
    // the tiles are streamed once per launch: p.tile_policy marks them evict-first when they are too big to stay in L2
    // from step to step anyway, so that they do not push the decoder weights (re-read by every step's GEMMs) out
    bulk_g2s_hint(st + sm.p_off, p.p_att + l * A, nrows * A * 2, bar, p.tile_policy);
    if (p.slab_map) {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(st),
                   "l"(reinterpret_cast<uint64_t>(&tmap_att)), "r"(bar), "r"(0), "r"(static_cast<int>(l)), "r"(0), "l"(p.tile_policy)
                   : "memory");
    } else {
      for (int sl = 0; sl < n_slabs; ++sl)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
                         st + sl * ATT_SLAB_BYTES),
                     "l"(reinterpret_cast<uint64_t>(&tmap_att)), "r"(bar), "r"(sl * 64), "r"(static_cast<int>(l)), "l"(p.tile_policy)
                     : "memory");
    }
    if (first) {
      const int nb = min(NB, p.beams - pr.grp * NB);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const long long row = static_cast<long long>(pr.img) * p.beams + pr.grp * NB + (j < nb ? j : 0);
        bulk_g2s(s_base + sm.off_F + ((pr.fbuf * NB + j) * A) * 4, p.att_h + row * p.ld_att_h, A * 4, bar);
      }
    }
  };
  // The ring starts full.  Afterwards batch b + 3 is requested once all warps have finished batch b; the att_h
  // buffer it may overwrite then belongs to an image whose regions have all been scored (f_bufs = 3 when every
  // image is a single batch, else two buffers suffice because a whole image spans at least two batches).
  // (every CTA starts at the same time: asking for one batch first, and for the next two when it has landed, lets
  // HBM deliver 296 first batches instead of 888 before anybody can start)
  if (threadIdx.x == 0) produce();
  att_advance(pr, p.nbpi, p.n_grp, p.f_bufs);
  ++pr_index;

  // ---- per-lane constants ---------------------------------------------------------------------------------
  // tanh(p + a) = 1 - 2 / (E F + 1) with E = exp(2 p) (bf16 tile) and F = exp(2 att_h), both capped at 2^60 by their
  // GEMM epilogues; the constant sum(w) drops out of the softmax, so the score is e = sum_a (-2 w_a) / (E_a F_a + 1).
  // One reciprocal serves a PAIR of units: w1/d1 + w2/d2 = (w1 d2 + w2 d1) / (d1 d2).  d <= 2^120 + 1 and the numerator
  // stay finite; when d1 d2 overflows the reciprocal is 0 and so is the term -- its limit (both tanh saturated at 1).
  const int g = lane >> 2, t = lane & 3;
  const int n_mtiles = (H + 15) >> 4;
  // Tile warp + 8 q is the (warp & 3)-th 16-column group of box 2 q + warp / 4.  ldmatrix address of this lane:
  // row k_in of the box, 16-byte chunk 2 (warp & 3) + m_in / 8, XOR-swizzled with the row (SWIZZLE_128B).
  const int k_in = (lane & 7) + ((lane >> 4) & 1) * 8, m_in = ((lane >> 3) & 1) * 8;
  const uint32_t a_off = (warp >> 2) * ATT_SLAB_BYTES + k_in * 128 + (((2 * (warp & 3) + (m_in >> 3)) ^ (k_in & 7)) << 4);

  float m_run = -INFINITY, s_run = 0.0f;  // of beam g (lanes with g >= NB idle along)
  float acc[MT][4];
#pragma unroll
  for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0f;

  // ---- scoring of one region of the batch under `sc` ---------------------------------------------------------
  // Scores NR = 1 or 2 regions (r and r + 8) of the batch under `sc`: the pair shares the att_h (F) loads and one
  // transposed butterfly (lanes 0-15 end up with the sums of the first region, lanes 16-31 with the second's).
  auto score_regions = [&](int r, auto nr_tag) {
    constexpr int NR = decltype(nr_tag)::value;
    const uint32_t prow = s_base + sc.stage * sm.stage_bytes + sm.p_off + r * A * 2;
    const float* Fimg = s_F + sc.fbuf * NB * A;
    float Ef[NR][CA * 8];
#pragma unroll
    for (int x = 0; x < NR; ++x)
#pragma unroll
      for (int c = 0; c < CA; ++c)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int u0 = c * 256 + h * 128 + lane * 4;
          uint32_t u[2];
          if (!EXA) u[0] = u[1] = 0;
          if (EXA || u0 < A)
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(u[0]), "=r"(u[1]) : "r"(prow + x * ATT_WARPS * A * 2 + u0 * 2));
#pragma unroll
          for (int k = 0; k < 2; ++k) {  // bf16 -> fp32 is a shift / a mask
            Ef[x][c * 8 + h * 4 + 2 * k] = __uint_as_float(u[k] << 16);
            Ef[x][c * 8 + h * 4 + 2 * k + 1] = __uint_as_float(u[k] & 0xffff0000u);
          }
        }
    float e[NR][NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float p0[NR], p1[NR];
#pragma unroll
      for (int x = 0; x < NR; ++x) p0[x] = p1[x] = 0.0f;
#pragma unroll
      for (int c = 0; c < CA; ++c)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int u0 = c * 256 + h * 128 + lane * 4;
          float4 f;
          if (!EXA) f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
          if (EXA || u0 < A) f = *reinterpret_cast<const float4*>(Fimg + j * A + u0);
          const int k = c * 8 + h * 4;
#pragma unroll
          for (int x = 0; x < NR; ++x) {
            {
              const float d1 = fmaf(Ef[x][k], f.x, 1.0f), d2 = fmaf(Ef[x][k + 1], f.y, 1.0f);
              p0[x] = fmaf(rcp_approx(d1 * d2), fmaf(w[k + 1], d1, w[k] * d2), p0[x]);
            }
            {
              const float d1 = fmaf(Ef[x][k + 2], f.z, 1.0f), d2 = fmaf(Ef[x][k + 3], f.w, 1.0f);
              p1[x] = fmaf(rcp_approx(d1 * d2), fmaf(w[k + 3], d1, w[k + 2] * d2), p1[x]);
            }
          }
        }
#pragma unroll
      for (int x = 0; x < NR; ++x) e[x][j] = p0[x] + p1[x];
    }
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
    if (jl < NB && (NR == 2 || lane < 16)) {
      float ej = v[0];
#pragma unroll
      for (int j = 1; j < NB; ++j) ej = (jl == j) ? v[j] : ej;
      s_e[(sc.stage * NB + jl) * ATT_BATCH + r + (lane >> 4) * ATT_WARPS] = ej;
    }
  };

  // ---- softmax update + context MMA (+ image finalisation) of the batch under `cx` ---------------------------
  AttCursor cx = sc;
  auto context = [&](bool range_end) {
    const int beam0 = cx.grp * NB;
    const int nb = min(NB, p.beams - beam0);
    const int l0 = cx.kb * ATT_BATCH;
    const int nrows = min(ATT_BATCH, L - l0);
    const int kk[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
    float mk[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    if (p.masks != nullptr) {
      const float* m_img = p.masks + static_cast<long long>(cx.img) * L + l0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (kk[q] < nrows) mk[q] = __ldg(m_img + kk[q]);
    }
    mbar_wait_addr_sleepy(bar_scored + cx.stage * 8, cx.parity);  // all eight warps have scored this batch
    ATT_TRACE(50 + pr_index);

    // online softmax at batch granularity: this lane's four regions of beam g
    const float* se = s_e + (cx.stage * NB + (g < NB ? g : 0)) * ATT_BATCH;
    const float2 ea = *reinterpret_cast<const float2*>(se + 2 * t);
    const float2 eb = *reinterpret_cast<const float2*>(se + 2 * t + 8);
    float ev[4] = {ea.x, ea.y, eb.x, eb.y};
    float mb = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (kk[q] >= nrows) ev[q] = -INFINITY;
      mb = fmaxf(mb, ev[q]);
    }
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    const float m_new = fmaxf(m_run, mb);  // finite: every batch has at least one region
    float scale = 1.0f, pl[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (g < NB) {
      constexpr float LOG2E = 1.4426950408889634f;
      const float off = -m_new * LOG2E;
      scale = att_ex2(fmaf(m_run, LOG2E, off));  // 0 for the first batch of an image (m_run = -inf)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (kk[q] < nrows) pl[q] = att_ex2(fmaf(ev[q], LOG2E, off)) * mk[q];
    }
    float ps = (pl[0] + pl[1]) + (pl[2] + pl[3]);
    ps += __shfl_xor_sync(0xffffffffu, ps, 1);
    ps += __shfl_xor_sync(0xffffffffu, ps, 2);
    s_run = fmaf(s_run, scale, ps);
    m_run = m_new;
    const uint32_t b0 = f2_to_bf16x2(pl[0], pl[1]), b1 = f2_to_bf16x2(pl[2], pl[3]);
    if (p.alpha != nullptr && warp == 0) {  // raw scores for the backward pass, normalised when the image is finished
      for (int idx = lane; idx < NB * ATT_BATCH; idx += 32) {
        const int j = idx / ATT_BATCH, r = idx - j * ATT_BATCH;
        if (j < nb && r < nrows)
          p.alpha[(static_cast<long long>(cx.img) * p.beams + beam0 + j) * L + l0 + r] = s_e[(cx.stage * NB + j) * ATT_BATCH + r];
      }
    }

    // context: D[m = column][n = beam] += A[m][k = region] * B[k][n].  Accumulator columns n = 2t, 2t+1 belong
    // to the beams whose running max lives in lanes 8t and 8t+4.
    const float sc0 = __shfl_sync(0xffffffffu, scale, 8 * t), sc1 = __shfl_sync(0xffffffffu, scale, 8 * t + 4);
#pragma unroll
    for (int q = 0; q < MT; ++q) {  // (unconditional: sixteen independent multiplies are cheaper than a vote and a branch)
      acc[q][0] *= sc0;
      acc[q][1] *= sc1;
      acc[q][2] *= sc0;
      acc[q][3] *= sc1;
    }
    const uint32_t arow = s_base + cx.stage * sm.stage_bytes + a_off;
#pragma unroll
    for (int q = 0; q < MT; ++q) {
      if (warp + ATT_WARPS * q < n_mtiles) {  // warp-uniform
        uint32_t a[4];
        ldmatrix_x4_trans(arow + q * 2 * ATT_SLAB_BYTES, a);
        mma_bf16_16816(acc[q], a, b0, b1);
      }
    }
    ATT_TRACE(70 + pr_index);
    // Hand the stage back.  The warps take turns (batch index mod 8) at waiting for the other seven and requesting
    // the batch three ahead into it, so the wait costs each warp one batch in eight.
    __syncwarp();
    if (lane == 0) mbar_arrive_addr(bar_consumed + cx.stage * 8);
    if (pr_index < nloc) {
      if ((pr_index & (ATT_WARPS - 1)) == warp) {
        mbar_wait_addr_sleepy(bar_consumed + cx.stage * 8, cx.parity);
        if (lane == 0) produce();
      }
      att_advance(pr, p.nbpi, p.n_grp, p.f_bufs);
      ++pr_index;
    }

    // ---- image finished (or the range ends inside it) -----------------------------------------------------
    if (cx.kb == p.nbpi - 1 || range_end) {
      const int nseg = p.segs, seg = item0 - cx.job * p.segs;  // (segs > 1: this CTA owns exactly one item)
      // statistics of the accumulator columns' beams
      float Mn[2] = {__shfl_sync(0xffffffffu, m_run, 8 * t), __shfl_sync(0xffffffffu, m_run, 8 * t + 4)};
      float Sn[2] = {__shfl_sync(0xffffffffu, s_run, 8 * t), __shfl_sync(0xffffffffu, s_run, 8 * t + 4)};
      const int n0 = 2 * t;
      bool finish = true;
      if (nseg > 1) {
        // Every lane owns a private record of 1 + MT float4: (M0, M1, S0, S1) and its accumulators.
        constexpr int REC = 4 + 4 * MT;
        float4* rec = reinterpret_cast<float4*>(p.ws_partial) +
                      (((static_cast<long long>(cx.job) * p.segs + seg) * ATT_WARPS + warp) * 32 + lane) * (REC / 4);
        if (n0 < NB) {
          rec[0] = make_float4(Mn[0], Mn[1], Sn[0], Sn[1]);
#pragma unroll
          for (int q = 0; q < MT; ++q) rec[1 + q] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        }
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) {
          int* cnt = p.ws_counter + cx.job * ATT_WARPS + warp;
          last = (atomicAdd(cnt, 1) == nseg - 1);
          if (last) *cnt = 0;  // ready for the next launch
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        finish = last != 0;
        if (finish && n0 < NB) {
          __threadfence();
          const float4* base = reinterpret_cast<const float4*>(p.ws_partial) +
                               ((static_cast<long long>(cx.job) * p.segs * ATT_WARPS + warp) * 32 + lane) * (REC / 4);
          const long long seg_stride = static_cast<long long>(ATT_WARPS) * 32 * (REC / 4);
          float mx0 = -INFINITY, mx1 = -INFINITY;
          for (int sgm = 0; sgm < nseg; ++sgm) {
            const float4 st = __ldcg(base + sgm * seg_stride);
            mx0 = fmaxf(mx0, st.x);
            mx1 = fmaxf(mx1, st.y);
          }
          float ss0 = 0.0f, ss1 = 0.0f;
#pragma unroll
          for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0f;
          for (int sgm = 0; sgm < nseg; ++sgm) {
            const float4* r4 = base + sgm * seg_stride;
            const float4 st = __ldcg(r4);
            float4 v[MT];
#pragma unroll
            for (int q = 0; q < MT; ++q) v[q] = __ldcg(r4 + 1 + q);
            const float f0 = __expf(st.x - mx0), f1 = __expf(st.y - mx1);
            ss0 = fmaf(st.z, f0, ss0);
            ss1 = fmaf(st.w, f1, ss1);
#pragma unroll
            for (int q = 0; q < MT; ++q) {
              acc[q][0] = fmaf(v[q].x, f0, acc[q][0]);
              acc[q][1] = fmaf(v[q].y, f1, acc[q][1]);
              acc[q][2] = fmaf(v[q].z, f0, acc[q][2]);
              acc[q][3] = fmaf(v[q].w, f1, acc[q][3]);
            }
          }
          Mn[0] = mx0;
          Mn[1] = mx1;
          Sn[0] = ss0;
          Sn[1] = ss1;
        }
      }
      if (finish) {
        // Element (q, e4) of the accumulators is column colb + 128 q + 8 (e4 >> 1) of beam n0 + (e4 & 1): one base
        // pointer per beam plus compile-time offsets.
        const int colb = warp * 16 + g;
        const float inv[2] = {1.0f / Sn[0], 1.0f / Sn[1]};
        const long long row0 = static_cast<long long>(cx.img) * p.beams + beam0;
        const long long r_n[2] = {row0 + (n0 < nb ? n0 : 0), row0 + (n0 + 1 < nb ? n0 + 1 : 0)};
        if (p.ctx_bf16 != nullptr) {
          __nv_bfloat16* o_n[2] = {p.ctx_bf16 + r_n[0] * p.ld_ctx_bf16 + colb, p.ctx_bf16 + r_n[1] * p.ld_ctx_bf16 + colb};
#pragma unroll
          for (int q = 0; q < MT; ++q)
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int off = 128 * q + 8 * (e4 >> 1);
              if (n0 + (e4 & 1) < nb && colb + off < H) o_n[e4 & 1][off] = __float2bfloat16_rn(acc[q][e4] * inv[e4 & 1]);
            }
        }
        if (p.ctx_f32 != nullptr) {
          float* o_n[2] = {p.ctx_f32 + r_n[0] * p.ld_ctx_f32 + colb, p.ctx_f32 + r_n[1] * p.ld_ctx_f32 + colb};
#pragma unroll
          for (int q = 0; q < MT; ++q)
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int off = 128 * q + 8 * (e4 >> 1);
              if (n0 + (e4 & 1) < nb && colb + off < H) o_n[e4 & 1][off] = acc[q][e4] * inv[e4 & 1];
            }
        }
        if (p.alpha != nullptr && warp == 0) {  // raw scores (possibly written by other CTAs) -> weights
          __syncwarp();
          const float* mfull = p.masks ? p.masks + static_cast<long long>(cx.img) * L : nullptr;
          for (int j = 0; j < nb; ++j) {
            // beam j's statistics live in accumulator column n = j, i.e. in the lanes with t == j / 2
            const float Mj = __shfl_sync(0xffffffffu, (j & 1) ? Mn[1] : Mn[0], (j >> 1));
            const float Sj = __shfl_sync(0xffffffffu, (j & 1) ? Sn[1] : Sn[0], (j >> 1));
            float* arow_g = p.alpha + (row0 + j) * L;
            for (int l = lane; l < L; l += 32) {
              const float mkl = mfull ? mfull[l] : 1.0f;
              arow_g[l] = __expf(__ldcg(arow_g + l) - Mj) * mkl / Sj;
            }
          }
        }
      }
      m_run = -INFINITY;
      s_run = 0.0f;
#pragma unroll
      for (int q = 0; q < MT; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0f;
    }
  };

  // ---- main loop: score batch i, finishing batch i-1 between its two regions -----------------------------------
  for (int i = 0; i <= nloc; ++i) {
    if (i < nloc) {
      const int nrows = min(ATT_BATCH, L - sc.kb * ATT_BATCH);
      mbar_wait_addr_sleepy(bar_full + sc.stage * 8, sc.parity);
      ATT_TRACE(1 + 4 * i);
      if (i == 0) {
        for (int k = 1; k < ATT_STAGES && k < nloc; ++k) {
          if (threadIdx.x == 0) produce();
          att_advance(pr, p.nbpi, p.n_grp, p.f_bufs);
          ++pr_index;
        }
      }
      if (warp + ATT_WARPS < nrows)
        score_regions(warp, std::integral_constant<int, 2>{});
      else if (warp < nrows)
        score_regions(warp, std::integral_constant<int, 1>{});
      __syncwarp();
      if (lane == 0) mbar_arrive_addr(bar_scored + sc.stage * 8);
      ATT_TRACE(2 + 4 * i);
    }
    // batch i - 1: every warp scored it a whole batch ago, nobody waits (finishing it BEFORE scoring batch i was
    // measured too: same time)
    if (i > 0) context(i == nloc);
    ATT_TRACE(3 + 4 * i);
    if (i < nloc) {
      cx = sc;
      att_advance(sc, p.nbpi, p.n_grp, p.f_bufs);
    }
  }
  ATT_TRACE(100);
  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 448) p.trace[129 + 2 * blockIdx.x] = att_now();
}

// ---- host side -------------------------------------------------------------------------------------------
static unsigned long long att_tile_policy(long long tile_bytes) {
  static int mode = -1;  // UIC_ATT_L2: 0 = no hint, 1 = evict-first always, default = by size
  if (mode < 0) {
    const char* e = getenv("UIC_ATT_L2");
    mode = e ? atoi(e) : 2;
  }
  if (mode == 0) return L2_EVICT_NORMAL;
  if (mode == 1) return L2_EVICT_FIRST;
  return tile_bytes > (56LL << 20) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
}

static int beams_per_group(int beams, int cap) {
  if (beams <= cap) return beams;
  const int groups = (beams + cap - 1) / cap;
  return (beams + groups - 1) / groups;
}

static AttPlan make_plan(int n_img, int beams, int L, int A, int H) {
  AttPlan pl;
  // Beams of an image that share one pass over its tiles: up to 3 in general (registers of the 2-CTA-per-SM kernels, the
  // v6 instantiations), up to 5 where the v7 kernel runs one CTA per SM anyway (att_v7_beam_cap)
  pl.nb = beams_per_group(beams, att_v7_beam_cap(n_img, beams, L, A, H));
  pl.groups = (beams + pl.nb - 1) / pl.nb;
  pl.nbpi = (L + ATT_BATCH - 1) / ATT_BATCH;
  const long long jobs = static_cast<long long>(n_img) * pl.groups;
  const int per_sm = ((A + 255) / 256 <= 2 && H <= 512) ? 2 : 1;  // matches the kernel's launch bounds
  const long long slots = 148LL * per_sm;
  // Whole jobs per CTA (no merging) unless that would leave more than half of the slots empty: then every job
  // is cut into `segs` segments of whole batches, one CTA each.
  long long segs = slots / (jobs > 0 ? jobs : 1);
  segs = segs < 1 ? 1 : (segs > pl.nbpi ? pl.nbpi : segs);
  pl.segs = static_cast<int>(segs);
  const long long items = jobs * segs;
  pl.items = static_cast<int>(items);
  pl.ctas = static_cast<int>(items < slots ? items : slots);
  pl.mt = H <= 512 ? 4 : 8;
  pl.f_bufs = pl.nbpi == 1 ? 3 : 2;
  return pl;
}

long long att_step_workspace_bytes(int n_img, int beams, int L, int A, int H) {
  const AttPlan pl = make_plan(n_img, beams, L, A, H);
  const long long jobs = static_cast<long long>(n_img) * pl.groups;
  const long long counters = ((jobs * ATT_WARPS * 4 + 255) / 256) * 256;
  const long long v6 = counters + (pl.segs > 1 ? jobs * pl.segs * ATT_WARPS * 32 * (4 + 4 * pl.mt) * 4 : 0);
  const int v7_ctas = att_v7_ctas(n_img, beams, L, A, H, pl);
  const long long v7 = v7_ctas > 0 ? att_v7_workspace_bytes(v7_ctas, pl.mt) : 0;
  return v6 > v7 ? v6 : v7;
}

template <int NB, int CA, int MT, bool EXA>
static int launch_att(AttParams& p, const AttPlan& pl, int n_img, cudaStream_t stream) {
  const size_t smem = att_smem_layout(p.A, p.H, NB, pl.f_bufs).total;
  auto kern = att_step_fwd_kernel<NB, CA, MT, EXA>;
  if (smem > 226 * 1024) return set_error(UIC_ERR_SHAPE, "att_step_fwd: A=%d H=%d need %zu bytes of shared memory", p.A, p.H, smem);
  if (smem > 40 * 1024)  // dynamic + the kernel's static shared memory may exceed the 48 KB default
    UIC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  // att as a 2-D tensor [n_img * L regions][H columns]: boxes of 16 regions x 64 columns, 128-byte swizzle
  CUtensorMap tm;
  p.slab_map = (p.H % 64 == 0);
  int rc = p.slab_map ? get_tensor_map_bf16_slabs(&tm, p.att, static_cast<long long>(n_img) * p.L, p.H, ATT_BATCH) : -1;
  if (rc) {  // H is not a multiple of 64 (or the driver refuses the slab view): one 2-D box per 64 columns
    p.slab_map = 0;
    rc = get_tensor_map_bf16(&tm, p.att, static_cast<long long>(n_img) * p.L, p.H, p.H, ATT_BATCH, 64);
    if (rc) return rc;
  }
  launch_begin("att_step_fwd", stream);
  UIC_CUDA_OK(launch_pdl(kern, dim3(pl.ctas), dim3(ATT_THREADS), smem, stream, tm, p));
  UIC_CUDA_OK(cudaGetLastError());
  launch_end(stream);
  return 0;
}

template <int CA, int MT>
static int dispatch_nb(AttParams& p, const AttPlan& pl, int n_img, cudaStream_t stream) {
  const bool exa = p.A == 256 * CA;
  switch (pl.nb) {
    case 1: return exa ? launch_att<1, CA, MT, true>(p, pl, n_img, stream) : launch_att<1, CA, MT, false>(p, pl, n_img, stream);
    case 2: return exa ? launch_att<2, CA, MT, true>(p, pl, n_img, stream) : launch_att<2, CA, MT, false>(p, pl, n_img, stream);
    default: return exa ? launch_att<3, CA, MT, true>(p, pl, n_img, stream) : launch_att<3, CA, MT, false>(p, pl, n_img, stream);
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
  if ((reinterpret_cast<uintptr_t>(att_h) & 15) || (ld_att_h % 4))
    return set_error(UIC_ERR_ALIGN, "att_step_fwd: att_h must be 16-byte aligned with a pitch that is a multiple of 4 floats");
  const AttPlan pl = make_plan(n_img, beams, L, A, H);
  const long long need = att_step_workspace_bytes(n_img, beams, L, A, H);
  if (workspace == nullptr || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15))
    return set_error(UIC_ERR_ARG, "att_step_fwd: workspace of %lld bytes (16-byte aligned, zeroed once) required, got %lld", need,
                     workspace_bytes);
  const long long jobs = static_cast<long long>(n_img) * pl.groups;
  if (jobs * pl.nbpi > 0x7fffffffLL) return set_error(UIC_ERR_SHAPE, "att_step_fwd: too many region batches");
  AttParams p{};
  p.att_h = att_h;
  p.ld_att_h = ld_att_h;
  p.p_att = static_cast<const __nv_bfloat16*>(p_att);
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
  p.ws_partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + ((jobs * ATT_WARPS * 4 + 255) / 256) * 256);
  p.beams = beams;
  p.L = L;
  p.A = A;
  p.H = H;
  p.n_grp = pl.groups;
  p.nbpi = pl.nbpi;
  p.segs = pl.segs;
  p.items = pl.items;
  p.f_bufs = pl.f_bufs;
  p.trace = gemm_trace_buffer();
  // Tiles that cannot stay L2-resident from one step to the next (> ~half of the 126 MB L2) are streamed evict-first;
  // smaller tile sets (36-region features) are worth keeping, they are re-read by the next step.
  const long long tile_bytes = static_cast<long long>(n_img) * L * (A + H) * 2;
  p.tile_policy = att_tile_policy(tile_bytes);
  // v7 (batch-balanced ranges, attention_v7.cu) where it applies; else the whole-job / equal-segment kernel of this file
  const int v7_ctas = (reinterpret_cast<uintptr_t>(w_alpha) & 15) ? 0 : att_v7_ctas(n_img, beams, L, A, H, pl);
  if (v7_ctas > 0) {
    p.n_batches = static_cast<int>(jobs * pl.nbpi);
    p.v7_flag = static_cast<int*>(workspace);
    p.v7_rec = reinterpret_cast<float4*>(static_cast<uint8_t*>(workspace) + ((static_cast<long long>(v7_ctas) * ATT_WARPS * 4 + 255) / 256) * 256);
    const int rc7 = att_step_fwd_v7(p, pl, n_img, v7_ctas, stream);
    if (rc7 <= 0) return rc7;  // 1: not launched, fall through
  }
  if (pl.nb > 3) return set_error(UIC_ERR_SHAPE, "att_step_fwd: %d beams per pass are planned for the v7 kernel, which did not launch", pl.nb);
  const int ca = (A + 255) / 256;
  if (H <= 512) {
    if (ca <= 1) return dispatch_nb<1, 4>(p, pl, n_img, stream);
    if (ca <= 2) return dispatch_nb<2, 4>(p, pl, n_img, stream);
    return dispatch_nb<4, 4>(p, pl, n_img, stream);
  }
  if (ca <= 2) return dispatch_nb<2, 8>(p, pl, n_img, stream);
  return dispatch_nb<4, 8>(p, pl, n_img, stream);
}

}  // namespace uic
