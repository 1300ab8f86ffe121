// Declarations shared by the two attention-step kernels (attention.cu: v6, whole jobs or equal segments per CTA;
// attention_v7.cu: batch-balanced ranges with owner-side merging).
#pragma once
#include <cstdlib>
#include <type_traits>

#include "uic_internal.h"
#include "uic_ptx.cuh"

namespace uic {

constexpr int ATT_WARPS = 8;
constexpr int ATT_THREADS = 32 * ATT_WARPS;
constexpr int ATT_STAGES = 3;
constexpr int ATT_BATCH = 16;  // regions per batch = the K of one MMA
constexpr int ATT_SLAB_BYTES = ATT_BATCH * 128;  // one TMA box: 16 regions x 64 bf16 columns, 128-byte swizzle

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
  float* ws_partial;   // [job][segment][warp][lane][4 + 4 MT]
  int* ws_counter;     // [job][ATT_WARPS], zero between launches
  int beams, L, A, H;
  int n_grp;           // beam groups per image (job = img * n_grp + grp)
  int nbpi;            // batches per image
  int segs;            // segments per job (1: CTAs own whole jobs, no merging; > 1: one CTA per segment)
  int items;           // jobs * segs
  int f_bufs;          // att_h buffers in shared memory (2, or 3 when every image is a single batch)
  int slab_map;        // the tensor map is the 3-D slab view: one TMA instruction stages a batch's att rows
  long long* trace;    // debug (uic_gemm_set_trace buffer): CTA 0 records globaltimer at its pipeline events
  unsigned long long tile_policy;  // L2 eviction hint of the feature-tile loads (uic_ptx.cuh)
  // v7 (attention_v7.cu): CTA c owns the batches [c n_batches / ctas, (c + 1) n_batches / ctas)
  int n_batches;       // jobs * nbpi
  float4* v7_rec;      // [cta][warp][1 + MT][lane]: partial (max, sum, accumulators) of the job a CTA starts in the middle of
  int* v7_flag;        // [cta][warp]: record published (zero between launches: the reader clears it)
};

struct AttPlan {
  int nb, groups, nbpi, ctas, segs, items, mt, f_bufs;
};
// v7 launcher (attention_v7.cu).  Returns 1 when the shape is not one of its instantiations (caller falls back to v6).
int att_step_fwd_v7(AttParams& p, const AttPlan& pl, int n_img, int ctas, cudaStream_t stream);
int att_v7_ctas(int n_img, int beams, int L, int A, int H, const AttPlan& pl);  // 0: shape not covered by v7
long long att_v7_workspace_bytes(int ctas, int mt);
// Largest number of beams of an image the v7 kernel scores in one pass for this problem (3: the general limit).
int att_v7_beam_cap(int n_img, int beams, int L, int A, int H);

__device__ __forceinline__ void bulk_g2s_hint(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_dst),
               "l"(gsrc), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float att_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ long long att_now() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ATT_TRACE(slot)                                       \
  do {                                                       \
    if (tracing && (slot) < 128) p.trace[slot] = att_now();  \
  } while (0)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttSmem {  // byte offsets from the 1024-byte aligned base of a CTA's dynamic shared memory
  uint32_t p_off, stage_bytes, off_F, off_e, off_bar, total;
};
__host__ __device__ inline AttSmem att_smem_layout(int A, int H, int NB, int f_bufs) {
  AttSmem s;
  s.p_off = (H + 63) / 64 * ATT_SLAB_BYTES;  // att boxes first (their swizzle atoms need 1024-byte alignment), p_att rows after
  s.stage_bytes = (s.p_off + ATT_BATCH * A * 2 + 1023) / 1024 * 1024;
  s.off_F = ATT_STAGES * s.stage_bytes;                       // [f_bufs][NB][A] fp32
  s.off_e = s.off_F + f_bufs * NB * A * 4;                    // [STAGES][NB][16] fp32
  s.off_bar = s.off_e + ATT_STAGES * NB * ATT_BATCH * 4;      // full / scored / consumed [STAGES] mbarriers
  s.total = s.off_bar + 3 * ATT_STAGES * 8 + 1024;            // + alignment slack
  return s;
}

}  // namespace uic
