// Micro-benchmark: issue rate of the MUFU variants the attention-step kernel could use (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rates mufu_rates.cu ; prints lane-ops per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 1) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 2) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 6) asm volatile("tanh.approx.f16 %0, %1;" : "=h"(*reinterpret_cast<uint16_t*>(&y)) : "h"(static_cast<uint16_t>(x)));
  if (OP == 7) asm volatile("fma.rn.f16x2 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  if (OP == 8) asm volatile("fma.rn.bf16x2 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  if (OP == 9) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t* out, int iters, long long* cycles) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003800u + threadIdx.x * 8 + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<OP>(a[i]) ^ (OP == 3 ? 1u : 0u);
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_op) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 2 * 512 * 4);
  cudaMalloc(&cyc, 148 * 2 * 8);
  const int iters = 2000;
  k<OP><<<148 * 2, 512>>>(out, iters, cyc);
  k<OP><<<148 * 2, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 296; ++i) avg += h[i];
  avg /= 296;
  // per SM: 2 CTAs x 512 threads x iters x 8 ops in `avg` cycles
  const double warp_instr_per_clk_sm = 2.0 * 16 * iters * 8 / avg;
  printf("%-22s %7.3f warp-instr/clk/SM  = %6.1f results/clk/SM  (%s)\n", name, warp_instr_per_clk_sm, warp_instr_per_clk_sm * 32 * per_op,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

// The scoring inner loop of the attention-step kernel in isolation: per unit pair 2 FFMA (denominators), FMUL, MUFU.RCP,
// FMUL + FFMA (numerator), FFMA (accumulate); NCH independent accumulator chains per thread.
template <int NCH>
__global__ void __launch_bounds__(256) pair_loop(float* out, int iters, long long* cycles, const float* in) {
  float E[NCH][2], acc[NCH];
  const float f0 = in[threadIdx.x & 7], f1 = in[(threadIdx.x + 1) & 7], w0 = in[2], w1 = in[3];
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    E[i][0] = in[i & 7] + threadIdx.x * 1e-3f;
    E[i][1] = in[(i + 3) & 7];
    acc[i] = 0.0f;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const float d1 = fmaf(E[i][0], f0, 1.0f), d2 = fmaf(E[i][1], f1, 1.0f);
      float r;
      asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d1 * d2));
      acc[i] = fmaf(r, fmaf(w1, d1, w0 * d2), acc[i]);
      E[i][0] = __uint_as_float(__float_as_uint(E[i][0]) ^ (it & 1));  // keeps the loop body from being hoisted (ALU pipe)
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int NCH>
void run_pair(int ctas_per_sm) {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 2 * 256 * 4);
  cudaMalloc(&in, 64);
  cudaMalloc(&cyc, 148 * 2 * 8);
  float hin[8] = {0.5f, 1.5f, 0.25f, -0.3f, 2.0f, 0.7f, 1.1f, 0.9f};
  cudaMemcpy(in, hin, 32, cudaMemcpyHostToDevice);
  const int iters = 4000;
  for (int r = 0; r < 2; ++r) pair_loop<NCH><<<148 * ctas_per_sm, 256>>>(out, iters, cyc, in);
  cudaDeviceSynchronize();
  long long h[296];
  cudaMemcpy(h, cyc, 148 * ctas_per_sm * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148 * ctas_per_sm; ++i) avg += h[i];
  avg /= 148 * ctas_per_sm;
  const double pairs_per_clk_sm = static_cast<double>(ctas_per_sm) * 256 * iters * NCH / avg;
  printf("pair loop, %d chains, %d warps/SM: %6.1f unit pairs/clk/SM = %5.1f us for 38.5 M pairs on 148 SMs at 1.965 GHz (%s)\n", NCH,
         ctas_per_sm * 8, pairs_per_clk_sm, 38.5e6 / (pairs_per_clk_sm * 148 * 1965.0), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run_pair<4>(2);
  run_pair<12>(2);
  run_pair<12>(1);
  run_pair<24>(2);
  run<0>("tanh.approx.f32", 1);
  run<1>("tanh.approx.f16x2", 2);
  run<2>("tanh.approx.bf16x2", 2);
  run<6>("tanh.approx.f16", 1);
  run<3>("rcp.approx.ftz.f32", 1);
  run<4>("ex2.approx.ftz.f32", 1);
  run<5>("ex2.approx.f16x2", 2);
  run<7>("fma.rn.f16x2", 2);
  run<8>("fma.rn.bf16x2", 2);
  run<9>("fma.rn.f32", 1);
  return 0;
}
