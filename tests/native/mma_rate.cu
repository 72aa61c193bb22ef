// Microbenchmark (hardware only): issue rate of tcgen05.mma kind::f16, M = 128, K = 16, as a function of N and of where
// the A operand lives (TMEM as in the fused depth step, or shared memory).  One CTA per SM, one elected thread issues
// `reps` back-to-back MMAs into one accumulator (operands are whatever shared memory / TMEM hold: timing only), commits,
// waits; cycles per MMA = clock64 delta / reps.   ./mma_rate
#include <cuda.h>
#include <cstdio>
#include <cstdint>
#include "../../chemprop_b200/csrc/tc_common.cuh"
using namespace dmpnn::tc;

template <int GROUP, bool ONE_THREAD = false>
__global__ void __launch_bounds__(128, 1) k_rate(int N, int a_in_tmem, int reps, int nd, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint64_t bar_store;
  __shared__ uint32_t s_tmem;
  const uint32_t bar = smem_u32(&bar_store);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = s_tmem;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t bdesc = umma_desc_sw128(sbase);                 // B: N rows x 64 bf16 (SW128 K-major)
    const uint64_t adesc = umma_desc_sw128(sbase + 65536);         // A (smem variant): 128 rows x 64 bf16
    long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    for (int pass = 0; pass < 2; ++pass) {                          // pass 0 warms up
      t0 = clock64();
      if (ONE_THREAD) {
        if ((threadIdx.x & 31) == 0) {
          for (int i = 0; i < reps; i += GROUP) {
            const uint32_t d = tb + (uint32_t)(((i / GROUP) % nd) * 80);
#pragma unroll
            for (int g = 0; g < GROUP; ++g) {
              if (a_in_tmem) umma_bf16_ts(d, tb + 320 + 8 * (g & 3), bdesc + 2 * (g & 3), idesc, 1u);
              else umma_bf16(d, adesc + 2 * (g & 3), bdesc + 2 * (g & 3), idesc, 1u);
            }
          }
        }
        __syncwarp();
      } else {
      for (int i = 0; i < reps; i += GROUP) {
          if (elect_one()) {
            const uint32_t d = tb + (uint32_t)(((i / GROUP) % nd) * 80);   // nd accumulators side by side (80-column pitch)
  #pragma unroll
            for (int g = 0; g < GROUP; ++g) {
              if (a_in_tmem) umma_bf16_ts(d, tb + 320 + 8 * (g & 3), bdesc + 2 * (g & 3), idesc, 1u);
              else umma_bf16(d, adesc + 2 * (g & 3), bdesc + 2 * (g & 3), idesc, 1u);
            }
          }
          __syncwarp();
        }
  }
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, ph); ph ^= 1;
      t1 = clock64();
    }
    if (threadIdx.x == 32 && blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// two issuing warps, independent accumulators (warp w: columns 160 * w ..), 12 MMAs per issue block each
__global__ void __launch_bounds__(128, 1) k_rate2(int N, int reps, int n_issuers, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint64_t bar_store[2];
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar_store[0]), 1); mbar_init(smem_u32(&bar_store[1]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = s_tmem;
  if (warp >= 1 && warp <= n_issuers) {
    const int w = warp - 1;
    const uint32_t bar = smem_u32(&bar_store[w]);
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t bdesc = umma_desc_sw128(sbase + w * 32768);
    long long t0 = 0, t1 = 0;
    uint32_t ph = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      for (int i = 0; i < reps; i += 12) {
        if (elect_one()) {
          const uint32_t d = tb + (uint32_t)(w * 160 + ((i / 12) & 1) * 80);
#pragma unroll
          for (int g = 0; g < 12; ++g) umma_bf16_ts(d, tb + 320 + 8 * (g & 3), bdesc + 2 * (g & 3), idesc, 1u);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar);
      __syncwarp();
      mbar_wait(bar, ph); ph ^= 1;
      t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[w] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  int dev = 0; cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  if (pr.major != 10) { printf("no sm_100 device\n"); return 0; }
  long long* d_out; cudaMalloc(&d_out, 16);
  const int smem = 65536 + 16384 + 1024;
  cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 4096;
  printf("tcgen05.mma kind::f16 M=128 K=16, %d back-to-back MMAs per CTA, %d CTAs (clock64 cycles per MMA)\n", reps, pr.multiProcessorCount);
  auto run = [&](int N, int a_tm, int group, int nd) {
    if (group == 1) k_rate<1><<<pr.multiProcessorCount, 128, smem>>>(N, a_tm, reps, nd, d_out);
    else if (group == 4) k_rate<4><<<pr.multiProcessorCount, 128, smem>>>(N, a_tm, reps, nd, d_out);
    else if (group == -1) k_rate<1, true><<<pr.multiProcessorCount, 128, smem>>>(N, a_tm, reps, nd, d_out);
    else if (group == -4) k_rate<4, true><<<pr.multiProcessorCount, 128, smem>>>(N, a_tm, reps, nd, d_out);
    else k_rate<12><<<pr.multiProcessorCount, 128, smem>>>(N, a_tm, reps / 12 * 12, nd, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("  A in %-4s N=%3d  %2d MMAs per issue block (negative: one thread runs the loop, no elect / syncwarp), %d accumulator(s): %7.1f cycles/MMA   (N/2 = %5.1f)  %s\n", a_tm ? "TMEM" : "smem", N,
           group, nd, (double)c / (group == 12 ? reps / 12 * 12 : reps), N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  };
  for (int a_tm = 1; a_tm >= 1; --a_tm)
    for (int group : {12})
      for (int nd : {1, 4})
        for (int N : {16, 80, 160, 256}) {
          if (nd == 4 && N > 80) continue;
          run(N, a_tm, group, nd);
        }
  cudaFuncSetAttribute(k_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n_iss : {1, 2})
    for (int N : {16, 64, 80}) {
      const int r12 = reps / 12 * 12;
      k_rate2<<<pr.multiProcessorCount, 128, smem>>>(N, r12, n_iss, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      long long c[2] = {0, 0}; cudaMemcpy(c, d_out, 16, cudaMemcpyDeviceToHost);
      printf("  %d issuing warp(s), N=%3d, 12 MMAs per block: %7.1f cycles per MMA of the CTA (each warp %d MMAs in %lld / %lld cycles)  %s\n", n_iss, N,
             (double)(c[0] > c[1] ? c[0] : c[1]) / (r12 * n_iss), r12, c[0], n_iss > 1 ? c[1] : 0LL, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
