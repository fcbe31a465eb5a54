// Microbenchmark: issue rate of tcgen05.mma (cta_group::1, M = 128) by kind (tf32 / bf16), operand form (A in shared
// memory / tensor memory) and N, on every SM at once.  Operands are zeros; only the timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I feartracker_b200/csrc -o tools/microbench/mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace fear::tc;

__device__ __forceinline__ void mma_ss(int bf16, uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (bf16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(int bf16, uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  if (bf16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc)
                 : "memory");
}

// style 0: `if (threadIdx.x == 0)` issues; style 1: the whole warp runs the loop, elect.sync picks the issuing lane per MMA.
// issuers: warps 0..issuers-1 each issue the full sequence into their own accumulator (N <= 128 / issuers columns apart).
__global__ void __launch_bounds__(128, 1) mma_rate2(int bf16, int N, int style, int issuers, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < issuers) {
    const uint32_t fmt = bf16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 16384;
    const uint32_t d = tm + warp * 128;
    const long long t0 = clock64();
    if (style == 0) {
      if (lane == 0) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_ss(bf16, d, umma_desc_k_sw128(a0 + j * 32), umma_desc_k_sw128(b0 + j * 32), idesc);
        }
        tc_commit(&bar[warp]);
      }
    } else {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (elect_one()) mma_ss(bf16, d, umma_desc_k_sw128(a0 + j * 32), umma_desc_k_sw128(b0 + j * 32), idesc);
      }
      if (elect_one()) tc_commit(&bar[warp]);
    }
    mbar_wait(&bar[warp], 0);
    const long long t1 = clock64();
    if (lane == 0 && warp == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

__global__ void __launch_bounds__(128, 1) mma_rate(int bf16, int ts, int N, int rotate, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x < 32) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    // idesc: D fp32; A/B format tf32 (2) or bf16 (1); N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t fmt = bf16 ? 1u : 2u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 16384;  // A 128 rows x 128 B, B up to 256 rows x 128 B
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t d = tm + (rotate ? ((r * 4 + j) & 1) * 256 : 0);
        if (ts) mma_ts(bf16, d, tm + 256 + 0 + j * 8, umma_desc_k_sw128(b0 + j * 32), idesc);
        else mma_ss(bf16, d, umma_desc_k_sw128(a0 + j * 32), umma_desc_k_sw128(b0 + j * 32), idesc);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int bf16 = 0; bf16 < 2; ++bf16)
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {32, 64, 128, 256})
        for (int rot = 0; rot < 2; ++rot) {
          if (rot && N > 128) continue;
          if (ts && rot) continue;
          mma_rate<<<148, 128, 100 * 1024>>>(bf16, ts, N, rot, 100, d_out);  // warm-up
          cudaEventRecord(e0);
          mma_rate<<<148, 128, 100 * 1024>>>(bf16, ts, N, rot, reps, d_out);
          cudaEventRecord(e1);
          if (cudaDeviceSynchronize() != cudaSuccess) {
            printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            return 1;
          }
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          long long h[148];
          cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (long long v : h) mx = v > mx ? v : mx;
          const double clk = (double)mx / (reps * 4), ns = ms * 1e6 / (reps * 4);
          const int K = bf16 ? 16 : 8;
          printf("%s %s N=%3d rotate=%d: %6.1f clk/MMA  %6.1f ns/MMA  %7.0f MAC/clk/SM  %6.1f TFLOP/s chip\n", bf16 ? "bf16" : "tf32",
                 ts ? "TS" : "SS", N, rot, clk, ns, 128.0 * N * K / clk, 2.0 * 128 * N * K * 148 / ns * 1e-3);
        }
  cudaFuncSetAttribute(mma_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int bf16 = 0; bf16 < 2; ++bf16)
    for (int N : {64, 128})
      for (int style = 0; style < 2; ++style)
        for (int issuers : {1, 2, 4}) {
          if (N * issuers > 512) continue;
          mma_rate2<<<148, 128, 100 * 1024>>>(bf16, N, style, issuers, 100, d_out);
          cudaEventRecord(e0);
          mma_rate2<<<148, 128, 100 * 1024>>>(bf16, N, style, issuers, reps, d_out);
          cudaEventRecord(e1);
          if (cudaDeviceSynchronize() != cudaSuccess) {
            printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            return 1;
          }
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          long long h[148];
          cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (long long v : h) mx = v > mx ? v : mx;
          const double n = (double)reps * 4 * issuers;
          printf("%s SS N=%3d style=%s issuers=%d: %6.1f clk/MMA (per SM)  %6.1f ns/MMA\n", bf16 ? "bf16" : "tf32", N,
                 style ? "elect" : "lane0", issuers, (double)mx / n, ms * 1e6 / n);
        }
  return 0;
}
