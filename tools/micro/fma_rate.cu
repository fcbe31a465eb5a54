// Microbenchmark: fp32 FMA issue rate on sm_100a -- scalar FFMA vs packed FFMA2, with and without a shared
// (reuse-cache friendly) multiplier, at 1 / 2 / 4 warps per SM sub-partition.  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  // 16 independent accumulators, 8 distinct a operands, 4 distinct b operands
  float acc[16], a[8], b[4];
  unsigned long long acc2[16], a2[8], b2[4];
  for (int i = 0; i < 16; ++i) { acc[i] = seed * i; acc2[i] = (unsigned long long)(threadIdx.x + i) * 0x3f8000003f800000ull; }
  for (int i = 0; i < 8; ++i) { a[i] = seed + i; a2[i] = 0x3f8000013f800001ull + i; }
  for (int i = 0; i < 4; ++i) { b[i] = seed * 0.5f + i; b2[i] = 0x3f0000003f000000ull + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) acc[i] = fmaf(a[(i + r) & 7], b[(i * 3 + r) & 3], acc[i]);       // all operands vary
        if (MODE == 1) acc[i] = fmaf(a[(i + r) & 7], b[r & 3], acc[i]);                 // b shared by 16 FMAs
        if (MODE == 2) ffma2(acc2[i], a2[(i + r) & 7], b2[(i * 3 + r) & 3]);
        if (MODE == 3) ffma2(acc2[i], a2[(i + r) & 7], b2[r & 3]);
      }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += acc[i] + (float)(acc2[i] & 0xff);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_smsp) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4);
  const int threads = warps_per_smsp * 4 * 32;
  const int iters = 20000;
  k<MODE><<<148, threads>>>(out, 100, 1.0f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double insts = (double)iters * 128 * warps_per_smsp;        // warp-instructions per SMSP
  const double fmas = insts * 32 * (MODE >= 2 ? 2 : 1) * 4 * 148;
  printf("%-28s warps/SMSP %d: %.3f ms  %.1f TFLOP/s  (%.3f inst/ns/SMSP)\n", name, warps_per_smsp, ms,
         2 * fmas / ms * 1e-9, insts / (ms * 1e6));
  cudaFree(out);
}

int main() {
  for (int w : {1, 2, 4}) {
    run<0>("FFMA distinct", w);
    run<1>("FFMA shared multiplier", w);
    run<2>("FFMA2 distinct", w);
    run<3>("FFMA2 shared multiplier", w);
  }
  return 0;
}
