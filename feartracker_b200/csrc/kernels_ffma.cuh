// CUDA-core (FFMA) kernels of the FEAR-XS hot path, channels-last (NHWC) fp32.
//
// These are the always-correct implementations every stage can run on; the tcgen05 kernels in
// kernels_tc.cuh replace the dense contractions (1x1 convs, correlation) where they apply.
// Layout: activations [B][H][W][C] fp32, C a multiple of 4 (float4 over channels); 1x1 weights
// torch-native [Cout][Cin] ("K-major"); depthwise weights [k*k][C]; stem weights [27][16].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fear_b200.h"

namespace fear {

// ------------------------------------------------------------------------------------------
// Stem: conv3x3 stride 2 pad 1, 3 -> 16, + folded BN bias + ReLU.   NCHW in -> NHWC out.
// (fbnet_c xif0_0; reference call site fear_net.py:58-61.)  One thread per output pixel, all 16
// output channels in registers; the 432 weights are broadcast from shared memory.
// ------------------------------------------------------------------------------------------
// U8 = true: img is a uint8 HWC crop (B,H,W,3) as the tracker holds it; the ImageNet normalisation of
// Tracker._preprocess_image (reference base_tracker.py:69-81,97-103: (x - mean*255) * (1/(std*255)), float32,
// subtract then multiply) is applied on the fly with the same two roundings, so the result is bit-identical
// to normalising on the host while the host->device copy shrinks 4x.
struct StemNorm {
  float mean[3], inv[3];
};
template <bool U8>
__global__ void __launch_bounds__(128) stem_conv3x3s2_kernel(const void* __restrict__ img_, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int B, int H, int W, StemNorm nrm) {
  __shared__ float sw[27 * 16];
  __shared__ float sb[16];
  for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 16) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  // CTA = 32 output columns x 4 output rows (one warp per row): vertically adjacent rows share an input
  // row, which now hits L1 instead of being re-read from L2 by another CTA.
  const int Ho = H >> 1, Wo = W >> 1;
  const int xg = (Wo + 31) / 32, yg = (Ho + 3) / 4;
  int t = blockIdx.x;
  const int bx = t % xg;
  t /= xg;
  const int by = t % yg;
  const int b = t / yg;
  const int ox = bx * 32 + (threadIdx.x & 31);
  const int oy = by * 4 + (threadIdx.x >> 5);
  if (ox >= Wo || oy >= Ho || b >= B) return;
  const long long idx = ((long long)b * Ho + oy) * Wo + ox;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = sb[c];
  const float* base = static_cast<const float*>(img_) + (long long)b * 3 * H * W;
  const uint8_t* base8 = static_cast<const uint8_t*>(img_) + (long long)b * 3 * H * W;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          if (U8)
            v = __fmul_rn(__fsub_rn((float)__ldg(base8 + ((long long)iy * W + ix) * 3 + ci), nrm.mean[ci]), nrm.inv[ci]);
          else
            v = __ldg(base + ((long long)ci * H + iy) * W + ix);
        }
        const float* wr = sw + (ci * 9 + ky * 3 + kx) * 16;
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
      }
    }
  }
  float4* o = reinterpret_cast<float4*>(out + idx * 16);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    o[q] = make_float4(fmaxf(acc[4 * q], 0.f), fmaxf(acc[4 * q + 1], 0.f), fmaxf(acc[4 * q + 2], 0.f),
                       fmaxf(acc[4 * q + 3], 0.f));
}

// ------------------------------------------------------------------------------------------
// Depthwise KxK conv (pad K/2, stride S) NHWC, optional bias / ReLU.  One thread per
// (pixel, 4 channels): neighbouring threads walk the channel dimension => 16-byte coalesced
// loads; the K*K taps of a pixel hit L1 (each input value is reused by up to K*K/S^2 outputs).
// Backbone dw (+BN+ReLU): mobile_cv IRF block; head dw: SepConv.depthwise, blocks.py:57-66.
// ------------------------------------------------------------------------------------------
template <int K, int S, bool RELU, bool BIAS>
__global__ void __launch_bounds__(256) dw_conv_nhwc_kernel(const float4* __restrict__ in, const float4* __restrict__ w,
                                                           const float4* __restrict__ bias, float4* __restrict__ out,
                                                           int B, int H, int W, int C4) {
  const int Ho = H / S, Wo = W / S;
  const long long total = (long long)B * Ho * Wo * C4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  const long long pix = idx / C4;
  const int ox = (int)(pix % Wo);
  const int oy = (int)((pix / Wo) % Ho);
  const int b = (int)(pix / ((long long)Wo * Ho));
  constexpr int P = K / 2;
  float4 acc = BIAS ? __ldg(bias + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* inb = in + (long long)b * H * W * C4;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const int iy = oy * S - P + ky;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const int ix = ox * S - P + kx;
      if (ix < 0 || ix >= W) continue;
      const float4 v = __ldg(inb + ((long long)iy * W + ix) * C4 + c4);
      const float4 k = __ldg(w + (ky * K + kx) * C4 + c4);
      acc.x = fmaf(v.x, k.x, acc.x);
      acc.y = fmaf(v.y, k.y, acc.y);
      acc.z = fmaf(v.z, k.z, acc.z);
      acc.w = fmaf(v.w, k.w, acc.w);
    }
  }
  if (RELU) {
    acc.x = fmaxf(acc.x, 0.f);
    acc.y = fmaxf(acc.y, 0.f);
    acc.z = fmaxf(acc.z, 0.f);
    acc.w = fmaxf(acc.w, 0.f);
  }
  out[idx] = acc;
}

// ------------------------------------------------------------------------------------------
// Depthwise KxK conv, register-strip version: one thread produces TX consecutive output pixels of a
// row for 4 channels.  Each input row segment (TX*S + K - S float4) and the K weights of that row are
// loaded once and reused by all TX outputs, cutting L1 wavefronts per output ~3x against the
// one-pixel-per-thread kernel above (which is L1-wavefront bound, not HBM bound).
// Requires Wo % TX == 0.
// ------------------------------------------------------------------------------------------
template <int K, int S, int TX, bool RELU, bool BIAS>
__global__ void __launch_bounds__(TX >= 16 ? 128 : 256) dw_conv_strip_kernel(const float4* __restrict__ in, const float4* __restrict__ w,
                                                            const float4* __restrict__ bias, float4* __restrict__ out,
                                                            int B, int H, int W, int C4) {
  const int Ho = H / S, Wo = W / S;
  const int strips = Wo / TX;
  const long long total = (long long)B * Ho * strips * C4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  long long rest = idx / C4;
  const int sx = (int)(rest % strips);
  rest /= strips;
  const int oy = (int)(rest % Ho);
  const int b = (int)(rest / Ho);
  constexpr int P = K / 2;
  constexpr int NIN = (TX - 1) * S + K;  // input columns feeding TX outputs
  const int ox0 = sx * TX;
  const int ix0 = ox0 * S - P;
  float4 acc[TX];
  const float4 b4 = BIAS ? __ldg(bias + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int t = 0; t < TX; ++t) acc[t] = b4;
  const float4* inb = in + (long long)b * H * W * C4 + c4;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const int iy = oy * S - P + ky;
    if (iy < 0 || iy >= H) continue;
    const float4* row = inb + (long long)iy * W * C4;
    float4 v[NIN];
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
      const int ix = ix0 + i;
      v[i] = (ix >= 0 && ix < W) ? __ldg(row + (long long)ix * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 k = __ldg(w + (ky * K + kx) * C4 + c4);
#pragma unroll
      for (int t = 0; t < TX; ++t) {
        const float4 x = v[t * S + kx];
        acc[t].x = fmaf(x.x, k.x, acc[t].x);
        acc[t].y = fmaf(x.y, k.y, acc[t].y);
        acc[t].z = fmaf(x.z, k.z, acc[t].z);
        acc[t].w = fmaf(x.w, k.w, acc[t].w);
      }
    }
  }
  float4* o = out + (((long long)b * Ho + oy) * Wo + ox0) * C4 + c4;
#pragma unroll
  for (int t = 0; t < TX; ++t) {
    float4 r = acc[t];
    if (RELU) {
      r.x = fmaxf(r.x, 0.f);
      r.y = fmaxf(r.y, 0.f);
      r.z = fmaxf(r.z, 0.f);
      r.w = fmaxf(r.w, 0.f);
    }
    o[(long long)t * C4] = r;
  }
}

// ------------------------------------------------------------------------------------------
// Register-strip depthwise conv with an L1-friendly thread layout.  Same arithmetic as
// dw_conv_strip_kernel, but a CTA now owns a compact patch -- 8 channel groups (32 channels) x 4 adjacent
// strips x 8 output rows -- instead of 256 consecutive channel groups of one strip.  Neighbouring rows and
// strips share (K-1)/K of their inputs, and with this layout those re-reads hit L1 (ncu showed the flat
// layout pulling ~2.5x the algorithmic bytes through L2 on the 5x5 layers).  A warp covers 8 channel
// groups of 4 strips in one row => each load instruction touches four 128-byte segments.
// ------------------------------------------------------------------------------------------
template <int K, int S, int TX, bool RELU, bool BIAS>
__global__ void __launch_bounds__(256) dw_conv_strip_blocked_kernel(const float4* __restrict__ in,
                                                                    const float4* __restrict__ w,
                                                                    const float4* __restrict__ bias,
                                                                    float4* __restrict__ out, int B, int H, int W,
                                                                    int C4) {
  const int Ho = H / S, Wo = W / S;
  const int strips = Wo / TX;
  const int sgroups = (strips + 3) / 4, rgroups = (Ho + 7) / 8, cgroups = (C4 + 7) / 8;
  int t = blockIdx.x;
  const int cg = t % cgroups;
  t /= cgroups;
  const int sg = t % sgroups;
  t /= sgroups;
  const int rg = t % rgroups;
  const int b = t / rgroups;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int c4 = cg * 8 + (lane & 7);
  const int sx = sg * 4 + (lane >> 3);
  const int oy = rg * 8 + wrp;
  if (c4 >= C4 || sx >= strips || oy >= Ho) return;
  constexpr int P = K / 2;
  constexpr int NIN = (TX - 1) * S + K;
  const int ox0 = sx * TX;
  const int ix0 = ox0 * S - P;
  float4 acc[TX];
  const float4 b4 = BIAS ? __ldg(bias + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < TX; ++i) acc[i] = b4;
  const float4* inb = in + (long long)b * H * W * C4 + c4;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const int iy = oy * S - P + ky;
    if (iy < 0 || iy >= H) continue;
    const float4* row = inb + (long long)iy * W * C4;
    float4 v[NIN];
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
      const int ix = ix0 + i;
      v[i] = (ix >= 0 && ix < W) ? __ldg(row + (long long)ix * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 k = __ldg(w + (ky * K + kx) * C4 + c4);
#pragma unroll
      for (int i = 0; i < TX; ++i) {
        const float4 x = v[i * S + kx];
        acc[i].x = fmaf(x.x, k.x, acc[i].x);
        acc[i].y = fmaf(x.y, k.y, acc[i].y);
        acc[i].z = fmaf(x.z, k.z, acc[i].z);
        acc[i].w = fmaf(x.w, k.w, acc[i].w);
      }
    }
  }
  float4* o = out + (((long long)b * Ho + oy) * Wo + ox0) * C4 + c4;
#pragma unroll
  for (int i = 0; i < TX; ++i) {
    float4 r = acc[i];
    if (RELU) {
      r.x = fmaxf(r.x, 0.f);
      r.y = fmaxf(r.y, 0.f);
      r.z = fmaxf(r.z, 0.f);
      r.w = fmaxf(r.w, 0.f);
    }
    o[(long long)i * C4] = r;
  }
}

// ------------------------------------------------------------------------------------------
// Depthwise KxK conv, rolling-window version.  One thread owns (4 channels, TX output columns) and
// walks ROWS output rows top to bottom: all K*K weights live in registers, every input row segment
// is loaded ONCE and scattered into a ring of ceil(K/S) live output-row accumulators, so a 5x5 conv
// issues ~3 loads per output float4 instead of ~16 (strip) / 50 (per-pixel).  That moves the kernel
// from L1-wavefront-bound to FFMA/HBM-bound.  Fully unrolled => all ring indices are compile time.
// Accumulation order per output (bias, then ky, kx ascending) is identical to the other two kernels.
// ------------------------------------------------------------------------------------------
template <int K, int S, int TX, int ROWS, bool RELU, bool BIAS>
__global__ void __launch_bounds__(128) dw_conv_roll_kernel(const float4* __restrict__ in, const float4* __restrict__ w,
                                                           const float4* __restrict__ bias, float4* __restrict__ out,
                                                           int B, int H, int W, int C4) {
  constexpr int P = K / 2;
  constexpr int NIN = (TX - 1) * S + K;
  constexpr int LIVE = (K + S - 1) / S;
  constexpr int NR = (ROWS - 1) * S + K;  // input rows feeding ROWS output rows
  const int Ho = H / S, Wo = W / S;
  const int strips = Wo / TX, segs = (Ho + ROWS - 1) / ROWS;
  const long long total = (long long)B * segs * strips * C4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c4 = (int)(idx % C4);
  long long rest = idx / C4;
  const int sx = (int)(rest % strips);
  rest /= strips;
  const int seg = (int)(rest % segs);
  const int b = (int)(rest / segs);
  const int oy0 = seg * ROWS, ox0 = sx * TX;
  const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;

  float4 wr[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) wr[i] = __ldg(w + i * C4 + c4);
  const float4 b4 = BIAS ? __ldg(bias + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc[LIVE][TX];
#pragma unroll
  for (int l = 0; l < LIVE; ++l)
#pragma unroll
    for (int t = 0; t < TX; ++t) acc[l][t] = b4;

  const float4* inb = in + (long long)b * H * W * C4 + c4;
  float4* outb = out + (long long)b * Ho * Wo * C4 + c4;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int iy = iy0 + r;
    if (iy >= 0 && iy < H) {
      const float4* row = inb + (long long)iy * W * C4;
      float4 v[NIN];
#pragma unroll
      for (int i = 0; i < NIN; ++i) {
        const int ix = ix0 + i;
        v[i] = (ix >= 0 && ix < W) ? __ldg(row + (long long)ix * C4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int ky = K - 1; ky >= 0; --ky) {  // descending ky = ascending output row; order per output unchanged
        if ((r - ky) >= 0 && (r - ky) % S == 0 && (r - ky) / S < ROWS) {
          constexpr int dummy = 0;
          (void)dummy;
          const int slot = ((r - ky) / S) % LIVE;
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const float4 k = wr[ky * K + kx];
#pragma unroll
            for (int t = 0; t < TX; ++t) {
              const float4 x = v[t * S + kx];
              acc[slot][t].x = fmaf(x.x, k.x, acc[slot][t].x);
              acc[slot][t].y = fmaf(x.y, k.y, acc[slot][t].y);
              acc[slot][t].z = fmaf(x.z, k.z, acc[slot][t].z);
              acc[slot][t].w = fmaf(x.w, k.w, acc[slot][t].w);
            }
          }
        }
      }
    }
    if (r >= K - 1 && (r - (K - 1)) % S == 0) {  // output row o is complete after input row r
      const int o = (r - (K - 1)) / S;
      const int slot = o % LIVE;
      if (oy0 + o < Ho) {
        float4* orow = outb + ((long long)(oy0 + o) * Wo + ox0) * C4;
#pragma unroll
        for (int t = 0; t < TX; ++t) {
          float4 res = acc[slot][t];
          if (RELU) {
            res.x = fmaxf(res.x, 0.f);
            res.y = fmaxf(res.y, 0.f);
            res.z = fmaxf(res.z, 0.f);
            res.w = fmaxf(res.w, 0.f);
          }
          orow[(long long)t * C4] = res;
        }
      }
#pragma unroll
      for (int t = 0; t < TX; ++t) acc[slot][t] = b4;
    }
  }
}

// ------------------------------------------------------------------------------------------
// 1x1 conv / correlation as a GEMM on CUDA cores:  C[M][N] = A[M][K] * Bw[N][K]^T (+bias)(+R)(ReLU)
// A rows = pixels (lda floats apart), Bw rows = output channels (ldb apart), both K-contiguous.
// Tile 128 x BN x 16, 256 threads as 32 (rows, 4 each) x 8 (cols, TN = BN/8 each).
// gridDim.z batches independent problems (per-frame correlation: A/C strided per frame, Bw
// strided per frame or shared when strideB == 0 -- template batch-1 broadcast, blocks.py:123).
// ------------------------------------------------------------------------------------------
template <int TN>
__global__ void __launch_bounds__(256) gemm_nt_ffma_kernel(const float* __restrict__ A, int lda, long long strideA,
                                                           const float* __restrict__ Bw, int ldb, long long strideB,
                                                           const float* __restrict__ bias, const float* __restrict__ R,
                                                           int ldr, float* __restrict__ C, int ldc, long long strideC,
                                                           int M, int N, int K, int relu) {
  constexpr int BM = 128, BK = 16, BN = 8 * TN, TM = 4;
  constexpr int AS = BM + 4;  // padded row length of the k-major A tile
  __shared__ __align__(16) float As[BK][AS];
  __shared__ __align__(16) float Bs[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  A += (long long)blockIdx.z * strideA;
  Bw += (long long)blockIdx.z * strideB;
  C += (long long)blockIdx.z * strideC;
  if (R) R += (long long)blockIdx.z * strideC;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: 128 rows x 16 k = 512 float4, two per thread; stored transposed (k-major).
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int id = tid + it * 256;
      const int r = id >> 2, kq = (id & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && k0 + kq < K) v = __ldg(reinterpret_cast<const float4*>(A + (long long)(m0 + r) * lda + k0 + kq));
      As[kq + 0][r] = v.x;
      As[kq + 1][r] = v.y;
      As[kq + 2][r] = v.z;
      As[kq + 3][r] = v.w;
    }
    // B tile: BN rows (output channels) x 16 k.
    for (int id = tid; id < BN * 4; id += 256) {
      const int n = id >> 2, kq = (id & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + n < N && k0 + kq < K) v = __ldg(reinterpret_cast<const float4*>(Bw + (long long)(n0 + n) * ldb + k0 + kq));
      Bs[kq + 0][n] = v.x;
      Bs[kq + 1][n] = v.y;
      Bs[kq + 2][n] = v.z;
      Bs[kq + 3][n] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
      const float a[TM] = {a4.x, a4.y, a4.z, a4.w};
      float bv[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + ty * TM + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int col = n0 + tx * TN + j;
      if (col >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + col);
      if (R) v += __ldg(R + (long long)row * ldr + col);
      if (relu) v = fmaxf(v, 0.f);
      C[(long long)row * ldc + col] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Depthwise KxK stride-1 conv, shared-memory tiled: one CTA = one frame x 16x16 output tile x 32-channel
// slab.  The (16+K-1)^2 x 32-channel input patch is staged once in shared memory (coalesced 128-byte
// rows, zero halo), so every input value is fetched from L2/HBM exactly once per tile instead of being
// re-requested through L1 by up to K*K neighbouring threads (the strip kernel moves ~2.5x the algorithmic
// bytes through L2 on the 16x16 stage).  Each thread owns 4 channels x 8 consecutive pixels of a row;
// per kernel row it reads 8+K-1 inputs and K weights from shared memory for 8*K FMA4.
// Accumulation order per output is the same as in the other depthwise kernels (bit-identical results).
// ------------------------------------------------------------------------------------------
template <int K, bool RELU, bool BIAS>
__global__ void __launch_bounds__(256, 2) dw_conv_tile_kernel(const float4* __restrict__ in, const float4* __restrict__ w,
                                                           const float4* __restrict__ bias, float4* __restrict__ out,
                                                           int H, int W, int C4) {
  constexpr int T = 16, TX = 8, P = K / 2, IT = T + K - 1, CS4 = 8;  // 32-channel slab = 8 float4
  extern __shared__ __align__(16) float4 dw_tile_smem[];
  float4* sIn = dw_tile_smem;              // [IT * IT][CS4]
  float4* sW = sIn + IT * IT * CS4;        // [K * K][CS4]
  const int tid = threadIdx.x;
  const int slabs = C4 / CS4;
  const int tiles_x = W / T, tiles_y = H / T;
  int t = blockIdx.x;
  const int slab = t % slabs;
  t /= slabs;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int c40 = slab * CS4;
  const int y0 = ty * T - P, x0 = tx * T - P;
  const float4* inb = in + (long long)b * H * W * C4 + c40;
  for (int i = tid; i < IT * IT * CS4; i += 256) {
    const int q = i % CS4, p = i / CS4;
    const int iy = y0 + p / IT, ix = x0 + p % IT;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(inb + ((long long)iy * W + ix) * C4 + q);
    sIn[i] = v;
  }
  for (int i = tid; i < K * K * CS4; i += 256) sW[i] = __ldg(w + (i / CS4) * C4 + c40 + (i % CS4));
  __syncthreads();

  const int q = tid % CS4;           // channel group inside the slab
  const int strip = tid / CS4;       // 32 strips: 16 rows x 2 half-rows
  const int oy = strip >> 1, ox0 = (strip & 1) * TX;
  const float4 b4 = BIAS ? __ldg(bias + c40 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc[TX];
#pragma unroll
  for (int i = 0; i < TX; ++i) acc[i] = b4;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    float4 v[TX + K - 1];
    const float4* row = sIn + ((oy + ky) * IT + ox0) * CS4 + q;
#pragma unroll
    for (int i = 0; i < TX + K - 1; ++i) v[i] = row[i * CS4];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 k = sW[(ky * K + kx) * CS4 + q];
#pragma unroll
      for (int i = 0; i < TX; ++i) {
        acc[i].x = fmaf(v[i + kx].x, k.x, acc[i].x);
        acc[i].y = fmaf(v[i + kx].y, k.y, acc[i].y);
        acc[i].z = fmaf(v[i + kx].z, k.z, acc[i].z);
        acc[i].w = fmaf(v[i + kx].w, k.w, acc[i].w);
      }
    }
  }
  float4* o = out + (long long)b * H * W * C4 + ((long long)(ty * T + oy) * W + tx * T + ox0) * C4 + c40 + q;
#pragma unroll
  for (int i = 0; i < TX; ++i) {
    float4 r = acc[i];
    if (RELU) {
      r.x = fmaxf(r.x, 0.f);
      r.y = fmaxf(r.y, 0.f);
      r.z = fmaxf(r.z, 0.f);
      r.w = fmaxf(r.w, 0.f);
    }
    o[(long long)i * C4] = r;
  }
}

template <int K>
constexpr int dw_tile_smem_bytes() {
  return 16 * ((16 + K - 1) * (16 + K - 1) * 8 + K * K * 8);
}

// ------------------------------------------------------------------------------------------
// Fused inverted-residual front half for the stride-2 blocks (xif2_0, xif3_0, xif4_0):
//     D = ReLU(dw_KxK_s2(ReLU(X * W1^T + b1)) + b2)
// The 6x-expanded tensor E = ReLU(X*W1^T + b1) (1.6 GB per 256-frame step for xif2_0 alone) never leaves
// the SM: each CTA takes a TH x TW tile of OUTPUT pixels, stages the (2TH+K-2) x (2TW+K-2) input patch of X,
// expands it into shared memory with CUDA cores (Cin is only 16..32 here, so this is ~1 FLOP per byte of
// the unfused traffic), then runs the depthwise stride-2 window over the resident tile and writes D.
// Pixels of the patch that fall outside the image are forced to 0 (the depthwise conv zero-pads E, and
// E(0-input) = ReLU(b1) != 0).
// ------------------------------------------------------------------------------------------
template <int CIN, int MID, int MSL, int K, int TH, int TW, int THREADS>
__global__ void __launch_bounds__(THREADS) fused_expand_dw_s2_kernel(
    const float* __restrict__ X, const float* __restrict__ w1, const float* __restrict__ b1,
    const float* __restrict__ wd, const float* __restrict__ bd, float* __restrict__ D, int H, int W) {
  // MSL = channels of the expanded tensor resident at a time (MID / MSL passes over the same input patch):
  // a smaller slice buys a larger spatial tile, i.e. less halo recomputation for the 5x5 blocks.
  static_assert(MID % MSL == 0 && MSL % 4 == 0 && CIN % 4 == 0, "channel slicing");
  constexpr int P = K / 2;
  constexpr int IH = 2 * TH + K - 2, IW = 2 * TW + K - 2, NPIX = IH * IW;
  constexpr int C4 = MSL / 4, K4 = CIN / 4;
  extern __shared__ __align__(16) float fsm[];
  float* sX = fsm;                        // [NPIX][CIN]
  float* sW1 = sX + NPIX * CIN;           // [CIN][MID]  (transposed: channel-contiguous per k)
  float* sB1 = sW1 + CIN * MID;           // [MID]
  float* sWd = sB1 + MID;                 // [K*K][MID]
  float* sBd = sWd + K * K * MID;         // [MID]
  float* sE = sBd + MID;                  // [NPIX][MSL]
  const int tid = threadIdx.x;
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_x = Wo / TW, tiles_y = Ho / TH;
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int b = t / tiles_y;
  const int oy0 = ty * TH, ox0 = tx * TW;
  const int iy0 = 2 * oy0 - P, ix0 = 2 * ox0 - P;

  // ---- stage weights and the input patch --------------------------------------------------
  for (int i = tid; i < CIN * MID; i += THREADS) {
    const int o = i / CIN, k = i - o * CIN;  // w1 is [MID][CIN]
    sW1[k * MID + o] = __ldg(w1 + i);
  }
  for (int i = tid; i < MID; i += THREADS) {
    sB1[i] = __ldg(b1 + i);
    sBd[i] = __ldg(bd + i);
  }
  for (int i = tid; i < K * K * MID; i += THREADS) sWd[i] = __ldg(wd + i);
  const float* Xb = X + (long long)b * H * W * CIN;
  for (int i = tid; i < NPIX * K4; i += THREADS) {
    const int p = i / K4, q = i - p * K4;
    const int iy = iy0 + p / IW, ix = ix0 + p % IW;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(reinterpret_cast<const float4*>(Xb + ((long long)iy * W + ix) * CIN) + q);
    reinterpret_cast<float4*>(sX)[i] = v;
  }
  __syncthreads();

  float* Db = D + (long long)b * Ho * Wo * MID;
  constexpr int PG = (NPIX + 3) / 4;
#pragma unroll 1
  for (int m0 = 0; m0 < MID; m0 += MSL) {
    // ---- expand: E[p][c] = ReLU(b1[c] + sum_k X[p][k] W1[c][k]), 4 pixels x 4 channels per work item ----
    for (int u = tid; u < PG * C4; u += THREADS) {
      const int c4 = u % C4, pg = u / C4;
      const int ch = m0 + 4 * c4;
      const float4 bb = *reinterpret_cast<const float4*>(sB1 + ch);
      float4 acc[4] = {bb, bb, bb, bb};
#pragma unroll
      for (int kq = 0; kq < K4; ++kq) {
        float4 xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = pg * 4 + j;
          xv[j] = (p < NPIX) ? *reinterpret_cast<const float4*>(sX + p * CIN + 4 * kq) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const float4 wv = *reinterpret_cast<const float4*>(sW1 + (4 * kq + kk) * MID + ch);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xs = kk == 0 ? xv[j].x : kk == 1 ? xv[j].y : kk == 2 ? xv[j].z : xv[j].w;
            acc[j].x = fmaf(xs, wv.x, acc[j].x);
            acc[j].y = fmaf(xs, wv.y, acc[j].y);
            acc[j].z = fmaf(xs, wv.z, acc[j].z);
            acc[j].w = fmaf(xs, wv.w, acc[j].w);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int p = pg * 4 + j;
        if (p < NPIX) {
          const int iy = iy0 + p / IW, ix = ix0 + p % IW;
          const bool inside = iy >= 0 && iy < H && ix >= 0 && ix < W;
          float4 e = make_float4(fmaxf(acc[j].x, 0.f), fmaxf(acc[j].y, 0.f), fmaxf(acc[j].z, 0.f), fmaxf(acc[j].w, 0.f));
          if (!inside) e = make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(sE + p * MSL + 4 * c4) = e;
        }
      }
    }
    __syncthreads();

    // ---- depthwise KxK stride 2 over the resident slice -------------------------------------
    for (int u = tid; u < TH * TW * C4; u += THREADS) {
      const int c4 = u % C4, op = u / C4;
      const int oy = op / TW, ox = op % TW;
      const int ch = m0 + 4 * c4;
      float4 acc = *reinterpret_cast<const float4*>(sBd + ch);
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float4 e = *reinterpret_cast<const float4*>(sE + ((2 * oy + ky) * IW + 2 * ox + kx) * MSL + 4 * c4);
          const float4 k = *reinterpret_cast<const float4*>(sWd + (ky * K + kx) * MID + ch);
          acc.x = fmaf(e.x, k.x, acc.x);
          acc.y = fmaf(e.y, k.y, acc.y);
          acc.z = fmaf(e.z, k.z, acc.z);
          acc.w = fmaf(e.w, k.w, acc.w);
        }
      acc.x = fmaxf(acc.x, 0.f);
      acc.y = fmaxf(acc.y, 0.f);
      acc.z = fmaxf(acc.z, 0.f);
      acc.w = fmaxf(acc.w, 0.f);
      *reinterpret_cast<float4*>(Db + ((long long)(oy0 + oy) * Wo + ox0 + ox) * MID + ch) = acc;
    }
    __syncthreads();  // the slice buffer is rewritten by the next pass
  }
}

template <int CIN, int MID, int MSL, int K, int TH, int TW>
constexpr int fused_expand_dw_smem_bytes() {
  constexpr int NPIX = (2 * TH + K - 2) * (2 * TW + K - 2);
  return 4 * (NPIX * CIN + CIN * MID + MID + K * K * MID + MID + NPIX * MSL);
}

// ------------------------------------------------------------------------------------------
// Tiny 1x1 convs (Cin, Cout <= 32: xif1_0.pwl 16->16, xif2_2/2_3.pwl 24->24) are pure streaming:
// ~1 FLOP per byte, millions of pixels.  A tensor-core tile pipeline only adds per-tile latency there
// (measured: 585 us vs the 125 us HBM time), so these run one pixel per thread on CUDA cores with the
// [Cin][Cout] weights broadcast from shared memory.   out = act(x * W^T + b (+ residual)).
// ------------------------------------------------------------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(256) pw_small_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, const float* __restrict__ res,
                                                       float* __restrict__ out, long long M, int relu) {
  __shared__ __align__(16) float sw[CIN * COUT];  // [k][o]
  __shared__ __align__(16) float sb[COUT];
  for (int i = threadIdx.x; i < CIN * COUT; i += blockDim.x) {
    const int o = i / CIN, k = i - o * CIN;  // global layout [o][k]
    sw[k * COUT + o] = w[i];
  }
  if (threadIdx.x < COUT) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float xin[CIN];
  const float4* xp = reinterpret_cast<const float4*>(x + m * CIN);
#pragma unroll
  for (int i = 0; i < CIN / 4; ++i) {
    const float4 v = __ldg(xp + i);
    xin[4 * i] = v.x;
    xin[4 * i + 1] = v.y;
    xin[4 * i + 2] = v.z;
    xin[4 * i + 3] = v.w;
  }
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = sb[o];
#pragma unroll
  for (int k = 0; k < CIN; ++k) {
#pragma unroll
    for (int o4 = 0; o4 < COUT / 4; ++o4) {
      const float4 wv = *reinterpret_cast<const float4*>(&sw[k * COUT + 4 * o4]);
      acc[4 * o4] = fmaf(xin[k], wv.x, acc[4 * o4]);
      acc[4 * o4 + 1] = fmaf(xin[k], wv.y, acc[4 * o4 + 1]);
      acc[4 * o4 + 2] = fmaf(xin[k], wv.z, acc[4 * o4 + 2]);
      acc[4 * o4 + 3] = fmaf(xin[k], wv.w, acc[4 * o4 + 3]);
    }
  }
  float4* op = reinterpret_cast<float4*>(out + m * COUT);
  const float4* rp = res ? reinterpret_cast<const float4*>(res + m * COUT) : nullptr;
#pragma unroll
  for (int o4 = 0; o4 < COUT / 4; ++o4) {
    float4 r = make_float4(acc[4 * o4], acc[4 * o4 + 1], acc[4 * o4 + 2], acc[4 * o4 + 3]);
    if (rp) {
      const float4 q = __ldg(rp + o4);
      r.x += q.x;
      r.y += q.y;
      r.z += q.z;
      r.w += q.w;
    }
    if (relu) {
      r.x = fmaxf(r.x, 0.f);
      r.y = fmaxf(r.y, 0.f);
      r.z = fmaxf(r.z, 0.f);
      r.w = fmaxf(r.w, 0.f);
    }
    op[o4] = r;
  }
}

// Same layer with the weights passed BY VALUE (kernel parameter = constant bank): every FFMA takes its weight as a
// uniform-register / constant operand instead of a shared-memory broadcast.  An LDS costs one LSU wavefront per
// 4 bytes even when all lanes read the same address, which made the smem version LSU-bound (CIN*COUT/4 LDS.128
// per pixel against CIN*COUT FMAs).  Same accumulation order => bit-identical results.
template <int CIN, int COUT>
struct PwSmallWeights {
  float w[CIN * COUT];  // [k][o]
  float b[COUT];
};
template <int CIN, int COUT>
__global__ void __launch_bounds__(256) pw_small_const_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                             float* __restrict__ out, long long M, int relu,
                                                             const __grid_constant__ PwSmallWeights<CIN, COUT> wts) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float xin[CIN];
  const float4* xp = reinterpret_cast<const float4*>(x + m * CIN);
#pragma unroll
  for (int i = 0; i < CIN / 4; ++i) {
    const float4 v = __ldg(xp + i);
    xin[4 * i] = v.x;
    xin[4 * i + 1] = v.y;
    xin[4 * i + 2] = v.z;
    xin[4 * i + 3] = v.w;
  }
  float4 q[COUT / 4];
  const float4* rp = res ? reinterpret_cast<const float4*>(res + m * COUT) : nullptr;
  if (rp) {
#pragma unroll
    for (int o4 = 0; o4 < COUT / 4; ++o4) q[o4] = __ldg(rp + o4);  // issued early: overlaps the FMAs
  }
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = wts.b[o];
#pragma unroll
  for (int k = 0; k < CIN; ++k)
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = fmaf(xin[k], wts.w[k * COUT + o], acc[o]);
  float4* op = reinterpret_cast<float4*>(out + m * COUT);
#pragma unroll
  for (int o4 = 0; o4 < COUT / 4; ++o4) {
    float4 r = make_float4(acc[4 * o4], acc[4 * o4 + 1], acc[4 * o4 + 2], acc[4 * o4 + 3]);
    if (rp) {
      r.x += q[o4].x;
      r.y += q[o4].y;
      r.z += q[o4].z;
      r.w += q[o4].w;
    }
    if (relu) {
      r.x = fmaxf(r.x, 0.f);
      r.y = fmaxf(r.y, 0.f);
      r.z = fmaxf(r.z, 0.f);
      r.w = fmaxf(r.w, 0.f);
    }
    op[o4] = r;
  }
}

// ------------------------------------------------------------------------------------------
// Batched 2-D transpose with leading dimensions: out[b][j][i] = in[b][i][j], i < R, j < Cn.
// Used for NCHW <-> NHWC at the API boundary (the reference API is NCHW, fear_net.py:58-96).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int ldin, long long strideIn,
                                                        float* __restrict__ out, int ldout, long long strideOut, int R,
                                                        int Cn) {
  __shared__ float tile[32][33];
  in += (long long)blockIdx.z * strideIn;
  out += (long long)blockIdx.z * strideOut;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    if (r < R && c < Cn) tile[i][tx] = in[(long long)r * ldin + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < Cn) out[(long long)c * ldout + r] = tile[tx][i];
  }
}

// ------------------------------------------------------------------------------------------
// Prediction 1x1 conv (256 -> NOUT, NOUT = 4 | 1) fused with the BoxTower epilogue
// (blocks.py:187-188,192):  bbox = exp(adjust * pred + bias), cls = 0.1 * pred.  adjust / 0.1 /
// biases are folded into w, b on the host, so this is  out = f(w . t + b).  One warp per pixel,
// NHWC in, NCHW out (B, NOUT, 16, 16) -- the layout FEARNet returns.
// ------------------------------------------------------------------------------------------
template <int NOUT, bool EXP>
__global__ void __launch_bounds__(256) pred_pw_kernel(const float* __restrict__ t, const float* __restrict__ w,
                                                      const float* __restrict__ b, float* __restrict__ out, int B) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B * 256) return;
  const float4* tp = reinterpret_cast<const float4*>(t + (long long)warp * 256);
  const float4 v0 = __ldg(tp + lane), v1 = __ldg(tp + 32 + lane);
  float acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    const float4* wp = reinterpret_cast<const float4*>(w + o * 256);
    const float4 w0 = __ldg(wp + lane), w1 = __ldg(wp + 32 + lane);
    float s = v0.x * w0.x;
    s = fmaf(v0.y, w0.y, s);
    s = fmaf(v0.z, w0.z, s);
    s = fmaf(v0.w, w0.w, s);
    s = fmaf(v1.x, w1.x, s);
    s = fmaf(v1.y, w1.y, s);
    s = fmaf(v1.z, w1.z, s);
    s = fmaf(v1.w, w1.w, s);
    acc[o] = s;
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], d);
  if (lane < NOUT) {
    float v = 0.f;
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (lane == o) v = acc[o];
    v += __ldg(b + lane);
    if (EXP) v = expf(v);
    const int frame = warp >> 8, p = warp & 255;
    out[((long long)frame * NOUT + lane) * 256 + p] = v;
  }
}

// ------------------------------------------------------------------------------------------
// Box decode (FEARTracker._postprocess + FEARBoxCoder.decode, fear_tracker.py:74-86,
// box_coder.py:75-107): score = sigmoid(cls) in fp32, argmax = first maximum in row-major
// order, box = [gx - l, gy - t, (gx + r) - (gx - l), (gy + b) - (gy - t)] evaluated in double
// (the reference's grid is float64, utils/utils.py:183-199, so torch promotes).  One 256-thread
// block per frame.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) decode_kernel(const float* __restrict__ bbox, const float* __restrict__ cls,
                                                     int apply_sigmoid, FearBox* __restrict__ boxes) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int f = blockIdx.x, t = threadIdx.x;
  float v = cls[(long long)f * 256 + t];
  if (apply_sigmoid) v = 1.0f / (1.0f + expf(-v));
  int i = t;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, d);
    const int oi = __shfl_xor_sync(0xffffffffu, i, d);
    if (ov > v || (ov == v && oi < i)) {
      v = ov;
      i = oi;
    }
  }
  if ((t & 31) == 0) {
    sv[t >> 5] = v;
    si[t >> 5] = i;
  }
  __syncthreads();
  if (t == 0) {
    for (int k = 1; k < 8; ++k)
      if (sv[k] > v || (sv[k] == v && si[k] < i)) {
        v = sv[k];
        i = si[k];
      }
    const int r = i >> 4, c = i & 15;
    const double gx = (double)((c - 8) * 16 + 128), gy = (double)((r - 8) * 16 + 128);
    const float* bb = bbox + (long long)f * 4 * 256 + i;
    const double x1 = gx - (double)bb[0], y1 = gy - (double)bb[256];
    const double x2 = gx + (double)bb[512], y2 = gy + (double)bb[768];
    FearBox o;
    o.x = x1;
    o.y = y1;
    o.w = x2 - x1;
    o.h = y2 - y1;
    o.score = v;
    o.row = r;
    o.col = c;
    o.flat = i;
    boxes[f] = o;
  }
}

}  // namespace fear
