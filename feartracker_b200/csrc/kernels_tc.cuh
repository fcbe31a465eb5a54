// tcgen05 / TMEM / TMA kernels of the FEAR-XS hot path (sm_100a only).
//
// corr_tc_kernel -- the pixel-wise template (x) search correlation (MobileCorrelation.forward's matmul,
// reference model_training/model/blocks.py:123) on the 5th-generation tensor cores:
//
//     s[b, p, k] = sum_c x[b, p, c] * z[b, k, c]        p in 256 search cells, k in 64 template cells, c in 256
//
// Channels-last operands are K-major GEMM operands as they lie in HBM: x = first 256 channels of the
// 320-channel concat buffer [B*256][320] (written there by the encode 1x1 conv), z = [Bz*64][256].
// The result goes straight into channels [256,320) of the same buffer, so the "torch.cat" of the
// reference costs nothing and the kernel moves exactly the algorithmic bytes (z + x in, s out).
//
// fp32 fidelity: kind::tf32 keeps 11 significand bits, which fails the 1e-3 parity bar (SURVEY.md
// section 7.4), so each operand is split on the fly into tf32 (hi, lo) pairs and three MMAs
// (hi*hi + lo*hi + hi*lo) accumulate in TMEM in fp32 -- error ~1e-6, still far above the FFMA rate.
//
// Structure (persistent, one CTA per SM, warp-specialised, mbarrier pipelines):
//   warp 0      TMA producer: per 32-channel chunk, x tile [128 p][32 c] + z tile [64 k][32 c] -> smem
//   warp 1      TMEM owner + MMA issuer: 4 K-steps x 3 MMAs (128x64x8) per chunk, tcgen05.commit
//   warps 2-5   operand split: raw fp32 tile -> tf32 hi (in place) and lo (second tile), same swizzled
//               positions, then fence.proxy.async so the tensor core sees the generic-proxy writes
//   warps 6-9   epilogue: tcgen05.ld 128x64 fp32 accumulator -> registers -> global (256 B per pixel)
// Two TMEM accumulators (2 x 64 columns) let the epilogue of tile t overlap the MMAs of tile t+1;
// a 4-stage smem ring (4 x 48 KB) keeps ~96 KB of loads in flight per SM.
#pragma once
#include <cuda_runtime.h>

#include "tc_common.cuh"

namespace fear {
namespace tc {

constexpr int kCorrStages = 4;
constexpr int kCorrChunk = 32;                 // channels per stage = one 128-byte swizzled row
constexpr int kCorrABytes = 128 * 128;         // [128 pixels][32 ch] fp32
constexpr int kCorrBBytes = 64 * 128;          // [64 template cells][32 ch] fp32
constexpr int kCorrStageBytes = 2 * (kCorrABytes + kCorrBBytes);
constexpr int kCorrSmemBytes = kCorrStages * kCorrStageBytes + 1024 /*align*/ + 256 /*barriers*/;
constexpr int kCorrThreads = 320;
constexpr int kCorrTmemCols = 128;

__global__ void __launch_bounds__(kCorrThreads, 1)
corr_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               float* __restrict__ cat, int num_frames, int z_broadcast) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCorrStages * kCorrStageBytes);
  uint64_t* full = bars;                       // [stages] TMA landed
  uint64_t* split = bars + kCorrStages;        // [stages] hi/lo tiles ready for the tensor core
  uint64_t* empty = bars + 2 * kCorrStages;    // [stages] MMAs reading the stage have completed
  uint64_t* acc_full = bars + 3 * kCorrStages; // [2] accumulator complete
  uint64_t* acc_empty = acc_full + 2;          // [2] accumulator drained by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = num_frames * 2;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kCorrTmemCols);
    tmem_relinquish();
  }
  if (threadIdx.x == 64) {
    for (int s = 0; s < kCorrStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 4);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto a_hi = [&](int s) { return smem + s * kCorrStageBytes; };
  auto a_lo = [&](int s) { return smem + s * kCorrStageBytes + kCorrABytes; };
  auto b_hi = [&](int s) { return smem + s * kCorrStageBytes + 2 * kCorrABytes; };
  auto b_lo = [&](int s) { return smem + s * kCorrStageBytes + 2 * kCorrABytes + kCorrBBytes; };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int frame = t >> 1, half = t & 1;
        const int arow = frame * 256 + half * 128;
        const int brow = z_broadcast ? 0 : frame * 64;
        for (int c = 0; c < 256 / kCorrChunk; ++c) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kCorrABytes + kCorrBBytes);
          tma_load_2d(a_hi(stage), &tmA, &full[stage], c * kCorrChunk, arow);
          tma_load_2d(b_hi(stage), &tmB, &full[stage], c * kCorrChunk, brow);
          if (++stage == kCorrStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, 64);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 64;
        for (int c = 0; c < 256 / kCorrChunk; ++c) {
          mbar_wait(&split[stage], phase);
          tc_fence_after();
          const uint32_t ah = smem_u32(a_hi(stage)), al = smem_u32(a_lo(stage));
          const uint32_t bh = smem_u32(b_hi(stage)), bl = smem_u32(b_lo(stage));
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 4 K-steps of 8 tf32 (32 B) inside the 128-B swizzle row
            const uint64_t dah = umma_desc_k_sw128(ah + j * 32), dal = umma_desc_k_sw128(al + j * 32);
            const uint64_t dbh = umma_desc_k_sw128(bh + j * 32), dbl = umma_desc_k_sw128(bl + j * 32);
            mma_tf32_ss(d, dah, dbh, idesc, (c | j) != 0);
            mma_tf32_ss(d, dal, dbh, idesc, 1);
            mma_tf32_ss(d, dah, dbl, idesc, 1);
          }
          tc_commit(&empty[stage]);  // stage reusable once these MMAs have read it
          if (++stage == kCorrStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(&acc_full[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp < 6) {
    // ===================================== operand split ====================================
    const int ts = threadIdx.x - 64;  // 0..127
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int c = 0; c < 256 / kCorrChunk; ++c) {
        mbar_wait(&full[stage], phase);
        float4* ah = reinterpret_cast<float4*>(a_hi(stage));
        float4* al = reinterpret_cast<float4*>(a_lo(stage));
#pragma unroll
        for (int i = 0; i < kCorrABytes / 16 / 128; ++i) {
          const float4 v = ah[ts + i * 128];
          float4 h, l;
          split_tf32(v.x, h.x, l.x);
          split_tf32(v.y, h.y, l.y);
          split_tf32(v.z, h.z, l.z);
          split_tf32(v.w, h.w, l.w);
          ah[ts + i * 128] = h;
          al[ts + i * 128] = l;
        }
        float4* bh = reinterpret_cast<float4*>(b_hi(stage));
        float4* bl = reinterpret_cast<float4*>(b_lo(stage));
#pragma unroll
        for (int i = 0; i < kCorrBBytes / 16 / 128; ++i) {
          const float4 v = bh[ts + i * 128];
          float4 h, l;
          split_tf32(v.x, h.x, l.x);
          split_tf32(v.y, h.y, l.y);
          split_tf32(v.z, h.z, l.z);
          split_tf32(v.w, h.w, l.w);
          bh[ts + i * 128] = h;
          bl[ts + i * 128] = l;
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&split[stage]);
        if (++stage == kCorrStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int frame = t >> 1, half = t & 1;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 64 + ((uint32_t)(q * 32) << 16);
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(taddr, r0);
      tmem_ld_32x32(taddr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);  // accumulator free for tile t+2
      const long long row = (long long)frame * 256 + half * 128 + q * 32 + lane;
      float4* dst = reinterpret_cast<float4*>(cat + row * 320 + 256);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_float4(__uint_as_float(r0[4 * i]), __uint_as_float(r0[4 * i + 1]), __uint_as_float(r0[4 * i + 2]),
                             __uint_as_float(r0[4 * i + 3]));
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[8 + i] = make_float4(__uint_as_float(r1[4 * i]), __uint_as_float(r1[4 * i + 1]),
                                 __uint_as_float(r1[4 * i + 2]), __uint_as_float(r1[4 * i + 3]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCorrTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static bool g_tc_ready = false;

inline int init() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return -1;
  g_num_sms = p.multiProcessorCount;
  if (resolve_driver()) return 0;  // tcgen05 path stays unavailable; the FFMA path still works
  if (cudaFuncSetAttribute(corr_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCorrSmemBytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  g_tc_ready = true;
  return 0;
}

inline bool available() { return g_tc_ready; }

inline int launch_corr(cudaStream_t s, const float* zt, int Bz, float* cat, int B) {
  if (!g_tc_ready) return -20;
  CUtensorMap tmA, tmB;
  int r = make_tmap_2d(&tmA, cat, (uint64_t)B * 256, 320, 320, 128, kCorrChunk);
  if (r) return r;
  r = make_tmap_2d(&tmB, zt, (uint64_t)Bz * 64, 256, 256, 64, kCorrChunk);
  if (r) return r;
  const int tiles = B * 2;
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  corr_tc_kernel<<<grid, kCorrThreads, kCorrSmemBytes, s>>>(tmA, tmB, cat, B, Bz == 1 ? 1 : 0);
  return 0;
}

inline bool pw_supported(int, int) { return false; }
inline int launch_pw(cudaStream_t, const float*, int, const float*, const float*, const float*, int, float*, int, int,
                     int, int, int) {
  return -1;
}

}  // namespace tc
}  // namespace fear
