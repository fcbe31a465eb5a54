// An "expand = 1" inverted-residual block (FBNet xif2_2 / xif2_3: depthwise 3x3 + ReLU -> 1x1 24 -> 24 -> + x) in ONE kernel.
//
// Unfused, the block is dw_tma_kernel (x in, depthwise map out) followed by pw_small_const_kernel (depthwise map + x in,
// y out): 5 x the block's tensor through HBM (500 MB at 256 frames of 64 x 64 x 24) for 0.2 GFMA -- both kernels sit at
// 60-65 % of the DRAM bandwidth.  Here the depthwise map stays in shared memory: x is read once, y written once.
//
// Same pipeline as dw_tma_kernel (persistent CTAs, a 4-stage ring of 18 x 18 x 32-channel TMA boxes whose out-of-image
// halo and channels 24..31 are zero-filled by the TMA unit, two consumer groups of four warps on alternate tiles), and per
// tile and group three phases separated by a 128-thread named barrier:
//   1. depthwise: thread = (4-channel group, 2 x 8 pixel block), FFMA2, bias-then-(ky, kx) order  -> D[pixel][28] (smem)
//   2. pointwise: thread = 2 pixels; 24 inputs from D (LDS.128, 112-byte pixel pitch: conflict free), weights by value in
//      the constant bank, k-outer / o-inner FMA order of pw_small_const_kernel                      -> D in place
//   3. output   : thread = (4-channel group, 2 x 8 block) again: D + x (the centre of the TMA box)  -> global, 96 B / pixel
// Arithmetic and its order are those of the two-kernel path: bit-identical (tests/test_gpu_parity.py).
#pragma once
#include "kernels_dw_tma.cuh"
#include "kernels_ffma.cuh"

namespace fear {
namespace tc {

constexpr int kDpC = 24;                // channels of the block
constexpr int kDpTH = 16, kDpTW = 16;   // output tile
constexpr int kDpStages = 4, kDpGroups = 2, kDpGW = 4;
constexpr int kDpPitch = 28;            // floats per pixel in D (24 + 4: LDS.128 of 8 consecutive pixels hit 8 different bank quads)
constexpr int kDpDBytes = kDpTH * kDpTW * kDpPitch * 4;
using DpTile = DwTile<3, 1, kDpTH, kDpTW>;
constexpr int kDpSmemBytes = kDpStages * DpTile::kStageBytes + kDpGroups * kDpDBytes + 2 * kDpStages * 8 + 128;

__global__ void __launch_bounds__(kDpGroups * kDpGW * 32, 1)
dw3_pw24_fused_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW,
                      const __grid_constant__ CUtensorMap tmBias, const DwTmaParams p,
                      const __grid_constant__ PwSmallWeights<kDpC, kDpC> wts) {
  using T = DpTile;
  constexpr int K = 3, TX = 8, TY = 2, IW = T::IW, NIN = TX + K - 1, NR = TY + K - 1, PX = kDpTW / TX;
  constexpr int STAGES = kDpStages, GROUPS = kDpGroups, GW = kDpGW;

  extern __shared__ uint8_t dp_smem_raw[];
  uint8_t* smem = dp_smem_raw + ((128u - (smem_u32(dp_smem_raw) & 127u)) & 127u);
  uint8_t* dbuf = smem + STAGES * T::kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(dbuf + GROUPS * kDpDBytes);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GW);
    }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();

  auto issue_tile = [&](int tile, int stage) {
    const int tx = tile % p.tiles_x;
    int rest = tile / p.tiles_x;
    const int ty = rest % p.tiles_y;
    const int b = rest / p.tiles_y;
    uint8_t* st = smem + stage * T::kStageBytes;
    mbar_arrive_expect_tx(&full[stage], T::kInBytes + T::kWBytes + T::kBiasBytes);
    tma_load_4d(st, &tmIn, &full[stage], 0, tx * kDpTW - 1, ty * kDpTH - 1, b);
    tma_load_2d(st + T::kInBytes, &tmW, &full[stage], 0, 0);
    tma_load_2d(st + T::kInBytes + T::kWBytes, &tmBias, &full[stage], 0, 0);
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      if (tile < p.num_tiles) issue_tile(tile, i);
    }
  }

  const int cg = lane & 7;
  const int group = warp / GW, gwarp = warp % GW;
  const int tg = threadIdx.x & (GW * 32 - 1);  // thread inside the group
  float* D = reinterpret_cast<float*>(dbuf + group * kDpDBytes);
  const int pos = gwarp * 4 + (lane >> 3);      // 16 positions of 2 x 8 pixels
  const int ox_l = (pos % PX) * TX, oy_l = (pos / PX) * TY;
  int it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    if (it % GROUPS != group) continue;
    const int s = it % STAGES;
    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
    const int tx = tile % p.tiles_x;
    int rest = tile / p.tiles_x;
    const int ty = rest % p.tiles_y;
    const int b = rest / p.tiles_y;
    mbar_wait(&full[s], ph);
    const F4* in4 = reinterpret_cast<const F4*>(smem + s * T::kStageBytes);
    const F4* w4 = reinterpret_cast<const F4*>(smem + s * T::kStageBytes + T::kInBytes);
    const F4* b4p = reinterpret_cast<const F4*>(smem + s * T::kStageBytes + T::kInBytes + T::kWBytes);

    // ---- 1. depthwise 3x3 + bias + ReLU (as dw_tma_kernel<3, 1, 16, 16, 8, 2>) -> D ----
    if (cg < kDpC / 4) {
      F4 acc[TY][TX];
      const F4 bias4 = b4p[cg];
#pragma unroll
      for (int y = 0; y < TY; ++y)
#pragma unroll
        for (int t = 0; t < TX; ++t) acc[y][t] = bias4;
      F4 wk[K][K];
      const F4* base = in4 + (oy_l * IW + ox_l) * (kDwCB / 4) + cg;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        F4 v[NIN];
#pragma unroll
        for (int i = 0; i < NIN; ++i) v[i] = base[(r * IW + i) * (kDwCB / 4)];
        if (r < K) {
#pragma unroll
          for (int kx = 0; kx < K; ++kx) wk[r < K ? r : 0][kx] = w4[(r * K + kx) * (kDwCB / 4) + cg];
        }
#pragma unroll
        for (int y = 0; y < TY; ++y) {
          const int ky = r - y;
          if (ky >= 0 && ky < K) {
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
              const F4 k = wk[(ky >= 0 && ky < K) ? ky : 0][kx];
#pragma unroll
              for (int t = 0; t < TX; ++t) ffma2(acc[y][t].lo, v[t + kx].lo, k.lo);
#pragma unroll
              for (int t = 0; t < TX; ++t) ffma2(acc[y][t].hi, v[t + kx].hi, k.hi);
            }
          }
        }
      }
#pragma unroll
      for (int y = 0; y < TY; ++y)
#pragma unroll
        for (int t = 0; t < TX; ++t) {
          float4 r4 = f4_to_float4(acc[y][t]);
          r4.x = fmaxf(r4.x, 0.f);
          r4.y = fmaxf(r4.y, 0.f);
          r4.z = fmaxf(r4.z, 0.f);
          r4.w = fmaxf(r4.w, 0.f);
          *reinterpret_cast<float4*>(D + ((oy_l + y) * kDpTW + ox_l + t) * kDpPitch + cg * 4) = r4;
        }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(GW * 32) : "memory");

    // ---- 2. pointwise 24 -> 24 (+ bias), two pixels per thread, in place ----
#pragma unroll 1
    for (int pp = tg; pp < kDpTH * kDpTW; pp += GW * 32) {
      float* dp = D + pp * kDpPitch;
      float xin[kDpC];
#pragma unroll
      for (int i = 0; i < kDpC / 4; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(dp + 4 * i);
        xin[4 * i] = v.x;
        xin[4 * i + 1] = v.y;
        xin[4 * i + 2] = v.z;
        xin[4 * i + 3] = v.w;
      }
      float acc[kDpC];
#pragma unroll
      for (int o = 0; o < kDpC; ++o) acc[o] = wts.b[o];
#pragma unroll
      for (int k = 0; k < kDpC; ++k)
#pragma unroll
        for (int o = 0; o < kDpC; ++o) acc[o] = fmaf(xin[k], wts.w[k * kDpC + o], acc[o]);
#pragma unroll
      for (int i = 0; i < kDpC / 4; ++i)
        *reinterpret_cast<float4*>(dp + 4 * i) = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(GW * 32) : "memory");

    // ---- 3. + x (box centre), store ----
    if (cg < kDpC / 4) {
      float4* o = reinterpret_cast<float4*>(p.out) +
                  (((long long)b * p.Ho + ty * kDpTH + oy_l) * p.Wo + tx * kDpTW + ox_l) * (kDpC / 4) + cg;
      const float4* xc = reinterpret_cast<const float4*>(in4) + ((oy_l + 1) * IW + ox_l + 1) * (kDwCB / 4) + cg;
#pragma unroll
      for (int y = 0; y < TY; ++y)
#pragma unroll
        for (int t = 0; t < TX; ++t) {
          float4 r = *reinterpret_cast<const float4*>(D + ((oy_l + y) * kDpTW + ox_l + t) * kDpPitch + cg * 4);
          const float4 q = xc[(y * IW + t) * (kDwCB / 4)];
          r.x += q.x;
          r.y += q.y;
          r.z += q.z;
          r.w += q.w;
          o[((long long)y * p.Wo + t) * (kDpC / 4)] = r;
        }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&empty[s]);
      if (gwarp == 0) {  // refill this stage with the tile STAGES iterations ahead once the whole group has left it
        const int next = tile + STAGES * gridDim.x;
        if (next < p.num_tiles) {
          mbar_wait(&empty[s], ph);
          issue_tile(next, s);
        }
      }
    }
    // (no barrier needed before the group's next tile: phase 1 rewrites exactly the D elements this thread has just read in
    //  phase 3 -- same (block, channel group) mapping -- and phase 2 of that tile starts behind its own named barrier)
    __syncwarp();
  }
}

// x, y: [B][H][W][24] channels-last; dw_w [9][24], dw_b [24]; pw weights by value.  Returns 0 on launch, 1 when the shape
// is not covered (caller runs the two-kernel path), < 0 on error.
inline int launch_dw3_pw24(cudaStream_t s, const float* x, const float* dw_w, const float* dw_b,
                           const PwSmallWeights<kDpC, kDpC>& wts, float* y, int B, int H, int W, int num_sms) {
  if (H % kDpTH || W % kDpTW) return 1;
  if (attr_needed(reinterpret_cast<const void*>(dw3_pw24_fused_kernel))) {
    if (cudaFuncSetAttribute(dw3_pw24_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDpSmemBytes) != cudaSuccess)
      return -30;
  }
  CUtensorMap tmIn, tmW, tmB;
  int r = make_tmap_nhwc(&tmIn, x, (uint64_t)B, (uint64_t)H, (uint64_t)W, (uint64_t)kDpC, kDwCB, DpTile::IW, DpTile::IH);
  if (r) return r;
  r = make_tmap_2d_plain(&tmW, dw_w, 9, (uint64_t)kDpC, 9, kDwCB);
  if (r) return r;
  r = make_tmap_2d_plain(&tmB, dw_b, 1, (uint64_t)kDpC, 1, kDwCB);
  if (r) return r;
  DwTmaParams p;
  p.out = y;
  p.C4 = kDpC / 4;
  p.Ho = H;
  p.Wo = W;
  p.tiles_x = W / kDpTW;
  p.tiles_y = H / kDpTH;
  p.cblocks = 1;
  p.num_tiles = B * p.tiles_x * p.tiles_y;
  const int grid = num_sms < p.num_tiles ? num_sms : p.num_tiles;
  if (launch_pdl(dw3_pw24_fused_kernel, dim3(grid), dim3(kDpGroups * kDpGW * 32), (size_t)kDpSmemBytes, s, tmIn, tmW, tmB, p,
                 wts) != cudaSuccess)
    return -31;
  return 0;
}

}  // namespace tc
}  // namespace fear
