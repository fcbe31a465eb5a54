// irf_s2_fused_kernel -- the inverted-residual block xif2_0 of fbnet_c as ONE kernel (sm_100a):
//
//     Y = W2 * relu(dw3x3_s2(relu(W1 * X + b1)) + bd) + b2          16 -> 96 -> 96 -> 24 channels, 2x down
//
// (mobile_cv IRF block pw -> dw -> pwl as restated in oracle/fbnet_c.py:95-110; call site reference
// model_training/model/blocks.py:27-35).  Unfused this block is three kernels that write and re-read the 6x
// expanded tensor E = relu(W1 X + b1) and the depthwise map: 4.3 GB of HBM traffic per 256-frame step for 0.37 GB
// of block input + output.  Here E and the depthwise output never leave the SM.
//
// One persistent CTA per SM walks 8 x 16 output-pixel tiles (= one M = 128 tile of the project GEMM).  Per tile:
//
//   loader warps (4)   read the 17 x 33 input pixels the tile needs (halo of the stride-2 3x3 window; 561 pixels,
//                      64 B each) straight from global memory one tile AHEAD, split them into tf32 (hi, lo) and park
//                      them in TENSOR MEMORY as the A operand of the expand GEMM (5 M-tiles x (16 + 16) columns):
//                      X never touches shared memory.  They also run the project epilogue (+b2, store Y).
//   MMA thread         expand: per 32-channel slab c and M-tile i, D[128 x 32] = A_i * W1[c]^T as 3xTF32 with the A
//                      operand in TMEM ("TS" form) into a ring of eight 32-column TMEM accumulators;
//                      project: acc2[128 x 24] += dw_c * W2[:, c]^T (A = depthwise output tile in smem, hi/lo).
//   worker groups (2x4 warps)  take alternate slabs: (1) expand epilogue TMEM -> +b1, ReLU, zero outside the image
//                      (the depthwise conv zero-pads E, and E(0) = relu(b1) != 0) -> 561 x 32-channel fp32 slab in
//                      shared memory; (2) depthwise 3x3 stride 2 out of that slab (packed FFMA2, same FMA order
//                      as the stand-alone depthwise kernels) -> +bd, ReLU -> (hi, lo) A tile of the project GEMM in
//                      the SWIZZLE_128B K-major layout.  Group A's TMEM reads overlap group B's FMAs.
//
// All MMA orders / operand splits / epilogue additions replicate pw_tc_kernel, so the block is bit-identical to the
// three-kernel path (tests/test_gpu_parity.py::test_fused_irf_block_is_bit_identical).
//
// TMEM (512 columns): [0,160) A operand (5 x (16 hi + 16 lo)); [160,416) 8 expand accumulators x 32;
// [416,448) project main, [448,480) project correction accumulator.
// smem: A2 (hi, lo) 32 KB | weights image 40.25 KB (W1 [hi|lo] rows, W2 [hi;lo] x 3 chunks, dw, biases) | barriers |
// 2 x 70.1 KB slabs = 214 KB.
#pragma once
#include "tc_common.cuh"

namespace fear {
namespace tc {

constexpr int kIrfCin = 16, kIrfMid = 96, kIrfCout = 24, kIrfCoutPad = 32;
constexpr int kIrfTH = 8, kIrfTW = 16;                          // output tile
constexpr int kIrfIH = 2 * kIrfTH + 1, kIrfIW = 2 * kIrfTW + 1;  // 17 x 33 input pixels
constexpr int kIrfPix = kIrfIH * kIrfIW;                        // 561
constexpr int kIrfMT = (kIrfPix + 127) / 128;                   // 5 M-tiles of the expand GEMM
constexpr int kIrfSlabs = kIrfMid / 32;                         // 3
constexpr int kIrfUnitsPerTile = kIrfSlabs * kIrfMT;            // 15
constexpr int kIrfRing = 8;                                     // expand accumulator ring (TMEM)
constexpr int kIrfThreads = 512;                                // 16 warps: MMA, 3 spare, 4 loader, 2 x 4 workers

// weights image (floats), copied verbatim into shared memory
constexpr int kIrfW1Floats = kIrfMid * 32;                        // [96 rows][hi 16 | lo 16], SWIZZLE_128B
constexpr int kIrfW2Floats = kIrfSlabs * 2 * kIrfCoutPad * 32;    // per chunk: [hi 32 rows ; lo 32 rows] x 32 k
constexpr int kIrfDwFloats = 9 * kIrfMid;
constexpr int kIrfImageFloats = kIrfW1Floats + kIrfW2Floats + kIrfDwFloats + kIrfMid + kIrfMid + kIrfCoutPad;

constexpr int kIrfOffA2 = 0;                                       // hi 16 KB, lo 16 KB
constexpr int kIrfOffImg = 32768;
constexpr int kIrfOffW1 = kIrfOffImg;
constexpr int kIrfOffW2 = kIrfOffW1 + kIrfW1Floats * 4;
constexpr int kIrfOffDw = kIrfOffW2 + kIrfW2Floats * 4;
constexpr int kIrfOffB1 = kIrfOffDw + kIrfDwFloats * 4;
constexpr int kIrfOffBd = kIrfOffB1 + kIrfMid * 4;
constexpr int kIrfOffB2 = kIrfOffBd + kIrfMid * 4;
constexpr int kIrfOffBars = ((kIrfOffB2 + kIrfCoutPad * 4 + 255) / 256) * 256;
constexpr int kIrfOffSlab = kIrfOffBars + 256;
constexpr int kIrfSlabBytes = kIrfPix * 128;
constexpr int kIrfSmemBytes = kIrfOffSlab + 2 * kIrfSlabBytes + 1024 /*alignment slack*/;
static_assert(kIrfSmemBytes <= 232448, "fused IRF kernel exceeds the 227 KB shared-memory limit");
static_assert(kIrfOffW1 % 1024 == 0 && kIrfOffW2 % 1024 == 0, "UMMA tiles must be 1024-byte aligned");

constexpr int kIrfColA = 0, kIrfColAcc = 160, kIrfColP = 416;  // TMEM column map
constexpr int kIrfTmemCols = 512;

struct IrfParams {
  const float* X;      // [B][H][W][16]
  float* Y;            // [B][H/2][W/2][24]
  const float* image;  // kIrfImageFloats packed weights (device)
  int B, H, W;         // input map
  int tiles_x, tiles_y, num_tiles;
};

// Host: build the shared-memory weights image.  w1_hi/w1_lo [96][16], w2_hi/w2_lo [24][96] (tf32-split copies the
// tensor-core path already keeps), dw [9][96] (tap-major), biases.
inline void irf_build_image(float* img, const float* w1_hi, const float* w1_lo, const float* w2_hi, const float* w2_lo,
                            const float* dw, const float* b1, const float* bd, const float* b2) {
  for (int i = 0; i < kIrfImageFloats; ++i) img[i] = 0.f;
  float* W1 = img;
  for (int r = 0; r < kIrfMid; ++r)
    for (int j = 0; j < 8; ++j) {  // logical 16-byte chunk j of row r: j < 4 -> hi[4j..], else lo[4(j-4)..]
      const float* src = (j < 4 ? w1_hi : w1_lo) + r * kIrfCin + 4 * (j & 3);
      float* dst = W1 + r * 32 + 4 * (j ^ (r & 7));
      for (int e = 0; e < 4; ++e) dst[e] = src[e];
    }
  float* W2 = img + kIrfW1Floats;
  for (int c = 0; c < kIrfSlabs; ++c)
    for (int part = 0; part < 2; ++part)
      for (int n = 0; n < kIrfCout; ++n)
        for (int j = 0; j < 8; ++j) {
          const float* src = (part == 0 ? w2_hi : w2_lo) + n * kIrfMid + 32 * c + 4 * j;
          float* dst = W2 + ((c * 2 + part) * kIrfCoutPad + n) * 32 + 4 * (j ^ (n & 7));
          for (int e = 0; e < 4; ++e) dst[e] = src[e];
        }
  float* p = img + kIrfW1Floats + kIrfW2Floats;
  for (int i = 0; i < kIrfDwFloats; ++i) p[i] = dw[i];
  p += kIrfDwFloats;
  for (int i = 0; i < kIrfMid; ++i) p[i] = b1[i];
  p += kIrfMid;
  for (int i = 0; i < kIrfMid; ++i) p[i] = bd[i];
  p += kIrfMid;
  for (int i = 0; i < kIrfCout; ++i) p[i] = b2[i];
}

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

__global__ void __launch_bounds__(kIrfThreads, 1) irf_s2_fused_kernel(const IrfParams p) {
  extern __shared__ uint8_t irf_smem_raw[];
  uint8_t* smem = irf_smem_raw + ((1024u - (smem_u32(irf_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kIrfOffBars);
  uint64_t* a_full = bars;                 // A operand of tile t is in TMEM            (4 loader warps)
  uint64_t* a_empty = bars + 1;            // expand MMAs of the tile have read it      (commit)
  uint64_t* acc_full = bars + 2;           // [8] expand accumulator complete           (commit)
  uint64_t* acc_empty = bars + 10;         // [8] drained by the worker group           (4 warps)
  uint64_t* a2_full = bars + 18;           // depthwise (hi, lo) tile written           (4 warps)
  uint64_t* a2_empty = bars + 19;          // project MMAs have read it                 (commit)
  uint64_t* acc2_full = bars + 20;         // project accumulator complete              (commit)
  uint64_t* acc2_empty = bars + 21;        // drained by the loader warps               (4 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(tmem_slot, kIrfTmemCols);
    tmem_relinquish();
  }
  if (threadIdx.x == 32) {
    mbar_init(a_full, 4);
    mbar_init(a_empty, 1);
    for (int s = 0; s < kIrfRing; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_init(a2_full, 4);
    mbar_init(a2_empty, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 4);
    fence_mbar_init();
  }
  {  // weights image -> shared memory (constant data: may be read before the previous kernel has finished)
    const float4* src = reinterpret_cast<const float4*>(p.image);
    float4* dst = reinterpret_cast<float4*>(smem + kIrfOffImg);
    for (int i = threadIdx.x; i < kIrfImageFloats / 4; i += kIrfThreads) dst[i] = __ldg(src + i);
    fence_proxy_async_smem();  // the tensor core reads W1 / W2 through the async proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();  // X is written by the previous kernel in the stream

  const int Ho = p.H >> 1, Wo = p.W >> 1;
  const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
  auto tile_coords = [&](int tl, int& b, int& oy0, int& ox0) {
    const int tile = blockIdx.x + tl * gridDim.x;
    const int tx = tile % p.tiles_x;
    const int rest = tile / p.tiles_x;
    const int ty = rest % p.tiles_y;
    b = rest / p.tiles_y;
    oy0 = ty * kIrfTH;
    ox0 = tx * kIrfTW;
  };

  if (warp == 0) {
    // ===================================== MMA issuer =====================================
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc_e = umma_idesc_tf32(128, 32);   // expand: N = 32 channel slab
      constexpr uint32_t idesc_p = umma_idesc_tf32(128, 32);   // project: a_lo x w_hi
      constexpr uint32_t idesc_p2 = umma_idesc_tf32(128, 64);  // project: a_hi x [w_hi ; w_lo] -> main | corr
      const uint32_t w1 = smem_u32(smem + kIrfOffW1), w2 = smem_u32(smem + kIrfOffW2);
      const uint32_t a2h = smem_u32(smem + kIrfOffA2), a2l = a2h + 16384;
      const int units = my_tiles * kIrfUnitsPerTile, projects = my_tiles * kIrfSlabs;
      auto do_project = [&](int pj) {
        const int tl = pj / kIrfSlabs, c = pj - tl * kIrfSlabs;
        if (c == 0) {
          mbar_wait(acc2_empty, (uint32_t)((tl & 1) ^ 1));
          tc_fence_after();
        }
        mbar_wait(a2_full, (uint32_t)(pj & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + kIrfColP;
        const uint32_t bh = w2 + c * 8192;  // [hi 32 rows ; lo 32 rows] x 128 B
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t dah = umma_desc_k_sw128(a2h + j * 32), dal = umma_desc_k_sw128(a2l + j * 32);
          const uint64_t dbh = umma_desc_k_sw128(bh + j * 32);
          mma_tf32_ss(d, dah, dbh, idesc_p2, (c | j) != 0);  // main += a_hi w_hi ; corr (+)= a_hi w_lo
          mma_tf32_ss(d + 32, dal, dbh, idesc_p, 1);         // corr += a_lo w_hi
        }
        tc_commit(a2_empty);
        if (c == kIrfSlabs - 1) tc_commit(acc2_full);
      };
      int next_project = 0;
      for (int u = 0; u < units; ++u) {
        // project pj runs right before expand unit 5 * pj + 13 (see the deadlock analysis in DESIGN.md)
        if (u >= 13 && (u - 13) % kIrfMT == 0 && next_project < projects) do_project(next_project++);
        const int tl = u / kIrfUnitsPerTile, ui = u - tl * kIrfUnitsPerTile;
        const int c = ui / kIrfMT, i = ui - c * kIrfMT;
        if (ui == 0) {
          mbar_wait(a_full, (uint32_t)(tl & 1));
          tc_fence_after();
        }
        const int slot = u % kIrfRing;
        mbar_wait(&acc_empty[slot], (uint32_t)(((u / kIrfRing) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d = tmem_base + kIrfColAcc + slot * 32;
        const uint32_t a_hi = tmem_base + kIrfColA + i * 32, a_lo = a_hi + 16;
        const uint32_t brow = w1 + c * 4096;  // rows [32c, 32c + 32) of W1: [hi 64 B | lo 64 B] per row
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // K = 16 = 2 K-steps; same product order as pw_tc_kernel's single-accumulator path
          const uint64_t dbh = umma_desc_k_sw128(brow + j * 32), dbl = umma_desc_k_sw128(brow + 64 + j * 32);
          mma_tf32_ts(d, a_hi + j * 8, dbh, idesc_e, j != 0);
          mma_tf32_ts(d, a_lo + j * 8, dbh, idesc_e, 1);
          mma_tf32_ts(d, a_hi + j * 8, dbl, idesc_e, 1);
        }
        tc_commit(&acc_full[slot]);
        if (ui == kIrfUnitsPerTile - 1) tc_commit(a_empty);
      }
      while (next_project < projects) do_project(next_project++);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================== loader + project epilogue =====================================
    const int q = warp & 3;
    const int row = q * 32 + lane;  // TMEM lane = row inside every M-tile
    float4 xr[kIrfMT][4];
    auto prefetch = [&](int tl) {
      int b, oy0, ox0;
      tile_coords(tl, b, oy0, ox0);
#pragma unroll
      for (int i = 0; i < kIrfMT; ++i) {
        const int pb = i * 128 + row;
        const int by = pb / kIrfIW, bx = pb - by * kIrfIW;
        const int iy = 2 * oy0 - 1 + by, ix = 2 * ox0 - 1 + bx;
        const bool ok = pb < kIrfPix && iy >= 0 && ix >= 0;
        const float4* src = reinterpret_cast<const float4*>(p.X + (((long long)b * p.H + iy) * p.W + ix) * kIrfCin);
#pragma unroll
        for (int j = 0; j < 4; ++j) xr[i][j] = ok ? __ldg(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto project_epilogue = [&](int tl) {
      int b, oy0, ox0;
      tile_coords(tl, b, oy0, ox0);
      mbar_wait(acc2_full, (uint32_t)(tl & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + kIrfColP + ((uint32_t)(q * 32) << 16);
      const int oy = oy0 + (row >> 4), ox = ox0 + (row & 15);
      float4* dst = reinterpret_cast<float4*>(p.Y + (((long long)b * Ho + oy) * Wo + ox) * kIrfCout);
      const float4* b2 = reinterpret_cast<const float4*>(smem + kIrfOffB2);
#pragma unroll
      for (int g = 0; g < kIrfCout; g += 8) {
        uint32_t m[8], s[8];
        tmem_ld_32x8(taddr + g, m);
        tmem_ld_32x8(taddr + 32 + g, s);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 bb = b2[g / 4 + h];
          float4 o;
          o.x = __uint_as_float(m[4 * h]) + __uint_as_float(s[4 * h]) + bb.x;
          o.y = __uint_as_float(m[4 * h + 1]) + __uint_as_float(s[4 * h + 1]) + bb.y;
          o.z = __uint_as_float(m[4 * h + 2]) + __uint_as_float(s[4 * h + 2]) + bb.z;
          o.w = __uint_as_float(m[4 * h + 3]) + __uint_as_float(s[4 * h + 3]) + bb.w;
          dst[g / 4 + h] = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
    };
    if (my_tiles > 0) prefetch(0);
    for (int tl = 0; tl < my_tiles; ++tl) {
      mbar_wait(a_empty, (uint32_t)((tl & 1) ^ 1));
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < kIrfMT; ++i) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v[4] = {xr[i][j].x, xr[i][j].y, xr[i][j].z, xr[i][j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t h = __float_as_uint(v[e]) & 0xFFFFE000u;  // = what kind::tf32 reads from the raw word
            hi[4 * j + e] = h;
            lo[4 * j + e] = __float_as_uint(v[e] - __uint_as_float(h));
          }
        }
        const uint32_t tdst = tmem_base + kIrfColA + i * 32 + ((uint32_t)(q * 32) << 16);
        tmem_st_32x16(tdst, hi);
        tmem_st_32x16(tdst + 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full);
      if (tl + 1 < my_tiles) prefetch(tl + 1);
      if (tl > 0) project_epilogue(tl - 1);
    }
    if (my_tiles > 0) project_epilogue(my_tiles - 1);
  } else if (warp >= 8) {
    // ===================================== worker groups =====================================
    const int group = (warp - 8) >> 2, gw = warp & 3;  // gw = TMEM lane quadrant of this warp
    uint8_t* slab = smem + kIrfOffSlab + group * kIrfSlabBytes;
    const int bar_id = 1 + group;
    const int total_slabs = my_tiles * kIrfSlabs;
    const int cg = lane & 7;
    for (int sc = group; sc < total_slabs; sc += 2) {
      const int tl = sc / kIrfSlabs, c = sc - tl * kIrfSlabs;
      int b, oy0, ox0;
      tile_coords(tl, b, oy0, ox0);
      // ---- (1) expand epilogue: TMEM -> +b1, ReLU, zero outside the image -> slab ----
      const float4* b1 = reinterpret_cast<const float4*>(smem + kIrfOffB1 + c * 128);
#pragma unroll 1
      for (int i = 0; i < kIrfMT; ++i) {
        const int u = sc * kIrfMT + i;
        const int slot = u % kIrfRing;
        mbar_wait(&acc_full[slot], (uint32_t)((u / kIrfRing) & 1));
        tc_fence_after();
        const int pb = i * 128 + gw * 32 + lane;
        if (i * 128 + gw * 32 < kIrfPix) {  // warp-uniform: this warp's 32 rows hold at least one real pixel
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + kIrfColAcc + slot * 32 + ((uint32_t)(gw * 32) << 16), r);
          tmem_ld_wait();
          if (pb < kIrfPix) {
            const int by = pb / kIrfIW, bx = pb - by * kIrfIW;
            const bool inside = (2 * oy0 - 1 + by) >= 0 && (2 * ox0 - 1 + bx) >= 0;
            float4* dst = reinterpret_cast<float4*>(slab + pb * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bb = b1[j];
              float4 o;
              o.x = inside ? fmaxf(__uint_as_float(r[4 * j]) + bb.x, 0.f) : 0.f;
              o.y = inside ? fmaxf(__uint_as_float(r[4 * j + 1]) + bb.y, 0.f) : 0.f;
              o.z = inside ? fmaxf(__uint_as_float(r[4 * j + 2]) + bb.z, 0.f) : 0.f;
              o.w = inside ? fmaxf(__uint_as_float(r[4 * j + 3]) + bb.w, 0.f) : 0.f;
              dst[j ^ (pb & 7)] = o;  // 16-byte chunk swizzle: conflict-free stores here and loads below
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // the slab is complete
      // ---- (2) depthwise 3x3 stride 2 + bd + ReLU -> (hi, lo) A tile of the project GEMM ----
      mbar_wait(a2_empty, (uint32_t)((sc & 1) ^ 1));
      {
        const F4* w4 = reinterpret_cast<const F4*>(smem + kIrfOffDw) + c * 8 + cg;  // tap t at + t * 24
        const F4 bias4 = reinterpret_cast<const F4*>(smem + kIrfOffBd)[c * 8 + cg];
        F4 wk[3][3];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) wk[ky][kx] = w4[(ky * 3 + kx) * (kIrfMid / 4)];
        uint8_t* ah = smem + kIrfOffA2;
        uint8_t* al = ah + 16384;
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
          // 32 blocks of 2 x 2 output pixels: block id = it * 16 + gw * 4 + (lane >> 3)
          const int blk = it * 16 + gw * 4 + (lane >> 3);
          const int oy_l = (blk >> 3) * 2, ox_l = (blk & 7) * 2;
          F4 acc[2][2];
#pragma unroll
          for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int x = 0; x < 2; ++x) acc[y][x] = bias4;
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            F4 v[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const int pb = (2 * oy_l + r) * kIrfIW + 2 * ox_l + i;
              v[i] = *reinterpret_cast<const F4*>(slab + pb * 128 + ((cg ^ (pb & 7)) << 4));
            }
#pragma unroll
            for (int y = 0; y < 2; ++y) {
              const int ky = r - 2 * y;
              if (ky >= 0 && ky < 3) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  const F4 k = wk[(ky >= 0 && ky < 3) ? ky : 0][kx];
#pragma unroll
                  for (int x = 0; x < 2; ++x) ffma2(acc[y][x].lo, v[2 * x + kx].lo, k.lo);
#pragma unroll
                  for (int x = 0; x < 2; ++x) ffma2(acc[y][x].hi, v[2 * x + kx].hi, k.hi);
                }
              }
            }
          }
#pragma unroll
          for (int y = 0; y < 2; ++y)
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              float4 v = f4_to_float4(acc[y][x]);
              v.x = fmaxf(v.x, 0.f);
              v.y = fmaxf(v.y, 0.f);
              v.z = fmaxf(v.z, 0.f);
              v.w = fmaxf(v.w, 0.f);
              float4 h, l;
              split_tf32_trunc(v.x, h.x, l.x);
              split_tf32_trunc(v.y, h.y, l.y);
              split_tf32_trunc(v.z, h.z, l.z);
              split_tf32_trunc(v.w, h.w, l.w);
              const int R = (oy_l + y) * kIrfTW + ox_l + x;           // A-tile row = pixel inside the 8 x 16 tile
              const int off = R * 128 + ((cg ^ (R & 7)) << 4);       // SWIZZLE_128B
              *reinterpret_cast<float4*>(ah + off) = v;              // raw fp32 = hi operand (hardware truncation)
              *reinterpret_cast<float4*>(al + off) = l;
            }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a2_full);
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // everyone has left the slab before it is rewritten
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kIrfTmemCols);
  }
}

// X [B][H][W][16] -> Y [B][H/2][W/2][24].  Returns 0 on launch, 1 when the shape is not covered, < 0 on error.
inline int launch_irf_s2(cudaStream_t s, const float* X, float* Y, const float* image, int B, int H, int W) {
  if (!available()) return 1;
  const int Ho = H / 2, Wo = W / 2;
  if (H % 2 || W % 2 || Ho % kIrfTH || Wo % kIrfTW) return 1;
  if (attr_needed(reinterpret_cast<const void*>(irf_s2_fused_kernel))) {
    if (cudaFuncSetAttribute(irf_s2_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kIrfSmemBytes) != cudaSuccess)
      return -30;
  }
  IrfParams p;
  p.X = X;
  p.Y = Y;
  p.image = image;
  p.B = B;
  p.H = H;
  p.W = W;
  p.tiles_x = Wo / kIrfTW;
  p.tiles_y = Ho / kIrfTH;
  p.num_tiles = B * p.tiles_x * p.tiles_y;
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  if (launch_pdl(irf_s2_fused_kernel, dim3(grid), dim3(kIrfThreads), (size_t)kIrfSmemBytes, s, p) != cudaSuccess) return -31;
  return 0;
}

}  // namespace tc
}  // namespace fear
