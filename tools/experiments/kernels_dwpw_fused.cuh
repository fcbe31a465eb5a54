// EXPERIMENT (round 2, not part of the library build): frame-level depthwise + 1x1 kernel with 16-channel chunks,
// separate box / A rings and [hi | lo] operands sharing a 128-byte row.  Bit-identical to the two-kernel path on B200,
// but SLOWER than pw_tc_kernel<DWK> (head SepConvs 79 us vs 65 us): the 64-byte rows of its TMA boxes (324 row requests
// per box) cost more than the deeper ring hides.  Kept as a record; see DESIGN.md section 5.

// dwpw_frame_kernel<DWK> -- depthwise DWK x DWK (stride 1, 16 x 16 maps) + 1x1 conv as ONE tcgen05 kernel (sm_100a):
//
//     out = act( dw(X) * W^T + bias (+ R) ),   dw(X) = [relu](depthwise(X) [+ bd])   never written to memory
//
// Covers the project half of the seven 16x16-stage IRF blocks of fbnet_c (dw 5x5 + BN + ReLU -> pwl 1x1 (+x);
// oracle/fbnet_c.py:95-110) and the ten SepConvs of the head (dw 3x3 -> 1x1 + BN + ReLU; reference
// model_training/model/blocks.py:45-72).  Unfused, the depthwise map of every such layer is written and re-read
// (2 x 0.18..0.69 MB per frame and layer).
//
// Second generation of pw_tc_kernel<DWK> (kernels_tc.cuh), restructured around what limited that one:
//   * K is walked in 16-CHANNEL chunks (not 32).  A chunk's depthwise input box (whole 16 x 16 frame + halo, 64 B per
//     pixel), its [W_hi | W_lo] tile and the (hi | lo) A tiles are half the size, so the TMA ring is 3 deep with the box
//     and the A operand in SEPARATE rings: the first version had 2 stages, each owned by one depthwise group, and its
//     chunk period was the exposed TMA latency (~3000 cycles) instead of the tensor-pipe / L2 time.
//   * One CTA tile = a whole FRAME (M = 256 = two M-tiles) x one N tile: the weight tile and the halo are fetched
//     once per 256 pixels instead of once per 128 -- these layers are bound by L2 -> SM bytes per output, not by HBM.
//   * hi and lo halves of an operand share a 128-byte row: row = [hi 16 ch | lo 16 ch] in the SWIZZLE_128B K-major
//     layout, so "the lo operand" is the same tile addressed 64 bytes further (the W side is pre-interleaved on the
//     host: fear_pack_weights builds [N][K/16][hi 16 | lo 16]).
//   * all 8 depthwise warps work on every chunk: thread = (channel pair, one column, the 8 rows of one M-tile); 16
//     consecutive lanes read two adjacent pixels = 128 contiguous bytes (conflict free without any swizzle), the K x K
//     weights of the pair stay in registers; FMA order (bias, ky, kx ascending) as in the stand-alone depthwise kernels.
// MMA order per K-step (hi*hi -> main; hi*lo, lo*hi -> correction accumulator) and the epilogue additions replicate
// pw_tc_kernel, so the layer is bit-identical to the two-kernel path.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-9 depthwise, warps 10-17 epilogue
// (warp -> (TMEM lane quadrant, M-tile)).  TMEM: 2 M-tiles x (main NT | correction NT) columns, single-buffered.
#pragma once
#include "tc_common.cuh"

namespace fear {
namespace tc {

struct DwpwParams {
  const float* bias;  // [N] or null
  const float* R;     // residual [M][ldr] or null
  float* C;
  int ldr, ldc;
  int frames, N, NT, num_n_tiles, num_chunks, relu;
  int dw_relu, dw_bias;
  int stages, stage_bytes, box_bytes, w_bytes, tmem_cols;
};

constexpr int kDpThreads = 576;
constexpr int kDpATile = 128 * 128;            // one M-tile of the A operand: [128 px][hi 16 | lo 16] fp32
constexpr int kDpASlot = 2 * kDpATile;         // both M-tiles of a frame
constexpr int kDpTail = 1024 /*barriers*/ + 8 * 2048 /*epilogue staging*/ + 1024 /*bias*/;
constexpr int kDpMaxSmem = 232448 - 1024;

template <int DWK>
__global__ void __launch_bounds__(kDpThreads, 1)
dwpw_frame_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmDW, const __grid_constant__ CUtensorMap tmDB, const DwpwParams p) {
  constexpr int K = DWK, P = K / 2, IW = 16 + K - 1, IH = 16 + K - 1;
  extern __shared__ uint8_t dp_smem_raw[];
  uint8_t* smem = dp_smem_raw + ((1024u - (smem_u32(dp_smem_raw) & 1023u)) & 1023u);
  const int S = p.stages;
  // layout: [A slot 0][A slot 1][stage 0 .. S-1: W tile | box | dw weights | dw bias][barriers][epilogue staging][bias]
  uint8_t* a_ring = smem;
  uint8_t* ring = smem + 2 * kDpASlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + S * p.stage_bytes);
  uint64_t* full = bars;             // [S] TMA landed
  uint64_t* empty = bars + 4;        // [S] box read by the 8 depthwise warps + W read by the MMAs (count 9)
  uint64_t* a_full = bars + 8;       // [2] A slot written (8 warps)
  uint64_t* a_empty = bars + 10;     // [2] MMAs have read it (commit)
  uint64_t* acc_full = bars + 12;    // accumulators complete (commit)
  uint64_t* acc_empty = bars + 13;   // drained by the 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(bars) + 1024;       // 8 warps x 2 KB
  float* sbias = reinterpret_cast<float*>(epi_stage + 8 * 2048);      // [NT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.frames * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    prefetch_tmap(&tmDW);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  if (threadIdx.x == 64) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 9);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&a_full[a], 8);
      mbar_init(&a_empty[a], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  auto st_w = [&](int s) { return ring + s * p.stage_bytes; };
  auto st_box = [&](int s) { return ring + s * p.stage_bytes + p.w_bytes; };
  auto st_dww = [&](int s) { return st_box(s) + p.box_bytes; };
  auto st_dwb = [&](int s) { return st_dww(s) + ((K * K * 64 + 127) & ~127); };  // (TMA destinations: 128-byte aligned)

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t bytes = (uint32_t)(p.w_bytes + p.box_bytes + K * K * 64 + (p.dw_bias ? 64 : 0));
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int f = t / p.num_n_tiles, nt = t - f * p.num_n_tiles;
        for (int c = 0; c < p.num_chunks; ++c) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], bytes);
          tma_load_4d(st_box(stage), &tmX, &full[stage], c * 16, -P, -P, f);       // whole frame + zero-filled halo
          tma_load_2d(st_dww(stage), &tmDW, &full[stage], c * 16, 0);
          if (p.dw_bias) tma_load_2d(st_dwb(stage), &tmDB, &full[stage], c * 16, 0);
          tma_load_2d(st_w(stage), &tmW, &full[stage], c * 32, nt * p.NT);         // [NT][hi 16 | lo 16]
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, p.NT);
      int stage = 0, q = 0;
      uint32_t phase = 0, tile_par = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(acc_empty, tile_par ^ 1);
        tc_fence_after();
        for (int c = 0; c < p.num_chunks; ++c, ++q) {
          const int slot = q & 1;
          mbar_wait(&full[stage], phase);            // W tile landed (the box landed with it)
          mbar_wait(&a_full[slot], (uint32_t)((q >> 1) & 1));
          tc_fence_after();
          const uint32_t w = smem_u32(st_w(stage));
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const uint32_t a = smem_u32(a_ring + slot * kDpASlot + m * kDpATile);
            const uint32_t d = tmem_base + m * 2 * p.NT;  // main; + NT = correction accumulator
#pragma unroll
            for (int j = 0; j < 2; ++j) {  // 16 channels = 2 K-steps; product order as in pw_tc_kernel (split accumulators)
              const uint64_t dah = umma_desc_k_sw128(a + j * 32), dal = umma_desc_k_sw128(a + 64 + j * 32);
              const uint64_t dwh = umma_desc_k_sw128(w + j * 32), dwl = umma_desc_k_sw128(w + 64 + j * 32);
              mma_tf32_ss(d, dah, dwh, idesc, (c | j) != 0);         // main += a_hi w_hi
              mma_tf32_ss(d + p.NT, dah, dwl, idesc, (c | j) != 0);  // corr += a_hi w_lo
              mma_tf32_ss(d + p.NT, dal, dwh, idesc, 1);             // corr += a_lo w_hi
            }
          }
          tc_commit(&a_empty[slot]);
          tc_commit(&empty[stage]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
        tc_commit(acc_full);
        tile_par ^= 1;
      }
    }
  } else if (warp < 10) {
    // ===================================== depthwise =====================================
    // thread = (channel PAIR cp of the 16-channel chunk, column x, M-tile mt): the 8 vertically adjacent output pixels
    // of one column of one half-frame, two channels (one packed FFMA2 lane pair).  16 consecutive lanes read 128
    // contiguous bytes (two adjacent pixels), so every LDS.64 wavefront is conflict free; the K x K weights of the
    // pair stay in registers (2 * K * K) -- with four channels per thread the 5x5 case would not fit.
    const int dwarp = warp - 2;
    const int cp = lane & 7, x = 4 * (dwarp & 3) + (lane >> 3), mt = dwarp >> 2;
    int stage = 0, q = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      for (int c = 0; c < p.num_chunks; ++c, ++q) {
        const int slot = q & 1;
        mbar_wait(&full[stage], phase);
        const unsigned long long* box =
            reinterpret_cast<const unsigned long long*>(st_box(stage)) + ((8 * mt) * IW + x) * 8 + cp;
        const unsigned long long* w2 = reinterpret_cast<const unsigned long long*>(st_dww(stage)) + cp;
        unsigned long long wk[K][K];
#pragma unroll
        for (int ky = 0; ky < K; ++ky)
#pragma unroll
          for (int kx = 0; kx < K; ++kx) wk[ky][kx] = w2[(ky * K + kx) * 8];
        unsigned long long acc[8];
        const unsigned long long bias2 = p.dw_bias ? reinterpret_cast<const unsigned long long*>(st_dwb(stage))[cp] : 0ull;
#pragma unroll
        for (int y = 0; y < 8; ++y) acc[y] = bias2;
#pragma unroll
        for (int r = 0; r < 8 + K - 1; ++r) {
          unsigned long long v[K];
#pragma unroll
          for (int kx = 0; kx < K; ++kx) v[kx] = box[(r * IW + kx) * 8];
#pragma unroll
          for (int y = 0; y < 8; ++y) {
            const int ky = r - y;
            if (ky >= 0 && ky < K) {
#pragma unroll
              for (int kx = 0; kx < K; ++kx) ffma2(acc[y], v[kx], wk[(ky >= 0 && ky < K) ? ky : 0][kx]);
            }
          }
        }
        mbar_wait(&a_empty[slot], (uint32_t)(((q >> 1) & 1) ^ 1));  // the MMAs of chunk q - 2 have read this slot
        uint8_t* at = a_ring + slot * kDpASlot + mt * kDpATile;
        const int hi_first = (x & 1) == 0;  // odd columns store lo first: the two pixels of a wavefront then hit different banks
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          float v0 = __uint_as_float((uint32_t)acc[y]), v1 = __uint_as_float((uint32_t)(acc[y] >> 32));
          if (p.dw_relu) {
            v0 = fmaxf(v0, 0.f);
            v1 = fmaxf(v1, 0.f);
          }
          float h0, l0, h1, l1;
          split_tf32_trunc(v0, h0, l0);
          split_tf32_trunc(v1, h1, l1);
          const int R = y * 16 + x;  // row inside the M-tile
          uint8_t* row = at + R * 128 + (cp & 1) * 8;
          const int ch_hi = ((cp >> 1) ^ (R & 7)) << 4, ch_lo = ((4 + (cp >> 1)) ^ (R & 7)) << 4;
          const float2 hv = make_float2(v0, v1), lv = make_float2(l0, l1);  // raw fp32 = hi operand (hardware truncation)
          *reinterpret_cast<float2*>(row + (hi_first ? ch_hi : ch_lo)) = hi_first ? hv : lv;
          *reinterpret_cast<float2*>(row + (hi_first ? ch_lo : ch_hi)) = hi_first ? lv : hv;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_full[slot]);
          mbar_arrive(&empty[stage]);  // this warp is done with the box
        }
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const int ew = warp - 10;
    const int q = warp & 3;     // TMEM lane quadrant this warp may access
    const int m = ew >> 2;      // M-tile of the frame handled by this warp
    const int etid = threadIdx.x - 320;
    uint32_t tile_par = 0;
    float4* stg = reinterpret_cast<float4*>(epi_stage + ew * 2048);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int f = t / p.num_n_tiles, nt = t - f * p.num_n_tiles;
      const int n0 = nt * p.NT;
      if (etid < p.NT) sbias[etid] = (p.bias && n0 + etid < p.N) ? __ldg(p.bias + n0 + etid) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(acc_full, tile_par);
      tc_fence_after();
      const uint32_t taddr = tmem_base + m * 2 * p.NT + ((uint32_t)(q * 32) << 16);
      const long long row0 = (long long)f * 256 + m * 128 + q * 32;
      for (int g = 0; g < p.NT; g += 16) {
        uint32_t r[16], rs[16];
        tmem_ld_32x16(taddr + g, r);
        tmem_ld_32x16(taddr + p.NT + g, rs);
        tmem_ld_wait();
        // per-warp 32 x 16 transpose through smem so that loads / stores touch 64 contiguous bytes per row
#pragma unroll
        for (int j = 0; j < 4; ++j)
          stg[lane * 4 + (j ^ ((lane >> 1) & 3))] =
              make_float4(__uint_as_float(r[4 * j]) + __uint_as_float(rs[4 * j]),
                          __uint_as_float(r[4 * j + 1]) + __uint_as_float(rs[4 * j + 1]),
                          __uint_as_float(r[4 * j + 2]) + __uint_as_float(rs[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]) + __uint_as_float(rs[4 * j + 3]));
        __syncwarp();
        const int j = lane & 3;
        const int col = n0 + g + j * 4;
        if (col < p.N) {
          const float4 b = *reinterpret_cast<const float4*>(sbias + g + 4 * j);
#pragma unroll
          for (int rb = 0; rb < 4; ++rb) {
            const int rl = rb * 8 + (lane >> 2);
            const long long grow = row0 + rl;
            float4 o = stg[rl * 4 + (j ^ ((rl >> 1) & 3))];
            if (p.R) {  // same association as pw_tc_kernel's residual path: (main + corr) + (bias + residual)
              const float4 rr = __ldg(reinterpret_cast<const float4*>(p.R + grow * p.ldr + col));
              o.x += b.x + rr.x;
              o.y += b.y + rr.y;
              o.z += b.z + rr.z;
              o.w += b.w + rr.w;
            } else {
              o.x += b.x;
              o.y += b.y;
              o.z += b.z;
              o.w += b.w;
            }
            if (p.relu) {
              o.x = fmaxf(o.x, 0.f);
              o.y = fmaxf(o.y, 0.f);
              o.z = fmaxf(o.z, 0.f);
              o.w = fmaxf(o.w, 0.f);
            }
            *reinterpret_cast<float4*>(p.C + grow * p.ldc + col) = o;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      tile_par ^= 1;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // sbias may be rewritten for the next tile
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Host: [N][K] hi / lo weight copies -> [N][K/16][hi 16 | lo 16] (the B operand layout of dwpw_frame_kernel).
inline void dwpw_interleave_weights(float* dst, const float* w_hi, const float* w_lo, int N, int K) {
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < K / 16; ++c)
      for (int e = 0; e < 16; ++e) {
        dst[((size_t)n * (K / 16) + c) * 32 + e] = w_hi[(size_t)n * K + c * 16 + e];
        dst[((size_t)n * (K / 16) + c) * 32 + 16 + e] = w_lo[(size_t)n * K + c * 16 + e];
      }
}

// X = [B][16][16][K] channels-last; out = act(dw_k(X) * W^T + bias (+R)), W given interleaved (see above).
// Returns 0 on launch, 1 when the shape is not covered (caller runs the two kernels separately), < 0 on error.
inline int launch_dwpw_frame(cudaStream_t s, const float* X, int B, int dw_k, const float* dw_w, const float* dw_b, int dw_relu,
                             const float* w_il, const float* bias, const float* R, int ldr, float* C, int ldc, int N, int K,
                             int relu) {
  if (!available()) return -20;
  if ((dw_k != 3 && dw_k != 5) || K % 16 || !w_il) return 1;
  DwpwParams p;
  p.bias = bias;
  p.R = R;
  p.C = C;
  p.ldr = ldr;
  p.ldc = ldc;
  p.frames = B;
  p.N = N;
  const int Np = (N + 15) & ~15;
  p.NT = Np <= 128 ? Np : (Np % 128 == 0 ? 128 : 0);
  if (!p.NT) return 1;
  p.num_n_tiles = Np / p.NT;
  p.num_chunks = K / 16;
  p.relu = relu;
  p.dw_relu = dw_relu;
  p.dw_bias = dw_b != nullptr;
  const int iw = 16 + dw_k - 1;
  p.box_bytes = iw * iw * 64;
  p.w_bytes = (p.NT * 128 + 1023) & ~1023;
  p.stage_bytes = (p.w_bytes + p.box_bytes + ((dw_k * dw_k * 64 + 127) & ~127) + 128 + 1023) & ~1023;
  p.stages = (kDpMaxSmem - 1024 - kDpTail - 2 * kDpASlot) / p.stage_bytes;
  if (p.stages > 4) p.stages = 4;
  if (p.stages < 2) return 1;
  int cols = 32;
  while (cols < 4 * p.NT) cols <<= 1;
  if (cols > 512) return 1;
  p.tmem_cols = cols;
  CUtensorMap tmX, tmW, tmDW, tmDB;
  int r = make_tmap_nhwc(&tmX, X, (uint64_t)B, 16, 16, (uint64_t)K, 16, iw, iw);
  if (r) return r;
  r = make_tmap_2d(&tmW, w_il, (uint64_t)N, (uint64_t)2 * K, (uint64_t)2 * K, p.NT, 32);
  if (r) return r;
  r = make_tmap_2d_plain(&tmDW, dw_w, (uint64_t)dw_k * dw_k, (uint64_t)K, dw_k * dw_k, 16);
  if (r) return r;
  if (dw_b) {
    r = make_tmap_2d_plain(&tmDB, dw_b, 1, (uint64_t)K, 1, 16);
    if (r) return r;
  } else {
    tmDB = tmDW;
  }
  const int tiles = B * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  const size_t smem_bytes = (size_t)2 * kDpASlot + (size_t)p.stages * p.stage_bytes + 1024 + kDpTail;
  auto k3 = dwpw_frame_kernel<3>;
  auto k5 = dwpw_frame_kernel<5>;
  if (attr_needed(reinterpret_cast<const void*>(k3))) {
    if (cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, kDpMaxSmem) != cudaSuccess ||
        cudaFuncSetAttribute(k5, cudaFuncAttributeMaxDynamicSharedMemorySize, kDpMaxSmem) != cudaSuccess)
      return -30;
  }
  cudaError_t e = dw_k == 5 ? launch_pdl(k5, dim3(grid), dim3(kDpThreads), smem_bytes, s, tmX, tmW, tmDW, tmDB, p)
                            : launch_pdl(k3, dim3(grid), dim3(kDpThreads), smem_bytes, s, tmX, tmW, tmDW, tmDB, p);
  return e == cudaSuccess ? 0 : -23;
}

}  // namespace tc
}  // namespace fear
