// libfear_b200.so -- executor + C ABI of the FEAR-XS hot path on B200 (sm_100a).
// See include/fear_b200.h for the contract and DESIGN.md for the data layout.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fear_b200.h"
#include "arch.h"
#include "kernels_ffma.cuh"
#include "kernels_tc.cuh"
#include "kernels_dw_tma.cuh"
#include "kernels_stem_fused.cuh"

using namespace fear;

// ------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";

static int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return set_err((int)_e, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define FEAR_TRY(expr)       \
  do {                       \
    int _r = (expr);         \
    if (_r != 0) return _r;  \
  } while (0)

// ------------------------------------------------------------------------------ stages
enum Stage {
  ST_STEM = 0,
  ST_BACKBONE_PW,
  ST_BACKBONE_DW,
  ST_NECK,
  ST_HEAD_DW,
  ST_HEAD_PW,
  ST_CORR,
  ST_PRED,
  ST_DECODE,
  ST_LAYOUT,
  ST_COUNT
};
static const char* kStageNames[ST_COUNT] = {"stem",    "backbone_pw", "backbone_dw", "neck",   "head_dw",
                                            "head_pw", "corr",        "pred",        "decode", "layout"};

enum Impl { IMPL_FFMA = 0, IMPL_TC = 1, IMPL_TS = 2, IMPL_TC2 = 3 };  // CUDA cores | tcgen05 A-from-smem | tcgen05 A-from-TMEM

struct Options {
  int corr = -1;  // -1 = auto: tcgen05 (A from smem) when the tensor-core path initialised, else CUDA cores
  int pw = -1;
  int dw_wide = 0;    // 1: 16-wide strips for 5x5 stride-1 depthwise
  int fuse_dwpw = 0;  // EXPERIMENTAL bit mask: 1 = 16x16-stage blocks, 2 = head SepConvs run depthwise + 1x1 as one tcgen05
                      // kernel (pw_tc_kernel<DWK>)
  int small_const = 1;  // 1: tiny 1x1 layers take their weights by value (constant bank) instead of via shared memory
  int fuse_stem = 1;  // 1: stem + xif1_0 in one kernel (stem_xif1_fused_kernel) when the map tiles by 16x32
  int fuse = 0;       // 1 = fused pw-expand + depthwise kernels for the stride-2 blocks (FFMA-bound: measured slower than the tcgen05 GEMM + strip dw pair)
  int early_sub = 0;  // > 0: run the high-resolution backbone blocks in sub-batches of this many frames
  int dw = 3;  // 3 = auto (default); 0 = one pixel per thread, 1 = register-strip kernel, 2 = rolling-window kernel,
               // 4 = smem tile, 5 = L1-blocked strip, 6 = TMA pipeline only where it applies (auto also uses it)
};
static Options g_default_options;
static inline int effective(int impl) { return impl >= 0 ? impl : (tc::available() ? IMPL_TC : IMPL_FFMA); }

struct PwW {
  const float* w = nullptr;  // [cout][cin]
  const float* w_hi = nullptr;  // tf32 split of w for the tcgen05 path: w ~= w_hi + w_lo
  const float* w_lo = nullptr;
  const float* b = nullptr;
  const float* h_w = nullptr;  // host copies (persistent): small layers pass their weights by value
  const float* h_b = nullptr;
  int cin = 0, cout = 0;
};
struct DwW {
  const float* w = nullptr;  // [k*k][c]
  const float* b = nullptr;
  int c = 0, k = 0;
};
struct BlockW {
  PwW pw, pwl;
  DwW dw;
};
struct BranchW {
  DwW enc_dw, corr_dw;
  PwW enc_pw, corr_pw;
};
struct TowerW {
  DwW dw[2];
  PwW pw[2];
};

struct EventPair {
  cudaEvent_t a, b;
  int stage;
};

struct FearContext {
  FsWeights fs;  // host copy of the stem + xif1_0 weights, passed by value to stem_xif1_fused_kernel
  int device = 0;
  Options opt;
  float* d_weights = nullptr;
  std::vector<float> h_weights;  // host mirror of d_weights (device layout)
  const float *stem_w = nullptr, *stem_b = nullptr;
  BlockW blocks[kNumBlocks];
  PwW neck;
  BranchW branch[2];  // 0 = cls, 1 = reg
  TowerW tower[2];    // 0 = bbox, 1 = cls
  DwW pred_dw[2];     // 0 = bbox, 1 = cls
  const float *pred_w[2] = {nullptr, nullptr}, *pred_b[2] = {nullptr, nullptr};

  int reserved = 0;
  float* ws = nullptr;
  // backbone ping-pong (per frame sizes in floats)
  float *bufX = nullptr, *bufY = nullptr, *bufE = nullptr, *bufD = nullptr;
  float* bufS = nullptr;  // output of the early (sub-batched) blocks: 32 x 32 x 32 per frame
  // head
  float *hF = nullptr, *hT = nullptr, *hCAT[2] = {nullptr, nullptr}, *hD[2] = {nullptr, nullptr}, *hP = nullptr;
  float* hQ[2] = {nullptr, nullptr};  // tower outputs: [0] = bbox tower (x_reg), [1] = cls tower
  float *zt = nullptr, *mapB = nullptr, *mapC = nullptr;
  float *zth = nullptr, *ztl = nullptr;  // tf32 (hi, lo) split of zt for the TS correlation

  int64_t launches = 0;
  bool profiling = false;
  std::vector<EventPair> events;
  size_t events_used = 0;
  double stage_ms[ST_COUNT] = {0};
  int64_t stage_launches[ST_COUNT] = {0};
};

static constexpr int64_t kActX = 128 * 128 * 16;  // largest block input / output per frame (floats)
static constexpr int64_t kActE = 128 * 128 * 96;  // largest expanded tensor (xif2_0.pw)
static constexpr int64_t kActD = 64 * 64 * 96;    // largest depthwise output (xif2_0.dw)

// RAII bracket around one kernel launch: counts it and, when profiling, records events.
struct LaunchScope {
  FearContext* c;
  cudaStream_t s;
  EventPair* ev = nullptr;
  LaunchScope(FearContext* c_, int stage, cudaStream_t s_) : c(c_), s(s_) {
    if (!c) return;
    c->launches++;
    c->stage_launches[stage]++;
    if (c->profiling && c->events_used < c->events.size()) {
      ev = &c->events[c->events_used++];
      ev->stage = stage;
      cudaEventRecord(ev->a, s);
    }
  }
  ~LaunchScope() {
    if (ev) cudaEventRecord(ev->b, s);
  }
};

static int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err((int)e, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------ launchers
static int launch_gemm_ffma(FearContext* c, int stage, cudaStream_t s, const float* A, int lda, long long sA,
                            const float* Bw, int ldb, long long sB, const float* bias, const float* R, int ldr,
                            float* C, int ldc, long long sC, int M, int N, int K, int relu, int batch) {
  LaunchScope scope(c, stage, s);
#define GEMM_CASE(TN_)                                                                                    \
  {                                                                                                       \
    dim3 grid((M + 127) / 128, (N + 8 * TN_ - 1) / (8 * TN_), batch);                                     \
    gemm_nt_ffma_kernel<TN_><<<grid, 256, 0, s>>>(A, lda, sA, Bw, ldb, sB, bias, R, ldr, C, ldc, sC, M, N, K, \
                                                  relu);                                                  \
  }
  if (N % 64 == 0) GEMM_CASE(8)
  else if (N % 56 == 0) GEMM_CASE(7)
  else if (N % 48 == 0) GEMM_CASE(6)
  else if (N % 32 == 0) GEMM_CASE(4)
  else if (N % 24 == 0) GEMM_CASE(3)
  else GEMM_CASE(2)
#undef GEMM_CASE
  return check_launch("gemm_nt_ffma_kernel");
}

// 1x1 conv over M pixels: out = act(A * W^T + b (+R)).
static int launch_pw(FearContext* c, int stage, cudaStream_t s, const float* A, int lda, const PwW& w, const float* R,
                     int ldr, float* C, int ldc, int M, int relu) {
  const int pw_impl = effective(c->opt.pw);
  if (pw_impl != IMPL_FFMA && lda == w.cin && ldc == w.cout && (!R || ldr == w.cout) &&
      ((w.cin == 16 && w.cout == 16) || (w.cin == 24 && w.cout == 24))) {
    // streaming layers: one pixel per thread on CUDA cores beats a tensor-core tile pipeline here
    LaunchScope scope(c, stage, s);
    const unsigned blocks = (unsigned)((M + 255) / 256);
    if (c->opt.small_const && w.h_w && w.h_b) {
      // weights by value in the constant bank (see pw_small_const_kernel)
      if (w.cin == 16) {
        PwSmallWeights<16, 16> pw;
        for (int o = 0; o < 16; ++o)
          for (int k = 0; k < 16; ++k) pw.w[k * 16 + o] = w.h_w[o * 16 + k];
        memcpy(pw.b, w.h_b, sizeof(pw.b));
        pw_small_const_kernel<16, 16><<<blocks, 256, 0, s>>>(A, R, C, M, relu, pw);
      } else {
        PwSmallWeights<24, 24> pw;
        for (int o = 0; o < 24; ++o)
          for (int k = 0; k < 24; ++k) pw.w[k * 24 + o] = w.h_w[o * 24 + k];
        memcpy(pw.b, w.h_b, sizeof(pw.b));
        pw_small_const_kernel<24, 24><<<blocks, 256, 0, s>>>(A, R, C, M, relu, pw);
      }
      return check_launch("pw_small_const_kernel");
    }
    if (w.cin == 16)
      pw_small_kernel<16, 16><<<blocks, 256, 0, s>>>(A, w.w, w.b, R, C, M, relu);
    else
      pw_small_kernel<24, 24><<<blocks, 256, 0, s>>>(A, w.w, w.b, R, C, M, relu);
    return check_launch("pw_small_kernel");
  }
  if (pw_impl == IMPL_TS && tc::pw_supported(w.cin, w.cout) && tc::ts_tile_n(w.cout)) {
    LaunchScope scope(c, stage, s);
    int r = tc::launch_gemm_ts(s, A, lda, w.w_hi, w.w_lo, (uint64_t)w.cout, w.b, R, ldr, C, ldc, M, w.cout, w.cin, relu,
                               0, 1, 0);
    if (r) return set_err(r, "tcgen05 (TS) pw launch failed (%d)", r);
    return check_launch("tc::gemm_ts");
  }
  if (pw_impl == IMPL_TC && tc::pw_supported(w.cin, w.cout)) {
    LaunchScope scope(c, stage, s);
    int r = tc::launch_pw(s, A, lda, w.w_hi, w.w_lo, w.b, R, ldr, C, ldc, M, w.cout, w.cin, relu);
    if (r) return set_err(r, "tcgen05 pw launch failed (%d)", r);
    return check_launch("tc::pw");
  }
  return launch_gemm_ffma(c, stage, s, A, lda, 0, w.w, w.cin, 0, w.b, R, ldr, C, ldc, 0, M, w.cout, w.cin, relu, 1);
}

static int launch_dw(FearContext* c, int stage, cudaStream_t s, const float* in, const DwW& w, float* out, int B, int H,
                     int W, int stride, bool relu) {
  LaunchScope scope(c, stage, s);
  const int C4 = w.c / 4;
  const int threads = 256;
  const float4* i4 = reinterpret_cast<const float4*>(in);
  const float4* w4 = reinterpret_cast<const float4*>(w.w);
  const float4* b4 = reinterpret_cast<const float4*>(w.b);
  float4* o4 = reinterpret_cast<float4*>(out);
  const bool bias = w.b != nullptr;
  const int Wo = W / stride;
  // TMA-fed shared-memory pipeline (kernels_dw_tma.cuh): stride 1, maps that are multiples of 16x16
  const bool want_tma = (c->opt.dw == 6 || c->opt.dw == 3) && tc::available();
  if (want_tma && stride == 2 && w.k == 5 && relu && bias) {
    // 5x5 stride 2: 8x8 output tiles (19x19 input pixels), 4x1 outputs per thread
    int r = tc::launch_dw_tma_t<5, 2, 8, 8, 4, 1, 4, 2, true, true>(s, in, w.w, w.b, out, B, H, W, w.c, tc::g_num_sms);
    if (r < 0) return set_err(FEAR_EINVAL, "TMA depthwise launch failed (%d)", r);
    if (r == 0) return check_launch("tc::dw_tma_kernel<5,2>");
  }
  // TMA-fed shared-memory pipeline (kernels_dw_tma.cuh): stride 1, maps that are multiples of 16x16
  if (want_tma && stride == 1 && w.c >= 24) {
    int r = 1;
#define DW_TMA(K_, RELU_, BIAS_) \
  tc::launch_dw_tma_t<K_, 1, 16, 16, 8, 2, 4, 2, RELU_, BIAS_>(s, in, w.w, w.b, out, B, H, W, w.c, tc::g_num_sms)
    if (w.k == 5 && relu && bias) r = DW_TMA(5, true, true);
    else if (w.k == 3 && relu && bias) r = DW_TMA(3, true, true);
    else if (w.k == 3 && !relu && !bias) r = DW_TMA(3, false, false);
#undef DW_TMA
    if (r < 0) return set_err(FEAR_EINVAL, "TMA depthwise launch failed (%d)", r);
    if (r == 0) return check_launch("tc::dw_tma_kernel");
  }
  // shared-memory tiled kernel: stride 1, maps that are multiples of 16x16, channels in 32-slabs
  const bool want_tile = c->opt.dw == 4 && stride == 1 && H % 16 == 0 && W % 16 == 0 &&
                         C4 % 8 == 0 && ((relu && bias) || (!relu && !bias));
  if (want_tile) {
    const unsigned blocks = (unsigned)(B * (H / 16) * (W / 16) * (C4 / 8));
    static bool attr_done = false;
    if (!attr_done) {
      CUDA_TRY(cudaFuncSetAttribute(dw_conv_tile_kernel<5, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    dw_tile_smem_bytes<5>()));
      attr_done = true;
    }
    if (w.k == 5 && relu)
      dw_conv_tile_kernel<5, true, true><<<blocks, 256, dw_tile_smem_bytes<5>(), s>>>(i4, w4, b4, o4, H, W, C4);
    else if (w.k == 3 && relu)
      dw_conv_tile_kernel<3, true, true><<<blocks, 256, dw_tile_smem_bytes<3>(), s>>>(i4, w4, b4, o4, H, W, C4);
    else if (w.k == 3)
      dw_conv_tile_kernel<3, false, false><<<blocks, 256, dw_tile_smem_bytes<3>(), s>>>(i4, w4, b4, o4, H, W, C4);
    else return set_err(FEAR_EINVAL, "unsupported tiled depthwise config k=%d", w.k);
    return check_launch("dw_conv_tile_kernel");
  }
  const bool want_roll = c->opt.dw == 2 || (c->opt.dw == 3 && w.k == 3 && stride == 1);
  if (want_roll && Wo % 4 == 0 && (H / stride) % 16 == 0) {
    // rolling-window kernels: TX output columns x 16 output rows per thread, weights in registers
    constexpr int ROWS = 16;
    const int Ho = H / stride;
#define ROLL_CASE(K_, S_, TX_, RELU_, BIAS_)                                                        \
  {                                                                                                  \
    const long long total = (long long)B * (Ho / ROWS) * (Wo / TX_) * C4;                            \
    const unsigned blocks = (unsigned)((total + 127) / 128);                                         \
    dw_conv_roll_kernel<K_, S_, TX_, ROWS, RELU_, BIAS_><<<blocks, 128, 0, s>>>(i4, w4, b4, o4, B, H, W, C4); \
  }
    if (w.k == 3 && stride == 1 && relu && bias) ROLL_CASE(3, 1, 4, true, true)
    else if (w.k == 3 && stride == 2 && relu && bias) ROLL_CASE(3, 2, 4, true, true)
    else if (w.k == 5 && stride == 1 && relu && bias) ROLL_CASE(5, 1, 2, true, true)
    else if (w.k == 5 && stride == 2 && relu && bias) ROLL_CASE(5, 2, 2, true, true)
    else if (w.k == 3 && stride == 1 && !relu && !bias) ROLL_CASE(3, 1, 4, false, false)
    else
      return set_err(FEAR_EINVAL, "unsupported depthwise config k=%d s=%d relu=%d bias=%d", w.k, stride, (int)relu,
                     (int)bias);
#undef ROLL_CASE
    return check_launch("dw_conv_roll_kernel");
  }
  if (c->opt.dw == 5 && Wo % 4 == 0) {
    // L1-blocked strip layout: CTA = 32 channels x 4 strips x 8 rows
    const int TX = stride == 1 ? 4 : 2;
    const int strips = Wo / TX;
    const unsigned blocks = (unsigned)((long long)B * (((H / stride) + 7) / 8) * ((strips + 3) / 4) * ((C4 + 7) / 8));
    if (w.k == 3 && stride == 1 && relu && bias)
      dw_conv_strip_blocked_kernel<3, 1, 4, true, true><<<blocks, 256, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 3 && stride == 2 && relu && bias)
      dw_conv_strip_blocked_kernel<3, 2, 2, true, true><<<blocks, 256, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 5 && stride == 1 && relu && bias)
      dw_conv_strip_blocked_kernel<5, 1, 4, true, true><<<blocks, 256, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 5 && stride == 2 && relu && bias)
      dw_conv_strip_blocked_kernel<5, 2, 2, true, true><<<blocks, 256, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 3 && stride == 1 && !relu && !bias)
      dw_conv_strip_blocked_kernel<3, 1, 4, false, false><<<blocks, 256, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else
      return set_err(FEAR_EINVAL, "unsupported depthwise config k=%d s=%d relu=%d bias=%d", w.k, stride, (int)relu,
                     (int)bias);
    return check_launch("dw_conv_strip_blocked_kernel");
  }
  if (c->opt.dw == 3 && w.k == 5 && stride == 1 && Wo % 8 == 0 && relu && bias) {
    // 5x5 stride 1: wide strips (fewer loads per FMA: the kernel is bound by L1 wavefronts, not by HBM)
    if (Wo % 16 == 0 && c->opt.dw_wide) {
      const long long total = (long long)B * H * (Wo / 16) * C4;
      dw_conv_strip_kernel<5, 1, 16, true, true><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(i4, w4, b4, o4, B, H, W,
                                                                                               C4);
      return check_launch("dw_conv_strip_kernel<5,1,16>");
    }
    const long long total = (long long)B * H * (Wo / 8) * C4;
    dw_conv_strip_kernel<5, 1, 8, true, true><<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(
        i4, w4, b4, o4, B, H, W, C4);
    return check_launch("dw_conv_strip_kernel<5,1,8>");
  }
  if (c->opt.dw >= 1 && Wo % 4 == 0) {
    // register-strip kernels: 4 outputs per thread (stride 1) / 2 outputs per thread (stride 2)
    const int TX = stride == 1 ? 4 : 2;
    const long long total = (long long)B * (H / stride) * (Wo / TX) * C4;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if (w.k == 3 && stride == 1 && relu && bias)
      dw_conv_strip_kernel<3, 1, 4, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 3 && stride == 2 && relu && bias)
      dw_conv_strip_kernel<3, 2, 2, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 5 && stride == 1 && relu && bias)
      dw_conv_strip_kernel<5, 1, 4, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 5 && stride == 2 && relu && bias)
      dw_conv_strip_kernel<5, 2, 2, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else if (w.k == 3 && stride == 1 && !relu && !bias)
      dw_conv_strip_kernel<3, 1, 4, false, false><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
    else
      return set_err(FEAR_EINVAL, "unsupported depthwise config k=%d s=%d relu=%d bias=%d", w.k, stride, (int)relu,
                     (int)bias);
    return check_launch("dw_conv_strip_kernel");
  }
  const long long total = (long long)B * (H / stride) * Wo * C4;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (w.k == 3 && stride == 1 && relu && bias)
    dw_conv_nhwc_kernel<3, 1, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
  else if (w.k == 3 && stride == 2 && relu && bias)
    dw_conv_nhwc_kernel<3, 2, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
  else if (w.k == 5 && stride == 1 && relu && bias)
    dw_conv_nhwc_kernel<5, 1, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
  else if (w.k == 5 && stride == 2 && relu && bias)
    dw_conv_nhwc_kernel<5, 2, true, true><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
  else if (w.k == 3 && stride == 1 && !relu && !bias)
    dw_conv_nhwc_kernel<3, 1, false, false><<<blocks, threads, 0, s>>>(i4, w4, b4, o4, B, H, W, C4);
  else
    return set_err(FEAR_EINVAL, "unsupported depthwise config k=%d s=%d relu=%d bias=%d", w.k, stride, (int)relu,
                   (int)bias);
  return check_launch("dw_conv_nhwc_kernel");
}

// Fused pw-expand + depthwise for the stride-2 blocks; returns 1 if this block shape has no fused kernel.
template <int CIN, int MID, int MSL, int K, int TH, int TW>
static int launch_fused_one(FearContext* c, cudaStream_t s, const float* X, const BlockW& bw, float* D, int B, int H,
                            int W) {
  constexpr int THREADS = 512;
  constexpr int smem = fused_expand_dw_smem_bytes<CIN, MID, MSL, K, TH, TW>();
  static bool attr_done = false;
  if (!attr_done) {
    CUDA_TRY(cudaFuncSetAttribute(fused_expand_dw_s2_kernel<CIN, MID, MSL, K, TH, TW, THREADS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  const int Ho = H / 2, Wo = W / 2;
  if (Ho % TH || Wo % TW) return 1;
  LaunchScope scope(c, ST_BACKBONE_DW, s);
  const unsigned grid = (unsigned)(B * (Ho / TH) * (Wo / TW));
  fused_expand_dw_s2_kernel<CIN, MID, MSL, K, TH, TW, THREADS><<<grid, THREADS, smem, s>>>(X, bw.pw.w, bw.pw.b, bw.dw.w,
                                                                                   bw.dw.b, D, H, W);
  return check_launch("fused_expand_dw_s2_kernel");
}

static int launch_fused_expand_dw(FearContext* c, cudaStream_t s, const IrfSpec& sp, const float* X, const BlockW& bw,
                                  float* D, int B, int H, int W) {
  if (sp.stride != 2 || !sp.has_pw()) return 1;
  // (tile, resident channel slice) per block: keep the halo recompute factor low and >= 2 CTAs per SM where possible
  if (sp.cin == 16 && sp.mid() == 96 && sp.k == 3) return launch_fused_one<16, 96, 96, 3, 8, 4>(c, s, X, bw, D, B, H, W);
  if (sp.cin == 24 && sp.mid() == 144 && sp.k == 5) {
    if ((H / 2) % 8 == 0) return launch_fused_one<24, 144, 48, 5, 8, 8>(c, s, X, bw, D, B, H, W);
    return launch_fused_one<24, 144, 48, 5, 4, 4>(c, s, X, bw, D, B, H, W);
  }
  if (sp.cin == 32 && sp.mid() == 192 && sp.k == 5) {
    if ((H / 2) % 8 == 0) return launch_fused_one<32, 192, 32, 5, 8, 8>(c, s, X, bw, D, B, H, W);
    return launch_fused_one<32, 192, 32, 5, 4, 4>(c, s, X, bw, D, B, H, W);
  }
  return 1;
}

static int launch_transpose(FearContext* c, cudaStream_t s, const float* in, int ldin, long long sIn, float* out,
                            int ldout, long long sOut, int R, int Cn, int batch) {
  LaunchScope scope(c, ST_LAYOUT, s);
  dim3 grid((Cn + 31) / 32, (R + 31) / 32, batch);
  transpose_kernel<<<grid, 256, 0, s>>>(in, ldin, sIn, out, ldout, sOut, R, Cn);
  return check_launch("transpose_kernel");
}

// cat[b, p, 256 + k] = sum_c zt[b, k, c] * cat[b, p, c]    (MobileCorrelation matmul, blocks.py:123)
// `groups` consecutive [B][256][320] buffers starting at cat share the templates (head: cls + reg branch).
static int launch_corr(FearContext* c, const Options& opt, cudaStream_t s, const float* zt, int Bz, float* cat, int B,
                       int groups, float* zth = nullptr, float* ztl = nullptr) {
  const int corr_impl = effective(opt.corr);
  if (corr_impl == IMPL_TS && zth && ztl) {
    {
      LaunchScope scope(c, ST_LAYOUT, s);  // template features -> tf32 (hi, lo) planes
      const long long n4 = (long long)Bz * kCorrC * kFeatC / 4;
      tc::split_hi_lo_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(
          reinterpret_cast<const float4*>(zt), reinterpret_cast<float4*>(zth), reinterpret_cast<float4*>(ztl), n4);
      FEAR_TRY(check_launch("split_hi_lo_kernel"));
    }
    LaunchScope scope(c, ST_CORR, s);
    int r = tc::launch_gemm_ts(s, cat, kCatC, zth, ztl, (uint64_t)Bz * kCorrC, nullptr, nullptr, 0, cat + kFeatC, kCatC,
                               B * groups * kScorePix, kCorrC, kFeatC, 0, 2, Bz == 1 ? 1 : B, kCorrC);
    if (r) return set_err(r, "tcgen05 (TS) corr launch failed (%d)", r);
    return check_launch("tc::gemm_ts(corr)");
  }
  if (corr_impl == IMPL_TC2 && zth && ztl) {
    {
      LaunchScope scope(c, ST_LAYOUT, s);  // template features -> tf32 (hi, lo) planes
      const long long n4 = (long long)Bz * kCorrC * kFeatC / 4;
      tc::split_hi_lo_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(
          reinterpret_cast<const float4*>(zt), reinterpret_cast<float4*>(zth), reinterpret_cast<float4*>(ztl), n4);
      FEAR_TRY(check_launch("split_hi_lo_kernel"));
    }
    LaunchScope scope(c, ST_CORR, s);
    int r = tc::launch_corr2(s, zth, ztl, Bz, cat, B, groups);
    if (r) return set_err(r, "tcgen05 (v2) corr launch failed (%d)", r);
    return check_launch("tc::corr2");
  }
  if (corr_impl == IMPL_TC || corr_impl == IMPL_TS || corr_impl == IMPL_TC2) {
    LaunchScope scope(c, ST_CORR, s);
    int r = tc::launch_corr(s, zt, Bz, cat, B, groups);
    if (r) return set_err(r, "tcgen05 corr launch failed (%d)", r);
    return check_launch("tc::corr");
  }
  for (int g = 0; g < groups; ++g) {
    float* cg = cat + (long long)g * B * kScorePix * kCatC;
    FEAR_TRY(launch_gemm_ffma(c, ST_CORR, s, cg, kCatC, (long long)kScorePix * kCatC, zt, kFeatC,
                              Bz == 1 ? 0 : (long long)kCorrC * kFeatC, nullptr, nullptr, 0, cg + kFeatC, kCatC,
                              (long long)kScorePix * kCatC, kScorePix, kCorrC, kFeatC, 0, B));
  }
  return 0;
}

// ------------------------------------------------------------------------------ executor
// img (B,3,H,W) NCHW -> NHWC backbone features [B][H/16 * W/16][112] left in *feat (a workspace buffer)
// Run backbone blocks [first, last) on NHWC activations X (B frames of h x w).  The last block's output goes to
// `final_out` when given (else into a ping-pong buffer); *out receives the output pointer, h/w are updated.
static int run_blocks(FearContext* c, cudaStream_t s, float* X, int B, int& h, int& w, int first, int last,
                      float* final_out, float** out) {
  float* Y = (X == c->bufX) ? c->bufY : c->bufX;
  for (int i = first; i < last; ++i) {
    const IrfSpec& sp = kBlocks[i];
    const BlockW& bw = c->blocks[i];
    const int M = B * h * w;
    int fused = 1;
    if (c->opt.fuse) {
      fused = launch_fused_expand_dw(c, s, sp, X, bw, c->bufD, B, h, w);
      if (fused < 0 || fused > 1) return fused;
    }
    if (fused == 1) {  // unfused: materialise the expanded tensor, then the depthwise conv
      const float* E = X;
      if (sp.has_pw()) {
        FEAR_TRY(launch_pw(c, ST_BACKBONE_PW, s, X, sp.cin, bw.pw, nullptr, 0, c->bufE, sp.mid(), M, 1));
        E = c->bufE;
      }
      if ((c->opt.fuse_dwpw & 1) && tc::available() && sp.stride == 1 && h == 16 && w == 16 && effective(c->opt.pw) == IMPL_TC) {
        // experimental: depthwise + project 1x1 in one tcgen05 kernel (the depthwise map is never written)
        float* dst = (i == last - 1 && final_out) ? final_out : Y;
        LaunchScope scope(c, ST_BACKBONE_PW, s);
        int r = tc::launch_pw_dw(s, E, B, sp.k, bw.dw.w, bw.dw.b, 1, bw.pwl.w_hi, bw.pwl.w_lo, bw.pwl.b,
                                 sp.residual() ? X : nullptr, sp.cout, dst, sp.cout, sp.cout, sp.mid(), 0);
        if (r < 0) return set_err(FEAR_EINVAL, "fused depthwise + 1x1 launch failed (%d)", r);
        if (r == 0) {
          FEAR_TRY(check_launch("tc::pw_tc_kernel<DWK>"));
          if (dst == Y) Y = (X == c->bufS) ? ((Y == c->bufX) ? c->bufY : c->bufX) : X;
          X = dst;
          continue;
        }
      }
      FEAR_TRY(launch_dw(c, ST_BACKBONE_DW, s, E, bw.dw, c->bufD, B, h, w, sp.stride, true));
    }
    h /= sp.stride;
    w /= sp.stride;
    float* dst = (i == last - 1 && final_out) ? final_out : Y;
    FEAR_TRY(launch_pw(c, ST_BACKBONE_PW, s, c->bufD, sp.mid(), bw.pwl, sp.residual() ? X : nullptr, sp.cout, dst,
                       sp.cout, B * h * w, 0));
    if (dst == Y) Y = (X == c->bufS) ? ((Y == c->bufX) ? c->bufY : c->bufX) : X;
    X = dst;
  }
  *out = X;
  return 0;
}

static StemNorm imagenet_norm() {
  // float32 arithmetic exactly as albumentations.Normalize does it (reference base_tracker.py:73)
  StemNorm n;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int i = 0; i < 3; ++i) {
    volatile float m = mean[i] * 255.0f, sd = stdv[i] * 255.0f;
    n.mean[i] = m;
    n.inv[i] = 1.0f / sd;
  }
  return n;
}

constexpr int kEarlyBlocks = 5;  // xif1_0 .. xif3_0: the high-resolution part (128^2 / 64^2 maps at 256^2 input)

// img (B,3,H,W) NCHW -> NHWC backbone features [B][H/16 * W/16][112] left in *feat (a workspace buffer).
// With opt.early_sub > 0 the high-resolution blocks run in sub-batches of that many frames so that their
// (6x expanded) intermediates stay resident in the 126 MB L2 instead of round-tripping through HBM.
static int run_backbone(FearContext* c, cudaStream_t s, const void* img, int B, int H, int W, const float** feat,
                        bool u8 = false) {
  const int sub = (c->opt.early_sub > 0 && c->opt.early_sub < B) ? c->opt.early_sub : B;
  const bool blocked = sub < B;
  const int eh = H / 8, ew = W / 8;                      // map size after the early blocks (stride 8)
  const long long per_frame_s = (long long)eh * ew * kBlocks[kEarlyBlocks - 1].cout;
  float* X = nullptr;
  int h = H / 2, w = W / 2;
  for (int b0 = 0; b0 < B; b0 += sub) {
    const int nb = (B - b0 < sub) ? B - b0 : sub;
    // (TMA needs 16-byte aligned image rows and base: W % 16 == 0 covers both layouts)
    const bool fuse_stem = c->opt.fuse_stem && tc::available() && (H / 2) % kFsTH == 0 && (W / 2) % kFsTW == 0 &&
                           (reinterpret_cast<uintptr_t>(img) & 15) == 0;
    if (fuse_stem) {
      // stem + xif1_0 (dw3x3 -> 1x1 + residual) in one pass over the image: the block output lands in bufX
      LaunchScope scope(c, ST_STEM, s);
      static bool attr_done = false;
      if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(stem_xif1_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFsSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(stem_xif1_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFsSmemBytes));
        attr_done = true;
      }
      const unsigned blocks = (unsigned)((long long)nb * ((H / 2) / kFsTH) * ((W / 2) / kFsTW));
      CUtensorMap tm;
      if (u8) {
        const uint8_t* base = static_cast<const uint8_t*>(img) + (long long)b0 * 3 * H * W;
        int r = tc::make_tmap_3d(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, base, (uint64_t)3 * W, (uint64_t)H, (uint64_t)nb,
                                 (uint64_t)3 * W, (uint64_t)3 * W * H, kFsRawPitch, kFsPH, 1);
        if (r) return set_err(FEAR_EINVAL, "tensor map for the uint8 image failed (%d)", r);
        stem_xif1_fused_kernel<true><<<blocks, kFsThreads, kFsSmemBytes, s>>>(tm, c->bufX, H, W, imagenet_norm(), c->fs);
      } else {
        const float* base = static_cast<const float*>(img) + (long long)b0 * 3 * H * W;
        int r = tc::make_tmap_3d(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, (uint64_t)W, (uint64_t)H, (uint64_t)3 * nb,
                                 (uint64_t)W * 4, (uint64_t)W * H * 4, kFsPP, kFsPH, 3);
        if (r) return set_err(FEAR_EINVAL, "tensor map for the float image failed (%d)", r);
        stem_xif1_fused_kernel<false><<<blocks, kFsThreads, kFsSmemBytes, s>>>(tm, c->bufX, H, W, StemNorm(), c->fs);
      }
      FEAR_TRY(check_launch("stem_xif1_fused_kernel"));
    } else {
      LaunchScope scope(c, ST_STEM, s);
      const unsigned blocks = (unsigned)((long long)nb * ((H / 2 + 3) / 4) * ((W / 2 + 31) / 32));
      if (u8)
        stem_conv3x3s2_kernel<true><<<blocks, 128, 0, s>>>(static_cast<const uint8_t*>(img) + (long long)b0 * 3 * H * W,
                                                           c->stem_w, c->stem_b, c->bufX, nb, H, W, imagenet_norm());
      else
        stem_conv3x3s2_kernel<false><<<blocks, 128, 0, s>>>(static_cast<const float*>(img) + (long long)b0 * 3 * H * W,
                                                            c->stem_w, c->stem_b, c->bufX, nb, H, W, StemNorm());
      FEAR_TRY(check_launch("stem_conv3x3s2_kernel"));
    }
    h = H / 2;
    w = W / 2;
    FEAR_TRY(run_blocks(c, s, c->bufX, nb, h, w, fuse_stem ? 1 : 0, kEarlyBlocks,
                        blocked ? c->bufS + b0 * per_frame_s : nullptr, &X));
  }
  if (blocked) X = c->bufS;
  float* out = nullptr;
  FEAR_TRY(run_blocks(c, s, X, B, h, w, kEarlyBlocks, kNumBlocks, nullptr, &out));
  *feat = out;
  return 0;
}

// img (B,3,H,W) NCHW -> out NHWC [B][H/16 * W/16][256]   (FEARNet.get_features, fear_net.py:63-66)
static int run_features(FearContext* c, cudaStream_t s, const void* img, int B, int H, int W, float* out,
                        bool u8 = false) {
  const float* X = nullptr;
  FEAR_TRY(run_backbone(c, s, img, B, H, W, &X, u8));
  return launch_pw(c, ST_NECK, s, X, kBackboneC, c->neck, nullptr, 0, out, kFeatC, B * (H / 16) * (W / 16), 0);
}

// F: NHWC search features [B][256][256]; zt: [Bz][64][256]; outputs NCHW maps.
// SepConv of the head (blocks.py:45-72): depthwise 3x3 (no bias, no activation) then 1x1 (+bias, ReLU).
// With the experimental "fuse_dwpw" bit 2 the pair runs as one tcgen05 kernel (the depthwise map is not written).
static int launch_sepconv(FearContext* c, cudaStream_t s, const float* X, const DwW& dw, const PwW& pw, float* out,
                          int ldc, int B) {
  const int M = B * kScorePix;
  if ((c->opt.fuse_dwpw & 2) && tc::available() && effective(c->opt.pw) == IMPL_TC) {
    LaunchScope scope(c, ST_HEAD_PW, s);
    int r = tc::launch_pw_dw(s, X, B, dw.k, dw.w, dw.b, 0, pw.w_hi, pw.w_lo, pw.b, nullptr, 0, out, ldc, pw.cout, pw.cin, 1);
    if (r < 0) return set_err(FEAR_EINVAL, "fused SepConv launch failed (%d)", r);
    if (r == 0) return check_launch("tc::pw_tc_kernel<3>");
  }
  FEAR_TRY(launch_dw(c, ST_HEAD_DW, s, X, dw, c->hT, B, kScore, kScore, 1, false));
  return launch_pw(c, ST_HEAD_PW, s, c->hT, pw.cin, pw, nullptr, 0, out, ldc, M, 1);
}

static int run_head(FearContext* c, cudaStream_t s, const float* zt, int Bz, const float* F, int B, float* bbox,
                    float* cls) {
  const int M = B * kScorePix;
  // the two concat buffers are laid out back to back for THIS batch so one correlation launch covers both
  c->hCAT[1] = c->hCAT[0] + (long long)B * kScorePix * kCatC;
  for (int br = 0; br < 2; ++br) {
    const BranchW& w = c->branch[br];
    // MatrixMobile: x -> dw3x3 -> 1x1 (+BN) -> ReLU, written into channels [0,256) of the concat buffer
    FEAR_TRY(launch_sepconv(c, s, F, w.enc_dw, w.enc_pw, c->hCAT[br], kCatC, B));
  }
  // pixel-wise correlation of both branches into channels [256,320) of their concat buffers
  FEAR_TRY(launch_corr(c, c->opt, s, zt, Bz, c->hCAT[0], B, 2, c->zth, c->ztl));
  for (int br = 0; br < 2; ++br) {
    const BranchW& w = c->branch[br];
    // MobileCorrelation.enc: dw3x3(320) -> 1x1 320->256 (+BN) -> ReLU
    FEAR_TRY(launch_sepconv(c, s, c->hCAT[br], w.corr_dw, w.corr_pw, c->hD[br], kFeatC, B));
  }
  // towers: tower[0] = bbox_tower on reg branch (hD[1]); tower[1] = cls_tower on cls branch (hD[0])
  for (int t = 0; t < 2; ++t) {
    const float* x = c->hD[t == 0 ? 1 : 0];
    float* outs[2] = {c->hP, c->hQ[t]};
    for (int i = 0; i < 2; ++i) {
      FEAR_TRY(launch_sepconv(c, s, x, c->tower[t].dw[i], c->tower[t].pw[i], outs[i], kFeatC, B));
      x = outs[i];
    }
    FEAR_TRY(launch_dw(c, ST_HEAD_DW, s, x, c->pred_dw[t], c->hT, B, kScore, kScore, 1, false));
    LaunchScope scope(c, ST_PRED, s);
    const unsigned blocks = (unsigned)((M * 32 + 255) / 256);
    if (t == 0)
      pred_pw_kernel<4, true><<<blocks, 256, 0, s>>>(c->hT, c->pred_w[0], c->pred_b[0], bbox, B);
    else
      pred_pw_kernel<1, false><<<blocks, 256, 0, s>>>(c->hT, c->pred_w[1], c->pred_b[1], cls, B);
    FEAR_TRY(check_launch("pred_pw_kernel"));
  }
  return 0;
}

static int run_decode(FearContext* c, cudaStream_t s, const float* bbox, const float* cls, int B, int apply_sigmoid,
                      FearBox* boxes) {
  LaunchScope scope(c, ST_DECODE, s);
  decode_kernel<<<B, 256, 0, s>>>(bbox, cls, apply_sigmoid, boxes);
  return check_launch("decode_kernel");
}

// ------------------------------------------------------------------------------ C ABI
static bool g_inited = false;
static int g_device = 0;

extern "C" int fear_abi_version(void) { return FEAR_ABI_VERSION; }
extern "C" const char* fear_last_error(void) { return g_err; }

extern "C" int fear_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return set_err(FEAR_ENODEV, "no CUDA device: %s", cudaGetErrorString(e));
  if (device < 0 || device >= n) return set_err(FEAR_EINVAL, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp p;
  CUDA_TRY(cudaGetDeviceProperties(&p, device));
  if (p.major != 10)
    return set_err(FEAR_ENODEV, "device %d is sm_%d%d; libfear_b200 is built for sm_100a only", device, p.major, p.minor);
  CUDA_TRY(cudaSetDevice(device));
  g_device = device;
  g_inited = true;
  return tc::init();
}

extern "C" int fear_weight_count(void) { return (int)weight_table().size(); }
extern "C" const char* fear_weight_name(int i) {
  if (i < 0 || i >= fear_weight_count()) return nullptr;
  return weight_table()[i].name.c_str();
}
extern "C" int64_t fear_weight_numel(int i) {
  if (i < 0 || i >= fear_weight_count()) return -1;
  return weight_table()[i].numel;
}
extern "C" int fear_stage_count(void) { return ST_COUNT; }
extern "C" const char* fear_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : nullptr; }

extern "C" int fear_pack_weights(const float* blob, const uint64_t* offsets, int n, FearContext** handle) {
  if (!g_inited) return set_err(FEAR_ESTATE, "fear_init() has not been called");
  if (!blob || !offsets || !handle) return set_err(FEAR_EINVAL, "null argument");
  const auto& table = weight_table();
  if (n != (int)table.size()) return set_err(FEAR_EINVAL, "expected %d tensors, got %d", (int)table.size(), n);
  for (int i = 0; i < n; ++i)
    if ((int64_t)(offsets[i + 1] - offsets[i]) != table[i].numel)
      return set_err(FEAR_EINVAL, "tensor %d (%s): expected %lld elements, got %lld", i, table[i].name.c_str(),
                     (long long)table[i].numel, (long long)(offsets[i + 1] - offsets[i]));

  // Device arena: every tensor 256-byte aligned; depthwise [C][k][k] -> [k*k][C], stem -> [27][16].
  std::vector<float> arena;
  std::vector<size_t> dev_off(n), hi_off(n, 0), lo_off(n, 0);
  auto is_gemm_weight = [&](const std::string& nm) {
    if (nm.rfind("bbox_pred", 0) == 0 || nm.rfind("cls_pred", 0) == 0) return false;
    auto ends = [&](const char* suf) {
      const size_t l = strlen(suf);
      return nm.size() >= l && nm.compare(nm.size() - l, l, suf) == 0;
    };
    return ends(".pw.w") || ends(".pwl.w") || nm == "neck.w";
  };
  for (int i = 0; i < n; ++i) {
    size_t o = (arena.size() + 63) & ~(size_t)63;
    arena.resize(o + table[i].numel, 0.f);
    dev_off[i] = o;
    const float* src = blob + offsets[i];
    const std::string& nm = table[i].name;
    const bool is_dw = nm.size() > 5 && nm.compare(nm.size() - 5, 5, ".dw.w") == 0;
    if (nm == "stem.w") {
      for (int co = 0; co < 16; ++co)
        for (int t = 0; t < 27; ++t) arena[o + t * 16 + co] = src[co * 27 + t];
    } else if (is_dw) {
      // numel = C * kk; kk is 9 or 25.  Find it from the matching bias / table neighbour: C divides numel.
      int kk = 9;
      const bool head = nm.find("xif") == std::string::npos;
      if (!head) {
        for (const IrfSpec& b : kBlocks)
          if (nm == std::string(b.name) + ".dw.w") kk = b.k * b.k;
      }
      const int64_t C = table[i].numel / kk;
      for (int64_t ch = 0; ch < C; ++ch)
        for (int t = 0; t < kk; ++t) arena[o + (int64_t)t * C + ch] = src[ch * kk + t];
    } else {
      memcpy(&arena[o], src, sizeof(float) * table[i].numel);
    }
    if (is_gemm_weight(nm)) {  // tf32 (hi, lo) split for the 3xTF32 tensor-core GEMM
      for (int part = 0; part < 2; ++part) {
        size_t po = (arena.size() + 63) & ~(size_t)63;
        arena.resize(po + table[i].numel, 0.f);
        (part == 0 ? hi_off : lo_off)[i] = po;
        for (int64_t e = 0; e < table[i].numel; ++e) {
          const float hi = tc::host_rna_tf32(src[e]);
          arena[po + e] = part == 0 ? hi : tc::host_rna_tf32(src[e] - hi);
        }
      }
    }
  }
  arena.resize((arena.size() + 63) & ~(size_t)63, 0.f);

  FearContext* c = new FearContext();
  c->device = g_device;
  c->opt = g_default_options;
  cudaError_t e = cudaMalloc(&c->d_weights, arena.size() * sizeof(float));
  if (e != cudaSuccess) {
    delete c;
    return set_err(FEAR_ENOMEM, "cudaMalloc(weights) failed: %s", cudaGetErrorString(e));
  }
  e = cudaMemcpy(c->d_weights, arena.data(), arena.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(c->d_weights);
    delete c;
    return set_err((int)e, "cudaMemcpy(weights) failed: %s", cudaGetErrorString(e));
  }
  int idx = 0;
  auto next = [&]() { return (const float*)(c->d_weights + dev_off[idx++]); };
  auto next_pw = [&](PwW& w, int cin, int cout) {  // weight (+ its hi/lo copies) followed by its bias
    w.w_hi = c->d_weights + hi_off[idx];
    w.w_lo = c->d_weights + lo_off[idx];
    w.w = next();
    w.b = next();
    w.h_w = reinterpret_cast<const float*>((w.w - c->d_weights));  // offsets for now; rebased onto h_weights below
    w.h_b = reinterpret_cast<const float*>((w.b - c->d_weights));
    w.cin = cin;
    w.cout = cout;
  };
  c->stem_w = next();
  c->stem_b = next();
  for (int i = 0; i < kNumBlocks; ++i) {
    const IrfSpec& sp = kBlocks[i];
    BlockW& b = c->blocks[i];
    if (sp.has_pw()) next_pw(b.pw, sp.cin, sp.mid());
    b.dw.w = next();
    b.dw.b = next();
    b.dw.c = sp.mid();
    b.dw.k = sp.k;
    next_pw(b.pwl, sp.mid(), sp.cout);
  }
  next_pw(c->neck, kBackboneC, kFeatC);
  for (int br = 0; br < 2; ++br) {
    BranchW& w = c->branch[br];
    w.enc_dw = {next(), nullptr, kFeatC, 3};
    next_pw(w.enc_pw, kFeatC, kFeatC);
    w.corr_dw = {next(), nullptr, kCatC, 3};
    next_pw(w.corr_pw, kCatC, kFeatC);
  }
  for (int t = 0; t < 2; ++t)
    for (int i = 0; i < 2; ++i) {
      c->tower[t].dw[i] = {next(), nullptr, kFeatC, 3};
      next_pw(c->tower[t].pw[i], kFeatC, kFeatC);
    }
  for (int t = 0; t < 2; ++t) {
    c->pred_dw[t] = {next(), nullptr, kFeatC, 3};
    c->pred_w[t] = next();
    c->pred_b[t] = next();
  }
  if (idx != n) {
    cudaFree(c->d_weights);
    delete c;
    return set_err(FEAR_ESTATE, "internal: weight table walk consumed %d of %d tensors", idx, n);
  }
  c->h_weights = std::move(arena);
  {
    auto rebase = [&](PwW& w) {
      w.h_w = c->h_weights.data() + reinterpret_cast<intptr_t>(w.h_w);
      w.h_b = c->h_weights.data() + reinterpret_cast<intptr_t>(w.h_b);
    };
    for (int i = 0; i < kNumBlocks; ++i) {
      if (kBlocks[i].has_pw()) rebase(c->blocks[i].pw);
      rebase(c->blocks[i].pwl);
    }
    rebase(c->neck);
    for (int br = 0; br < 2; ++br) {
      rebase(c->branch[br].enc_pw);
      rebase(c->branch[br].corr_pw);
    }
    for (int t = 0; t < 2; ++t)
      for (int i = 0; i < 2; ++i) rebase(c->tower[t].pw[i]);
  }
  {
    auto host_of = [&](const float* dptr) { return c->h_weights.data() + (dptr - c->d_weights); };
    const BlockW& b0 = c->blocks[0];
    memcpy(c->fs.sw, host_of(c->stem_w), sizeof(c->fs.sw));
    memcpy(c->fs.sb, host_of(c->stem_b), sizeof(c->fs.sb));
    memcpy(c->fs.dw, host_of(b0.dw.w), sizeof(c->fs.dw));
    memcpy(c->fs.db, host_of(b0.dw.b), sizeof(c->fs.db));
    const float* pw = host_of(b0.pwl.w);  // [o][k]
    for (int o = 0; o < 16; ++o)
      for (int k = 0; k < 16; ++k) c->fs.pw[k * 16 + o] = pw[o * 16 + k];
    memcpy(c->fs.pb, host_of(b0.pwl.b), sizeof(c->fs.pb));
  }
  *handle = c;
  int r = fear_reserve(c, 1);
  if (r) {
    fear_free(c);
    *handle = nullptr;
  }
  return r;
}

extern "C" int fear_reserve(FearContext* c, int max_batch) {
  if (!c) return set_err(FEAR_ESTATE, "null handle");
  if (max_batch < 1) return set_err(FEAR_EINVAL, "max_batch must be >= 1");
  if (max_batch <= c->reserved) return 0;
  CUDA_TRY(cudaDeviceSynchronize());
  if (c->ws) cudaFree(c->ws);
  c->ws = nullptr;
  c->reserved = 0;
  const int64_t per_frame[] = {
      kActX, kActX, kActE, kActD,                                   // bufX bufY bufE bufD
      (int64_t)kScorePix * kFeatC,                                  // hF
      (int64_t)kScorePix * kCatC,                                   // hT
      (int64_t)kScorePix * kCatC, (int64_t)kScorePix * kCatC,       // hCAT[2]
      (int64_t)kScorePix * kFeatC, (int64_t)kScorePix * kFeatC,     // hD[2]
      (int64_t)kScorePix * kFeatC,                                  // hP
      (int64_t)kScorePix * kFeatC, (int64_t)kScorePix * kFeatC,     // hQ[2]
      (int64_t)kTmplPix * kFeatC,                                   // zt
      4 * kScorePix, kScorePix,                                     // mapB mapC
      (int64_t)kTmplPix * kFeatC, (int64_t)kTmplPix * kFeatC,       // zth ztl
      32 * 32 * 32,                                                 // bufS
  };
  int64_t total = 0;
  std::vector<int64_t> offs;
  for (int64_t pf : per_frame) {
    offs.push_back(total);
    total += ((pf * max_batch + 63) / 64) * 64;
  }
  cudaError_t e = cudaMalloc(&c->ws, (size_t)total * sizeof(float));
  if (e != cudaSuccess)
    return set_err(FEAR_ENOMEM, "workspace cudaMalloc(%lld MB) failed: %s", (long long)(total * 4 >> 20),
                   cudaGetErrorString(e));
  float* p = c->ws;
  c->bufX = p + offs[0];
  c->bufY = p + offs[1];
  c->bufE = p + offs[2];
  c->bufD = p + offs[3];
  c->hF = p + offs[4];
  c->hT = p + offs[5];
  c->hCAT[0] = p + offs[6];
  c->hCAT[1] = p + offs[7];
  c->hD[0] = p + offs[8];
  c->hD[1] = p + offs[9];
  c->hP = p + offs[10];
  c->hQ[0] = p + offs[11];
  c->hQ[1] = p + offs[12];
  c->zt = p + offs[13];
  c->mapB = p + offs[14];
  c->mapC = p + offs[15];
  c->zth = p + offs[16];
  c->ztl = p + offs[17];
  c->bufS = p + offs[18];
  c->reserved = max_batch;
  return 0;
}

extern "C" void fear_free(FearContext* c) {
  if (!c) return;
  cudaDeviceSynchronize();
  for (auto& ev : c->events) {
    cudaEventDestroy(ev.a);
    cudaEventDestroy(ev.b);
  }
  if (c->ws) cudaFree(c->ws);
  if (c->d_weights) cudaFree(c->d_weights);
  delete c;
}

static int check_ctx(FearContext* c) {
  if (!c || !c->d_weights || !c->ws) return set_err(FEAR_ESTATE, "handle not initialised");
  return 0;
}

extern "C" int fear_get_features(FearContext* c, const float* d_img, int B, int H, int W, float* d_feat, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_img || !d_feat || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (H % 16 || W % 16 || H < 16 || W < 16 || H > 256 || W > 256)
    return set_err(FEAR_EINVAL, "H, W must be multiples of 16 in [16, 256] (got %dx%d)", H, W);
  cudaStream_t s = (cudaStream_t)stream;
  const int P = (H / 16) * (W / 16);
  for (int b0 = 0; b0 < B; b0 += c->reserved) {
    const int nb = (B - b0 < c->reserved) ? B - b0 : c->reserved;
    FEAR_TRY(run_features(c, s, d_img + (long long)b0 * 3 * H * W, nb, H, W, c->hF));
    FEAR_TRY(launch_transpose(c, s, c->hF, kFeatC, (long long)P * kFeatC, d_feat + (long long)b0 * kFeatC * P, P,
                              (long long)kFeatC * P, P, kFeatC, nb));
  }
  return 0;
}

extern "C" int fear_backbone(FearContext* c, const float* d_img, int B, int H, int W, float* d_feat, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_img || !d_feat || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (H % 16 || W % 16 || H < 16 || W < 16 || H > 256 || W > 256)
    return set_err(FEAR_EINVAL, "H, W must be multiples of 16 in [16, 256] (got %dx%d)", H, W);
  cudaStream_t s = (cudaStream_t)stream;
  const int P = (H / 16) * (W / 16);
  for (int b0 = 0; b0 < B; b0 += c->reserved) {
    const int nb = (B - b0 < c->reserved) ? B - b0 : c->reserved;
    const float* X = nullptr;
    FEAR_TRY(run_backbone(c, s, d_img + (long long)b0 * 3 * H * W, nb, H, W, &X));
    FEAR_TRY(launch_transpose(c, s, X, kBackboneC, (long long)P * kBackboneC, d_feat + (long long)b0 * kBackboneC * P,
                              P, (long long)kBackboneC * P, P, kBackboneC, nb));
  }
  return 0;
}

// zfeat NCHW (Bz,256,8,8) -> c->zt chunk [nz][64][256]
static int stage_template(FearContext* c, cudaStream_t s, const float* d_zfeat, int nz) {
  return launch_transpose(c, s, d_zfeat, kTmplPix, (long long)kFeatC * kTmplPix, c->zt, kFeatC,
                          (long long)kTmplPix * kFeatC, kFeatC, kTmplPix, nz);
}

extern "C" int fear_head(FearContext* c, const float* d_zfeat, int Bz, const float* d_xfeat, int B, float* d_bbox,
                         float* d_cls, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_zfeat || !d_xfeat || !d_bbox || !d_cls || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (Bz != 1 && Bz != B) return set_err(FEAR_EINVAL, "template batch must be 1 or B (got %d vs %d)", Bz, B);
  cudaStream_t s = (cudaStream_t)stream;
  if (Bz == 1) FEAR_TRY(stage_template(c, s, d_zfeat, 1));
  for (int b0 = 0; b0 < B; b0 += c->reserved) {
    const int nb = (B - b0 < c->reserved) ? B - b0 : c->reserved;
    if (Bz != 1) FEAR_TRY(stage_template(c, s, d_zfeat + (long long)b0 * kFeatC * kTmplPix, nb));
    FEAR_TRY(launch_transpose(c, s, d_xfeat + (long long)b0 * kFeatC * kScorePix, kScorePix,
                              (long long)kFeatC * kScorePix, c->hF, kFeatC, (long long)kScorePix * kFeatC, kFeatC,
                              kScorePix, nb));
    FEAR_TRY(run_head(c, s, c->zt, Bz == 1 ? 1 : nb, c->hF, nb, d_bbox + (long long)b0 * 4 * kScorePix,
                      d_cls + (long long)b0 * kScorePix));
  }
  return 0;
}

static int track_impl(FearContext* c, cudaStream_t s, const float* d_template, const void* d_search,
                      const float* d_zfeat, int Bz, int B, float* d_bbox, float* d_cls, FearBox* d_boxes,
                      bool search_u8 = false) {
  if (d_zfeat && Bz == 1) FEAR_TRY(stage_template(c, s, d_zfeat, 1));
  for (int b0 = 0; b0 < B; b0 += c->reserved) {
    const int nb = (B - b0 < c->reserved) ? B - b0 : c->reserved;
    int nz = nb;
    if (d_template) {
      // template branch writes NHWC [nb][64][256] straight into zt (= the correlation kernel's layout)
      FEAR_TRY(run_features(c, s, d_template + (long long)b0 * 3 * 128 * 128, nb, 128, 128, c->zt));
    } else if (Bz != 1) {
      FEAR_TRY(stage_template(c, s, d_zfeat + (long long)b0 * kFeatC * kTmplPix, nb));
    } else {
      nz = 1;
    }
    const void* sp = search_u8 ? (const void*)(static_cast<const uint8_t*>(d_search) + (long long)b0 * 3 * 256 * 256)
                               : (const void*)(static_cast<const float*>(d_search) + (long long)b0 * 3 * 256 * 256);
    FEAR_TRY(run_features(c, s, sp, nb, 256, 256, c->hF, search_u8));
    float* bb = d_bbox ? d_bbox + (long long)b0 * 4 * kScorePix : c->mapB;
    float* cc = d_cls ? d_cls + (long long)b0 * kScorePix : c->mapC;
    FEAR_TRY(run_head(c, s, c->zt, nz, c->hF, nb, bb, cc));
    if (d_boxes) FEAR_TRY(run_decode(c, s, bb, cc, nb, 1, d_boxes + b0));
  }
  return 0;
}

extern "C" int fear_track(FearContext* c, const float* d_search, const float* d_zfeat, int Bz, int B, float* d_bbox,
                          float* d_cls, FearBox* d_boxes, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_search || !d_zfeat || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (Bz != 1 && Bz != B) return set_err(FEAR_EINVAL, "template batch must be 1 or B (got %d vs %d)", Bz, B);
  if (!d_boxes && (!d_bbox || !d_cls)) return set_err(FEAR_EINVAL, "no output requested");
  return track_impl(c, (cudaStream_t)stream, nullptr, d_search, d_zfeat, Bz, B, d_bbox, d_cls, d_boxes);
}

extern "C" int fear_track_u8(FearContext* c, const uint8_t* d_search_u8, const float* d_zfeat, int Bz, int B,
                             float* d_bbox, float* d_cls, FearBox* d_boxes, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_search_u8 || !d_zfeat || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (Bz != 1 && Bz != B) return set_err(FEAR_EINVAL, "template batch must be 1 or B (got %d vs %d)", Bz, B);
  if (!d_boxes && (!d_bbox || !d_cls)) return set_err(FEAR_EINVAL, "no output requested");
  return track_impl(c, (cudaStream_t)stream, nullptr, d_search_u8, d_zfeat, Bz, B, d_bbox, d_cls, d_boxes, true);
}

extern "C" int fear_get_features_u8(FearContext* c, const uint8_t* d_img_u8, int B, int H, int W, float* d_feat,
                                    void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_img_u8 || !d_feat || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (H % 16 || W % 16 || H < 16 || W < 16 || H > 256 || W > 256)
    return set_err(FEAR_EINVAL, "H, W must be multiples of 16 in [16, 256] (got %dx%d)", H, W);
  cudaStream_t s = (cudaStream_t)stream;
  const int P = (H / 16) * (W / 16);
  for (int b0 = 0; b0 < B; b0 += c->reserved) {
    const int nb = (B - b0 < c->reserved) ? B - b0 : c->reserved;
    FEAR_TRY(run_features(c, s, d_img_u8 + (long long)b0 * 3 * H * W, nb, H, W, c->hF, true));
    FEAR_TRY(launch_transpose(c, s, c->hF, kFeatC, (long long)P * kFeatC, d_feat + (long long)b0 * kFeatC * P, P,
                              (long long)kFeatC * P, P, kFeatC, nb));
  }
  return 0;
}

extern "C" int fear_forward(FearContext* c, const float* d_template, const float* d_search, int B, float* d_bbox,
                            float* d_cls, FearBox* d_boxes, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_template || !d_search || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (!d_boxes && (!d_bbox || !d_cls)) return set_err(FEAR_EINVAL, "no output requested");
  return track_impl(c, (cudaStream_t)stream, d_template, d_search, nullptr, B, B, d_bbox, d_cls, d_boxes);
}

extern "C" int fear_decode(const float* d_bbox, const float* d_cls, int B, int apply_sigmoid, FearBox* d_boxes,
                           void* stream) {
  if (!d_bbox || !d_cls || !d_boxes || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  return run_decode(nullptr, (cudaStream_t)stream, d_bbox, d_cls, B, apply_sigmoid, d_boxes);
}

// second scratch (hi/lo template planes of the handle-less channels-last correlation)
static float* g_scratch2 = nullptr;
static size_t g_scratch2_floats = 0;
static int ensure_scratch2(size_t need) {
  if (need <= g_scratch2_floats) return 0;
  CUDA_TRY(cudaDeviceSynchronize());
  if (g_scratch2) cudaFree(g_scratch2);
  g_scratch2 = nullptr;
  g_scratch2_floats = 0;
  CUDA_TRY(cudaMalloc(&g_scratch2, need * sizeof(float)));
  g_scratch2_floats = need;
  return 0;
}

extern "C" int fear_corr_nhwc_f32(const float* d_zt, int Bz, float* d_cat, int B, void* stream) {
  if (!d_zt || !d_cat || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (Bz != 1 && Bz != B) return set_err(FEAR_EINVAL, "template batch must be 1 or B (got %d vs %d)", Bz, B);
  float *zth = nullptr, *ztl = nullptr;
  if (effective(g_default_options.corr) == IMPL_TS || effective(g_default_options.corr) == IMPL_TC2) {
    const size_t need = 2 * (size_t)Bz * kCorrC * kFeatC;
    FEAR_TRY(ensure_scratch2(need));
    zth = g_scratch2;
    ztl = g_scratch2 + (size_t)Bz * kCorrC * kFeatC;
  }
  return launch_corr(nullptr, g_default_options, (cudaStream_t)stream, d_zt, Bz, d_cat, B, 1, zth, ztl);
}

// Scratch for the handle-less NCHW wrapper; grows (cudaMalloc) only when a larger batch arrives.
static float* g_scratch = nullptr;
static size_t g_scratch_floats = 0;

extern "C" int fear_corr_concat_f32(const float* d_z, int Bz, const float* d_x, int B, float* d_out, void* stream) {
  if (!d_z || !d_x || !d_out || B < 1) return set_err(FEAR_EINVAL, "bad argument");
  if (Bz != 1 && Bz != B) return set_err(FEAR_EINVAL, "template batch must be 1 or B (got %d vs %d)", Bz, B);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t need = (size_t)B * kScorePix * kCatC + 3 * (size_t)Bz * kCorrC * kFeatC;
  if (need > g_scratch_floats) {
    CUDA_TRY(cudaDeviceSynchronize());
    if (g_scratch) cudaFree(g_scratch);
    g_scratch = nullptr;
    g_scratch_floats = 0;
    CUDA_TRY(cudaMalloc(&g_scratch, need * sizeof(float)));
    g_scratch_floats = need;
  }
  float* cat = g_scratch;
  float* zt = g_scratch + (size_t)B * kScorePix * kCatC;
  // z [c][k] -> zt [k][c];  x [c][p] -> cat[p][0:256]
  FEAR_TRY(launch_transpose(nullptr, s, d_z, kCorrC, (long long)kFeatC * kCorrC, zt, kFeatC, (long long)kCorrC * kFeatC,
                            kFeatC, kCorrC, Bz));
  FEAR_TRY(launch_transpose(nullptr, s, d_x, kScorePix, (long long)kFeatC * kScorePix, cat, kCatC,
                            (long long)kScorePix * kCatC, kFeatC, kScorePix, B));
  FEAR_TRY(launch_corr(nullptr, g_default_options, s, zt, Bz, cat, B, 1, zt + (size_t)Bz * kCorrC * kFeatC,
                       zt + 2 * (size_t)Bz * kCorrC * kFeatC));
  // cat [p][320] -> out [320][p]
  return launch_transpose(nullptr, s, cat, kCatC, (long long)kScorePix * kCatC, d_out, kScorePix,
                          (long long)kCatC * kScorePix, kScorePix, kCatC, B);
}

// ---- debug / introspection of intermediates (tests localise a mismatch with these) ----------
extern "C" int fear_debug_backbone_prefix(FearContext* c, const float* d_img, int B, int H, int W, int nblocks,
                                          float* d_out, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!d_img || !d_out || B < 1 || B > c->reserved || nblocks < 0 || nblocks > kNumBlocks)
    return set_err(FEAR_EINVAL, "bad argument (B must be <= reserved batch)");
  cudaStream_t s = (cudaStream_t)stream;
  {
    LaunchScope scope(c, ST_STEM, s);
    const unsigned blocks = (unsigned)((long long)B * ((H / 2 + 3) / 4) * ((W / 2 + 31) / 32));
    stem_conv3x3s2_kernel<false><<<blocks, 128, 0, s>>>(d_img, c->stem_w, c->stem_b, c->bufX, B, H, W, StemNorm());
    FEAR_TRY(check_launch("stem_conv3x3s2_kernel"));
  }
  int h = H / 2, w = W / 2, ch = kStemC;
  float *X = c->bufX, *Y = c->bufY;
  for (int i = 0; i < nblocks; ++i) {
    const IrfSpec& sp = kBlocks[i];
    const BlockW& bw = c->blocks[i];
    const float* E = X;
    if (sp.has_pw()) {
      FEAR_TRY(launch_pw(c, ST_BACKBONE_PW, s, X, sp.cin, bw.pw, nullptr, 0, c->bufE, sp.mid(), B * h * w, 1));
      E = c->bufE;
    }
    FEAR_TRY(launch_dw(c, ST_BACKBONE_DW, s, E, bw.dw, c->bufD, B, h, w, sp.stride, true));
    h /= sp.stride;
    w /= sp.stride;
    FEAR_TRY(launch_pw(c, ST_BACKBONE_PW, s, c->bufD, sp.mid(), bw.pwl, sp.residual() ? X : nullptr, sp.cout, Y,
                       sp.cout, B * h * w, 0));
    float* t = X;
    X = Y;
    Y = t;
    ch = sp.cout;
  }
  const int P = h * w;
  return launch_transpose(c, s, X, ch, (long long)P * ch, d_out, P, (long long)ch * P, P, ch, B);
}

// Copy a head intermediate of the LAST run (first B frames) out as NCHW (B, C, 16, 16).
extern "C" int fear_debug_head_tensor(FearContext* c, const char* name, int B, float* d_out, void* stream) {
  FEAR_TRY(check_ctx(c));
  if (!name || !d_out || B < 1 || B > c->reserved) return set_err(FEAR_EINVAL, "bad argument");
  const float* src = nullptr;
  int ch = kFeatC;
  if (!strcmp(name, "cat_cls")) src = c->hCAT[0], ch = kCatC;
  else if (!strcmp(name, "cat_reg")) src = c->hCAT[1], ch = kCatC;
  else if (!strcmp(name, "cls_dw")) src = c->hD[0];
  else if (!strcmp(name, "reg_dw")) src = c->hD[1];
  else if (!strcmp(name, "x_reg")) src = c->hQ[0];
  else if (!strcmp(name, "cls_tower")) src = c->hQ[1];
  else if (!strcmp(name, "search_features")) src = c->hF;
  else return set_err(FEAR_EINVAL, "unknown head tensor '%s'", name);
  return launch_transpose(c, (cudaStream_t)stream, src, ch, (long long)kScorePix * ch, d_out, kScorePix,
                          (long long)ch * kScorePix, kScorePix, ch, B);
}

extern "C" int fear_set_option(FearContext* c, const char* key, const char* value) {
  if (!key || !value) return set_err(FEAR_EINVAL, "null option");
  Options& o = c ? c->opt : g_default_options;
  if (!strcmp(key, "pdl")) {  // process-wide: programmatic dependent launch for the TMA / tcgen05 kernels
    tc::pdl_enabled() = atoi(value) != 0;
    return 0;
  }
  if (!strcmp(key, "fuse_dwpw")) {
    o.fuse_dwpw = atoi(value) & 3;  // bit 0: 16x16-stage backbone blocks, bit 1: the head's SepConvs
    return 0;
  }
  if (!strcmp(key, "small_const")) {
    o.small_const = atoi(value) != 0;
    return 0;
  }
  if (!strcmp(key, "fuse_stem")) {
    o.fuse_stem = atoi(value) != 0;
    return 0;
  }
  if (!strcmp(key, "dw_wide")) {
    o.dw_wide = atoi(value) != 0;
    return 0;
  }
  if (!strcmp(key, "fuse")) {
    o.fuse = atoi(value) != 0;
    return 0;
  }
  if (!strcmp(key, "early_sub")) {
    o.early_sub = atoi(value);
    return 0;
  }
  if (!strcmp(key, "dw")) {
    if (!strcmp(value, "pixel")) o.dw = 0;
    else if (!strcmp(value, "strip")) o.dw = 1;
    else if (!strcmp(value, "roll")) o.dw = 2;
    else if (!strcmp(value, "auto")) o.dw = 3;  // measured best per shape: TMA pipeline where it applies, else rolling window (3x3 s1) / register strip
    else if (!strcmp(value, "tile")) o.dw = 4;
    else if (!strcmp(value, "blocked")) o.dw = 5;
    else if (!strcmp(value, "tma")) o.dw = 6;
    else return set_err(FEAR_EINVAL, "unknown depthwise implementation '%s' (pixel | strip | roll | tile | blocked | tma | auto)", value);
    return 0;
  }
  int impl;
  if (!strcmp(value, "ffma")) impl = IMPL_FFMA;
  else if (!strcmp(value, "tcgen05")) impl = IMPL_TC;
  else if (!strcmp(value, "tcgen05ts")) impl = IMPL_TS;
  else if (!strcmp(value, "tcgen05v2")) impl = IMPL_TC2;
  else if (!strcmp(value, "auto")) impl = -1;
  else return set_err(FEAR_EINVAL, "unknown implementation '%s' (auto | ffma | tcgen05 | tcgen05ts)", value);
  if (impl > IMPL_FFMA && !tc::available()) return set_err(FEAR_EINVAL, "tcgen05 kernels not available in this build");
  if (!strcmp(key, "corr")) o.corr = impl;
  else if (!strcmp(key, "pw")) o.pw = impl;
  else return set_err(FEAR_EINVAL, "unknown option '%s' (corr | pw)", key);
  return 0;
}

extern "C" int64_t fear_launch_count(const FearContext* c) { return c ? c->launches : 0; }

extern "C" int fear_profile(FearContext* c, int enable) {
  FEAR_TRY(check_ctx(c));
  if (enable && c->events.empty()) {
    c->events.resize(8192);
    for (auto& ev : c->events) {
      CUDA_TRY(cudaEventCreate(&ev.a));
      CUDA_TRY(cudaEventCreate(&ev.b));
    }
  }
  c->profiling = enable != 0;
  c->events_used = 0;
  for (int i = 0; i < ST_COUNT; ++i) {
    c->stage_ms[i] = 0;
    c->stage_launches[i] = 0;
  }
  return 0;
}

extern "C" int fear_stage_ms(FearContext* c, int i, float* ms, int64_t* launches) {
  FEAR_TRY(check_ctx(c));
  if (i < 0 || i >= ST_COUNT) return set_err(FEAR_EINVAL, "stage index out of range");
  if (c->events_used) {
    CUDA_TRY(cudaDeviceSynchronize());
    for (size_t k = 0; k < c->events_used; ++k) {
      float t = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&t, c->events[k].a, c->events[k].b));
      c->stage_ms[c->events[k].stage] += t;
    }
    c->events_used = 0;
  }
  if (ms) *ms = (float)c->stage_ms[i];
  if (launches) *launches = c->stage_launches[i];
  return 0;
}
