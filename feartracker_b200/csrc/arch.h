// FEAR-XS architecture tables shared by the weight packer and the executor.
//
// Backbone = fbnet_c blocks 0..17 as executed by the reference (model_training/model/blocks.py:27-35,
// fear_net.py:58-61 with max_layer=4); the block parameters follow the checkpoint's tensor shapes
// (SURVEY.md section 8(a)).  xif2_1 is an identity in fbnet_c and is skipped.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace fear {

struct IrfSpec {
  const char* name;
  int cin, cout, k, stride, expand;
  int mid() const { return cin * expand; }
  bool residual() const { return stride == 1 && cin == cout; }
  bool has_pw() const { return expand != 1; }
};

static const IrfSpec kBlocks[] = {
    {"xif1_0", 16, 16, 3, 1, 1},   {"xif2_0", 16, 24, 3, 2, 6},   {"xif2_2", 24, 24, 3, 1, 1},
    {"xif2_3", 24, 24, 3, 1, 1},   {"xif3_0", 24, 32, 5, 2, 6},   {"xif3_1", 32, 32, 5, 1, 3},
    {"xif3_2", 32, 32, 5, 1, 6},   {"xif3_3", 32, 32, 3, 1, 6},   {"xif4_0", 32, 64, 5, 2, 6},
    {"xif4_1", 64, 64, 5, 1, 3},   {"xif4_2", 64, 64, 5, 1, 6},   {"xif4_3", 64, 64, 5, 1, 6},
    {"xif4_4", 64, 112, 5, 1, 6},  {"xif4_5", 112, 112, 5, 1, 6}, {"xif4_6", 112, 112, 5, 1, 6},
    {"xif4_7", 112, 112, 5, 1, 3},
};
constexpr int kNumBlocks = sizeof(kBlocks) / sizeof(kBlocks[0]);

constexpr int kStemC = 16;
constexpr int kBackboneC = 112;  // encoder_channels["layer1"], blocks.py:16
constexpr int kFeatC = 256;      // adjust_channels, fear_net.py:21
constexpr int kCorrC = 64;       // num_corr_channels, blocks.py:113
constexpr int kCatC = kFeatC + kCorrC;
constexpr int kScore = 16;
constexpr int kScorePix = kScore * kScore;
constexpr int kTmplPix = 64;

struct WeightInfo {
  std::string name;
  int64_t numel;
};

// Canonical order of the BN-folded tensors the host packs (see include/fear_b200.h).
inline const std::vector<WeightInfo>& weight_table() {
  static const std::vector<WeightInfo> table = [] {
    std::vector<WeightInfo> t;
    auto add = [&](const std::string& n, int64_t e) { t.push_back({n, e}); };
    add("stem.w", 16 * 27);
    add("stem.b", 16);
    for (const IrfSpec& b : kBlocks) {
      std::string n = b.name;
      int mid = b.mid();
      if (b.has_pw()) {
        add(n + ".pw.w", (int64_t)mid * b.cin);
        add(n + ".pw.b", mid);
      }
      add(n + ".dw.w", (int64_t)mid * b.k * b.k);
      add(n + ".dw.b", mid);
      add(n + ".pwl.w", (int64_t)b.cout * mid);
      add(n + ".pwl.b", b.cout);
    }
    add("neck.w", (int64_t)kFeatC * kBackboneC);
    add("neck.b", kFeatC);
    for (const char* br : {"cls", "reg"}) {
      std::string n = br;
      add(n + "_encode.dw.w", kFeatC * 9);
      add(n + "_encode.pw.w", (int64_t)kFeatC * kFeatC);
      add(n + "_encode.pw.b", kFeatC);
      add(n + "_dw.dw.w", kCatC * 9);
      add(n + "_dw.pw.w", (int64_t)kFeatC * kCatC);
      add(n + "_dw.pw.b", kFeatC);
    }
    for (const char* tw : {"bbox_tower", "cls_tower"}) {
      for (int i = 0; i < 2; ++i) {
        std::string n = std::string(tw) + "." + std::to_string(i);
        add(n + ".dw.w", kFeatC * 9);
        add(n + ".pw.w", (int64_t)kFeatC * kFeatC);
        add(n + ".pw.b", kFeatC);
      }
    }
    add("bbox_pred.dw.w", kFeatC * 9);
    add("bbox_pred.pw.w", 4 * kFeatC);
    add("bbox_pred.pw.b", 4);
    add("cls_pred.dw.w", kFeatC * 9);
    add("cls_pred.pw.w", kFeatC);
    add("cls_pred.pw.b", 1);
    return t;
  }();
  return table;
}

}  // namespace fear
