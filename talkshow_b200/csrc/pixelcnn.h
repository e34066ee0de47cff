// talkshow_b200 — incremental gated-PixelCNN sampler: execution plan shared by the packer
// (host), the CUDA executor (pixelcnn.cu) and the CPU plan interpreter used by the tests.
#pragma once
#include "kernels.h"

namespace ts {

constexpr int PIX_D = 256;      // hidden width (config/body_pixel.json: convert_to_6d=false)
constexpr int PIX_MB = 64;      // batch tile: every activation row is [channel][64 samples]
constexpr int PIX_SEG = PIX_D * PIX_MB;  // floats in one [256][64] activation segment
constexpr int PIX_THREADS = 256;
constexpr int PIX_MAXROWS = 16; // weight rows one CTA handles per stage
constexpr int PIX_WBUF = 18560; // floats per weight staging buffer (74.2 KB), two buffers
constexpr int PIX_NCODE = 2048;

enum PixEpi {
  EPI_IDLE = 0,
  EPI_VERT0 = 1,   // layer-0 vertical stack (mask A), one output column
  EPI_VERT = 2,    // layers >= 1 vertical stack, one output column
  EPI_V2H = 3,     // vert_to_horiz of one layer, both columns (2 passes)
  EPI_FUSEV = 4,   // fusion_v (x_v half), both columns (2 passes)
  EPI_HGATE = 5,   // horizontal stack + gate, one column
  EPI_HRES = 6,    // horiz_resid (+ residual)
  EPI_FUSEH = 7,   // fusion_h (x_h half)
  EPI_OUT1 = 8,    // output_conv.0 + ReLU
  EPI_OUT2 = 9,    // output_conv.2 -> logits
  EPI_SAMPLE = 10, // softmax + categorical draw + embedding gather (K = 1: also emits layer-0 G of column 1)
  // fused plan (52 stages): linear stages folded into the gate matmul that consumes them
  EPI_HRESF = 11,  // (fusion_h . horiz_resid_0) G_0 + audio term -> x_h[1]
  EPI_HGATE2 = 12, // gate of layer l on (horiz_stack_l . horiz_resid_{l-1}) G_{l-1} + horiz_stack_l x_h[l-1] (+ column-0 tap)
  EPI_OUT1F = 13,  // output_conv.0 on (W horiz_resid_{L-1}) G_{L-1} + W x_h[L-1], ReLU
  // schedule 2: vert_to_horiz leaves the vertical stages and runs one column at a time beside the horizontal stage
  // that precedes its consumer
  EPI_V2H1 = 14    // vert_to_horiz of one layer, column t.col only
};

struct PixTask {  // one CTA's work in one stage (8 ints)
  int epi, layer, col, row0, nrows, wofs, K, rpad;
};

struct PixLayout {  // arena offsets in floats
  int E, XV1P, XV, HV, V2H, G, XHP, XH, Y, LOG, CLS, total;
};

struct PixelPlan {
  int L = 0, ncta = 0, nstages = 0, nclasses = 4;
  int D = PIX_D;               // hidden width of the checkpoint (the grid-wide executor is built for 256 only)
  bool has_v1 = false;         // stage table / blob of the grid-wide executor built (D == 256)
  int cl = 1;                  // CTAs per work unit
  bool fused = false;          // 52-stage plan (EPI_HRESF / EPI_HGATE2 / EPI_OUT1F), else the plain 84-stage plan
  int sched = 1;               // 0 plain, 1 fused, 2 fused + vert_to_horiz moved into the horizontal pass (EPI_V2H1)
  PixLayout lay;
  std::vector<PixTask> table;  // [nstages][ncta]
  std::vector<float> blob;     // packed per-task weights: [K][rpad] then bias [rpad]
  int64_t row_bytes = 0;       // algorithmic weight bytes per latent row (dense fp32 weights, once)
  int64_t staged_row_bytes = 0;  // bytes the kernel actually stages per row (blob incl. padding/duplication)
  // device copies
  PixTask* d_table = nullptr;
  float* d_blob = nullptr;
  float* d_emb = nullptr;      // [2048][256]
  float* d_cls = nullptr;      // [L][4][512]
  float* d_arena = nullptr;
  unsigned* d_barrier = nullptr;
  Layer emb_aud, fuse_v_a, fuse_h_a;  // audio terms (precomputed per call for all rows)
  void* p3 = nullptr;          // Plan3 of the cluster-resident executor (pixelcnn3.inc)
  unsigned long long* d_trace = nullptr;  // stage trace buffer (ts_pixelcnn_trace)
  int trace_row = -1;
  bool timing = false, pending = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t timed_rows = 0, timed_launches = 0;
};

// aud: AudioEncoder output, channel-last [B, T0+T, 256].
void pixelcnn_generate_act(ts_engine* e, const Act3& aud, const int64_t* label, const float* noise, int64_t* idx_out,
                           float* logits_out, int B, int T, const int64_t* pre, int T0, cudaStream_t s,
                           bool logits_all = false);
// idx [B,T,2] -> idx_c [2][B][T]; optionally copies idx to codes_out
void split_codes(ts_engine* e, const int64_t* idx, int64_t* idx_c, int B, int T, int64_t* codes_out, cudaStream_t s);
void pixel_destroy(ts_engine* e);
void face_destroy(ts_engine* e);
void mfcc_destroy(ts_engine* e);
void smplx_destroy(ts_engine* e);
void nccl_destroy(ts_engine* e);

}  // namespace ts
