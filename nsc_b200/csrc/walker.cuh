// The codec topology (neural_speech_coding_module.py:152-260) written ONCE and shared by inference (codec.cu) and
// training (train.cu).  struct Walker walks encoder / decoder and, per conv layer, either
//   * records the layer in the parameter layout (dry walk  -> flat parameter image, TF creation order),
//   * launches it into rotating activation buffers (inference), or
//   * launches it into a bump arena and records it on a tape (training: every activation is kept for backward).
#pragma once
#include <stdlib.h>

#include <vector>

#include "conv.cuh"

namespace nsc {
namespace {

constexpr int kFrameLen = NSC_FRAME_LENGTH;
// frames per internal pass (bounds the activation workspace); NSC_CHUNK_FRAMES overrides it for experiments
static int64_t chunk_frames() {
  static const int64_t v = [] {
    const char* e = getenv("NSC_CHUNK_FRAMES");
    const long long n = e ? atoll(e) : 0;
    return (int64_t)(n >= 16 && n <= 65536 ? n : 2048);
  }();
  return v;
}
#define kChunkFrames (::nsc::chunk_frames())

struct LayerInfo {
  int k, cin, cout, separable;
  int64_t off;  // first float of the layer inside the flat parameter image
};

int validate_cfg(const nsc_codec_cfg* c) {
  NSC_CHECK_ARG(c != nullptr, "codec cfg is null");
  NSC_CHECK_ARG(c->wide >= 1 && c->narrow >= 1 && c->k_plain >= 1 && c->k_dilated >= 1, "codec cfg: bad sizes");
  NSC_CHECK_ARG(c->n_blocks >= 1 && c->n_blocks <= NSC_MAX_BLOCKS, "codec cfg: n_blocks=%d", c->n_blocks);
  NSC_CHECK_ARG(c->n_strides >= 1 && c->n_strides <= NSC_MAX_STRIDES, "codec cfg: n_strides=%d", c->n_strides);
  NSC_CHECK_ARG(c->resnet_type == 0 || c->resnet_type == 1, "codec cfg: resnet_type=%d", c->resnet_type);
  NSC_CHECK_ARG(c->num_bins >= 1 && c->num_bins <= 256, "codec cfg: num_bins=%d", c->num_bins);
  NSC_CHECK_ARG(c->precision >= 0 && c->precision <= 2, "codec cfg: precision=%d", c->precision);
  int L = kFrameLen, C = c->wide;
  for (int i = 0; i < c->n_strides; ++i) {
    NSC_CHECK_ARG(c->strides[i] >= 1 && L % c->strides[i] == 0, "codec cfg: stride %d does not divide %d", c->strides[i], L);
    L /= c->strides[i];
  }
  for (int i = 0; i < c->n_strides; ++i) {
    NSC_CHECK_ARG(C % c->strides[i] == 0, "codec cfg: decoder channels %d not divisible by stride %d", C, c->strides[i]);
    C /= c->strides[i];
  }
  for (int i = 0; i < c->n_blocks; ++i) NSC_CHECK_ARG(c->dilations[i] >= 1, "codec cfg: dilation[%d]=%d", i, c->dilations[i]);
  return NSC_OK;
}

int code_length(const nsc_codec_cfg& c) {
  int L = kFrameLen;
  for (int i = 0; i < c.n_strides; ++i) L /= c.strides[i];
  return L;
}

// ---- workspace carving -------------------------------------------------------------------------
struct Carver {
  char* base;
  int64_t used = 0, cap;
  Carver(void* p, int64_t c) : base(static_cast<char*>(p)), cap(c) {}
  float* take(int64_t floats) {
    const int64_t bytes = align_up(floats * (int64_t)sizeof(float), 256);
    float* r = reinterpret_cast<float*>(base + used);
    used += bytes;
    return r;
  }
};



// One layer as executed (training tape)
enum TapeOp { OP_CONV = 0, OP_MUL = 1, OP_DEPTHWISE = 2, OP_POINTWISE = 3 };
struct ConvRec {
  int layer;                 // index into the layer table (parameter offsets)
  const float* x;
  float* y;
  const float* res;
  int Lin, Cin, Cout, K, dil, stride, act, res_mode, post_act, shuffle;
  // OP_CONV: a conv layer.  OP_MUL: y = x * res elementwise (the gate product of a gated block, kept apart from the tanh gate's conv
  // in training so that both factors are on the tape).  OP_DEPTHWISE / OP_POINTWISE: the two halves of a separable conv (ONE table
  // entry: taps (K, Cin), then the pointwise (Cin, Cout) kernel, then the bias); for OP_POINTWISE K is 1 and `kdw` the depthwise K.
  int op = OP_CONV;
  int kdw = 0;
};

struct Walker {
  nsc_codec_cfg cfg;
  bool dry = true;
  const float* params = nullptr;
  int64_t B = 0;
  cudaStream_t st = nullptr;
  std::vector<LayerInfo> layers;
  size_t cursor = 0;
  int64_t off = 0;
  float* wide[3] = {nullptr, nullptr, nullptr};
  float* nar[3] = {nullptr, nullptr, nullptr};
  void* wpack = nullptr;   // scratch for the tensor engine's packed weights
  Carver* arena = nullptr;             // training: every layer output is bump-allocated and kept
  bool launch = true;                  // false: replay the allocation / tape only (backward pass re-walk)
  std::vector<ConvRec>* tape = nullptr;
  int rc = NSC_OK;

  const LayerInfo* next_layer(int k, int cin, int cout, int separable) {
    if (dry) {
      LayerInfo li{k, cin, cout, separable, off};
      off += separable ? ((int64_t)k * cin + (int64_t)cin * cout + cout) : ((int64_t)k * cin * cout + cout);
      layers.push_back(li);
      return nullptr;
    }
    const LayerInfo* li = &layers[cursor++];
    if (li->k != k || li->cin != cin || li->cout != cout || li->separable != separable) {
      set_error("internal: layer table mismatch at %zu", cursor - 1);
      rc = NSC_E_INVALID;
    }
    return li;
  }

  void conv(const float* x, float** yslot, int Lin, int Cin, int Cout, int K, int dil, int stride, int act,
            const float* res = nullptr, int res_mode = RES_NONE, int post_act = NSC_ACT_NONE, int shuffle = 1) {
    const LayerInfo* li = next_layer(K, Cin, Cout, 0);
    if (dry || rc != NSC_OK) return;
    if (arena) {
      int Lo, pl;
      same_padding(Lin, K, dil, stride, &Lo, &pl);
      *yslot = arena->take(B * (int64_t)Cout * Lo);
      if (arena->used > arena->cap) { set_error("training arena overflow"); rc = NSC_E_WORKSPACE; return; }
    }
    float* y = *yslot;
    if (tape) tape->push_back(ConvRec{(int)cursor - 1, x, y, res, Lin, Cin, Cout, K, dil, stride, act, res_mode, post_act, shuffle});
    ConvArgs a;
    a.x = x; a.y = y;
    a.w = params + li->off;
    a.bias = a.w + (int64_t)K * Cin * Cout;
    a.res = res; a.res_mode = res_mode; a.post_act = post_act; a.shuffle = shuffle;
    a.B = B; a.Lin = Lin; a.Cin = Cin; a.Cout = Cout; a.K = K; a.dil = dil; a.stride = stride; a.act = act;
    if (!launch) return;
    if (cfg.precision > 0 && wpack != nullptr && tc_conv_supported(a)) rc = launch_conv_tc(a, cfg.precision, wpack, st);
    else rc = launch_conv(a, st);   // 1-channel stem / heads and odd shapes stay on the FFMA engine
  }

  // training only: y = a * b elementwise (n floats), both factors kept on the tape
  void mul(const float* a, const float* b, float** yslot, int L, int C) {
    if (dry || rc != NSC_OK) return;
    *yslot = arena->take(B * (int64_t)C * L);
    if (arena->used > arena->cap) { set_error("training arena overflow"); rc = NSC_E_WORKSPACE; return; }
    if (tape) {
      ConvRec r{-1, a, *yslot, b, L, C, C, 0, 1, 1, NSC_ACT_NONE, RES_MUL, NSC_ACT_NONE, 1};
      r.op = OP_MUL;
      tape->push_back(r);
    }
    if (launch) rc = launch_mul(*yslot, a, b, B * (int64_t)C * L, st);
  }

  // Keras SeparableConv1D: depthwise (k, cin, 1) -> pointwise (1, cin, cout) + bias + activation
  void sepconv(const float* x, float* tmp, float** yslot, int Lin, int Cin, int Cout, int K, int act, int shuffle) {
    const LayerInfo* li = next_layer(K, Cin, Cout, 1);
    if (dry || rc != NSC_OK) return;
    if (arena) {   // training: both halves' outputs are kept
      tmp = arena->take(B * (int64_t)Cin * Lin);
      *yslot = arena->take(B * (int64_t)Cout * Lin);
      if (arena->used > arena->cap) { set_error("training arena overflow"); rc = NSC_E_WORKSPACE; return; }
      if (tape) {
        ConvRec d{(int)cursor - 1, x, tmp, nullptr, Lin, Cin, Cin, K, 1, 1, NSC_ACT_NONE, RES_NONE, NSC_ACT_NONE, 1};
        d.op = OP_DEPTHWISE;
        tape->push_back(d);
        ConvRec q{(int)cursor - 1, tmp, *yslot, nullptr, Lin, Cin, Cout, 1, 1, 1, act, RES_NONE, NSC_ACT_NONE, shuffle};
        q.op = OP_POINTWISE; q.kdw = K;
        tape->push_back(q);
      }
      if (!launch) return;
    }
    float* y = *yslot;
    const float* dw = params + li->off;
    const float* pw = dw + (int64_t)K * Cin;
    const float* bias = pw + (int64_t)Cin * Cout;
    rc = launch_depthwise(x, dw, tmp, B, Lin, Cin, K, 1, 1, 0, 0, st);
    if (rc != NSC_OK) return;
    ConvArgs a;
    a.x = tmp; a.y = y; a.w = pw; a.bias = bias;
    a.B = B; a.Lin = Lin; a.Cin = Cin; a.Cout = Cout; a.K = 1; a.act = act; a.shuffle = shuffle;
    rc = launch_conv(a, st);
  }

  // _stack_bottleneck_blocks (nscm.py:183-217).  `in` lives in wide[in_idx] or is external (in_idx = -1).
  int stack(const float* in, int in_idx, int& C, int L) {
    const int wide_layer = (C == 1) ? cfg.wide : C;   // nscm.py:189-192 (is_post_up_samling is always False)
    const float* cur = in;
    int cur_idx = in_idx;
    for (int i = 0; i < cfg.n_blocks; ++i) {
      const bool flat = (i == cfg.n_blocks - 1);     // `flag`, nscm.py:196
      const int out_idx = cur_idx < 0 ? 0 : (cur_idx + 1) % 3;
      const int d = cfg.dilations[i];
      const int rmode = (C == wide_layer) ? RES_ADD : RES_ADD_BCAST;  // 1-channel input broadcasts (:77)
      const int post = flat ? NSC_ACT_NONE : NSC_ACT_LRELU;
      if (cfg.resnet_type == 0) {   // the_bottleneck, nn_core_operator.py:57-79
        conv(cur, &nar[0], L, C, cfg.narrow, cfg.k_plain, 1, 1, NSC_ACT_LRELU);
        conv(nar[0], &nar[1], L, cfg.narrow, cfg.narrow, cfg.k_dilated, d, 1, NSC_ACT_LRELU);
        conv(nar[1], &wide[out_idx], L, cfg.narrow, wide_layer, cfg.k_plain, 1, 1, NSC_ACT_NONE, cur, rmode, post);
      } else {                      // gated_bottleneck, nn_core_operator.py:82-112 (gate kernel 15 hard-coded)
        conv(cur, &nar[0], L, C, cfg.narrow, 1, 1, 1, NSC_ACT_LRELU);
        conv(nar[0], &nar[1], L, cfg.narrow, cfg.narrow, 15, d, 1, NSC_ACT_NONE);
        if (arena) {   // training: the tanh gate's output and the product are separate tape entries
          float* gate = nullptr;
          conv(nar[0], &gate, L, cfg.narrow, cfg.narrow, 15, d, 1, NSC_ACT_TANH);
          mul(nar[1], gate, &nar[2], L, cfg.narrow);
        } else {
          conv(nar[0], &nar[2], L, cfg.narrow, cfg.narrow, 15, d, 1, NSC_ACT_TANH, nar[1], RES_MUL);
        }
        conv(nar[2], &wide[out_idx], L, cfg.narrow, wide_layer, cfg.k_plain, 1, 1, NSC_ACT_NONE, cur, rmode, post);
      }
      cur = wide[out_idx];
      cur_idx = out_idx;
      C = wide_layer;
    }
    return cur_idx;
  }

  // _the_encoder_in_each_module (nscm.py:219-237): x (B,512) -> floating code (B,Lc), tanh
  void encoder(const float* x, float** fcode) {
    int L = kFrameLen, C = cfg.wide;
    conv(x, &wide[0], L, 1, cfg.wide, 55, 1, 1, NSC_ACT_LRELU);
    int cur = 0;
    for (int s = 0; s < cfg.n_strides; ++s) {
      cur = stack(wide[cur], cur, C, L);
      const int o = (cur + 1) % 3;
      conv(wide[cur], &wide[o], L, C, cfg.wide, 9, 1, cfg.strides[s], NSC_ACT_LRELU);   // _down_sampling_mod :152-156
      cur = o;
      C = cfg.wide;
      L = (L + cfg.strides[s] - 1) / cfg.strides[s];
    }
    cur = stack(wide[cur], cur, C, L);
    conv(wide[cur], fcode, L, C, 1, 55, 1, 1, NSC_ACT_TANH);
  }

  // _the_decoder_in_each_module (nscm.py:239-260): code (B,Lc) -> out (B,512)
  void decoder(const float* code, float** out) {
    int L = code_length(cfg), C = 1;
    const float* in = code;
    int cur = -1;
    for (int s = 0; s < cfg.n_strides; ++s) {
      cur = stack(in, cur, C, L);
      const int o = (cur + 1) % 3, t = (cur + 2) % 3;
      const int r = cfg.strides[s];
      if (cfg.resnet_type == 0) conv(wide[cur], &wide[o], L, C, C, 9, 1, 1, NSC_ACT_LRELU, nullptr, RES_NONE, NSC_ACT_NONE, r);
      else sepconv(wide[cur], wide[t], &wide[o], L, C, C, 9, NSC_ACT_LRELU, r);   // _up_sampling_mod :169-181
      cur = o;
      in = wide[cur];
      C /= r;
      L *= r;
    }
    cur = stack(in, cur, C, L);
    conv(wide[cur], out, L, C, 1, 55, 1, 1, NSC_ACT_NONE);
  }
};

struct CodecLayout {
  std::vector<LayerInfo> layers;
  int64_t conv_floats = 0;   // alpha sits at conv_floats, bins at conv_floats + 1
  int code_len = 0;
};

CodecLayout make_layout(const nsc_codec_cfg& cfg) {
  Walker w;
  w.cfg = cfg;
  w.dry = true;
  float* none = nullptr;
  w.encoder(nullptr, &none);
  w.decoder(nullptr, &none);
  CodecLayout l;
  l.layers = w.layers;
  l.conv_floats = w.off;
  l.code_len = code_length(cfg);
  return l;
}

}  // namespace
}  // namespace nsc
