// The codec topology (neural_speech_coding_module.py:152-260) as a program of plane-engine layers (plane.cuh): every
// activation between the 1-channel input and the 1-channel code / output stays an fp16 plane image in HBM, written by
// one layer's epilogue in exactly the form the next layer's bulk copies stage.  Host side only; included by codec.cu.
//
// Covered: resnet_type 'bottleneck' AND 'gln' (the reference's shipped default, constants.py:14: gated blocks with k15 gates and the
// separable up-conv), one stride-2 stage, narrow = 20, k = 9, dilations <= 2.  Anything else keeps the layer-by-layer engines
// (walker.cuh).
#pragma once
#include <stdlib.h>

#include <vector>

#include "plane.cuh"
#include "walker.cuh"

namespace nsc {
namespace {

bool plane_codec_supported(const nsc_codec_cfg& c) {
  static const bool off = [] { const char* e = getenv("NSC_PLANE"); return e && e[0] == '0'; }();
  if (off) return false;
  if (c.precision != 1 && c.precision != 2) return false;
  if ((c.resnet_type != 0 && c.resnet_type != 1) || c.n_strides != 1 || c.strides[0] != 2) return false;
  if (c.narrow != 20 || c.k_plain != 9 || c.k_dilated != 9) return false;
  if (c.wide < 66 || c.wide > 128 || (c.wide & 1)) return false;   // wide and wide/2 both use unpacked (>32 channel) images
  for (int i = 0; i < c.n_blocks; ++i)
    if (c.dilations[i] < 1 || c.dilations[i] > 2) return false;
  return true;
}

// frames per pass of the plane path: a whole number of frames per SM for every kernel
int64_t plane_chunk_frames() {
  static const int64_t v = [] {
    const char* e = getenv("NSC_PLANE_CHUNK");
    const long long n = e ? atoll(e) : 0;
    return (int64_t)(n >= 1 && n <= 65536 ? n : 14LL * sm_count());
  }();
  return v;
}

// PB_N0D / PB_M0D: de-interleaved twins of PB_N0 / PB_M0 ('gln': input of a dilation-2 gate conv); PB_H2: third half-length wide
// buffer ('gln': the depthwise half of the separable up-conv)
enum PBuf { PB_W0 = 0, PB_W1, PB_WD, PB_H0, PB_H1, PB_N0, PB_N1, PB_M0, PB_M1, PB_C0, PB_C1, PB_N0D, PB_M0D, PB_H2, PB_COUNT };

struct PlaneCodecPlan {
  int planes = 2;
  int Lc = 0;
  PlaneTensor buf[PB_COUNT];             // bases relative to the activation region (filled by bind)
  int64_t buf_off[PB_COUNT];
  int64_t act_bytes_per_frame = 0;
  std::vector<PlaneConv> enc, dec;
  std::vector<int> enc_layer, dec_layer;  // index into the codec's layer table (parameter offsets; a gated layer also owns the next entry)
  std::vector<int> enc_sep, dec_sep;      // 0 plain conv; 1 / 2: depthwise / pointwise half of a separable layer (one table entry)
  struct Io { int in, out, res; };        // activation buffers of a layer (PBuf ids, -1 = the 1-channel vector at the edge)
  std::vector<Io> enc_io, dec_io;
  std::vector<int64_t> w_off;             // packed-weight offset of every layer (enc then dec)
  int64_t wpack_bytes = 0;
  std::vector<int> enc_block, dec_block;  // per layer: 1 = first conv of a bottleneck block whose three convs can run as ONE fused launch
  uint32_t* flags = nullptr;              // frame flags of the fused blocks (plane_block_flag_words(flag_frames) words per fused block)
  int64_t flag_frames = 0;                // frames per pass the flag words were sized for
  int n_fused = 0;
};

// Lays the codec out as plane layers.  Tensors carry geometry only (base = nullptr) until plane_bind() attaches a workspace.
PlaneCodecPlan make_plane_plan(const nsc_codec_cfg& c) {
  PlaneCodecPlan pl;
  pl.planes = c.precision == 1 ? 2 : 1;
  const int P = pl.planes, L = kFrameLen, H = kFrameLen / 2, W = c.wide, Nn = c.narrow, Wd = c.wide / 2;
  pl.Lc = H;
  pl.buf[PB_W0] = make_plane_tensor(nullptr, L, W, P, 0);
  pl.buf[PB_W1] = pl.buf[PB_W0];
  pl.buf[PB_WD] = make_plane_tensor(nullptr, L, W, P, 1);
  pl.buf[PB_H0] = make_plane_tensor(nullptr, H, W, P, 0);
  pl.buf[PB_H1] = pl.buf[PB_H0];
  pl.buf[PB_N0] = make_plane_tensor(nullptr, L, Nn, P, 0);
  pl.buf[PB_N1] = pl.buf[PB_N0];
  pl.buf[PB_M0] = make_plane_tensor(nullptr, H, Nn, P, 0);
  pl.buf[PB_M1] = pl.buf[PB_M0];
  pl.buf[PB_C0] = make_plane_tensor(nullptr, L, Wd, P, 0);
  pl.buf[PB_C1] = pl.buf[PB_C0];
  const bool gln = c.resnet_type == 1;
  // (buffers only the gated topology uses stay empty for 'bottleneck')
  if (gln) {
    pl.buf[PB_N0D] = make_plane_tensor(nullptr, L, Nn, P, 1);
    pl.buf[PB_M0D] = make_plane_tensor(nullptr, H, Nn, P, 1);
    pl.buf[PB_H2] = pl.buf[PB_H0];
  }
  int64_t off = 0;
  for (int i = 0; i < PB_COUNT; ++i) { pl.buf_off[i] = off; off += pl.buf[i].frame_bytes; }
  pl.act_bytes_per_frame = off;

  // narrow -> narrow convs (20 -> 20): taps-in-N by default; the tap-shift kernel is within 5 % here (measured 5.6 vs 5.3 ms per
  // step: nine N = 32 MMAs per K step issue-bound vs the tap-sum epilogue) and can be selected for experiments
  const int kNarrowKind = plane_narrow_kind();
  int layer = 0;   // creation-order layer index (same walk as Walker::encoder / decoder)
  // sep: 0 plain conv (one table entry), 1 depthwise half (keeps the entry for the pointwise half that follows), 2 pointwise half
  auto add = [&](std::vector<PlaneConv>& v, std::vector<int>& vl, int kind, int Lin, int Cin, int Cout, int K, int dil, int stride,
                 int act, int in, int out, int res, int res_mode, int post, int shuffle, int sep = 0) -> PlaneConv& {
    (&v == &pl.enc ? pl.enc_io : pl.dec_io).push_back(PlaneCodecPlan::Io{in, out, res});
    (&v == &pl.enc ? pl.enc_sep : pl.dec_sep).push_back(sep);
    PlaneConv pc;
    pc.kind = kind; pc.Lin = Lin; pc.Cin = Cin; pc.Cout = Cout; pc.K = K; pc.dil = dil; pc.stride = stride;
    pc.act = act; pc.post_act = post; pc.res_mode = res_mode; pc.shuffle = shuffle; pc.planes = P;
    if (in >= 0) pc.in = pl.buf[in];
    if (out >= 0) pc.out = pl.buf[out];
    if (res >= 0) pc.res = pl.buf[res];
    v.push_back(pc);
    vl.push_back(layer);
    if (sep != 1) ++layer;
    return v.back();
  };
  // the two k15 gate convs of a gated block as ONE layer (plane.cuh): input n0 (or its de-interleaved twin n0d for dilation 2,
  // convolved as two independent half-length frames with dilation 1), output the plain packed image n1.  Owns TWO table entries.
  auto add_gates = [&](std::vector<PlaneConv>& v, std::vector<int>& vl, int Ls, int d, int n0, int n0d, int n1) {
    if (d == 2) {
      PlaneConv& g = add(v, vl, PK_X, Ls / 2, Nn, 2 * Nn, 15, 1, 1, NSC_ACT_NONE, n0d, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
      g.glu = 1; g.ileave = 1; g.bmul = 2;
      g.in.deint = 0; g.in.rows = Ls / 2; g.in.frame_bytes = pl.buf[n0d].frame_bytes / 2;   // view: one sub-image = one frame
    } else {
      PlaneConv& g = add(v, vl, PK_X, Ls, Nn, 2 * Nn, 15, d, 1, NSC_ACT_NONE, n0, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
      g.glu = 1;
    }
    ++layer;   // the tanh gate's entry
  };
  // one stack of bottleneck blocks (nscm.py:183-217) on `cur`; returns the buffer that holds the result
  auto stack = [&](std::vector<PlaneConv>& v, std::vector<int>& vl, int Ls, int Cw, int cur, int b0, int b1, int n0, int n1, int last_out,
                   bool vec_in) {
    for (int i = 0; i < c.n_blocks; ++i) {
      const bool flat = i == c.n_blocks - 1;
      int out = (cur == b0) ? b1 : b0;
      if (flat && last_out >= 0) out = last_out;
      const int post = flat ? NSC_ACT_NONE : NSC_ACT_LRELU;
      if (gln) {   // gated_bottleneck (nn_core_operator.py:82-112): k1 -> [k15 gate * tanh(k15 gate)] -> k9 + residual
        const int d = c.dilations[i];
        const int n0d = n0 == PB_N0 ? PB_N0D : PB_M0D;
        const int k1_out = d == 2 ? n0d : n0;
        const bool vec = vec_in && i == 0;
        add(v, vl, vec ? PK_GEN : PK_X, Ls, vec ? 1 : Cw, Nn, 1, 1, 1, NSC_ACT_LRELU, vec ? -1 : cur, k1_out, -1, RES_NONE, NSC_ACT_NONE, 1);
        add_gates(v, vl, Ls, d, n0, n0d, n1);
        add(v, vl, PK_X, Ls, Nn, Cw, c.k_plain, 1, 1, NSC_ACT_NONE, n1, out, vec ? -1 : cur, vec ? RES_ADD_BCAST : RES_ADD, post, 1);
        cur = out;
        continue;
      }
      if (vec_in && i == 0) {
        add(v, vl, PK_GEN, Ls, 1, Nn, c.k_plain, 1, 1, NSC_ACT_LRELU, -1, n0, -1, RES_NONE, NSC_ACT_NONE, 1);
        add(v, vl, kNarrowKind, Ls, Nn, Nn, c.k_dilated, c.dilations[i], 1, NSC_ACT_LRELU, n0, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
        add(v, vl, PK_X, Ls, Nn, Cw, c.k_plain, 1, 1, NSC_ACT_NONE, n1, out, -1, RES_ADD_BCAST, post, 1);
      } else {
        (&v == &pl.enc ? pl.enc_block : pl.dec_block).resize(v.size() + 1, 0);
        (&v == &pl.enc ? pl.enc_block : pl.dec_block)[v.size()] = 1;
        ++pl.n_fused;
        add(v, vl, PK_T, Ls, Cw, Nn, c.k_plain, 1, 1, NSC_ACT_LRELU, cur, n0, -1, RES_NONE, NSC_ACT_NONE, 1);
        add(v, vl, kNarrowKind, Ls, Nn, Nn, c.k_dilated, c.dilations[i], 1, NSC_ACT_LRELU, n0, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
        add(v, vl, PK_X, Ls, Nn, Cw, c.k_plain, 1, 1, NSC_ACT_NONE, n1, out, cur, RES_ADD, post, 1);
      }
      cur = out;
    }
    return cur;
  };
  // encoder (nscm.py:219-237)
  add(pl.enc, pl.enc_layer, PK_GEN, L, 1, W, 55, 1, 1, NSC_ACT_LRELU, -1, PB_W0, -1, RES_NONE, NSC_ACT_NONE, 1);
  int cur = stack(pl.enc, pl.enc_layer, L, W, PB_W0, PB_W0, PB_W1, PB_N0, PB_N1, PB_WD, false);
  add(pl.enc, pl.enc_layer, PK_X, L, W, W, 9, 1, 2, NSC_ACT_LRELU, cur, PB_H0, -1, RES_NONE, NSC_ACT_NONE, 1);
  cur = stack(pl.enc, pl.enc_layer, H, W, PB_H0, PB_H0, PB_H1, PB_M0, PB_M1, -1, false);
  add(pl.enc, pl.enc_layer, PK_T, H, W, 1, 55, 1, 1, NSC_ACT_TANH, cur, -1, -1, RES_NONE, NSC_ACT_NONE, 1);
  // decoder (nscm.py:239-260)
  cur = stack(pl.dec, pl.dec_layer, H, W, PB_H1, PB_H0, PB_H1, PB_M0, PB_M1, -1, true);   // first block writes the buffer that is not `cur`
  if (gln) {   // separable up-conv (nscm.py:175-177): depthwise k9 -> pointwise + bias + leaky ReLU -> sub-pixel shuffle
    add(pl.dec, pl.dec_layer, PK_DW, H, W, W, 9, 1, 1, NSC_ACT_NONE, cur, PB_H2, -1, RES_NONE, NSC_ACT_NONE, 1, 1);
    add(pl.dec, pl.dec_layer, PK_X, H, W, W, 1, 1, 1, NSC_ACT_LRELU, PB_H2, PB_C0, -1, RES_NONE, NSC_ACT_NONE, 2, 2);
  } else {
    add(pl.dec, pl.dec_layer, PK_X, H, W, W, 9, 1, 1, NSC_ACT_LRELU, cur, PB_C0, -1, RES_NONE, NSC_ACT_NONE, 2);
  }
  cur = stack(pl.dec, pl.dec_layer, L, Wd, PB_C0, PB_C0, PB_C1, PB_N0, PB_N1, -1, false);
  add(pl.dec, pl.dec_layer, PK_T, L, Wd, 1, 55, 1, 1, NSC_ACT_NONE, cur, -1, -1, RES_NONE, NSC_ACT_NONE, 1);

  int64_t woff = 0;
  auto size_w = [&](std::vector<PlaneConv>& v) {
    for (auto& pc : v) {
      PlaneConv t = pc;
      const float dummy = 0.f;
      if (t.Cin == 1) t.xvec = &dummy;
      if (t.res_mode == RES_ADD_BCAST) t.resvec = &dummy;
      pl.w_off.push_back(woff);
      const int64_t b = plane_wpack_bytes(t);
      woff += align_up(b < 0 ? 0 : b, 1024);
      if (b < 0) pl.wpack_bytes = -1;
    }
  };
  size_w(pl.enc);
  size_w(pl.dec);
  if (pl.wpack_bytes == 0) pl.wpack_bytes = woff;
  pl.enc_block.resize(pl.enc.size(), 0);
  pl.dec_block.resize(pl.dec.size(), 0);
  return pl;
}

// frame flags of the fused blocks of one codec pass
int64_t plane_codec_flag_bytes(const PlaneCodecPlan& pl, int64_t Bc) {
  return align_up((int64_t)pl.n_fused * plane_block_flag_words(Bc) * (int64_t)sizeof(uint32_t), 1024);
}

int64_t plane_codec_act_bytes(const PlaneCodecPlan& pl, int64_t Bc) {
  int64_t total = 0;
  for (int i = 0; i < PB_COUNT; ++i) total += align_up(pl.buf[i].frame_bytes * Bc, 1024);
  return total + 1024;
}

// Resolves the buffer ids to addresses inside `act` (sized for Bc frames), attaches parameters and packed weights.
void plane_bind(PlaneCodecPlan& pl, const CodecLayout& lay, const float* params, void* act, int64_t Bc, void* wpack) {
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(act) + 1023) & ~(uintptr_t)1023);
  uint8_t* addr[PB_COUNT];
  for (int i = 0; i < PB_COUNT; ++i) { addr[i] = base; base += align_up(pl.buf[i].frame_bytes * Bc, 1024); }
  size_t li = 0;
  auto fix = [&](std::vector<PlaneConv>& v, std::vector<int>& vl, const std::vector<PlaneCodecPlan::Io>& io, const std::vector<int>& sep) {
    for (size_t i = 0; i < v.size(); ++i, ++li) {
      PlaneConv& pc = v[i];
      pc.in.base = io[i].in >= 0 ? addr[io[i].in] : nullptr;
      pc.out.base = io[i].out >= 0 ? addr[io[i].out] : nullptr;
      pc.res.base = io[i].res >= 0 ? addr[io[i].res] : nullptr;
      const LayerInfo& info = lay.layers[vl[i]];
      pc.w = params + info.off;
      pc.bias = pc.w + (int64_t)info.k * info.cin * info.cout;
      if (sep[i] == 1) pc.bias = nullptr;                                   // depthwise half: (k, cin) taps, no bias
      if (sep[i] == 2) {                                                    // pointwise half: (1, cin, cout) after the taps, then the bias
        pc.w = params + info.off + (int64_t)info.k * info.cin;
        pc.bias = pc.w + (int64_t)info.cin * info.cout;
      }
      if (pc.glu) {                                                         // the tanh gate is the next table entry
        const LayerInfo& g2 = lay.layers[vl[i] + 1];
        pc.w2 = params + g2.off;
        pc.bias2 = pc.w2 + (int64_t)g2.k * g2.cin * g2.cout;
      }
      pc.wpack = static_cast<uint8_t*>(wpack) + pl.w_off[li];
    }
  };
  fix(pl.enc, pl.enc_layer, pl.enc_io, pl.enc_sep);
  fix(pl.dec, pl.dec_layer, pl.dec_io, pl.dec_sep);
}

int plane_codec_pack(PlaneCodecPlan& pl, cudaStream_t st) {
  for (auto* v : {&pl.enc, &pl.dec})
    for (auto& pc : *v) {
      PlaneConv t = pc;
      const float dummy = 0.f;
      if (t.Cin == 1) t.xvec = &dummy;
      if (t.res_mode == RES_ADD_BCAST) t.resvec = &dummy;
      NSC_TRY(plane_pack_weights(t, st));
    }
  return NSC_OK;
}

// Runs layers [i, i + 3) as one fused block launch when the fused kernel covers them; `*fused_no` counts the codec's fused blocks
// (each has its own flag words, cleared once per pass by plane_clear_flags).
bool plane_try_block(PlaneCodecPlan& pl, std::vector<PlaneConv>& v, const std::vector<int>& is_block, size_t i, int64_t nb, int* fused_no,
                     cudaStream_t st, int* rc) {
  if (!plane_block_default_on() || !is_block[i] || i + 2 >= v.size() || pl.flags == nullptr) return false;
  const int no = (*fused_no)++;
  PlaneBlock b;
  b.c1 = v[i]; b.c2 = v[i + 1]; b.c3 = v[i + 2];
  b.c1.B = b.c2.B = b.c3.B = nb;
  int ring = plane_block_ring_frames();
  while (ring > 2 && ring > nb) ring /= 2;
  b.ring = ring;
  b.flags = pl.flags + (int64_t)no * plane_block_flag_words(pl.flag_frames);
  if (!plane_block_supported(b)) return false;
  *rc = plane_block_launch(b, st);
  return true;
}

int plane_clear_flags(PlaneCodecPlan& pl, cudaStream_t st) {
  if (!plane_block_default_on() || pl.flags == nullptr || pl.n_fused == 0) return NSC_OK;
  NSC_CUDA_OK(cudaMemsetAsync(pl.flags, 0, (size_t)pl.n_fused * plane_block_flag_words(pl.flag_frames) * sizeof(uint32_t), st));
  return NSC_OK;
}

// encoder: x (nb, 512) -> fcode (nb, Lc);  decoder: code (nb, Lc) -> out (nb, 512)
int plane_run_encoder(PlaneCodecPlan& pl, const float* x, int64_t nb, float* fcode, cudaStream_t st) {
  NSC_TRY(plane_clear_flags(pl, st));
  int fused_no = 0;
  for (size_t i = 0; i < pl.enc.size(); ++i) {
    int rc = NSC_OK;
    if (plane_try_block(pl, pl.enc, pl.enc_block, i, nb, &fused_no, st, &rc)) { NSC_TRY(rc); i += 2; continue; }
    PlaneConv t = pl.enc[i];
    t.B = nb * t.bmul;
    if (t.Cin == 1) t.xvec = x;
    if (t.Cout == 1) t.yvec = fcode;
    NSC_TRY(plane_launch(t, st));
  }
  return NSC_OK;
}

int plane_run_decoder(PlaneCodecPlan& pl, const float* code, int64_t nb, float* out, cudaStream_t st) {
  NSC_TRY(plane_clear_flags(pl, st));
  int fused_no = 0;
  for (size_t i = 0; i < pl.dec.size(); ++i) {
    int rc = NSC_OK;
    if (plane_try_block(pl, pl.dec, pl.dec_block, i, nb, &fused_no, st, &rc)) { NSC_TRY(rc); i += 2; continue; }
    PlaneConv t = pl.dec[i];
    t.B = nb * t.bmul;
    if (t.Cin == 1) t.xvec = code;
    if (t.res_mode == RES_ADD_BCAST) t.resvec = code;
    if (t.Cout == 1) t.yvec = out;
    NSC_TRY(plane_launch(t, st));
  }
  return NSC_OK;
}

}  // namespace
}  // namespace nsc
