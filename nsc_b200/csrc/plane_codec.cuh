// The codec topology (neural_speech_coding_module.py:152-260) as a program of plane-engine layers (plane.cuh): every
// activation between the 1-channel input and the 1-channel code / output stays an fp16 plane image in HBM, written by
// one layer's epilogue in exactly the form the next layer's bulk copies stage.  Host side only; included by codec.cu.
//
// Covered: resnet_type 'bottleneck' AND 'gln' (the reference's shipped default, constants.py:14: gated blocks with k15 gates and the
// separable up-conv), one or two stride-2 stages (the_strides '2' / '4', cmrl.py:804), narrow = 20, k = 9, dilations <= 2.  Anything
// else keeps the layer-by-layer engines (walker.cuh).
#pragma once
#include <stdio.h>
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
  if ((c.resnet_type != 0 && c.resnet_type != 1) || c.n_strides < 1 || c.n_strides > 2) return false;
  for (int i = 0; i < c.n_strides; ++i)
    if (c.strides[i] != 2) return false;
  if (c.narrow != 20 || c.k_plain != 9 || c.k_dilated != 9) return false;
  if (c.wide < 66 || c.wide > 128 || (c.wide & 1)) return false;   // wide and wide/2 both use unpacked (>32 channel) images
  // two stages (the_strides = '4', cmrl.py:804): the decoder's last level has wide/4 channels, which must still be an unpacked image
  // (more than 20 channels with hi/lo planes; the one-plane mode would pack it)
  if (c.n_strides == 2 && (c.precision != 1 || (c.wide & 3) || c.wide / 4 <= 20)) return false;
  for (int i = 0; i < c.n_blocks; ++i)
    if (c.dilations[i] < 1 || c.dilations[i] > 2) return false;
  return true;
}

// frames per pass of the plane path: a whole number of frames per SM for every kernel.  28 per SM (4,144): re-measured at the end of
// round 2 against 14 per SM -- 105.7 vs 108.6 ms per 32,768 frames (half the launches, each kernel's fill and drain paid half as often)
int64_t plane_chunk_frames() {
  static const int64_t v = [] {
    const char* e = getenv("NSC_PLANE_CHUNK");
    const long long n = e ? atoll(e) : 0;
    return (int64_t)(n >= 1 && n <= 65536 ? n : 28LL * sm_count());
  }();
  return v;
}

// zeroed bytes in front of the first and behind the last activation image of a region: a layer with 16 halo rows reads 8 rows
// (1 KB) beyond an image's own zero rows -- normally the neighbouring image's zero rows, here at the region's ends
constexpr int64_t kPlaneGuard = 2048;

struct PlaneCodecPlan {
  int planes = 2;
  int Lc = 0;
  std::vector<PlaneTensor> buf;          // activation images; bases relative to the activation region (filled by bind)
  std::vector<PlaneConv> enc, dec;
  std::vector<int> enc_layer, dec_layer;  // index into the codec's layer table (parameter offsets; a gated layer also owns the next entry)
  std::vector<int> enc_sep, dec_sep;      // 0 plain conv; 1 / 2: depthwise / pointwise half of a separable layer (one table entry)
  struct Io { int in, out, res; };        // activation buffers of a layer (ids into buf, -1 = the 1-channel vector at the edge)
  std::vector<Io> enc_io, dec_io;
  std::vector<int64_t> w_off;             // packed-weight offset of every layer (enc then dec)
  int64_t wpack_bytes = 0;
  std::vector<int> enc_block, dec_block;  // per layer: 1 = first conv of a bottleneck block whose three convs can run as ONE fused launch
  uint32_t* flags = nullptr;              // frame flags of the fused blocks (plane_block_flag_words(flag_frames) words per fused block)
  int64_t flag_frames = 0;                // frames per pass the flag words were sized for
  int n_fused = 0;
};

// Lays the codec out as plane layers.  Tensors carry geometry only (base = nullptr) until plane_bind() attaches a workspace.
// Resolution levels: level l has 512 >> l positions; the encoder walks down the levels with 100-channel images, the decoder walks
// back up halving the channels at every sub-pixel stage (nscm.py:152-181, :219-260).  Every level owns two wide images (ping-pong
// of the blocks), a de-interleaved one (input of the stride-2 conv below it), two narrow images -- plus, for 'gln', the de-interleaved
// twin of the first narrow image (input of a dilation-2 gate conv) and a third wide image (depthwise half of the separable up-conv) --
// and the decoder's own pair of wide images where its channel count differs from the encoder's.
PlaneCodecPlan make_plane_plan(const nsc_codec_cfg& c) {
  PlaneCodecPlan pl;
  pl.planes = c.precision == 1 ? 2 : 1;
  const int P = pl.planes, W = c.wide, Nn = c.narrow, NS = c.n_strides;
  const bool gln = c.resnet_type == 1;
  pl.Lc = kFrameLen >> NS;
  auto newbuf = [&](int L, int C, int deint) { pl.buf.push_back(make_plane_tensor(nullptr, L, C, P, deint)); return (int)pl.buf.size() - 1; };
  // bottleneck blocks: the 20 -> 20 conv on FOLDED images (plane.cuh; NSC_PLANE_FOLD2=0 keeps the taps-in-N kernel).  Needs 128 folded
  // rows per (sub-)frame: dilation 1 from 256 positions, dilation 2 (folded per parity) from 512.
  static const bool fold2_knob = [] { const char* e = getenv("NSC_PLANE_FOLD2"); return !(e && e[0] == '0'); }();
  const bool fold2_on = fold2_knob && !gln && P == 2 && !plane_block_default_on();
  struct Level { int L, w0, w1, wd, n0, n1, n0d, w2, c0, c1, Cdec, nf, nfd; };
  std::vector<Level> lv(NS + 1);
  for (int l = 0; l <= NS; ++l) {
    Level& v = lv[l];
    v.L = kFrameLen >> l;
    v.Cdec = W >> (NS - l);                       // decoder channels at this level
    v.w0 = newbuf(v.L, W, 0);
    v.w1 = newbuf(v.L, W, 0);
    v.wd = l < NS ? newbuf(v.L, W, 1) : -1;
    v.n0 = newbuf(v.L, Nn, 0);
    v.n1 = newbuf(v.L, Nn, 0);
    v.n0d = gln ? newbuf(v.L, Nn, 1) : -1;
    v.nf = (fold2_on && v.L / 2 >= 128) ? newbuf(v.L / 2, kFoldC, 0) : -1;      // folded image of a block's first narrow tensor
    v.nfd = (fold2_on && v.L / 4 >= 128) ? newbuf(v.L / 2, kFoldC, 1) : -1;     // ... folded per position parity (dilation 2)
    v.c0 = v.c1 = v.w2 = -1;
    if (l < NS) { v.c0 = newbuf(v.L, v.Cdec, 0); v.c1 = newbuf(v.L, v.Cdec, 0); }
    if (l > 0 && gln) v.w2 = newbuf(v.L, W >> (NS - l), 0);     // depthwise result of the up-conv that leaves this level
  }

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
    if (d == 2 && Ls / 2 >= 128) {
      PlaneConv& g = add(v, vl, PK_X, Ls / 2, Nn, 2 * Nn, 15, 1, 1, NSC_ACT_NONE, n0d, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
      g.glu = 1; g.ileave = 1; g.bmul = 2;
      g.in.deint = 0; g.in.rows = Ls / 2; g.in.frame_bytes = pl.buf[n0d].frame_bytes / 2;   // view: one sub-image = one frame
    } else {
      // dilation 1 -- or dilation 2 at 128 positions, where a sub-image would be shorter than the 128-row MMA tile: the layer then
      // stages 16 halo rows per side on the plain image (the rows beyond an image's own 8 zero rows are the neighbouring image's
      // zero rows; zeroed guard bands stand in for them at both ends of the activation region, plane_bind)
      PlaneConv& g = add(v, vl, PK_X, Ls, Nn, 2 * Nn, 15, d, 1, NSC_ACT_NONE, n0, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
      g.glu = 1;
    }
    ++layer;   // the tanh gate's entry
  };
  // one stack of bottleneck blocks (nscm.py:183-217) on `cur` at level `lvl` with ping-pong images b0 / b1; returns the buffer that
  // holds the result
  auto stack = [&](std::vector<PlaneConv>& v, std::vector<int>& vl, const Level& lvl, int Cw, int cur, int b0, int b1, int last_out, bool vec_in) {
    const int Ls = lvl.L, n0 = lvl.n0, n1 = lvl.n1, n0d = lvl.n0d;
    for (int i = 0; i < c.n_blocks; ++i) {
      const bool flat = i == c.n_blocks - 1;
      int out = (cur == b0) ? b1 : b0;
      if (flat && last_out >= 0) out = last_out;
      const int post = flat ? NSC_ACT_NONE : NSC_ACT_LRELU;
      if (gln) {   // gated_bottleneck (nn_core_operator.py:82-112): k1 -> [k15 gate * tanh(k15 gate)] -> k9 + residual
        const int d = c.dilations[i];
        const int k1_out = (d == 2 && Ls / 2 >= 128) ? n0d : n0;
        const bool vec = vec_in && i == 0;
        add(v, vl, vec ? PK_GEN : PK_X, Ls, vec ? 1 : Cw, Nn, 1, 1, 1, NSC_ACT_LRELU, vec ? -1 : cur, k1_out, -1, RES_NONE, NSC_ACT_NONE, 1);
        add_gates(v, vl, Ls, d, n0, n0d, n1);
        add(v, vl, PK_X, Ls, Nn, Cw, c.k_plain, 1, 1, NSC_ACT_NONE, n1, out, vec ? -1 : cur, vec ? RES_ADD_BCAST : RES_ADD, post, 1);
        cur = out;
        continue;
      }
      // the block's 20 -> 20 conv: folded (48 -> 48 k5 on pairs of positions, plane.cuh) where the level is long enough, else taps-in-N
      const int d2 = c.dilations[i];
      // dilation 2: per position parity where a parity sub-frame still has 128 folded rows, else the block-diagonal k9 form on the
      // plain folded image
      const bool by_parity = d2 == 2 && lvl.nfd >= 0;
      const int nfold = (d2 == 1 || !by_parity) ? lvl.nf : lvl.nfd;
      const bool folded = nfold >= 0 && c.k_dilated == 9;
      const int fold_out = folded ? (by_parity ? 2 : 1) : 0;
      auto narrow_conv = [&]() {
        if (!folded) {
          add(v, vl, kNarrowKind, Ls, Nn, Nn, c.k_dilated, d2, 1, NSC_ACT_LRELU, n0, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
          return;
        }
        if (d2 == 2 && !by_parity) {
          add(v, vl, PK_X, Ls / 2, kFoldC, kFoldC, 9, 1, 1, NSC_ACT_LRELU, nfold, n1, -1, RES_NONE, NSC_ACT_NONE, 1).fold2 = 2;
          return;
        }
        PlaneConv& f = add(v, vl, PK_X, Ls / (2 * d2), kFoldC, kFoldC, kFoldK, 1, 1, NSC_ACT_LRELU, nfold, n1, -1, RES_NONE, NSC_ACT_NONE, 1);
        f.fold2 = 1;
        if (d2 == 2) {   // the two parities are independent frames of a quarter of the length; the epilogue interleaves them back
          f.ileave = 1; f.bmul = 2;
          f.in.deint = 0; f.in.rows = Ls / 4; f.in.frame_bytes = pl.buf[nfold].frame_bytes / 2;
        }
      };
      if (vec_in && i == 0) {
        add(v, vl, PK_GEN, Ls, 1, Nn, c.k_plain, 1, 1, NSC_ACT_LRELU, -1, folded ? nfold : n0, -1, RES_NONE, NSC_ACT_NONE, 1).fold_out = fold_out;
        narrow_conv();
        add(v, vl, PK_X, Ls, Nn, Cw, c.k_plain, 1, 1, NSC_ACT_NONE, n1, out, -1, RES_ADD_BCAST, post, 1);
      } else {
        (&v == &pl.enc ? pl.enc_block : pl.dec_block).resize(v.size() + 1, 0);
        (&v == &pl.enc ? pl.enc_block : pl.dec_block)[v.size()] = 1;
        ++pl.n_fused;
        add(v, vl, PK_T, Ls, Cw, Nn, c.k_plain, 1, 1, NSC_ACT_LRELU, cur, folded ? nfold : n0, -1, RES_NONE, NSC_ACT_NONE, 1).fold_out = fold_out;
        narrow_conv();
        add(v, vl, PK_X, Ls, Nn, Cw, c.k_plain, 1, 1, NSC_ACT_NONE, n1, out, cur, RES_ADD, post, 1);
      }
      cur = out;
    }
    return cur;
  };
  // encoder (nscm.py:219-237): stem, then per stage [blocks, stride-2 conv], blocks, code head
  add(pl.enc, pl.enc_layer, PK_GEN, kFrameLen, 1, W, 55, 1, 1, NSC_ACT_LRELU, -1, lv[0].w0, -1, RES_NONE, NSC_ACT_NONE, 1);
  int cur = lv[0].w0;
  for (int s = 0; s < NS; ++s) {
    cur = stack(pl.enc, pl.enc_layer, lv[s], W, cur, lv[s].w0, lv[s].w1, lv[s].wd, false);     // last block writes the de-interleaved image
    add(pl.enc, pl.enc_layer, PK_X, lv[s].L, W, W, 9, 1, 2, NSC_ACT_LRELU, cur, lv[s + 1].w0, -1, RES_NONE, NSC_ACT_NONE, 1);
    cur = lv[s + 1].w0;
  }
  cur = stack(pl.enc, pl.enc_layer, lv[NS], W, cur, lv[NS].w0, lv[NS].w1, -1, false);
  add(pl.enc, pl.enc_layer, PK_T, lv[NS].L, W, 1, 55, 1, 1, NSC_ACT_TANH, cur, -1, -1, RES_NONE, NSC_ACT_NONE, 1);
  // decoder (nscm.py:239-260): per stage [blocks, up-conv + sub-pixel shuffle], blocks, output head
  cur = stack(pl.dec, pl.dec_layer, lv[NS], W, lv[NS].w1, lv[NS].w0, lv[NS].w1, -1, true);   // first block writes the buffer that is not `cur`
  int C = W;
  for (int l = NS; l > 0; --l) {
    const int Ll = lv[l].L;
    if (gln) {   // separable up-conv (nscm.py:175-177): depthwise k9 -> pointwise + bias + leaky ReLU -> sub-pixel shuffle
      add(pl.dec, pl.dec_layer, PK_DW, Ll, C, C, 9, 1, 1, NSC_ACT_NONE, cur, lv[l].w2, -1, RES_NONE, NSC_ACT_NONE, 1, 1);
      add(pl.dec, pl.dec_layer, PK_X, Ll, C, C, 1, 1, 1, NSC_ACT_LRELU, lv[l].w2, lv[l - 1].c0, -1, RES_NONE, NSC_ACT_NONE, 2, 2);
    } else {
      add(pl.dec, pl.dec_layer, PK_X, Ll, C, C, 9, 1, 1, NSC_ACT_LRELU, cur, lv[l - 1].c0, -1, RES_NONE, NSC_ACT_NONE, 2);
    }
    C /= 2;
    cur = stack(pl.dec, pl.dec_layer, lv[l - 1], C, lv[l - 1].c0, lv[l - 1].c0, lv[l - 1].c1, -1, false);
  }
  add(pl.dec, pl.dec_layer, PK_T, kFrameLen, C, 1, 55, 1, 1, NSC_ACT_NONE, cur, -1, -1, RES_NONE, NSC_ACT_NONE, 1);

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
      if (b < 0) {
        pl.wpack_bytes = -1;
        if (getenv("NSC_PLANE_DEBUG")) fprintf(stderr, "plane plan: no launch plan for kind %d k%d d%d s%d %d->%d L%d shuffle %d glu %d\n", pc.kind, pc.K, pc.dil, pc.stride, pc.Cin, pc.Cout, pc.Lin, pc.shuffle, pc.glu);
      }
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
  for (const PlaneTensor& t : pl.buf) total += align_up(t.frame_bytes * Bc, 1024);
  return total + 1024 + 2 * kPlaneGuard;
}

// Resolves the buffer ids to addresses inside `act` (sized for Bc frames), attaches parameters and packed weights.
void plane_bind(PlaneCodecPlan& pl, const CodecLayout& lay, const float* params, void* act, int64_t Bc, void* wpack) {
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(act) + 1023) & ~(uintptr_t)1023) + kPlaneGuard;
  std::vector<uint8_t*> addr(pl.buf.size());
  for (size_t i = 0; i < pl.buf.size(); ++i) { addr[i] = base; base += align_up(pl.buf[i].frame_bytes * Bc, 1024); }
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

// encoder: x' = xscale * (x - xsub) (nb, 512) -> fcode (nb, Lc) [may be null when the quantiser is folded into the head];
// decoder: code (nb, Lc) -> out (nb, 512) [may be null when the cascade accumulation is folded into the head]
int plane_run_encoder(PlaneCodecPlan& pl, const float* x, const float* xsub, float xscale, int64_t nb, float* fcode, const HeadFold& fold,
                      cudaStream_t st) {
  NSC_TRY(plane_clear_flags(pl, st));
  int fused_no = 0;
  for (size_t i = 0; i < pl.enc.size(); ++i) {
    int rc = NSC_OK;
    if (plane_try_block(pl, pl.enc, pl.enc_block, i, nb, &fused_no, st, &rc)) { NSC_TRY(rc); i += 2; continue; }
    PlaneConv t = pl.enc[i];
    t.B = nb * t.bmul;
    if (t.Cin == 1) { t.xvec = x; t.xsub = xsub; t.xscale = xscale; }     // the stem applies the cascade's input arithmetic
    if (t.Cout == 1) { t.yvec = fcode; t.fold = fold; }
    NSC_TRY(plane_launch(t, st));
  }
  return NSC_OK;
}

int plane_run_decoder(PlaneCodecPlan& pl, const float* code, int64_t nb, float* out, const HeadFold& fold, cudaStream_t st) {
  NSC_TRY(plane_clear_flags(pl, st));
  int fused_no = 0;
  for (size_t i = 0; i < pl.dec.size(); ++i) {
    int rc = NSC_OK;
    if (plane_try_block(pl, pl.dec, pl.dec_block, i, nb, &fused_no, st, &rc)) { NSC_TRY(rc); i += 2; continue; }
    PlaneConv t = pl.dec[i];
    t.B = nb * t.bmul;
    if (t.Cin == 1) t.xvec = code;
    if (t.res_mode == RES_ADD_BCAST) t.resvec = code;
    if (t.Cout == 1) { t.yvec = out; t.fold = fold; }
    NSC_TRY(plane_launch(t, st));
  }
  return NSC_OK;
}

}  // namespace
}  // namespace nsc
