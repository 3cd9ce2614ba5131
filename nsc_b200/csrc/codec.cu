// K6 (host side) -- one neural codec (neural_speech_coding_module.py:152-335), the CMRL cascade
// (cmrl.py:513-543, :806-830) and the collaborative-quantisation feed-forward (cmrl.py:770-858) as a stream of
// kernel launches.  The topology is written ONCE (struct Walker): a dry walk enumerates the conv layers in the
// order TensorFlow would create their variables -- which defines the flat parameter image -- and the live walk
// issues the launches.  Frames are processed in chunks so the activation workspace stays bounded.
#include <string.h>

#include <mutex>

#include "plane_codec.cuh"
#include "walker.cuh"

namespace nsc {

namespace {

int64_t codec_ws_floats_per_frame(const nsc_codec_cfg& c) {
  return 3LL * c.wide * kFrameLen + 3LL * c.narrow * kFrameLen + 2LL * code_length(c);
}

int64_t codec_ws_bytes(const nsc_codec_cfg& c, int64_t Bc) {
  // 8 carved buffers, each rounded up to 256 bytes, plus the tensor engine's weight scratch
  return codec_ws_floats_per_frame(c) * Bc * (int64_t)sizeof(float) + 9 * 256 + kTcWpackBytes;
}

struct CodecBuffers {
  float* wide[3];
  float* nar[3];
  float* fcode;
  float* code;
  void* wpack;
};

CodecBuffers carve_codec(Carver& cv, const nsc_codec_cfg& c, int64_t Bc) {
  CodecBuffers b;
  for (int i = 0; i < 3; ++i) b.wide[i] = cv.take(Bc * c.wide * kFrameLen);
  for (int i = 0; i < 3; ++i) b.nar[i] = cv.take(Bc * c.narrow * kFrameLen);
  b.fcode = cv.take(Bc * code_length(c));
  b.code = cv.take(Bc * code_length(c));
  b.wpack = cv.take(kTcWpackBytes / (int64_t)sizeof(float));
  return b;
}

// one chunk of frames through one codec; any of fcode/idx/code/soft/hist/qloss may be null.
// which: bit0 = encoder+quantiser, bit1 = decoder
int run_codec_chunk(const nsc_codec_cfg& cfg, const CodecLayout& lay, const float* params, const CodecBuffers& buf,
                    const float* x, int64_t Bc, float iq, int use_soft, float* fcode, uint8_t* idx, float* code,
                    float* out, float* soft, float* hist, float* qloss, int which, cudaStream_t st) {
  Walker w;
  w.cfg = cfg;
  w.dry = false;
  w.params = params;
  w.B = Bc;
  w.st = st;
  w.layers = lay.layers;
  for (int i = 0; i < 3; ++i) { w.wide[i] = buf.wide[i]; w.nar[i] = buf.nar[i]; }
  w.wpack = buf.wpack;
  float* fc = fcode ? fcode : buf.fcode;
  float* cd = code ? code : buf.code;
  if (which & 1) {
    w.encoder(x, &fc);
    NSC_TRY(w.rc);
    const float* alpha = params + lay.conv_floats;
    const float* bins = alpha + 1;
    NSC_TRY(launch_quantize(fc, Bc, lay.code_len, bins, cfg.num_bins, alpha, iq, use_soft, cd, idx, soft, hist, qloss, st));
  } else {
    // skip the encoder's layers in the table
    Walker d;
    d.cfg = cfg;
    d.dry = true;
    float* none = nullptr;
    d.encoder(nullptr, &none);
    w.cursor = d.layers.size();
  }
  if (which & 2) {
    w.decoder(cd, &out);
    NSC_TRY(w.rc);
  }
  return NSC_OK;
}


// ---- prepared workspaces (nsc_prepare / nsc_release) ------------------------------------------------------------
// The once-per-weights part of a plane-path call -- clearing the activation images' zero rows and packing every layer's fp16
// operand slabs (one small launch per layer) -- costs more than the convs themselves at streaming batch sizes.  nsc_prepare runs
// it once into a caller-owned workspace and registers (workspace, entry point, configurations, parameter pointers, pass size);
// a later call that matches all of them skips it.  Nothing else is cached: the packed weights live in the caller's workspace.
struct PreparedWs {
  const void* ws;
  int kind, n;
  int64_t Bc;
  nsc_codec_cfg cfgs[NSC_MAX_CODECS];
  const float* params[NSC_MAX_CODECS];
};
static std::mutex g_prep_mu;
static std::vector<PreparedWs> g_prepared;

static bool prepared_same(const PreparedWs& e, int kind, const nsc_codec_cfg* cfgs, const float* const* params, int n, int64_t Bc) {
  if (e.kind != kind || e.n != n || e.Bc != Bc) return false;
  for (int i = 0; i < n; ++i)
    if (memcmp(&e.cfgs[i], &cfgs[i], sizeof(nsc_codec_cfg)) != 0 || e.params[i] != params[i]) return false;
  return true;
}
// true: `ws` holds this call's zero rows and packed weights.  A registered workspace that is used differently is forgotten (the
// call is about to overwrite it).
static bool prepared_lookup(const void* ws, int kind, const nsc_codec_cfg* cfgs, const float* const* params, int n, int64_t Bc) {
  std::lock_guard<std::mutex> lk(g_prep_mu);
  for (size_t i = 0; i < g_prepared.size(); ++i)
    if (g_prepared[i].ws == ws) {
      if (prepared_same(g_prepared[i], kind, cfgs, params, n, Bc)) return true;
      g_prepared.erase(g_prepared.begin() + (long)i);
      return false;
    }
  return false;
}
static void prepared_register(const void* ws, int kind, const nsc_codec_cfg* cfgs, const float* const* params, int n, int64_t Bc) {
  std::lock_guard<std::mutex> lk(g_prep_mu);
  for (size_t i = 0; i < g_prepared.size(); ++i)
    if (g_prepared[i].ws == ws) { g_prepared.erase(g_prepared.begin() + (long)i); break; }
  PreparedWs e;
  e.ws = ws; e.kind = kind; e.n = n; e.Bc = Bc;
  for (int i = 0; i < n; ++i) { e.cfgs[i] = cfgs[i]; e.params[i] = params[i]; }
  g_prepared.push_back(e);
}

// ---- plane path (plane_codec.cuh): per-call state of a codec or a cascade of codecs -------------------------------
struct PlaneCascade {
  std::vector<PlaneCodecPlan> plans;
  std::vector<int> region;      // activation region used by codec i (codecs with the same image geometry share one)
  int n_regions = 0;
  int64_t region_bytes = 0;     // bytes of one activation region
  float* fcode = nullptr;       // (Bc, Lc) scratch, largest Lc
  float* code = nullptr;

  // every codec's topology is covered AND every one of its layers has a launch plan (e.g. 'gln' with two stride-2 stages is not: its
  // dilation-2 gate conv at 128 positions would run on 64-row sub-images, below the 128-row MMA tile)
  static bool supported(const nsc_codec_cfg* cfgs, int n) {
    for (int i = 0; i < n; ++i)
      if (!plane_codec_supported(cfgs[i]) || make_plane_plan(cfgs[i]).wpack_bytes < 0) return false;
    return n >= 1;
  }
  static void layout(const nsc_codec_cfg* cfgs, int n, std::vector<int>* region, int* n_regions) {
    region->assign(n, 0);
    *n_regions = 0;
    for (int i = 0; i < n; ++i) {
      int r = -1;
      for (int j = 0; j < i; ++j)
        if (cfgs[j].wide == cfgs[i].wide && cfgs[j].precision == cfgs[i].precision && cfgs[j].resnet_type == cfgs[i].resnet_type &&
            cfgs[j].n_strides == cfgs[i].n_strides) { r = (*region)[j]; break; }   // same image geometry: the zero rows stay zero
      (*region)[i] = r >= 0 ? r : (*n_regions)++;
    }
  }
  static int64_t bytes(const nsc_codec_cfg* cfgs, int n, int64_t Bc) {
    std::vector<int> region;
    int nr = 0;
    layout(cfgs, n, &region, &nr);
    int64_t act = 0, w = 0;
    for (int i = 0; i < n; ++i) {
      const PlaneCodecPlan pl = make_plane_plan(cfgs[i]);
      const int64_t a = plane_codec_act_bytes(pl, Bc);
      if (a > act) act = a;
      w += align_up(pl.wpack_bytes, 1024) + plane_codec_flag_bytes(pl, Bc);
    }
    return act * nr + w + 2 * align_up(Bc * (kFrameLen / 2) * (int64_t)sizeof(float), 1024) + 2048;
  }
  // carves `ws`, clears the images' zero rows, binds every layer and packs all weights (once per call; `prepared`: the workspace
  // already holds the zero rows and the packed weights of exactly this call, see nsc_prepare)
  int setup(const nsc_codec_cfg* cfgs, const CodecLayout* lays, const float* const* params, int n, int64_t Bc, void* ws,
            int64_t ws_bytes, cudaStream_t st, bool prepared = false) {
    if (ws_bytes < bytes(cfgs, n, Bc)) {
      set_error("codec (plane path): workspace %lld < %lld bytes", (long long)ws_bytes, (long long)bytes(cfgs, n, Bc));
      return NSC_E_WORKSPACE;
    }
    layout(cfgs, n, &region, &n_regions);
    plans.clear();
    region_bytes = 0;
    for (int i = 0; i < n; ++i) {
      plans.push_back(make_plane_plan(cfgs[i]));
      if (plans[i].wpack_bytes < 0) { set_error("codec (plane path): a layer is not covered by the plane engine"); return NSC_E_INVALID; }
      const int64_t a = plane_codec_act_bytes(plans[i], Bc);
      if (a > region_bytes) region_bytes = a;
    }
    uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
    uint8_t* act0 = p;
    p += region_bytes * n_regions;
    if (!prepared) NSC_CUDA_OK(cudaMemsetAsync(act0, 0, (size_t)(region_bytes * n_regions), st));
    for (int i = 0; i < n; ++i) {
      plane_bind(plans[i], lays[i], params[i], act0 + region_bytes * region[i], Bc, p);
      p += align_up(plans[i].wpack_bytes, 1024);
      plans[i].flags = reinterpret_cast<uint32_t*>(p);
      plans[i].flag_frames = Bc;
      p += plane_codec_flag_bytes(plans[i], Bc);
      if (!prepared) NSC_TRY(plane_codec_pack(plans[i], st));
    }
    fcode = reinterpret_cast<float*>(p);
    p += align_up(Bc * (kFrameLen / 2) * (int64_t)sizeof(float), 1024);
    code = reinterpret_cast<float*>(p);
    return NSC_OK;
  }
};

// Cascade arithmetic around one codec of the plane path, folded into its first and last kernels (cmrl.py:522-531, :810-830):
// the stem reads in = in_scale * (x - in_sub) directly, the output head adds out / out_div to `acc` (and writes the quotient).
struct CascadeFold {
  const float* in_sub = nullptr;
  float in_scale = 1.f;
  float* acc = nullptr;
  float* quot = nullptr;
  float out_div = 1.f;
  int acc_first = 0;
};

static bool plane_fold_on() {
  static const bool on = [] { const char* e = getenv("NSC_PLANE_FOLD"); return !(e && e[0] == '0'); }();
  return on;
}

// plane-path twin of run_codec_chunk
int run_codec_chunk_plane(PlaneCascade& pcs, int i, const nsc_codec_cfg& cfg, const CodecLayout& lay, const float* params,
                          const float* x, int64_t nb, float iq, int use_soft, float* fcode, uint8_t* idx, float* code,
                          float* out, float* soft, float* hist, float* qloss, int which, cudaStream_t st,
                          const CascadeFold* cf = nullptr) {
  float* fc = fcode ? fcode : pcs.fcode;
  float* cd = code ? code : pcs.code;
  if (which & 1) {
    const float* alpha = params + lay.conv_floats;
    // hard codes and no statistics asked for: the quantiser runs in the code head's epilogue (NSC_PLANE_FOLD=0: stand-alone kernel)
    const bool qfold = plane_fold_on() && !use_soft && soft == nullptr && hist == nullptr && qloss == nullptr;
    HeadFold hf;
    if (qfold) { hf.q_bins = alpha + 1; hf.q_alpha = alpha; hf.q_n = cfg.num_bins; hf.q_iq = iq; hf.q_idx = idx; hf.q_code = cd; }
    NSC_TRY(plane_run_encoder(pcs.plans[i], x, cf ? cf->in_sub : nullptr, cf ? cf->in_scale : 1.f, nb, (qfold && !fcode) ? nullptr : fc, hf, st));
    if (!qfold) NSC_TRY(launch_quantize(fc, nb, lay.code_len, alpha + 1, cfg.num_bins, alpha, iq, use_soft, cd, idx, soft, hist, qloss, st));
  }
  if (which & 2) {
    HeadFold hf;
    if (cf && cf->acc) { hf.acc = cf->acc; hf.quot = cf->quot; hf.div = cf->out_div; hf.acc_first = cf->acc_first; }
    NSC_TRY(plane_run_decoder(pcs.plans[i], cd, nb, out, hf, st));
  }
  return NSC_OK;
}

}  // namespace
}  // namespace nsc

using nsc::kFrameLen;

// frames per internal pass: the plane path wants a whole number of frames per SM
static int64_t chunk_for(const nsc_codec_cfg* cfgs, int n) {
  return nsc::PlaneCascade::supported(cfgs, n) ? nsc::plane_chunk_frames() : kChunkFrames;
}

extern "C" {

int64_t nsc_codec_param_count(const nsc_codec_cfg* cfg) {
  if (nsc::validate_cfg(cfg) != NSC_OK) return -1;
  return nsc::make_layout(*cfg).conv_floats + 1 + cfg->num_bins;
}

int32_t nsc_codec_layer_info(const nsc_codec_cfg* cfg, int32_t i, int32_t* k, int32_t* cin, int32_t* cout,
                             int32_t* separable, int64_t* offset) {
  if (nsc::validate_cfg(cfg) != NSC_OK) return -1;
  nsc::CodecLayout l = nsc::make_layout(*cfg);
  if (i < 0) return (int32_t)l.layers.size();
  if (i >= (int32_t)l.layers.size()) {
    nsc::set_error("nsc_codec_layer_info: layer %d out of range (%zu)", i, l.layers.size());
    return -1;
  }
  if (k) *k = l.layers[i].k;
  if (cin) *cin = l.layers[i].cin;
  if (cout) *cout = l.layers[i].cout;
  if (separable) *separable = l.layers[i].separable;
  if (offset) *offset = l.layers[i].off;
  return (int32_t)l.layers.size();
}

int32_t nsc_codec_on_plane_engine(const nsc_codec_cfg* cfg) {
  if (nsc::validate_cfg(cfg) != NSC_OK) return -1;
  return nsc::PlaneCascade::supported(cfg, 1) ? 1 : 0;
}

int64_t nsc_codec_workspace_bytes(const nsc_codec_cfg* cfg, int64_t B) {
  if (nsc::validate_cfg(cfg) != NSC_OK) return -1;
  const int64_t chunk = chunk_for(cfg, 1);
  const int64_t Bc = B < chunk ? (B < 1 ? 1 : B) : chunk;
  if (nsc::PlaneCascade::supported(cfg, 1)) return nsc::PlaneCascade::bytes(cfg, 1, Bc);
  return nsc::codec_ws_bytes(*cfg, Bc);
}

static int codec_run(const nsc_codec_cfg* cfg, const float* params, const float* x, const float* code_in, int64_t B,
                     float iq, int32_t use_soft, float* fcode, uint8_t* idx, float* code, float* out, float* soft,
                     float* hist, float* qloss, void* workspace, int64_t workspace_bytes, void* stream, int which) {
  NSC_TRY(nsc::validate_cfg(cfg));
  NSC_CHECK_ARG(params != nullptr && workspace != nullptr, "codec: null params/workspace");
  NSC_CHECK_ARG(B >= 0, "codec: negative batch");
  if (B == 0) return NSC_OK;
  const int64_t chunk = chunk_for(cfg, 1);
  const int64_t Bc = B < chunk ? B : chunk;
  const nsc::CodecLayout lay = nsc::make_layout(*cfg);
  const int Lc = lay.code_len, n = cfg->num_bins;
  cudaStream_t st = (cudaStream_t)stream;
  if (nsc::PlaneCascade::supported(cfg, 1)) {
    nsc::PlaneCascade pcs;
    NSC_TRY(pcs.setup(cfg, &lay, &params, 1, Bc, workspace, workspace_bytes, st, nsc::prepared_lookup(workspace, 0, cfg, &params, 1, Bc)));
    for (int64_t b0 = 0; b0 < B; b0 += Bc) {
      const int64_t nb = (B - b0) < Bc ? (B - b0) : Bc;
      if (which == 2)
        NSC_TRY(nsc::run_codec_chunk_plane(pcs, 0, *cfg, lay, params, nullptr, nb, iq, use_soft, nullptr, nullptr,
                                           const_cast<float*>(code_in + b0 * Lc), out + b0 * kFrameLen, nullptr, nullptr, nullptr, 2, st));
      else
        NSC_TRY(nsc::run_codec_chunk_plane(pcs, 0, *cfg, lay, params, x + b0 * kFrameLen, nb, iq, use_soft,
                                           fcode ? fcode + b0 * Lc : nullptr, idx ? idx + b0 * Lc : nullptr,
                                           code ? code + b0 * Lc : nullptr, out ? out + b0 * kFrameLen : nullptr,
                                           soft ? soft + b0 * Lc * n : nullptr, hist, qloss ? qloss + b0 : nullptr, which, st));
    }
    return NSC_OK;
  }
  if (workspace_bytes < nsc::codec_ws_bytes(*cfg, Bc)) {
    nsc::set_error("codec: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)nsc::codec_ws_bytes(*cfg, Bc));
    return NSC_E_WORKSPACE;
  }
  for (int64_t b0 = 0; b0 < B; b0 += Bc) {
    const int64_t nb = (B - b0) < Bc ? (B - b0) : Bc;
    nsc::Carver cv(workspace, workspace_bytes);
    nsc::CodecBuffers buf = nsc::carve_codec(cv, *cfg, Bc);
    const float* code_src = code_in ? code_in + b0 * Lc : nullptr;
    float* code_dst = code ? code + b0 * Lc : nullptr;
    if (which == 2) {
      // decoder only: the code comes from the caller
      NSC_TRY(nsc::run_codec_chunk(*cfg, lay, params, buf, nullptr, nb, iq, use_soft, nullptr, nullptr,
                                   const_cast<float*>(code_src), out + b0 * kFrameLen, nullptr, nullptr, nullptr, 2, st));
    } else {
      NSC_TRY(nsc::run_codec_chunk(*cfg, lay, params, buf, x + b0 * kFrameLen, nb, iq, use_soft,
                                   fcode ? fcode + b0 * Lc : nullptr, idx ? idx + b0 * Lc : nullptr, code_dst,
                                   out ? out + b0 * kFrameLen : nullptr, soft ? soft + b0 * Lc * n : nullptr, hist,
                                   qloss ? qloss + b0 : nullptr, which, st));
    }
  }
  return NSC_OK;
}

int nsc_codec_forward(const nsc_codec_cfg* cfg, const float* params, const float* x, int64_t B, float is_quan_on,
                      int32_t use_soft, float* floating_code, uint8_t* idx, float* code, float* out, float* soft,
                      float* hist, float* qloss, void* workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch
  NSC_CHECK_ARG(x != nullptr && out != nullptr, "nsc_codec_forward: null x/out");
  return codec_run(cfg, params, x, nullptr, B, is_quan_on, use_soft, floating_code, idx, code, out, soft, hist, qloss,
                   workspace, workspace_bytes, stream, 3);
}

int nsc_codec_encode(const nsc_codec_cfg* cfg, const float* params, const float* x, int64_t B, float is_quan_on,
                     int32_t use_soft, float* floating_code, uint8_t* idx, float* code, float* soft, float* hist,
                     float* qloss, void* workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch
  NSC_CHECK_ARG(x != nullptr, "nsc_codec_encode: null x");
  return codec_run(cfg, params, x, nullptr, B, is_quan_on, use_soft, floating_code, idx, code, nullptr, soft, hist,
                   qloss, workspace, workspace_bytes, stream, 1);
}

int nsc_codec_decode(const nsc_codec_cfg* cfg, const float* params, const float* code, int64_t B, float* out,
                     void* workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch
  NSC_CHECK_ARG(code != nullptr && out != nullptr, "nsc_codec_decode: null code/out");
  return codec_run(cfg, params, nullptr, code, B, 1.0f, 0, nullptr, nullptr, nullptr, out, nullptr, nullptr, nullptr,
                   workspace, workspace_bytes, stream, 2);
}

// ---- cascade -------------------------------------------------------------------------------------
static int64_t cascade_ws_bytes(const nsc_codec_cfg* cfgs, int32_t n, int64_t Bc) {
  if (nsc::PlaneCascade::supported(cfgs, n))
    return nsc::PlaneCascade::bytes(cfgs, n, Bc) + 2 * (Bc * kFrameLen * (int64_t)sizeof(float) + 256);
  int64_t worst = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t b = nsc::codec_ws_bytes(cfgs[i], Bc);
    if (b > worst) worst = b;
  }
  // + codec input, codec output (pre-division)
  return worst + 2 * (Bc * kFrameLen * (int64_t)sizeof(float) + 256);
}

int64_t nsc_cascade_workspace_bytes(const nsc_codec_cfg* cfgs, int32_t n_codecs, int64_t B) {
  if (cfgs == nullptr || n_codecs < 1 || n_codecs > NSC_MAX_CODECS) return -1;
  for (int i = 0; i < n_codecs; ++i)
    if (nsc::validate_cfg(&cfgs[i]) != NSC_OK) return -1;
  const int64_t chunk = chunk_for(cfgs, n_codecs);
  const int64_t Bc = B < chunk ? (B < 1 ? 1 : B) : chunk;
  return cascade_ws_bytes(cfgs, n_codecs, Bc);
}

// one chunk of the cascade; decoded (Bc,512) is also the running sum of the codec outputs
static int cascade_chunk(const nsc_codec_cfg* cfgs, const nsc::CodecLayout* lays, int32_t n,
                         const float* const* params, const float* x, int64_t b0, int64_t nb, int64_t Bc,
                         float res_scalar, int lpc_variant, float iq, int use_soft, uint8_t* const* idx,
                         float* const* hist, float* const* qloss, float* const* outs, float* decoded,
                         void* workspace, int64_t workspace_bytes, cudaStream_t st, nsc::PlaneCascade* pcs = nullptr) {
  const int64_t nfl = nb * kFrameLen;
  for (int i = 0; i < n; ++i) {
    nsc::Carver cv(workspace, workspace_bytes);
    float* cin = cv.take(Bc * kFrameLen);
    float* cout = cv.take(Bc * kFrameLen);
    nsc::CodecBuffers buf{};
    if (!pcs) buf = nsc::carve_codec(cv, cfgs[i], Bc);
    const float* xin = x;
    const int Lc = lays[i].code_len;
    const bool divide = (i > 0) || lpc_variant;       // codec 0 of the plain cascade is not divided (cmrl.py:522-528)
    const float d = divide ? res_scalar : 1.0f;
    if (pcs && nsc::plane_fold_on()) {
      // plane path: the stem applies the input arithmetic, the output head accumulates -- no stand-alone cascade kernels, and the
      // codec's input / raw output never exist as tensors in HBM
      nsc::CascadeFold cf;
      if (i == 0) { cf.in_scale = lpc_variant ? res_scalar : 1.0f; }          // res_x * res_scalar, cmrl.py:810
      else { cf.in_sub = decoded; cf.in_scale = res_scalar; }                  // res_scalar * (x - sum_{j<i} out_j), cmrl.py:529-531 / :822-823
      cf.acc = decoded; cf.acc_first = i == 0 ? 1 : 0; cf.out_div = d;
      cf.quot = (outs && outs[i]) ? outs[i] + b0 * kFrameLen : nullptr;
      NSC_TRY(nsc::run_codec_chunk_plane(*pcs, i, cfgs[i], lays[i], params[i], x, nb, iq, use_soft, nullptr,
                                         (idx && idx[i]) ? idx[i] + b0 * Lc : nullptr, nullptr, nullptr, nullptr,
                                         (hist && hist[i]) ? hist[i] : nullptr, (qloss && qloss[i]) ? qloss[i] + b0 : nullptr, 3, st, &cf));
      continue;
    }
    if (i == 0) {
      if (lpc_variant && res_scalar != 1.0f) {       // res_x * res_scalar, cmrl.py:810
        NSC_TRY(nsc::launch_axpby(cin, x, res_scalar, nullptr, 0.f, nfl, st));
        xin = cin;
      }
    } else {                                          // res_scalar * (x - sum_{j<i} out_j), cmrl.py:529-531 / :822-823
      NSC_TRY(nsc::launch_axpby(cin, x, res_scalar, decoded, 1.0f, nfl, st));
      xin = cin;
    }
    if (pcs)
      NSC_TRY(nsc::run_codec_chunk_plane(*pcs, i, cfgs[i], lays[i], params[i], xin, nb, iq, use_soft, nullptr,
                                         (idx && idx[i]) ? idx[i] + b0 * Lc : nullptr, nullptr, cout, nullptr,
                                         (hist && hist[i]) ? hist[i] : nullptr, (qloss && qloss[i]) ? qloss[i] + b0 : nullptr, 3, st));
    else
      NSC_TRY(nsc::run_codec_chunk(cfgs[i], lays[i], params[i], buf, xin, nb, iq, use_soft, nullptr,
                                   (idx && idx[i]) ? idx[i] + b0 * Lc : nullptr, nullptr, cout, nullptr,
                                   (hist && hist[i]) ? hist[i] : nullptr, (qloss && qloss[i]) ? qloss[i] + b0 : nullptr, 3, st));
    if (outs && outs[i]) NSC_TRY(nsc::launch_div(outs[i] + b0 * kFrameLen, cout, d, nfl, st));
    NSC_TRY(nsc::launch_accum_div(decoded, cout, d, i == 0 ? 1 : 0, nfl, st));
  }
  return NSC_OK;
}

static int check_cascade_args(const nsc_codec_cfg* cfgs, int32_t n, const float* const* params, const float* x,
                              float* decoded, void* workspace, float res_scalar) {
  NSC_CHECK_ARG(cfgs != nullptr && n >= 1 && n <= NSC_MAX_CODECS, "cascade: n_codecs=%d", n);
  NSC_CHECK_ARG(params != nullptr && x != nullptr && decoded != nullptr && workspace != nullptr, "cascade: null pointer");
  NSC_CHECK_ARG(res_scalar != 0.0f, "cascade: res_scalar is zero");
  for (int i = 0; i < n; ++i) {
    NSC_TRY(nsc::validate_cfg(&cfgs[i]));
    NSC_CHECK_ARG(params[i] != nullptr, "cascade: params[%d] is null", i);
  }
  return NSC_OK;
}

int nsc_cascade_forward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                        const float* x, int64_t B, float res_scalar, int32_t lpc_variant, float is_quan_on,
                        int32_t use_soft, uint8_t* const* idx_ptrs_host, float* const* hist_ptrs_host,
                        float* const* qloss_ptrs_host, float* const* outs_ptrs_host, float* decoded,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_TRY(check_cascade_args(cfgs, n_codecs, params_ptrs_host, x, decoded, workspace, res_scalar));
  const int64_t chunk = chunk_for(cfgs, n_codecs);
  const int64_t Bc = B < chunk ? B : chunk;
  if (workspace_bytes < cascade_ws_bytes(cfgs, n_codecs, Bc)) {
    nsc::set_error("cascade: workspace %lld < %lld bytes", (long long)workspace_bytes,
                   (long long)cascade_ws_bytes(cfgs, n_codecs, Bc));
    return NSC_E_WORKSPACE;
  }
  std::vector<nsc::CodecLayout> lays;
  for (int i = 0; i < n_codecs; ++i) lays.push_back(nsc::make_layout(cfgs[i]));
  nsc::PlaneCascade pcs;
  const bool plane = nsc::PlaneCascade::supported(cfgs, n_codecs);
  if (plane) {
    const int64_t head = 2 * nsc::align_up(Bc * kFrameLen * (int64_t)sizeof(float), 256);   // cascade_chunk's codec input / output
    NSC_TRY(pcs.setup(cfgs, lays.data(), params_ptrs_host, n_codecs, Bc, static_cast<char*>(workspace) + head, workspace_bytes - head,
                      (cudaStream_t)stream, nsc::prepared_lookup(workspace, 1, cfgs, params_ptrs_host, n_codecs, Bc)));
  }
  for (int64_t b0 = 0; b0 < B; b0 += Bc) {
    const int64_t nb = (B - b0) < Bc ? (B - b0) : Bc;
    NSC_TRY(cascade_chunk(cfgs, lays.data(), n_codecs, params_ptrs_host, x + b0 * kFrameLen, b0, nb, Bc, res_scalar,
                          lpc_variant, is_quan_on, use_soft, idx_ptrs_host, hist_ptrs_host, qloss_ptrs_host,
                          outs_ptrs_host, decoded + b0 * kFrameLen, workspace, workspace_bytes, (cudaStream_t)stream,
                          plane ? &pcs : nullptr));
  }
  return NSC_OK;
}

// ---- collaborative quantisation feed-forward -----------------------------------------------------
static int64_t cq_extra_bytes(int64_t Bc) {
  // quantised LSF (16), poly (17), residual (512) per frame
  return (Bc * NSC_LPC_ORDER + Bc * (NSC_LPC_ORDER + 1) + Bc * kFrameLen) * (int64_t)sizeof(float) + 3 * 256;
}

int64_t nsc_pass_frames(const nsc_codec_cfg* cfgs, int32_t n_codecs) {
  if (cfgs == nullptr || n_codecs < 1 || n_codecs > NSC_MAX_CODECS) return -1;
  return chunk_for(cfgs, n_codecs);
}

int64_t nsc_cq_workspace_bytes(const nsc_codec_cfg* cfgs, int32_t n_codecs, int64_t B) {
  const int64_t c = nsc_cascade_workspace_bytes(cfgs, n_codecs, B);
  if (c < 0) return c;
  const int64_t chunk = chunk_for(cfgs, n_codecs);
  const int64_t Bc = B < chunk ? (B < 1 ? 1 : B) : chunk;
  return c + cq_extra_bytes(Bc);
}

int nsc_cq_forward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                   const float* lsf_params, int32_t n_lsf_bins, const float* x, const float* lsf, int64_t B,
                   float res_scalar, float is_quan_on, int32_t use_soft, uint8_t* lsf_idx, float* lsf_hist,
                   float* lsf_qloss, uint8_t* const* idx_ptrs_host, float* const* hist_ptrs_host,
                   float* const* qloss_ptrs_host, float* poly, float* res_x, float* decoded, float* synthesized,
                   void* workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_TRY(check_cascade_args(cfgs, n_codecs, params_ptrs_host, x, decoded, workspace, res_scalar));
  NSC_CHECK_ARG(lsf_params != nullptr && lsf != nullptr, "nsc_cq_forward: null LSF input");
  NSC_CHECK_ARG(n_lsf_bins >= 1 && n_lsf_bins <= 256, "nsc_cq_forward: n_lsf_bins=%d", n_lsf_bins);
  const int64_t chunk = chunk_for(cfgs, n_codecs);
  const int64_t Bc = B < chunk ? B : chunk;
  const int64_t need = cascade_ws_bytes(cfgs, n_codecs, Bc) + cq_extra_bytes(Bc);
  if (workspace_bytes < need) {
    nsc::set_error("nsc_cq_forward: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return NSC_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<nsc::CodecLayout> lays;
  for (int i = 0; i < n_codecs; ++i) lays.push_back(nsc::make_layout(cfgs[i]));
  const int P = NSC_LPC_ORDER;
  nsc::PlaneCascade pcs;
  const bool plane = nsc::PlaneCascade::supported(cfgs, n_codecs);
  if (plane) {
    nsc::Carver cv(workspace, workspace_bytes);
    cv.take(Bc * P); cv.take(Bc * (P + 1)); cv.take(Bc * kFrameLen);
    const int64_t head = cv.used + 2 * nsc::align_up(Bc * kFrameLen * (int64_t)sizeof(float), 256);
    NSC_TRY(pcs.setup(cfgs, lays.data(), params_ptrs_host, n_codecs, Bc, static_cast<char*>(workspace) + head, workspace_bytes - head, st,
                      nsc::prepared_lookup(workspace, 2, cfgs, params_ptrs_host, n_codecs, Bc)));
  }
  // The LPC front end (LSF codebook -> lsf2poly -> residual) does not depend on the codecs: when the caller keeps poly and res_x
  // for the whole batch it runs ONCE over all B frames instead of once per pass (three launches per call instead of three per
  // 2k frames, each at full occupancy).  The quantised LSFs (B x 16) borrow the head of `synthesized`, which is written last.
  const bool front_once = poly != nullptr && res_x != nullptr && synthesized != nullptr && B > Bc;
  if (front_once) {
    float* qlsf = synthesized;
    NSC_TRY(nsc::launch_quantize(lsf, B, P, lsf_params + 1, n_lsf_bins, lsf_params, is_quan_on, use_soft, qlsf, lsf_idx, nullptr,
                                 lsf_hist, lsf_qloss, st));                                   // cmrl.py:782-788
    NSC_TRY(nsc_lsf2poly(qlsf, B, poly, nullptr, stream));                                   // cmrl.py:793
    NSC_TRY(nsc_lpc_residual(x, poly, B, res_x, stream));                                     // cmrl.py:796
  }
  for (int64_t b0 = 0; b0 < B; b0 += Bc) {
    const int64_t nb = (B - b0) < Bc ? (B - b0) : Bc;
    nsc::Carver cv(workspace, workspace_bytes);
    float* qlsf = cv.take(Bc * P);
    float* poly_ws = cv.take(Bc * (P + 1));
    float* res_ws = cv.take(Bc * kFrameLen);
    void* rest = static_cast<char*>(workspace) + cv.used;
    const int64_t rest_bytes = workspace_bytes - cv.used;
    float* poly_c = poly ? poly + b0 * (P + 1) : poly_ws;
    float* res_c = res_x ? res_x + b0 * kFrameLen : res_ws;
    if (!front_once) {
      // LSF codebook: scalar_softmax_quantization(lpc_x, alpha, lpc_bins, ...) cmrl.py:782-788
      NSC_TRY(nsc::launch_quantize(lsf + b0 * P, nb, P, lsf_params + 1, n_lsf_bins, lsf_params, is_quan_on, use_soft, qlsf,
                                   lsf_idx ? lsf_idx + b0 * P : nullptr, nullptr, lsf_hist,
                                   lsf_qloss ? lsf_qloss + b0 : nullptr, st));
      NSC_TRY(nsc_lsf2poly(qlsf, nb, poly_c, nullptr, stream));                              // cmrl.py:793
      NSC_TRY(nsc_lpc_residual(x + b0 * kFrameLen, poly_c, nb, res_c, stream));               // cmrl.py:796
    }
    NSC_TRY(cascade_chunk(cfgs, lays.data(), n_codecs, params_ptrs_host, res_c, b0, nb, Bc, res_scalar, 1, is_quan_on,
                          use_soft, idx_ptrs_host, hist_ptrs_host, qloss_ptrs_host, nullptr, decoded + b0 * kFrameLen,
                          rest, rest_bytes, st, plane ? &pcs : nullptr));
    if (synthesized && poly == nullptr)                                                      // cmrl.py:843 (per chunk: poly lives in the workspace)
      NSC_TRY(nsc_lpc_synth(poly_c, decoded + b0 * kFrameLen, nb, synthesized + b0 * kFrameLen, stream));
  }
  // one thread per frame, 512 dependent steps: latency-bound, so it runs ONCE over the whole batch when the caller keeps poly
  if (synthesized && poly != nullptr) NSC_TRY(nsc_lpc_synth(poly, decoded, B, synthesized, stream));
  return NSC_OK;
}

// ---- prepared workspaces ---------------------------------------------------------------------------------------
int nsc_prepare(int32_t kind, const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host, int64_t B,
                void* workspace, int64_t workspace_bytes, void* stream) {
  NSC_CHECK_ARG(kind >= 0 && kind <= 2, "nsc_prepare: kind %d (0 codec, 1 cascade, 2 collaborative quantisation)", kind);
  NSC_CHECK_ARG(cfgs && params_ptrs_host && workspace && n_codecs >= 1 && n_codecs <= NSC_MAX_CODECS && (kind != 0 || n_codecs == 1) && B >= 1,
                "nsc_prepare: bad arguments");
  for (int i = 0; i < n_codecs; ++i) {
    NSC_TRY(nsc::validate_cfg(&cfgs[i]));
    NSC_CHECK_ARG(params_ptrs_host[i] != nullptr, "nsc_prepare: params[%d] is null", i);
  }
  if (!nsc::PlaneCascade::supported(cfgs, n_codecs)) return NSC_OK;      // the layer-by-layer engines keep no per-weights state
  const int64_t chunk = chunk_for(cfgs, n_codecs);
  const int64_t Bc = B < chunk ? B : chunk;
  const int64_t need = kind == 0 ? nsc::PlaneCascade::bytes(cfgs, 1, Bc)
                                 : cascade_ws_bytes(cfgs, n_codecs, Bc) + (kind == 2 ? cq_extra_bytes(Bc) : 0);
  if (workspace_bytes < need) {
    nsc::set_error("nsc_prepare: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return NSC_E_WORKSPACE;
  }
  // the same carve as the entry point of this kind
  int64_t head = 0;
  if (kind >= 1) head = 2 * nsc::align_up(Bc * kFrameLen * (int64_t)sizeof(float), 256);
  if (kind == 2) {
    nsc::Carver cv(workspace, workspace_bytes);
    cv.take(Bc * NSC_LPC_ORDER); cv.take(Bc * (NSC_LPC_ORDER + 1)); cv.take(Bc * kFrameLen);
    head += cv.used;
  }
  std::vector<nsc::CodecLayout> lays;
  for (int i = 0; i < n_codecs; ++i) lays.push_back(nsc::make_layout(cfgs[i]));
  nsc::PlaneCascade pcs;
  NSC_TRY(pcs.setup(cfgs, lays.data(), params_ptrs_host, n_codecs, Bc, static_cast<char*>(workspace) + head, workspace_bytes - head,
                    (cudaStream_t)stream, false));
  nsc::prepared_register(workspace, kind, cfgs, params_ptrs_host, n_codecs, Bc);
  return NSC_OK;
}

int nsc_release(const void* workspace) {
  std::lock_guard<std::mutex> lk(nsc::g_prep_mu);
  for (size_t i = 0; i < nsc::g_prepared.size(); ++i)
    if (nsc::g_prepared[i].ws == workspace) { nsc::g_prepared.erase(nsc::g_prepared.begin() + (long)i); return 1; }
  return 0;
}

// ---- single blocks on channels-last tensors (operator surface of nn_core_operator.py) ---------------
int64_t nsc_block_workspace_bytes(int64_t B, int32_t L, int32_t wide, int32_t narrow) {
  (void)wide;
  return 3 * (nsc::align_up(B * (int64_t)L * narrow * (int64_t)sizeof(float), 256));
}

int nsc_bottleneck_block(const float* x, const float* params, float* y, int64_t B, int32_t L, int32_t Cin,
                         int32_t wide, int32_t narrow, int32_t k_plain, int32_t k_dilated, int32_t dilation,
                         int32_t is_last_flat, int32_t gated, void* workspace, int64_t workspace_bytes, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(x && params && y && workspace, "nsc_bottleneck_block: null pointer");
  NSC_CHECK_ARG(Cin == wide || Cin == 1, "nsc_bottleneck_block: residual add needs Cin == wide or Cin == 1 (got %d vs %d)", Cin, wide);
  NSC_CHECK_ARG(workspace_bytes >= nsc_block_workspace_bytes(B, L, wide, narrow), "nsc_bottleneck_block: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  nsc::Carver cv(workspace, workspace_bytes);
  float* n0 = cv.take(B * (int64_t)L * narrow);
  float* n1 = cv.take(B * (int64_t)L * narrow);
  float* n2 = cv.take(B * (int64_t)L * narrow);
  const int rmode = (Cin == wide) ? nsc::RES_ADD : nsc::RES_ADD_BCAST;
  const int post = is_last_flat ? NSC_ACT_NONE : NSC_ACT_LRELU;
  const float* p = params;
  auto conv = [&](const float* in, float* out, int cin, int cout, int K, int dil, int act, const float* res, int res_mode,
                  int post_act, int in_cl, int out_cl, int r_cl) -> int {
    nsc::ConvArgs a;
    a.x = in; a.y = out; a.w = p; a.bias = p + (int64_t)K * cin * cout;
    p += (int64_t)K * cin * cout + cout;
    a.B = B; a.Lin = L; a.Cin = cin; a.Cout = cout; a.K = K; a.dil = dil; a.act = act;
    a.res = res; a.res_mode = res_mode; a.post_act = post_act; a.x_cl = in_cl; a.y_cl = out_cl; a.res_cl = r_cl;
    return nsc::launch_conv(a, st);
  };
  if (!gated) {
    NSC_TRY(conv(x, n0, Cin, narrow, k_plain, 1, NSC_ACT_LRELU, nullptr, nsc::RES_NONE, NSC_ACT_NONE, 1, 0, 0));
    NSC_TRY(conv(n0, n1, narrow, narrow, k_dilated, dilation, NSC_ACT_LRELU, nullptr, nsc::RES_NONE, NSC_ACT_NONE, 0, 0, 0));
    NSC_TRY(conv(n1, y, narrow, wide, k_plain, 1, NSC_ACT_NONE, x, rmode, post, 0, 1, 1));
  } else {
    NSC_TRY(conv(x, n0, Cin, narrow, 1, 1, NSC_ACT_LRELU, nullptr, nsc::RES_NONE, NSC_ACT_NONE, 1, 0, 0));
    NSC_TRY(conv(n0, n1, narrow, narrow, 15, dilation, NSC_ACT_NONE, nullptr, nsc::RES_NONE, NSC_ACT_NONE, 0, 0, 0));
    NSC_TRY(conv(n0, n2, narrow, narrow, 15, dilation, NSC_ACT_TANH, n1, nsc::RES_MUL, NSC_ACT_NONE, 0, 0, 0));
    NSC_TRY(conv(n2, y, narrow, wide, k_plain, 1, NSC_ACT_NONE, x, rmode, post, 0, 1, 1));
  }
  return NSC_OK;
}

}  // extern "C"
