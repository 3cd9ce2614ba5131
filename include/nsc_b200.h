/*
 * nsc_b200.h -- C ABI of libnsc_b200.so: the B200 (sm_100a) implementation of NSC's batched
 * frame-wise codec pass.
 *
 * The reference (/root/reference) has no FFI: its operator surface is plain Python functions
 * (nn_core_operator.py, lpc_utilities.py, loss_terms_and_measures.py) whose arithmetic runs inside
 * TensorFlow / audiolazy / spectrum.  Each entry point below names the reference function (file:line)
 * it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory,
 *     including workspaces (size queried with the matching *_workspace_bytes call);
 *   - all tensors are dense float32 unless stated; "frames" B is the batch dimension;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work and never synchronise;
 *   - return value: 0 = ok, <0 = error (NSC_E_*), message via nsc_last_error() (thread local);
 *   - the library keeps no global mutable state besides per-kernel attribute caches.
 */
#ifndef NSC_B200_H_
#define NSC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSC_OK 0
#define NSC_E_INVALID (-1)      /* bad argument / unsupported configuration */
#define NSC_E_CUDA (-2)         /* a CUDA runtime call failed               */
#define NSC_E_WORKSPACE (-3)    /* workspace too small                      */

#define NSC_FRAME_LENGTH 512    /* constants.py:25 */
#define NSC_LPC_ORDER 16        /* neural_speech_coding_module.py:50 */
#define NSC_MAX_BLOCKS 8
#define NSC_MAX_STRIDES 4
#define NSC_MAX_CODECS 8

int nsc_version(void);
const char* nsc_last_error(void);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches evidence). */
long long nsc_launch_count(void);
/* Optional per-launch timing for bench.py's roofline line: between begin and end every launch records a
 * CUDA-event pair on its stream together with its algorithmic flops and bytes.  nsc_profile_end synchronises on
 * the recorded events and returns up to `cap` records: names (cap x 32 chars), milliseconds, flops, bytes. */
int nsc_profile_begin(int32_t max_records);
int nsc_profile_end(int32_t* n_records, char* names, float* ms, double* flops, double* bytes, int32_t cap);

/* ------------------------------------------------------------------------------------------------
 * Activations of nn_core_operator.py: None, tf.nn.tanh, activation_func (= leaky_relu 0.2, :24-31)
 * ---------------------------------------------------------------------------------------------- */
enum { NSC_ACT_NONE = 0, NSC_ACT_TANH = 1, NSC_ACT_LRELU = 2 };

/* ------------------------------------------------------------------------------------------------
 * conv1d  (nn_core_operator.py:6-14; also change_channel :45-54)
 *   x (B, Lin, Cin) channels-last, w (k, Cin, Cout) [TF kernel layout], b (Cout) -> y (B, ceil(Lin/stride), Cout)
 *   SAME padding, cross-correlation, bias always, optional fused activation.
 * ---------------------------------------------------------------------------------------------- */
int nsc_conv1d(const float* x, const float* w, const float* b, float* y, int64_t B, int32_t Lin, int32_t Cin,
               int32_t Cout, int32_t k, int32_t dilation, int32_t stride, int32_t activation, void* stream);

/* PRECISION MODES of the tensor-core conv path (nsc_codec_cfg.precision, `precision` arguments):
 *   1  fp16 hi/lo split: every activation and weight is x = hi + lo (two fp16 values, 22 mantissa bits), products hi*wh + lo*wh +
 *      hi*wl accumulate in fp32 in TMEM.  fp32-CLASS: measured per-frame error of the whole encoder 3-5e-6 of the frame's code peak
 *      for input levels 1e-4 ... 1 (the same as a float32 evaluation) and within the 1e-4 parity bar for |x| <= 64 -- the SUPPORTED
 *      INPUT DOMAIN (the reference feeds unit-variance or 1/33.46-scaled frames, constants.py:16).  Beyond it (|x| ~ 1e3) sums that
 *      cancel over three decades expose the 22-bit mantissa: ~4x the float32 error (tests/test_gpu_parity_depth.py).
 *   2  plain fp16 inputs: REDUCED precision (~1e-3), stated separately, never the headline.
 *   0  fp32 FFMA on the CUDA cores. */
/* conv1d on the tcgen05 tensor cores ("plane engine", the codec's conv path) with the fused epilogue of the codec's
 * layers, channels-last fp32 tensors at the edge:
 *   y = post_act( act(conv1d(x) + b) (+ res) ), optionally sub-pixel shuffled: y[b, shuffle*l + r, c] = t[b, l, shuffle*c + r]
 *   (nscm.py:158-167).  res_mode 0 = none, 1 = res (B, Lout, Cout), 2 = res (B, Lout) broadcast over channels
 *   (nn_core_operator.py:77).  precision 1 = fp16 hi/lo split, 3 MMAs, fp32-class; 2 = fp16 inputs (reduced).
 *   Covers the codec's layer shapes (Lout a multiple of 128, stride 1 or 2, k <= 9 taps for multi-channel inputs, k55
 *   1-channel stems / 1-channel heads); anything else returns NSC_E_INVALID -- there is no fallback. */
int64_t nsc_conv1d_tc_workspace_bytes(int64_t B, int32_t Lin, int32_t Cin, int32_t Cout, int32_t k, int32_t dilation,
                                      int32_t stride, int32_t res_mode, int32_t shuffle, int32_t precision);
/* How nsc_conv1d_tc would launch this layer (host logic only, no device work) -- for tests and tuning.  out[0..11] =
 *   kernel family (0 taps-in-N, 1 tap-shift, 2 Toeplitz), staged epilogue, CTA pair (cta_group::2), M tiles per work unit,
 *   MMA-issuing threads, weights resident, weight-ring slots, input stages in flight, dynamic shared memory bytes,
 *   TMEM columns, grid size, work units (tiles, or frames for taps-in-N).  Returns NSC_OK or NSC_E_INVALID. */
int nsc_conv1d_tc_plan_info(int64_t B, int32_t Lin, int32_t Cin, int32_t Cout, int32_t k, int32_t dilation, int32_t stride,
                            int32_t res_mode, int32_t shuffle, int32_t precision, int64_t* out12);
/* Launch plan of the second conv of a bottleneck block (k9 20 -> 20, nn_core_operator.py:64-68) as the codec program runs it at L
 * positions: on FOLDED images (pairs of positions in the channel axis: a 48 -> 48 k5 conv, or block-diagonal k9 for dilation 2 at 256
 * positions) where the frame has 128 folded rows, else the taps-in-N kernel.  out12 as above; out12[1] = 5 / 6 for the two folded forms. */
int nsc_narrow_conv_plan_info(int64_t B, int32_t L, int32_t dilation, int64_t* out12);
int nsc_conv1d_tc(const float* x, const float* w, const float* b, const float* res, float* y, int64_t B, int32_t Lin,
                  int32_t Cin, int32_t Cout, int32_t k, int32_t dilation, int32_t stride, int32_t activation,
                  int32_t res_mode, int32_t post_activation, int32_t shuffle, int32_t precision, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* conv1d_depth (nn_core_operator.py:17-21): Keras SeparableConv1D, depth multiplier 1.
 *   dw (k, Cin, 1), pw (1, Cin, Cout), b (Cout); tmp is a (B, Lout, Cin) scratch buffer. */
int nsc_conv1d_depth(const float* x, const float* dw, const float* pw, const float* b, float* tmp, float* y,
                     int64_t B, int32_t Lin, int32_t Cin, int32_t Cout, int32_t k, int32_t dilation,
                     int32_t stride, int32_t activation, void* stream);

/* the_bottleneck (nn_core_operator.py:57-79) and gated_bottleneck (:82-112) on channels-last tensors.
 *   params: the block's conv parameters back to back in creation order, each kernel (k,cin,cout) then bias.
 *   the_bottleneck : 3 convs;  gated_bottleneck : 4 convs (1x1, k15 left, k15 right(tanh), k_plain).
 *   x (B, L, Cin) with Cin == wide or Cin == 1 (residual add broadcasts, :77);  y (B, L, wide).
 *   workspace: nsc_block_workspace_bytes(B, L, wide, narrow) bytes. */
int64_t nsc_block_workspace_bytes(int64_t B, int32_t L, int32_t wide, int32_t narrow);
int nsc_bottleneck_block(const float* x, const float* params, float* y, int64_t B, int32_t L, int32_t Cin,
                         int32_t wide, int32_t narrow, int32_t k_plain, int32_t k_dilated, int32_t dilation,
                         int32_t is_last_flat, int32_t gated, void* workspace, int64_t workspace_bytes,
                         void* stream);

/* the_bottleneck on the tcgen05 tensor cores (the codec's path): the block's three convs run as ONE persistent launch whose CTAs take
 * three roles (conv 1 / conv 2 / conv 3 + residual) and pass frames through L2-resident rings, so the 20-channel intermediates never
 * reach HBM (nn_core_operator.py:57-79).  x (B, L, wide) channels-last -> y (B, L, wide); params as nsc_bottleneck_block.
 *   Covers narrow = 20, k = 9, dilation 1 or 2, 32 < wide <= 128, L a multiple of 128; precision as nsc_conv1d_tc.
 *   *fused (may be NULL) reports whether the fused kernel ran (it needs an even number of 128/256-position tiles; otherwise the same
 *   three tensor-core kernels run one launch each; *fused == -1 on entry forces that form -- the two are bit-identical). */
int64_t nsc_bottleneck_block_tc_workspace_bytes(int64_t B, int32_t L, int32_t wide, int32_t narrow, int32_t precision);
int nsc_bottleneck_block_tc(const float* x, const float* params, float* y, int64_t B, int32_t L, int32_t wide, int32_t narrow,
                            int32_t k_plain, int32_t k_dilated, int32_t dilation, int32_t is_last_flat, int32_t precision,
                            int32_t* fused, void* workspace, int64_t workspace_bytes, void* stream);

/* gated_bottleneck (nn_core_operator.py:82-112; the reference's shipped block type, constants.py:14) on the tcgen05 tensor cores:
 * k1 conv + leaky ReLU, the two k15 gate convs as ONE layer whose epilogue writes gate * tanh(gate), k9 conv + residual
 * (+ leaky ReLU unless is_last_flat) -- the three launches of the codec program's gated block.  x (B, L, wide) channels-last ->
 * y (B, L, wide); params in creation order (w_1x1, b), (w_left, b), (w_right, b), (w_out, b).
 *   Covers narrow = 20, k_plain = 9, dilation 1 or 2, 32 < wide <= 128, L a multiple of 128; precision as nsc_conv1d_tc. */
int64_t nsc_gated_block_tc_workspace_bytes(int64_t B, int32_t L, int32_t wide, int32_t narrow, int32_t dilation, int32_t precision);
int nsc_gated_block_tc(const float* x, const float* params, float* y, int64_t B, int32_t L, int32_t wide, int32_t narrow,
                       int32_t k_plain, int32_t dilation, int32_t is_last_flat, int32_t precision, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* Tuning aid: per-CTA counters of the most recent fused block launch, recorded when the process runs with NSC_BLOCK_STATS=1
 * (8 words per CTA: role, epilogue-loop cycles, cycles waited for a free ring slot, cycles waited for a ready frame, work units,
 * cycles the storer waited for its own stores, 2 spare).  Synchronises the device.  Returns the number of CTAs copied. */
int nsc_debug_block_stats(unsigned long long* out_host, int32_t max_ctas);

/* ------------------------------------------------------------------------------------------------
 * scalar_softmax_quantization (nn_core_operator.py:140-164) + quan_loss (loss_terms_and_measures.py:257-259)
 * + the soft histogram that entropy_coding_loss (:262-267) reduces.
 *   x (rows = B*L) floating codes; bins (n); alpha: device scalar.
 *   idx[r]   = first argmax_k fp32(alpha * fp32(|x[r]-bins[k]|))   (uint8 when n <= 256, stored as int32 otherwise)
 *   use_soft = the_share (1: value path uses the soft assignment; 0: one-hot)
 *   out[r]   = (1-is_quan_on)*x[r] + is_quan_on*q[r]
 *   soft     (rows, n) or NULL;   hist (n) += sum_r soft[r,:] or NULL (caller zeroes it);
 *   qloss    (B) = mean_L sum_k sqrt(soft+1e-20) or NULL.
 * ---------------------------------------------------------------------------------------------- */
int nsc_quantize_scalar(const float* x, int64_t B, int32_t L, const float* bins, int32_t n, const float* alpha,
                        float is_quan_on, int32_t use_soft, float* out, uint8_t* idx, float* soft, float* hist,
                        float* qloss, void* stream);

/* decode side of a hard code: out[r] = bins[idx[r]] */
int nsc_dequantize_scalar(const uint8_t* idx, int64_t rows, const float* bins, int32_t n, float* out, void* stream);

/* quan_loss (loss_terms_and_measures.py:257-259) on a materialised soft assignment (B, L, n) -> (B). */
int nsc_quan_loss(const float* soft, int64_t B, int32_t L, int32_t n, float* qloss, void* stream);
/* hist (n) += column sums of a materialised soft assignment (rows, n); the caller zeroes hist. */
int nsc_soft_histogram(const float* soft, int64_t rows, int32_t n, float* hist, void* stream);

/* entropy_coding_loss (loss_terms_and_measures.py:262-267) from a soft histogram: one scalar (bits). */
int nsc_entropy_from_hist(const float* hist, int32_t n, float* entropy, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LPC  (lpc_utilities.py)
 * ---------------------------------------------------------------------------------------------- */
/* lpc_analysis_at_test loop body (:112-124): windows (N,1024) -> LSF (N,16) radians ascending.
 *   window = [rising half-Hann(512), 512 ones, falling half-Hann(512)], autocorrelation LPC order 16,
 *   poly2lsf.  float64 arithmetic; lsf_out is float64 like the reference's np.empty array, lsf_out_f32 its float32
 *   cast (what the `lpc_x` float32 placeholder receives, cmrl.py:699-702); either may be NULL.
 *   status (int32, may be NULL) is incremented for every frame that is not analysable (zero energy /
 *   non-minimum-phase); such rows are filled with NaN (the reference raises). */
int nsc_lpc_analyze(const float* windows, int64_t N, double* lsf_out, float* lsf_out_f32, int32_t* status,
                    void* stream);

/* lpc_analysis_at_train (:14-25): frames (B,512) -> highpass, pre-emphasis (zero state), LPC, LSF (B,16). */
int nsc_lpc_analyze_train(const float* frames, int64_t B, double* lsf_out, float* lsf_out_f32, int32_t* status,
                          void* stream);

/* lsf2poly_after_quan (:28-33): lsf (B,16) float32 -> poly (B,17) float32, a[0] = 1.
 *   rows with an LSF outside [0, pi] become NaN and bump *status (the reference raises ValueError). */
int nsc_lsf2poly(const float* lsf, int64_t B, float* poly, int32_t* status, void* stream);

/* lpc_analysis_get_residual (:37-77): x (B,512), poly (B,17) -> residual (B,512);
 *   7 zero-state sub-frame FIRs with Hann overlap-add, float64 accumulation, float32 result. */
int nsc_lpc_residual(const float* x, const float* poly, int64_t B, float* res, void* stream);

/* lpc_synthesizer_tr (:137-156): poly (B,17), res (B,512) -> y (B,512); zero-state all-pole IIR in float64. */
int nsc_lpc_synth(const float* poly, const float* res, int64_t B, float* y, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Losses (loss_terms_and_measures.py)
 * ---------------------------------------------------------------------------------------------- */
/* mel filterbank of mfcc_transform (:130-148): fills melw (257 x 184) = 4 HTK banks {8,16,32,128} side by side. */
#define NSC_MEL_BINS 257
#define NSC_MEL_TOTAL 184
/* melw buffer: 257*184 float weights followed by int32 lo[184], hi[184] (row support of every column) */
#define NSC_MEL_BUFFER_FLOATS (NSC_MEL_BINS * NSC_MEL_TOTAL + 2 * NSC_MEL_TOTAL)
int nsc_mel_filterbank(float* melw, void* stream);
/* mse_loss (:77-79) -> time_loss (B);  mfcc_loss (:151-175) -> freq_loss (B).  Either output may be NULL. */
int nsc_losses_forward(const float* decoded, const float* original, int64_t B, const float* melw,
                       float* time_loss, float* freq_loss, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One codec (neural_speech_coding_module.py:219-335) and the CMRL cascade (cmrl.py:513-543, :770-858)
 * ---------------------------------------------------------------------------------------------- */
typedef struct nsc_codec_cfg {
  int32_t k_dilated;                 /* bottleneck_kernel_and_dilation[0]                         */
  int32_t k_plain;                   /* [1]                                                      */
  int32_t wide;                      /* [2]                                                      */
  int32_t narrow;                    /* [3]                                                      */
  int32_t n_blocks;                  /* len(list) - 4                                            */
  int32_t dilations[NSC_MAX_BLOCKS]; /* [4:]                                                     */
  int32_t n_strides;                 /* the_strides expanded: '2' -> {2}, '4' -> {2,2}           */
  int32_t strides[NSC_MAX_STRIDES];
  int32_t resnet_type;               /* 0 = 'bottleneck', 1 = 'gln' (constants.py:13-14)         */
  int32_t num_bins;                  /* num_bins_for_follower[i]                                 */
  int32_t precision;                 /* conv arithmetic: 0 = fp32 FFMA (CUDA cores, exact fp32)
                                      *   1 = tcgen05 tensor cores, fp16 hi/lo split (3 MMAs, fp32-class results)
                                      *   2 = tcgen05 tensor cores, fp16 inputs / fp32 accumulate (REDUCED precision) */
} nsc_codec_cfg;

/* Flat parameter image of one codec: conv parameters in TF creation order (kernel (k,cin,cout) then bias,
 * separable convs: depthwise (k,cin,1), pointwise (1,cin,cout), bias), then alpha (1), then bins (num_bins). */
int64_t nsc_codec_param_count(const nsc_codec_cfg* cfg);
/* Describes conv layer `i` (creation order): fills k,cin,cout,separable, offset of its first float. Returns
 * the number of conv layers when i < 0. */
int32_t nsc_codec_layer_info(const nsc_codec_cfg* cfg, int32_t i, int32_t* k, int32_t* cin, int32_t* cout,
                             int32_t* separable, int64_t* offset);
int64_t nsc_codec_workspace_bytes(const nsc_codec_cfg* cfg, int64_t B);
/* 1 when this configuration's conv stack runs on the plane engine (tcgen05, fp16 plane images between layers): precision 1 / 2,
 * resnet_type 'bottleneck' or 'gln' (constants.py:13-14), one stride-2 stage (or two with precision 1), narrow 20, k 9,
 * dilations <= 2; 0 when it keeps the layer-by-layer engines; -1 on an invalid configuration. */
int32_t nsc_codec_on_plane_engine(const nsc_codec_cfg* cfg);

/* Whole codec, computational_graph_end2end_quan_on[_lpc]:
 *   x (B,512) -> floating_code (B,Lc) [may be NULL], idx (B,Lc) uint8 [may be NULL], code (B,Lc) [may be NULL],
 *   out (B,512).  soft/hist/qloss as in nsc_quantize_scalar (may be NULL).
 *   in_scale multiplies the input, out_scale the output (res_scalar handling of cmrl.py:810-830). */
int nsc_codec_forward(const nsc_codec_cfg* cfg, const float* params, const float* x, int64_t B, float is_quan_on,
                      int32_t use_soft, float* floating_code, uint8_t* idx, float* code, float* out, float* soft,
                      float* hist, float* qloss, void* workspace, int64_t workspace_bytes, void* stream);
/* Encoder only (x -> floating code -> idx/code) and decoder only (code -> out). */
int nsc_codec_encode(const nsc_codec_cfg* cfg, const float* params, const float* x, int64_t B, float is_quan_on,
                     int32_t use_soft, float* floating_code, uint8_t* idx, float* code, float* soft, float* hist,
                     float* qloss, void* workspace, int64_t workspace_bytes, void* stream);
int nsc_codec_decode(const nsc_codec_cfg* cfg, const float* params, const float* code, int64_t B, float* out,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* Prepared workspaces (serving loops at a fixed batch size).  On the plane engine a call first clears the zero rows of the
 * activation images and packs every layer's weights into fp16 operand slabs -- at streaming batch sizes that costs more than the
 * convs.  nsc_prepare does it once into `workspace` and registers (workspace, kind, cfgs, params pointers, pass size); a later
 * nsc_codec_forward / _encode / _decode (kind 0), nsc_cascade_forward (kind 1) or nsc_cq_forward (kind 2) call with the same
 * workspace, configurations and parameter POINTERS and the same pass size (B below one pass: the same B) skips that part.
 * The caller prepares again after changing the weights in place, and calls nsc_release(workspace) before freeing or reusing
 * the buffer for anything else (returns 1 if it was registered).  Configurations outside the plane engine: no-op.
 * No reference counterpart: TensorFlow keeps its variables resident between sess.run calls (cmrl.py:698-708). */
int nsc_prepare(int32_t kind, const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host, int64_t B,
                void* workspace, int64_t workspace_bytes, void* stream);
int nsc_release(const void* workspace);

/* CMRL cascade, all_modules_feedforward (lpc_variant = 0) / loop of all_modules_feedforward_lpc (= 1):
 *   in_0 = x (times res_scalar if lpc_variant); in_i = res_scalar * (x - sum_{j<i} out_j);
 *   out_i = dec_i(Q(enc_i(in_i))) / res_scalar (codec 0 undivided when lpc_variant = 0); decoded = sum_i out_i.
 *   params_ptrs_host: host array of n_codecs device pointers; cfgs: host array of n_codecs configs.
 *   idx / hist / qloss / outs: host arrays of n_codecs device pointers (array or entries may be NULL). */
int64_t nsc_cascade_workspace_bytes(const nsc_codec_cfg* cfgs, int32_t n_codecs, int64_t B);
int nsc_cascade_forward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                        const float* x, int64_t B, float res_scalar, int32_t lpc_variant, float is_quan_on,
                        int32_t use_soft, uint8_t* const* idx_ptrs_host, float* const* hist_ptrs_host,
                        float* const* qloss_ptrs_host, float* const* outs_ptrs_host, float* decoded,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* Collaborative-quantisation feed-forward (cmrl.py:770-858): LSF codebook -> lsf2poly -> residual -> cascade ->
 * synthesis.  lsf (B,16) float32 (from nsc_lpc_analyze, cast), lsf_params = {alpha, bins[n_lsf_bins]}.
 *   outputs: lsf_idx (B,16) uint8, poly (B,17), res_x (B,512), decoded (B,512), synthesized (B,512); any may be NULL
 *   except decoded.  */
int64_t nsc_cq_workspace_bytes(const nsc_codec_cfg* cfgs, int32_t n_codecs, int64_t B);
/* Frames the engines process per pass for these configurations (larger batches are walked in passes of this size; a workspace is
 * sized for min(B, pass) frames, and nsc_prepare covers every call of at least one pass).  -1 on bad arguments. */
int64_t nsc_pass_frames(const nsc_codec_cfg* cfgs, int32_t n_codecs);
int nsc_cq_forward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                   const float* lsf_params, int32_t n_lsf_bins, const float* x, const float* lsf, int64_t B,
                   float res_scalar, float is_quan_on, int32_t use_soft, uint8_t* lsf_idx, float* lsf_hist,
                   float* lsf_qloss, uint8_t* const* idx_ptrs_host, float* const* hist_ptrs_host,
                   float* const* qloss_ptrs_host, float* poly, float* res_x, float* decoded, float* synthesized,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step of the CQ cascade (loss assembly + optimiser of nscm.py:1033-1059 / cmrl.py:464-490; 'bottleneck'
 * codecs).  Per-quantiser arrays have n_codecs + 1 entries: index 0 = LSF codebook, 1..n = codecs.
 *   nsc_train_forward : soft path (the_share = True) with every activation kept in the workspace.  res_x is FED
 *                       like in the reference's training loop (nscm.py:586-595).  Returns decoded (B,512),
 *                       time_loss / freq_loss (B), quan_loss per quantiser (B each) and the LOCAL soft histograms
 *                       (caller zeroes them; under data parallelism they are all-reduced before the backward call).
 *   nsc_train_backward: gradients of  sum_b [c0*time_b + c1*freq_b + c2*sum_q quan_w[q]*quan_q,b] +
 *                       global_B * tau * sum_q ent_w[q]*H_q(global hist)  w.r.t. every codec's flat parameter image
 *                       (grad_ptrs, same layout as the parameters) and the LSF codebook {alpha, bins} (lsf_grad).
 *                       Must be called with the same workspace right after nsc_train_forward.
 *                       loss_coeff_host = {c0, c1, c2, tau}; trainable_host[q] = 0 skips that quantiser/codec.
 *   nsc_adam_step     : TF1 AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps), t >= 1.
 * ---------------------------------------------------------------------------------------------- */
int64_t nsc_train_workspace_bytes(const nsc_codec_cfg* cfgs, int32_t n_codecs, int64_t B);
int nsc_train_forward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                      const float* lsf_params, int32_t n_lsf_bins, const float* res_x, const float* lsf, int64_t B,
                      float res_scalar, float is_quan_on, const float* melw, float* decoded, float* time_loss,
                      float* freq_loss, float* const* qloss_ptrs_host, float* const* hist_ptrs_host, void* workspace,
                      int64_t workspace_bytes, void* stream);
int nsc_train_backward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                       const float* lsf_params, int32_t n_lsf_bins, const float* res_x, const float* lsf, int64_t B,
                       float res_scalar, float is_quan_on, const float* melw, const float* decoded,
                       const float* loss_coeff_host, const float* quan_w_host, const float* ent_w_host, int64_t global_B,
                       const float* const* hist_global_ptrs_host, const int32_t* trainable_host,
                       float* const* grad_ptrs_host, float* lsf_grad, void* workspace, int64_t workspace_bytes, void* stream);
int nsc_adam_step(float* params, const float* grad, float* m, float* v, int64_t n, float lr, int64_t t, float beta1,
                  float beta2, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Either side of the codec pass: framing, overlap-add, utterance-level filters, code packing
 * ---------------------------------------------------------------------------------------------- */
/* utterance_to_segment (utilities.py:25-39): frames of 512 at hop 480 starting at `offset` (cmrl.py:695 uses 256 on the
 *   LPC path); count = len(range(0, T - offset - 512, 480)) = nsc_segment_count(T - offset).  post_window = 1 copies,
 *   0 multiplies by the trapezoid-Hann `the_window`.  segments (N, 512). */
int64_t nsc_segment_count(int64_t T);
int nsc_utterance_to_segment(const float* utterance, int64_t T, int64_t offset, int32_t post_window, float* segments,
                             void* stream);
/* The 1024-sample analysis windows lpc_analysis_at_test cuts at hop 512 out of the FLATTENED (N, 512) hop-480 frame
 *   matrix (lpc_utilities.py:98-104; the 32 repeated samples per frame boundary are reproduced).  windows (Nw, 1024),
 *   Nw = nsc_lpc_window_count(nsc_segment_count(T)); feed them to nsc_lpc_analyze. */
int64_t nsc_lpc_window_count(int64_t n_segments);
int nsc_lpc_windows(const float* utterance, int64_t T, float* windows, void* stream);
/* hann_process + overlap-add (utilities.py:7-22; cmrl.py:595-597, :710-716): out[480 j : 480 j + 512] += w_j * frames[j]
 *   for j < n_used, w_j = first / last / middle window chosen by (j, seg_amount) exactly like hann_process(seg, j, seg_amount)
 *   (the LPC path passes seg_amount = N but only runs j < N - 2, so its last window is never used).  out (out_len). */
int nsc_overlap_add(const float* frames, int64_t n_used, int64_t seg_amount, float* out, int64_t out_len, void* stream);
/* The same three steps over a BATCH of n_signals equal-length utterances (row-major (n_signals, T)), one launch each:
 *   segments (n_signals, n_take, 512) with n_take <= nsc_segment_count(T - offset) frames per utterance;
 *   windows (n_signals, n_take, 1024) with n_take <= nsc_lpc_window_count(nsc_segment_count(T));
 *   frames (n_signals, n_used, 512) -> out (n_signals, out_len). */
int nsc_utterances_to_segments(const float* utterances, int64_t T, int64_t n_signals, int64_t offset, int32_t post_window,
                               int64_t n_take, float* segments, void* stream);
int nsc_lpc_windows_batch(const float* utterances, int64_t T, int64_t n_signals, int64_t n_take, float* windows, void* stream);
int nsc_overlap_add_batch(const float* frames, int64_t n_signals, int64_t n_used, int64_t seg_amount, float* out, int64_t out_len,
                          void* stream);
/* Zero-state second-order recursive filter over n_signals signals of length T (audiolazy ZFilter call semantics), float64
 *   arithmetic as a chunked parallel scan: y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2]; a[0] must be 1.
 *   highpass_filter / empha_filter / 1/empha_filter (lpc_utilities.py:8-11, cmrl.py:671, :735) are instances.
 *   y_f32 and/or y_f64 receive the result. */
int64_t nsc_iir_workspace_bytes(int64_t T, int64_t n_signals);
int nsc_iir_biquad(const float* x, int64_t T, int64_t n_signals, const double* b_host, const double* a_host, float* y_f32,
                   double* y_f64, void* workspace, int64_t workspace_bytes, void* stream);
/* Fixed-width packing of hard codes: rows of L indices (< 2^bits) -> rows of nsc_packed_row_bytes(L, bits) bytes,
 *   little-endian bit order.  (The reference has no bitstream; it estimates bitrate from entropy,
 *   loss_terms_and_measures.py:63-67.) */
int32_t nsc_packed_row_bytes(int32_t L, int32_t bits);
int nsc_pack_codes(const uint8_t* idx, int64_t rows, int32_t L, int32_t bits, uint8_t* packed, void* stream);
int nsc_unpack_codes(const uint8_t* packed, int64_t rows, int32_t L, int32_t bits, uint8_t* idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSC_B200_H_ */
