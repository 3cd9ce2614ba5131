// Plane engine interface (plane_conv.cu): the conv layers of the codec on tcgen05 with activations kept in HBM as
// fp16 "plane images" -- the exact shared-memory image (K-major, SWIZZLE_128B rows of 64 channels) the tensor core
// reads, so a layer's input tile is staged by ONE bulk copy per 64-channel slab and no thread ever converts or
// transposes an activation on the way in.
//
//   image of one frame = slabs back to back; one slab = (rows + 16) x 128 bytes: 8 zero rows, `rows` positions,
//   8 zero rows (SAME padding and tap halos come from the zero rows); 16-byte chunk c of row r sits at chunk
//   position c ^ (r & 7) (r counted from the start of the slab).
//   planes = 2: value = hi + lo, both fp16 (fp32-class products with 3 MMAs: hi*hi + hi*lo + lo*hi);  planes = 1: hi only.
//   "packed" (narrow tensors): one slab, one 128-byte row per position.
//     planes = 2, C <= 20: the row's 64 halves form ONE K axis on which x * w = hi*wh + lo*wh + hi*wl is a single MMA chain of 4 K
//       steps (instead of three products of 2 K steps each): 16-byte chunks
//         0,1: hi[0:16]   2,3: lo[0:16]   4,5: hi[0:16]   6: hi[16:20] | lo[16:20]   7: hi[16:20] | 0
//       against weight rows   wh[0:16]   wh[0:16]   wl[0:16]   wh[16:20] | wh[16:20]   wl[16:20] | 0   (plane_pack_kernel).
//     planes = 1, C <= 32: hi in bytes 0-63.
//   otherwise: plane-major, ceil(C/64) slabs per plane.
//   deint: two sub-images by position parity (position p -> sub-image p & 1, row p >> 1) -- the input layout of a
//   stride-2 conv, for which every tap is again a pure row shift.
#pragma once
#include "conv.cuh"

namespace nsc {

struct PlaneTensor {
  uint8_t* base = nullptr;   // image of frame 0
  int64_t frame_bytes = 0;
  int rows = 0;              // positions per (sub-)image
  int spp = 1;               // slabs per plane
  int planes = 1;
  int packed = 0;
  int deint = 0;
  int ring = 0;              // > 0 (a power of two): the tensor is a ring of `ring` frame slots, frame f lives in slot f & (ring - 1)
};

// byte offset of frame f's image
__host__ __device__ inline int64_t pt_frame_off(const PlaneTensor& t, int64_t f) {
  return (t.ring > 0 ? (f & (int64_t)(t.ring - 1)) : f) * t.frame_bytes;
}

__host__ __device__ inline int pt_slab_bytes(const PlaneTensor& t) { return (t.rows + 16) * 128; }
__host__ __device__ inline int pt_slab_index(const PlaneTensor& t, int sub, int plane, int s) {
  return (sub * (t.packed ? 1 : t.planes) + plane) * t.spp + s;
}
__host__ __device__ inline int pt_n_slabs(const PlaneTensor& t) { return (t.deint ? 2 : 1) * (t.packed ? 1 : t.planes * t.spp); }

inline PlaneTensor make_plane_tensor(void* base, int L, int C, int planes, int deint) {
  PlaneTensor t;
  const int cpad = (C + 15) & ~15;
  t.base = static_cast<uint8_t*>(base);
  t.rows = deint ? L / 2 : L;
  t.packed = (planes == 2 ? C <= 20 : cpad <= 32) ? 1 : 0;
  t.spp = t.packed ? 1 : (cpad + 63) / 64;
  t.planes = planes;
  t.deint = deint;
  t.frame_bytes = (int64_t)pt_n_slabs(t) * pt_slab_bytes(t);
  return t;
}
inline int64_t plane_tensor_frame_bytes(int L, int C, int planes, int deint) {
  return make_plane_tensor(nullptr, L, C, planes, deint).frame_bytes;
}

enum PlaneKind { PK_T = 0, PK_X = 1, PK_GEN = 2, PK_DW = 3 };

// Work folded into the epilogue of a 1-channel k55 head (PK_T, Cout = 1), bit-identical to the stand-alone kernels it replaces:
//   code head:   the HARD scalar quantiser (nn_core_operator.py:140-164; quantize_kernel's arithmetic): index and blended code
//                straight from the tanh output -- the floating code only goes to HBM if the caller asks for it;
//   output head: the cascade's accumulation decoded (+)= out / res_scalar (cmrl.py:522-531, :822-830; accum_div_kernel) and the
//                per-codec quotient.
struct HeadFold {
  const float* q_bins = nullptr;   // n bins (device); null = no quantiser fold
  const float* q_alpha = nullptr;  // device scalar
  int q_n = 0;
  float q_iq = 1.f;                // is_quan_on
  uint8_t* q_idx = nullptr;        // (B, L) indices, may be null
  float* q_code = nullptr;         // (B, L) code = (1 - iq) x + iq bins[idx], may be null
  float* acc = nullptr;            // (B, L) running sum; null = no accumulation fold
  float* quot = nullptr;           // (B, L) out / div, may be null
  float div = 1.f;
  int acc_first = 0;               // 1: overwrite acc instead of adding
};

// One conv layer of the plane engine.
//   PK_T   "taps in N": P[row, (tap, co)] = sum_ci X[row, ci] W[tap, ci, co] is ONE MMA chain per tile (N = taps * Cout)
//          and the tap sum y[row] = sum_t P[row + shift_t, t] is done across TMEM lanes with warp shuffles.  For the
//          narrow-output layers (Cout = 20, k = 9) and the k55 heads (Cout = 1), where one MMA per tap would be bound by
//          the ~44-cycle issue floor of an N = 32 instruction.
//   PK_X   one MMA per (tap, 16-channel K step): tap = row shift of the A descriptor.  Wide-output layers.
//   PK_GEN PK_X on the Toeplitz matrix of a 1-channel fp32 signal (stem k55 1->100, decoder k9 1->20), built in
//          shared memory by producer warps; taps become the K dimension.
//   PK_DW  depthwise FIR over the rows of a wide image on CUDA cores (the depthwise half of the 'gln' separable up-conv,
//          nscm.py:175-177): w = (K, C) fp32, no bias, no activation; the pointwise half is a PK_X layer with K = 1.
//
// Gated linear unit (gated_bottleneck, nn_core_operator.py:82-112): the two k15 gate convs of a block share their input, so they run
// as ONE PK_X layer with Cout = 40 -- columns [0, 20) the linear gate (w, bias), [20, 40) the tanh gate (w2, bias2) -- whose epilogue
// writes (a + b_a) * tanh(g + b_g) as the 20-channel packed image (`glu`).  A dilation-2 gate conv needs 14-row halos, more than the
// 8 zero rows of an image; it runs instead on the DE-INTERLEAVED image of its input (written that way by the k1 conv in front of it):
// the two sub-images by position parity are independent "frames" of half the length on which the conv has dilation 1 (`bmul` = 2
// frames per codec frame), and the epilogue interleaves the result back (`ileave`: frame f', row r -> frame f' >> 1, position
// 2 r + (f' & 1)).
struct PlaneConv {
  int kind = PK_X;
  int Lin = 0, Cin = 0, Cout = 0, K = 1, dil = 1, stride = 1;
  int act = NSC_ACT_NONE, post_act = NSC_ACT_NONE;
  int res_mode = RES_NONE;       // RES_ADD: `res` planes; RES_ADD_BCAST: `resvec` (B, Lout) fp32 broadcast over channels
  int shuffle = 1;               // sub-pixel factor of the output (1 or 2)
  int planes = 2;
  PlaneTensor in;                // Cin > 1
  const float* xvec = nullptr;   // Cin == 1: (B, Lin) fp32;  x' = xscale * (xvec - xsub)  (xsub may be null)
  const float* xsub = nullptr;
  float xscale = 1.f;
  PlaneTensor out;               // Cout > 1
  float* yvec = nullptr;         // Cout == 1: (B, Lout) fp32
  PlaneTensor res;
  const float* resvec = nullptr;
  const float* w = nullptr;      // (K, Cin, Cout) fp32
  const float* bias = nullptr;   // (Cout)
  void* wpack = nullptr;         // plane_wpack_bytes() bytes, filled by plane_pack_weights()
  int64_t B = 0;
  int glu = 0;                   // PK_X: gated linear unit, Cout = 40 = [linear | tanh] gate, 20-channel packed output
  const float* w2 = nullptr;     // glu: the tanh gate's (K, Cin, 20) kernel and bias
  const float* bias2 = nullptr;
  int ileave = 0;                // the output image interleaves pairs of input frames (see above)
  int bmul = 1;                  // frames of this layer per codec frame (2 for a layer on de-interleaved sub-images)
  HeadFold fold;                 // PK_T heads (Cout = 1)
  // FOLDED narrow images (the 20 -> 20 conv of a bottleneck block as a 48 -> 48 k5 conv on pairs of positions, see below)
  int fold_out = 0;              // PK_T / PK_GEN with Cout = 20: `out` is the folded image of the 20-channel result (1; 2 = its de-interleaved form)
  int fold2 = 0;                 // PK_X 48 -> 48 on a folded image: w / bias are the (9, 20, 20) kernel it is folded from, `out` the plain
                                 // packed image (the epilogue unfolds: row r, channel 24 ph + c -> position 2 r + ph).  1: k5 (dilation 1, or
                                 // dilation 2 per parity with `ileave`); 2: k9 block-diagonal (dilation 2 on the plain folded image, for
                                 // frames too short to split by parity)
};

// ---- the narrow -> narrow conv of a bottleneck block on FOLDED images --------------------------------------------------------
// A k9 20 -> 20 conv has N = 20: one MMA per tap runs at the ~44-cycle issue floor of narrow instructions, and putting the taps in N
// (plane_t_kernel) pays for it with a tap sum across TMEM lanes that is bound by instruction issue.  Folding PAIRS of positions into
// the channel axis turns it into an ordinary conv with more work per instruction and no lane crossing:
//     X'[r, 24 ph' + ci] = x[2 r + ph', ci]            (rows = L / 2, 48 channels: 2 x (20 + 4 zero))
//     Y'[r, 24 ph + co]  = sum_{s = 0..4} sum_{ph', ci} X'[r + s - 2, 24 ph' + ci] W[2 s + ph' - ph, ci, co]     (taps outside 0..8: zero)
// i.e. a 48 -> 48 k5 conv on half-length frames (45 MMAs of N = 48 per 256 positions instead of 36 N = 32 MMAs per 128), run by
// plane_x_kernel.  The folded image is an ordinary unpacked plane image (one 64-channel slab per plane, rows = L / 2) written by the
// epilogue of the conv in front (fold_out); the 48 -> 48 conv's epilogue writes the plain packed image the third conv reads (fold2).
// Dilation 2: the same on the two position parities separately (fold_out = 2 writes the de-interleaved folded image; the conv sees 2 B
// frames of length L / 4 and interleaves the result back, `ileave`).
constexpr int kFoldC = 48, kFoldK = 5;
inline int64_t plane_fold2_scratch_bytes(int K) { return ((int64_t)(K * kFoldC * kFoldC + kFoldC) * 4 + 1023) & ~(int64_t)1023; }

// kernel family of the narrow -> narrow convs (20 -> 20): NSC_PLANE_NARROW=T|X overrides the default
int plane_narrow_kind();
bool plane_conv_supported(const PlaneConv& c);
int64_t plane_wpack_bytes(const PlaneConv& c);
int plane_pack_weights(const PlaneConv& c, cudaStream_t st);
int plane_launch(const PlaneConv& c, cudaStream_t st);
// launch plan of a layer (what plane_launch would do), see nsc_conv1d_tc_plan_info
bool plane_plan_info(const PlaneConv& c, int64_t* out12);

// ---- fused bottleneck block (nn_core_operator.py:57-79): ONE persistent launch runs the block's three convs as three CTA roles
// that stream frames to each other through ring buffers of a few dozen frames (L2-resident: the 20-channel intermediates never
// reach HBM, the block's input is read from HBM once -- the residual read of the third conv hits L2 -- and its output written once).
struct PlaneBlock {
  PlaneConv c1, c2, c3;          // wide -> narrow (PK_T), narrow -> narrow (PK_T grouped), narrow -> wide + residual (PK_X staged, CTA pairs)
  uint32_t* flags = nullptr;     // plane_block_flag_words(B) words, zeroed before the launch
  int ring = 0;                  // frame slots of the two rings (c1.out / c2.in and c2.out / c3.in carry ring = this)
};
bool plane_block_supported(const PlaneBlock& b);
bool plane_block_default_on();   // NSC_BLOCK_FUSED=1: the codec program launches its blocks fused (default: one launch per conv)
int64_t plane_block_flag_words(int64_t B);
int plane_block_ring_frames();
int plane_block_launch(const PlaneBlock& b, cudaStream_t st);

// fp32 <-> plane images (API edges and tests)
int plane_from_f32(const float* x, int x_cl, int64_t B, int L, int C, const PlaneTensor& t, cudaStream_t st);
int plane_to_f32(const PlaneTensor& t, float* y, int y_cl, int64_t B, int L, int C, cudaStream_t st);

}  // namespace nsc
