// Internal interface of the fp32 conv engine (conv.cu) used by the codec program (codec.cu).
#pragma once
#include "common.cuh"

namespace nsc {

enum { RES_NONE = 0, RES_ADD = 1, RES_ADD_BCAST = 2, RES_MUL = 3 };

// One conv layer with its fused epilogue.  Internal activations are channel-major planes ("NCL":
// [frame][channel][position]) so that position tiles are contiguous for both the smem staging loads and
// the vectorised stores; the *_cl flags switch a tensor to the reference's channels-last (B, L, C)
// addressing for API-level buffers.
struct ConvArgs {
  const float* x = nullptr;     // input
  const float* w = nullptr;     // (K, Cin, Cout)  TF kernel layout
  const float* bias = nullptr;  // (Cout)
  const float* res = nullptr;   // residual / gate operand (RES_*), shape of the conv output before shuffling
  float* y = nullptr;           // output
  int64_t B = 0;
  int Lin = 0, Cin = 0, Cout = 0, K = 1, dil = 1, stride = 1;
  int act = NSC_ACT_NONE;       // conv1d's own activation (after bias)
  int res_mode = RES_NONE;
  int post_act = NSC_ACT_NONE;  // after the residual add (activation_func(y + x), nn_core_operator.py:76-79)
  int shuffle = 1;              // sub-pixel factor r: out[b, c/r, l*r + c%r] (nscm.py:158-167)
  int x_cl = 0, y_cl = 0, res_cl = 0;
};

int launch_conv(const ConvArgs& a, cudaStream_t st);

// Tensor-core engine (tc_conv.cu): same tensors and epilogue, tcgen05 implicit GEMM.
//   precision 1 = fp16 hi/lo split, 3 MMAs, fp32-class results; 2 = fp16 inputs (reduced precision).
bool tc_conv_supported(const ConvArgs& a);
int64_t tc_wpack_bytes(int K, int Cin, int Cout);
int launch_conv_tc(const ConvArgs& a, int precision, void* wpack, cudaStream_t st);
constexpr int64_t kTcWpackBytes = 1 << 20;   // scratch for one layer's packed fp16 weights

// depthwise part of SeparableConv1D: y[b,c,p] = sum_t x[b,c,p*s + t*d - padL] * dw[t,c]   (no bias)
int launch_depthwise(const float* x, const float* dw, float* y, int64_t B, int Lin, int C, int K, int dil,
                     int stride, int x_cl, int y_cl, cudaStream_t st);

// y = a * (x - b*z)  (z may be null -> y = a*x), elementwise over n floats
int launch_axpby(float* y, const float* x, float a, const float* z, float b, int64_t n, cudaStream_t st);
// y = x / d
int launch_div(float* y, const float* x, float d, int64_t n, cudaStream_t st);
// weight gradient of a stride-1 conv on the tensor cores (wgrad_tc.cu): dw (K, Cin, Cout) += sum_{b,p} x[b,ci,p + t*dil - padL] g[b,co,p],
// db (Cout) += sum g (db may be null).  x (B, Cin, L * stride), g (B, Cout, L) fp32 NCL; L = output positions (stride 1 or 2).  scratch: wgrad_tc_scratch_floats() floats.
bool wgrad_tc_supported(int64_t B, int Lin, int Lout, int Cin, int Cout, int K, int dil, int stride, int padL);
int64_t wgrad_tc_scratch_floats();
int launch_wgrad_tc(const float* x, const float* g, float* dw, float* db, int64_t B, int L, int Cin, int Cout, int K, int dil, int padL,
                    float* scratch, cudaStream_t st, int stride = 1);

// y = a * b elementwise (the gate product of a gated block on the training tape)
int launch_mul(float* y, const float* a, const float* b, int64_t n, cudaStream_t st);
// acc (+)= x / d ; first = 1 overwrites
int launch_accum_div(float* acc, const float* x, float d, int first, int64_t n, cudaStream_t st);

int launch_quantize(const float* x, int64_t B, int L, const float* bins, int n, const float* alpha, float iq,
                    int use_soft, float* out, uint8_t* idx, float* soft, float* hist, float* qloss,
                    cudaStream_t st);

}  // namespace nsc
