// K3/K4 (fp32 engine) -- 1-D SAME convolution with fused bias / activation / residual / sub-pixel epilogue
// (nn_core_operator.py:6-14 and the epilogues of :57-112, nscm.py:152-181).
//
// This is the exact-fp32 path: FFMA on CUDA cores, results equal to the fp32 reference up to summation order.
// Tiling: one CTA owns TP output positions of one frame x all output channels.  The input rows it needs
// ((TP-1)*stride + (K-1)*dil + 1 of them) and the weights of a chunk of CC input channels are staged in shared
// memory; each thread owns a PT x CT register tile and, per input channel, pulls ONE sliding window of
// (PT-1)*stride + (K-1)*dil + 1 activations into registers with 128-bit loads and reuses it across all K taps
// (K, dil and stride are template parameters so the window is indexed statically).  Per input channel a thread
// issues ~(WIN/4 + K*CT/4) LDS.128 for K*PT*CT FFMA (k9, 8x8 tile: 24 loads for 576 FFMA).
// Activations between layers are channel-major planes [frame][channel][position]: tile rows are contiguous
// in HBM/L2 for the staging loads and every thread stores runs of PT consecutive positions.
#include "conv.cuh"

namespace nsc {

namespace {

struct ConvLaunch {
  ConvArgs a;
  int Lout, padL;
  int txn, tyn;     // thread grid inside the CTA: txn * tyn threads are active
  int tiles;        // position tiles per frame
  int cc;           // input channels staged per chunk
  int xr;           // padded row length of the staged input tile (floats, multiple of 4)
  int64_t xs_b, xs_c, xs_l;  // input strides (frame, channel, position)
  int64_t ys_b, ys_c, ys_l;  // output strides
  int64_t rs_b, rs_c, rs_l;  // residual strides
  int Lout_y, Cout_y;        // output tensor dims after the sub-pixel shuffle
};

template <int PT, int CT>
__device__ __forceinline__ void conv_epilogue(const ConvLaunch& p, float (&acc)[PT][CT], int64_t b, int p0, int co0) {
  const ConvArgs& a = p.a;
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    const int co = co0 + j;
    if (co >= a.Cout) continue;
    const float bv = a.bias ? a.bias[co] : 0.f;
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      const int pos = p0 + i;
      if (pos >= p.Lout) continue;
      float v = apply_act(acc[i][j] + bv, a.act);
      if (a.res_mode != RES_NONE) {
        const int rc = a.res_mode == RES_ADD_BCAST ? 0 : co;
        const float r = a.res[b * p.rs_b + rc * p.rs_c + pos * p.rs_l];
        v = a.res_mode == RES_MUL ? v * r : v + r;
      }
      acc[i][j] = apply_act(v, a.post_act);
    }
  }
  // ---- store
  const bool vec_ok = (p.ys_l == 1) && (p0 + PT <= p.Lout) && (PT % 4 == 0);
  if (a.shuffle == 1) {
    if (vec_ok && ((p.Lout_y & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < CT; ++j) {
        const int co = co0 + j;
        if (co >= a.Cout) continue;
        float4* dst = reinterpret_cast<float4*>(a.y + b * p.ys_b + co * p.ys_c + p0);
#pragma unroll
        for (int i = 0; i < PT; i += 4) dst[i / 4] = make_float4(acc[i][j], acc[i + 1][j], acc[i + 2][j], acc[i + 3][j]);
      }
    } else if (p.ys_c == 1 && CT % 4 == 0 && (a.Cout & 3) == 0) {
      // channels-last output: CT consecutive channels per position
#pragma unroll
      for (int i = 0; i < PT; ++i) {
        const int pos = p0 + i;
        if (pos >= p.Lout) continue;
#pragma unroll
        for (int j = 0; j < CT; j += 4) {
          if (co0 + j >= a.Cout) continue;
          *reinterpret_cast<float4*>(a.y + b * p.ys_b + pos * p.ys_l + co0 + j) =
              make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < CT; ++j) {
        const int co = co0 + j;
        if (co >= a.Cout) continue;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
          const int pos = p0 + i;
          if (pos < p.Lout) a.y[b * p.ys_b + co * p.ys_c + pos * p.ys_l] = acc[i][j];
        }
      }
    }
  } else {
    // sub-pixel: out[b, co / r, pos * r + co % r]
    const int r = a.shuffle;
    if (r == 2 && vec_ok && CT % 2 == 0 && ((p.Lout_y & 3) == 0)) {
#pragma unroll
      for (int j = 0; j < CT; j += 2) {
        const int co = co0 + j;
        if (co >= a.Cout) continue;  // Cout is even here, so co + 1 is valid too
        float4* dst = reinterpret_cast<float4*>(a.y + b * p.ys_b + (co >> 1) * p.ys_c + 2 * p0);
#pragma unroll
        for (int i = 0; i < PT; i += 2)
          dst[i / 2] = make_float4(acc[i][j], acc[i][j + 1], acc[i + 1][j], acc[i + 1][j + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < CT; ++j) {
        const int co = co0 + j;
        if (co >= a.Cout) continue;
#pragma unroll
        for (int i = 0; i < PT; ++i) {
          const int pos = p0 + i;
          if (pos < p.Lout) a.y[b * p.ys_b + (co / r) * p.ys_c + (int64_t)(pos * r + co % r) * p.ys_l] = acc[i][j];
        }
      }
    }
  }
}

// Stage the input rows and the weight slice of input channels [ci0, ci0+ccur) into shared memory.
__device__ __forceinline__ void conv_stage(const ConvLaunch& p, float* xs, float* ws, int64_t b, int in0, int ci0,
                                           int ccur, int WN) {
  const ConvArgs& a = p.a;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* xb = a.x + b * p.xs_b;
  if (p.xs_l == 1) {
    for (int idx = tid; idx < ccur * p.xr; idx += nt) {
      const int c = idx / p.xr, r = idx - c * p.xr;
      const int g = in0 + r;
      xs[idx] = (g >= 0 && g < a.Lin) ? xb[(int64_t)(ci0 + c) * p.xs_c + g] : 0.f;
    }
  } else {  // channels-last input: walk channels fastest for coalescing
    for (int idx = tid; idx < ccur * p.xr; idx += nt) {
      const int r = idx / ccur, c = idx - r * ccur;
      const int g = in0 + r;
      xs[c * p.xr + r] = (g >= 0 && g < a.Lin) ? xb[(int64_t)g * p.xs_l + (ci0 + c) * p.xs_c] : 0.f;
    }
  }
  const int per_tap = ccur * WN;
  for (int idx = tid; idx < a.K * per_tap; idx += nt) {
    const int t = idx / per_tap, rem = idx - t * per_tap;
    const int c = rem / WN, n = rem - c * WN;
    ws[(t * p.cc + c) * WN + n] = n < a.Cout ? a.w[((int64_t)t * a.Cin + ci0 + c) * a.Cout + n] : 0.f;
  }
}

template <int K, int DIL, int STRIDE, int PT, int CT>
__global__ void __launch_bounds__(256) conv_tile_kernel(const __grid_constant__ ConvLaunch p) {
  constexpr int WIN = (PT - 1) * STRIDE + (K - 1) * DIL + 1;
  constexpr int WINP = (WIN + 3) & ~3;
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);
  float* ws = xs + p.cc * p.xr;
  const ConvArgs& a = p.a;
  const int WN = p.txn * CT;
  const int tile = blockIdx.x % p.tiles;
  const int64_t b = blockIdx.x / p.tiles;
  const int tid = threadIdx.x;
  const int tx = tid % p.txn, ty = tid / p.txn;
  const bool active = ty < p.tyn;
  const int TP = p.tyn * PT;
  const int p0 = tile * TP + ty * PT;
  const int in0 = tile * TP * STRIDE - p.padL;

  float acc[PT][CT];
#pragma unroll
  for (int i = 0; i < PT; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;

  for (int ci0 = 0; ci0 < a.Cin; ci0 += p.cc) {
    const int ccur = min(p.cc, a.Cin - ci0);
    conv_stage(p, xs, ws, b, in0, ci0, ccur, WN);
    __syncthreads();
    if (active) {
      for (int c = 0; c < ccur; ++c) {
        float win[WINP];
        const float4* src = reinterpret_cast<const float4*>(xs + c * p.xr + ty * PT * STRIDE);
#pragma unroll
        for (int q = 0; q < WINP / 4; ++q) {
          const float4 v = src[q];
          win[4 * q] = v.x; win[4 * q + 1] = v.y; win[4 * q + 2] = v.z; win[4 * q + 3] = v.w;
        }
        const float* wrow = ws + c * WN + tx * CT;
#pragma unroll
        for (int t = 0; t < K; ++t) {
          float wv[CT];
          const float* wp = wrow + t * p.cc * WN;
          if (CT % 4 == 0) {
#pragma unroll
            for (int q = 0; q < CT / 4; ++q) {
              const float4 v = reinterpret_cast<const float4*>(wp)[q];
              wv[4 * q] = v.x; wv[4 * q + 1] = v.y; wv[4 * q + 2] = v.z; wv[4 * q + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CT; ++j) wv[j] = wp[j];
          }
#pragma unroll
          for (int i = 0; i < PT; ++i)
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(win[i * STRIDE + t * DIL], wv[j], acc[i][j]);
        }
      }
    }
    __syncthreads();
  }
  if (active) conv_epilogue<PT, CT>(p, acc, b, p0, tx * CT);
}

// Any (K, dil, stride): same staging, taps indexed at run time (no register window).
template <int PT, int CT>
__global__ void __launch_bounds__(256) conv_tile_generic_kernel(const __grid_constant__ ConvLaunch p) {
  extern __shared__ float4 smem4[];
  float* xs = reinterpret_cast<float*>(smem4);
  float* ws = xs + p.cc * p.xr;
  const ConvArgs& a = p.a;
  const int WN = p.txn * CT;
  const int tile = blockIdx.x % p.tiles;
  const int64_t b = blockIdx.x / p.tiles;
  const int tid = threadIdx.x;
  const int tx = tid % p.txn, ty = tid / p.txn;
  const bool active = ty < p.tyn;
  const int TP = p.tyn * PT;
  const int p0 = tile * TP + ty * PT;
  const int in0 = tile * TP * a.stride - p.padL;
  float acc[PT][CT];
#pragma unroll
  for (int i = 0; i < PT; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[i][j] = 0.f;
  for (int ci0 = 0; ci0 < a.Cin; ci0 += p.cc) {
    const int ccur = min(p.cc, a.Cin - ci0);
    conv_stage(p, xs, ws, b, in0, ci0, ccur, WN);
    __syncthreads();
    if (active) {
      for (int c = 0; c < ccur; ++c) {
        const float* xrow = xs + c * p.xr + ty * PT * a.stride;
        for (int t = 0; t < a.K; ++t) {
          const float* wp = ws + (t * p.cc + c) * WN + tx * CT;
          float wv[CT];
#pragma unroll
          for (int j = 0; j < CT; ++j) wv[j] = wp[j];
#pragma unroll
          for (int i = 0; i < PT; ++i) {
            const float xv = xrow[i * a.stride + t * a.dil];
#pragma unroll
            for (int j = 0; j < CT; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (active) conv_epilogue<PT, CT>(p, acc, b, p0, tx * CT);
}

template <typename KernelT>
int launch_one(KernelT kernel, const ConvLaunch& p, int threads, size_t smem, cudaStream_t st) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(e));
      return NSC_E_CUDA;
    }
  }
  const int64_t grid = (int64_t)p.tiles * p.a.B;
  if (grid > 0x7fffffffLL) {
    set_error("conv grid too large (%lld CTAs)", (long long)grid);
    return NSC_E_INVALID;
  }
  char name[32];
  snprintf(name, sizeof(name), "conv_k%dd%ds%d_c%dto%d", p.a.K, p.a.dil, p.a.stride, p.a.Cin, p.a.Cout);
  const double macs = (double)p.a.B * p.Lout * p.a.K * p.a.Cin * p.a.Cout;
  const double bytes = 4.0 * ((double)p.a.B * ((double)p.a.Lin * p.a.Cin + (double)p.Lout * p.a.Cout) +
                              (double)p.a.K * p.a.Cin * p.a.Cout);
  ProfScope prof(st, name, 2.0 * macs, bytes);
  kernel<<<(unsigned)grid, threads, smem, st>>>(p);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

__global__ void depthwise_kernel(const float* __restrict__ x, const float* __restrict__ dw, float* __restrict__ y,
                                 int64_t B, int Lin, int Lout, int C, int K, int dil, int stride, int padL,
                                 int x_cl, int y_cl) {
  const int64_t total = B * C * (int64_t)Lout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b;
    int c, p;
    if (y_cl) { c = (int)(i % C); p = (int)((i / C) % Lout); b = i / ((int64_t)C * Lout); }
    else { p = (int)(i % Lout); c = (int)((i / Lout) % C); b = i / ((int64_t)C * Lout); }
    float acc = 0.f;
    for (int t = 0; t < K; ++t) {
      const int g = p * stride + t * dil - padL;
      if (g < 0 || g >= Lin) continue;
      const float xv = x_cl ? x[(b * Lin + g) * C + c] : x[(b * C + c) * Lin + g];
      acc = fmaf(xv, dw[t * C + c], acc);
    }
    y[i] = acc;
  }
}

// y = a * (x - b*z)  (cascade input, cmrl.py:529-531 / :822-823), or a * x when z is null
// y and z may be the SAME buffer (train.cu accumulates in place): neither is declared __restrict__
__global__ void axpby_kernel(float* y, const float* __restrict__ x, float a, const float* z, float b, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = z ? a * (x[i] - z[i] * b) : a * x[i];
}

__global__ void accum_div_kernel(float* __restrict__ acc, const float* __restrict__ x, float d, int first, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i] / d;
    acc[i] = first ? v : acc[i] + v;
  }
}

__global__ void mul_kernel(float* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = a[i] * b[i];
}

__global__ void div_kernel(float* __restrict__ y, const float* __restrict__ x, float d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] / d;
}

inline unsigned ew_grid(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

int launch_conv(const ConvArgs& a, cudaStream_t st) {
  if (a.B == 0) return NSC_OK;   // empty batch
  NSC_CHECK_ARG(a.x && a.w && a.y, "conv: null pointer");
  NSC_CHECK_ARG(a.Lin >= 1 && a.Cin >= 1 && a.Cout >= 1 && a.K >= 1 && a.dil >= 1 && a.stride >= 1,
                "conv: bad shape Lin=%d Cin=%d Cout=%d K=%d dil=%d stride=%d", a.Lin, a.Cin, a.Cout, a.K, a.dil, a.stride);
  NSC_CHECK_ARG(a.shuffle >= 1 && a.Cout % a.shuffle == 0, "conv: Cout=%d not divisible by shuffle=%d", a.Cout, a.shuffle);
  NSC_CHECK_ARG(a.res_mode == RES_NONE || a.res != nullptr, "conv: residual mode without residual pointer");
  if (a.B == 0) return NSC_OK;
  ConvLaunch p;
  p.a = a;
  same_padding(a.Lin, a.K, a.dil, a.stride, &p.Lout, &p.padL);
  p.Lout_y = p.Lout * a.shuffle;
  p.Cout_y = a.Cout / a.shuffle;
  if (a.x_cl) { p.xs_b = (int64_t)a.Lin * a.Cin; p.xs_l = a.Cin; p.xs_c = 1; }
  else { p.xs_b = (int64_t)a.Lin * a.Cin; p.xs_c = a.Lin; p.xs_l = 1; }
  if (a.y_cl) { p.ys_b = (int64_t)p.Lout_y * p.Cout_y; p.ys_l = p.Cout_y; p.ys_c = 1; }
  else { p.ys_b = (int64_t)p.Lout_y * p.Cout_y; p.ys_c = p.Lout_y; p.ys_l = 1; }
  const int Cres = a.res_mode == RES_ADD_BCAST ? 1 : a.Cout;
  if (a.res_cl) { p.rs_b = (int64_t)p.Lout * Cres; p.rs_l = Cres; p.rs_c = 1; }
  else { p.rs_b = (int64_t)p.Lout * Cres; p.rs_c = p.Lout; p.rs_l = 1; }

  const int CT = a.Cout == 1 ? 1 : (a.Cout <= 32 ? 4 : 8);
  const int PT = CT == 1 ? 4 : 8;
  p.txn = ceil_div(a.Cout, CT);
  NSC_CHECK_ARG(p.txn <= 256, "conv: Cout=%d too wide for the tile kernel", a.Cout);
  int tyn = 1;
  while (tyn * 2 * p.txn <= 256 && tyn * PT < p.Lout) tyn *= 2;
  p.tyn = tyn;
  const int TP = tyn * PT;
  p.tiles = ceil_div(p.Lout, TP);
  const int WN = p.txn * CT;
  const int WIN = (PT - 1) * a.stride + (a.K - 1) * a.dil + 1;
  const int WINP = (WIN + 3) & ~3;
  p.xr = (((tyn - 1) * PT * a.stride + WINP) + 3) & ~3;
  int cc = a.Cin;
  const int cc_w = (40 * 1024) / (a.K * WN * 4);
  const int cc_x = (24 * 1024) / (p.xr * 4);
  if (cc > cc_w) cc = cc_w;
  if (cc > cc_x) cc = cc_x;
  if (cc < 1) cc = 1;
  // prefer an even split of Cin
  const int nchunk = ceil_div(a.Cin, cc);
  cc = ceil_div(a.Cin, nchunk);
  p.cc = cc;
  const size_t smem = sizeof(float) * ((size_t)cc * p.xr + (size_t)a.K * cc * WN);
  NSC_CHECK_ARG(smem <= 200 * 1024, "conv: tile needs %zu bytes of shared memory", smem);
  const int threads = ((p.txn * p.tyn + 31) / 32) * 32;

#define NSC_CONV_CASE(KK, DD, SS)                                                                             \
  if (a.K == KK && a.dil == DD && a.stride == SS) {                                                           \
    if (CT == 8) return launch_one(conv_tile_kernel<KK, DD, SS, 8, 8>, p, threads, smem, st);                  \
    if (CT == 4) return launch_one(conv_tile_kernel<KK, DD, SS, 8, 4>, p, threads, smem, st);                  \
    return launch_one(conv_tile_kernel<KK, DD, SS, 4, 1>, p, threads, smem, st);                               \
  }
  NSC_CONV_CASE(9, 1, 1)
  NSC_CONV_CASE(9, 2, 1)
  NSC_CONV_CASE(9, 1, 2)
  NSC_CONV_CASE(55, 1, 1)
  NSC_CONV_CASE(15, 1, 1)
  NSC_CONV_CASE(15, 2, 1)
  NSC_CONV_CASE(1, 1, 1)
  NSC_CONV_CASE(5, 1, 1)    // data gradient of the stride-2 k9 conv as a sub-pixel conv (train.cu)
#undef NSC_CONV_CASE
  if (CT == 8) return launch_one(conv_tile_generic_kernel<8, 8>, p, threads, smem, st);
  if (CT == 4) return launch_one(conv_tile_generic_kernel<8, 4>, p, threads, smem, st);
  return launch_one(conv_tile_generic_kernel<4, 1>, p, threads, smem, st);
}

int launch_depthwise(const float* x, const float* dw, float* y, int64_t B, int Lin, int C, int K, int dil,
                     int stride, int x_cl, int y_cl, cudaStream_t st) {
  int Lout, padL;
  same_padding(Lin, K, dil, stride, &Lout, &padL);
  const int64_t total = B * C * (int64_t)Lout;
  if (total == 0) return NSC_OK;
  ProfScope prof(st, "depthwise", 2.0 * (double)total * K, 8.0 * (double)total);
  depthwise_kernel<<<ew_grid(total), 256, 0, st>>>(x, dw, y, B, Lin, Lout, C, K, dil, stride, padL, x_cl, y_cl);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int launch_axpby(float* y, const float* x, float a, const float* z, float b, int64_t n, cudaStream_t st) {
  if (n == 0) return NSC_OK;
  ProfScope prof(st, "cascade_input", 2.0 * (double)n, (z ? 12.0 : 8.0) * (double)n);
  axpby_kernel<<<ew_grid(n), 256, 0, st>>>(y, x, a, z, b, n);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int launch_mul(float* y, const float* a, const float* b, int64_t n, cudaStream_t st) {
  if (n == 0) return NSC_OK;
  ProfScope prof(st, "gate_product", (double)n, 12.0 * (double)n);
  mul_kernel<<<ew_grid(n), 256, 0, st>>>(y, a, b, n);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int launch_div(float* y, const float* x, float d, int64_t n, cudaStream_t st) {
  if (n == 0) return NSC_OK;
  ProfScope prof(st, "cascade_div", (double)n, 8.0 * (double)n);
  div_kernel<<<ew_grid(n), 256, 0, st>>>(y, x, d, n);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int launch_accum_div(float* acc, const float* x, float d, int first, int64_t n, cudaStream_t st) {
  if (n == 0) return NSC_OK;
  ProfScope prof(st, "cascade_accum", 2.0 * (double)n, (first ? 8.0 : 12.0) * (double)n);
  accum_div_kernel<<<ew_grid(n), 256, 0, st>>>(acc, x, d, first, n);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // namespace nsc

extern "C" {

int nsc_conv1d(const float* x, const float* w, const float* b, float* y, int64_t B, int32_t Lin, int32_t Cin,
               int32_t Cout, int32_t k, int32_t dilation, int32_t stride, int32_t activation, void* stream) {
  NSC_CHECK_ARG(!(dilation > 1 && stride > 1), "nsc_conv1d: strides > 1 with dilation_rate > 1 is not supported by TF either");
  nsc::ConvArgs a;
  a.x = x; a.w = w; a.bias = b; a.y = y;
  a.B = B; a.Lin = Lin; a.Cin = Cin; a.Cout = Cout; a.K = k; a.dil = dilation; a.stride = stride;
  a.act = activation;
  a.x_cl = 1; a.y_cl = 1;
  return nsc::launch_conv(a, (cudaStream_t)stream);
}

int nsc_conv1d_depth(const float* x, const float* dw, const float* pw, const float* b, float* tmp, float* y,
                     int64_t B, int32_t Lin, int32_t Cin, int32_t Cout, int32_t k, int32_t dilation,
                     int32_t stride, int32_t activation, void* stream) {
  NSC_CHECK_ARG(x && dw && pw && tmp && y, "nsc_conv1d_depth: null pointer");
  NSC_TRY(nsc::launch_depthwise(x, dw, tmp, B, Lin, Cin, k, dilation, stride, 1, 1, (cudaStream_t)stream));
  int Lout, padL;
  nsc::same_padding(Lin, k, dilation, stride, &Lout, &padL);
  nsc::ConvArgs a;
  a.x = tmp; a.w = pw; a.bias = b; a.y = y;
  a.B = B; a.Lin = Lout; a.Cin = Cin; a.Cout = Cout; a.K = 1; a.dil = 1; a.stride = 1;
  a.act = activation;
  a.x_cl = 1; a.y_cl = 1;
  return nsc::launch_conv(a, (cudaStream_t)stream);
}

}  // extern "C"
