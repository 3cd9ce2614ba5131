// K9 -- the training step of the CQ cascade: forward that keeps every activation,
// reverse-mode backward of every kernel on the path, and the TF1-style Adam update.
//
// Semantics reproduced from the reference graphs (nscm.py:1033-1059, cmrl.py:464-490):
//   * soft value path (the_share = True), is_quan_on blend, res_scalar in / out scaling;
//   * per-frame losses are VECTORS of shape (B,) and `minimize` differentiates their SUM over the batch; the scalar
//     entropy term is broadcast, i.e. counted B times (B = global batch across ranks);
//   * no gradient flows through lsf2poly / residual / synthesis (tf.py_func): the LSF codebook only sees quan_loss
//     and entropy_coding_loss of its own soft assignment; res_x is FED (nscm.py:586-595), exactly like here;
//   * Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t * m / (sqrt(v) + eps)   (epsilon outside the bias correction).
// Data-parallel split: nsc_train_forward returns the LOCAL soft histograms; the caller all-reduces them (tiny) so
// that the batch-global entropy is exact, calls nsc_train_backward, all-reduces (SUM) the flat gradient buffers and
// calls nsc_adam_step.  The library itself never communicates.
//
// Backward kernels: the data gradient of a stride-1 conv is the forward conv engine run on flipped / transposed
// weights; the weight gradient is an im2col-on-the-fly SGEMM reduced over (frame, position) with one atomic flush
// per CTA; activation / residual / sub-pixel epilogues are undone by one elementwise kernel that reads the sign of
// the stored OUTPUT (leaky-ReLU and tanh derivatives are functions of the output).
#include <limits.h>

#include "walker.cuh"

namespace nsc {
namespace {

// ------------------------------------------------------------------------------------------------ elementwise
// gy (grad w.r.t. the stored layer output y, NCL, shuffled layout when r > 1) -> gpre (grad w.r.t. conv + bias, [Cout][L]);
// the residual branch receives gz (before the conv's own activation is undone it equals the post-activation grad).
__global__ void epilogue_backward_kernel(const float* __restrict__ gy, const float* __restrict__ y, float* __restrict__ gpre,
                                         float* __restrict__ gres, int64_t B, int L, int Cout, int act, int res_mode,
                                         int post_act, int r) {
  const int64_t total = B * (int64_t)L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / L;
    const int p = (int)(i - b * L);
    const int Ly = L * r, Cy = Cout / r;
    float bsum = 0.f;
    for (int co = 0; co < Cout; ++co) {
      const int64_t yi = r == 1 ? (b * Cout + co) * (int64_t)L + p : (b * Cy + co / r) * (int64_t)Ly + (int64_t)p * r + (co % r);
      const float yv = y[yi];
      float g = gy[yi];
      if (post_act == NSC_ACT_LRELU) g *= (yv > 0.f ? 1.f : kLeakySlope);
      else if (post_act == NSC_ACT_TANH) g *= (1.f - yv * yv);
      if (res_mode == RES_ADD) gres[(b * Cout + co) * (int64_t)L + p] += g;
      else if (res_mode == RES_ADD_BCAST) bsum += g;
      // the conv's own activation (only present on layers without a residual, so y = act(pre))
      if (act == NSC_ACT_LRELU) g *= (yv > 0.f ? 1.f : kLeakySlope);
      else if (act == NSC_ACT_TANH) g *= (1.f - yv * yv);
      gpre[(b * Cout + co) * (int64_t)L + p] = g;
    }
    if (res_mode == RES_ADD_BCAST) gres[b * (int64_t)L + p] += bsum;
  }
}

// Wf[t][co][ci] = W[K-1-t][ci][co]: the kernel of the data-gradient conv (stride 1, odd K => same SAME padding)
__global__ void flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wf, int K, int Cin, int Cout) {
  const int total = K * Cin * Cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin, co = (i / Cin) % Cout, t = i / (Cin * Cout);
    wf[i] = w[((int64_t)(K - 1 - t) * Cin + ci) * Cout + co];
  }
}

// data gradient of a strided conv (the down-sampling layer): gx[ci][u] += sum_{t,co : (u+padL-t*d) % s == 0} g[co][(u+padL-t*d)/s] W[t][ci][co]
__global__ void dgrad_strided_kernel(const float* __restrict__ g, const float* __restrict__ w, float* __restrict__ gx,
                                     int64_t B, int Lin, int Lout, int Cin, int Cout, int K, int dil, int stride, int padL) {
  const int64_t total = B * (int64_t)Cin * Lin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(i % Lin);
    const int ci = (int)((i / Lin) % Cin);
    const int64_t b = i / ((int64_t)Lin * Cin);
    float acc = 0.f;
    for (int t = 0; t < K; ++t) {
      const int num = u + padL - t * dil;
      if (num < 0 || num % stride != 0) continue;
      const int p = num / stride;
      if (p >= Lout) continue;
      const float* gp = g + b * (int64_t)Cout * Lout + p;
      const float* wp = w + ((int64_t)t * Cin + ci) * Cout;
      for (int co = 0; co < Cout; ++co) acc = fmaf(gp[(int64_t)co * Lout], wp[co], acc);
    }
    gx[i] += acc;
  }
}

// The data gradient of the stride-2 k9 conv (padL = 3) as a SUB-PIXEL conv on the forward engine: input position u = 2 v + par takes
// g[co][v + q] W[par + padL - 2 q][ci][co] for q in [-2, 2], i.e. a stride-1 k5 conv from Cout to 2 Cin channels (c' = 2 ci + par)
// whose output is pixel-shuffled by 2.  Wt[j][co][2 ci + par] = W[par + padL - 2 (j - 2)][ci][co], zero outside the kernel.
__global__ void subpixel_dgrad_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int Cin, int Cout, int padL) {
  const int total = 5 * Cout * 2 * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c2 = i % (2 * Cin), co = (i / (2 * Cin)) % Cout, j = i / (2 * Cin * Cout);
    const int ci = c2 >> 1, par = c2 & 1;
    const int t = par + padL - 2 * (j - 2);
    wt[i] = (t >= 0 && t < K) ? w[((int64_t)t * Cin + ci) * Cout + co] : 0.f;
  }
}

// The same sub-pixel data gradient as TWO dense k5 convs Cout -> Cin, one per input-position parity (the tensor engine's layers have at
// most 128 output channels): wt[par][j][co][ci] = W[par + padL - 2 (j - 2)][ci][co]; their outputs are interleaved into the gradient.
__global__ void subpixel_dgrad_weights_split_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int Cin, int Cout, int padL) {
  const int total = 2 * 5 * Cout * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ci = i % Cin, co = (i / Cin) % Cout, j = (i / (Cin * Cout)) % 5, par = i / (5 * Cin * Cout);
    const int t = par + padL - 2 * (j - 2);
    wt[i] = (t >= 0 && t < K) ? w[((int64_t)t * Cin + ci) * Cout + co] : 0.f;
  }
}
// dst (B, C, 2 L)[b][c][2 v + par] += src[par] (B, C, L)[b][c][v]
__global__ void add_interleaved_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t rows, int L) {
  const int64_t total = rows * 2 * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (2 * L);
    const int u = (int)(i - row * 2 * L);
    dst[i] += src[(int64_t)(u & 1) * rows * L + row * L + (u >> 1)];
  }
}

// gate product y = a * b (gated_bottleneck, nn_core_operator.py:102): ga += gy * b, gb += gy * a
__global__ void mul_backward_kernel(const float* __restrict__ gy, const float* __restrict__ a, const float* __restrict__ b,
                                    float* __restrict__ ga, float* __restrict__ gb, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = gy[i];
    ga[i] += g * b[i];
    gb[i] += g * a[i];
  }
}

// depthwise half of the separable up-conv (y[b,c,p] = sum_t x[b,c,p + t - padL] w[t,c]): dw[t,c] += sum_{b,p} g[b,c,p] x[b,c,p + t - padL].
// grid (C, splits); every thread keeps the K partial sums of its (frame, position) share, the block reduces and adds them atomically.
constexpr int kDwMaxK = 17;
__global__ void __launch_bounds__(256) depthwise_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dw,
                                                              int64_t B, int L, int C, int K, int padL, int frames_per_split) {
  const int c = blockIdx.x;
  const int64_t b0 = (int64_t)blockIdx.y * frames_per_split;
  const int64_t b1 = b0 + frames_per_split < B ? b0 + frames_per_split : B;
  float acc[kDwMaxK];
#pragma unroll
  for (int t = 0; t < kDwMaxK; ++t) acc[t] = 0.f;
  for (int64_t b = b0; b < b1; ++b) {
    const float* xr = x + (b * C + c) * (int64_t)L;
    const float* gr = g + (b * C + c) * (int64_t)L;
    for (int p = threadIdx.x; p < L; p += blockDim.x) {
      const float gv = gr[p];
#pragma unroll
      for (int t = 0; t < kDwMaxK; ++t) {
        const int q = p + t - padL;
        if (t < K && q >= 0 && q < L) acc[t] = fmaf(gv, xr[q], acc[t]);
      }
    }
  }
  __shared__ float red[8][kDwMaxK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < kDwMaxK; ++t) {
    float v = acc[t];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float v = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) v += red[w8][threadIdx.x];
    atomicAdd(dw + (int64_t)threadIdx.x * C + c, v);
  }
}

// taps of the depthwise data gradient: wf[t][c] = w[K - 1 - t][c] (odd K: same SAME padding)
__global__ void flip_taps_kernel(const float* __restrict__ w, float* __restrict__ wf, int K, int C) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * C; i += gridDim.x * blockDim.x) {
    const int t = i / C, c = i - t * C;
    wf[i] = w[(int64_t)(K - 1 - t) * C + c];
  }
}

__global__ void add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = reinterpret_cast<const float4*>(src)[i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[(t,ci)][co] += sum_{b,p} x[b][ci][p*s + t*d - padL] * g[b][co][p] : a (K*Cin) x Cout x (B*Lout) GEMM whose A operand is
// gathered on the fly.  A CTA owns a (8 TMW) x (NT TNW) tile of dW, thread (tm, tn) an 8 x NT block of it, and walks a range of
// 16-position chunks of the (frame, position) reduction; one atomicAdd per output element per CTA at the end (measured: the
// flush is < 1 % of the kernel).
//   * Software pipeline: the NEXT chunk's operands are fetched from global memory into registers before the current chunk is
//     multiplied out of shared memory, then stored into the other shared-memory buffer -- one barrier per chunk and the global
//     latency hidden behind the FMAs (the first version staged, synchronised and multiplied in turn: 6 TFLOP/s, bound by exposed
//     load latency).
//   * Tile shapes follow the codec's layers instead of powers of two: the weight matrices are 180 / 450 / 900 rows by 20 / 50 /
//     100 columns, which a 128 x 32 / 128 x 64 tile fills to 44-69 %.  25 x 10 threads (200 rows; 20 or 50 columns) and 12 x 20
//     threads (96 rows; 100 columns) fill 75-94 %.
//   * COLS_STRIDED: thread tn owns columns tn + TNW j (scalar shared loads, conflict-free) when NT is not 2 or 4.
// blockIdx.x == 0 CTAs also reduce the bias gradient.
constexpr int kWgR = 16;   // positions per chunk

template <int TMW, int TNW, int NT>
struct WgShape {
  static constexpr int kRows = 8 * TMW, kCols = NT * TNW, kActive = TMW * TNW;
  static constexpr int kXK = (kRows * kWgR + 255) / 256;     // A-tile elements a thread stages per chunk
  static constexpr int kGJ = (kCols * kWgR + 255) / 256;     // gradient-tile elements
  static constexpr bool kStrided = !(NT == 2 || NT == 4);
  static_assert(kActive <= 256, "wgrad tile: more owners than threads");
};

template <int TMW, int TNW, int NT>
__global__ void __launch_bounds__(256, 2)
wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dw, float* __restrict__ db,
             int64_t B, int Lin, int Lout, int Cin, int Cout, int K, int dil, int stride, int padL, int64_t chunks_per_split) {
  using S = WgShape<TMW, TNW, NT>;
  constexpr int kRows = S::kRows, TN = S::kCols, XK = S::kXK, GJ = S::kGJ;
  __shared__ __align__(16) float xs[2][kWgR][kRows + 4];
  __shared__ __align__(16) float gs[2][kWgR][TN + 4];
  const int M = K * Cin;
  const int m0 = blockIdx.x * kRows, n0 = blockIdx.y * TN;
  const int cpf = (Lout + kWgR - 1) / kWgR;                     // chunks per frame
  const int64_t c_lo = (int64_t)blockIdx.z * chunks_per_split;
  int64_t c_hi = c_lo + chunks_per_split;
  if (c_hi > B * cpf) c_hi = B * cpf;
  const int tid = threadIdx.x;
  const bool owner = tid < S::kActive;
  const int tm = owner ? tid / TNW : 0, tn = owner ? tid % TNW : 0;
  // staging map, fixed for the whole CTA: this thread fetches position (tid & 15) of rows (tid >> 4) + 16 k of the A tile and of
  // columns (tid >> 4) + 16 j of the gradient tile; lanes walk positions (contiguous in NCL)
  const int sr = tid & 15, sm = tid >> 4;
  int xoff[XK], xshift[XK];      // ci * Lin + shift, shift = t * dil - padL (INT_MIN / 2: row beyond the tile or M)
#pragma unroll
  for (int k = 0; k < XK; ++k) {
    const int mm = sm + 16 * k, m = m0 + mm;
    if (mm < kRows && m < M) {
      const int t = m / Cin, ci = m - t * Cin;
      xshift[k] = t * dil - padL;
      xoff[k] = ci * Lin + xshift[k];
    } else {
      xshift[k] = INT_MIN / 2;
      xoff[k] = 0;
    }
  }
  float xr[XK], gr[GJ];
  auto fetch = [&](int64_t c) {
    const int64_t b = c / cpf;
    const int p = (int)(c - b * cpf) * kWgR + sr;
    const float* xb = x + b * (int64_t)Cin * Lin;
    const float* gb = g + b * (int64_t)Cout * Lout;
    const int ps = p * stride;
#pragma unroll
    for (int k = 0; k < XK; ++k) {
      const int u = ps + xshift[k];
      xr[k] = (p < Lout && u >= 0 && u < Lin) ? __ldg(xb + xoff[k] + ps) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < GJ; ++j) {
      const int nn = sm + 16 * j, n = n0 + nn;
      gr[j] = (nn < TN && n < Cout && p < Lout) ? __ldg(gb + (int64_t)n * Lout + p) : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int k = 0; k < XK; ++k)
      if (sm + 16 * k < kRows) xs[buf][sr][sm + 16 * k] = xr[k];
#pragma unroll
    for (int j = 0; j < GJ; ++j)
      if (sm + 16 * j < TN) gs[buf][sr][sm + 16 * j] = gr[j];
  };
  float acc[8][NT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  float bacc = 0.f;   // bias partial: thread tid < TN of the blockIdx.x == 0 CTAs owns column n0 + tid
  const bool do_bias = db != nullptr && blockIdx.x == 0 && tid < TN;

  if (c_lo < c_hi) {
    fetch(c_lo);
    stash(0);
  }
  __syncthreads();
  for (int64_t c = c_lo; c < c_hi; ++c) {
    const int buf = (int)((c - c_lo) & 1);
    if (c + 1 < c_hi) fetch(c + 1);                // in flight while this chunk is multiplied
    if (owner) {
#pragma unroll
      for (int r = 0; r < kWgR; ++r) {
        const float4 a0 = *reinterpret_cast<const float4*>(&xs[buf][r][tm * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&xs[buf][r][tm * 8 + 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float cc[NT];
        if constexpr (NT == 4) {
          const float4 cv = *reinterpret_cast<const float4*>(&gs[buf][r][tn * 4]);
          cc[0] = cv.x; cc[1] = cv.y; cc[2] = cv.z; cc[3] = cv.w;
        } else if constexpr (NT == 2) {
          const float2 cv = *reinterpret_cast<const float2*>(&gs[buf][r][tn * 2]);
          cc[0] = cv.x; cc[1] = cv.y;
        } else {
#pragma unroll
          for (int j = 0; j < NT; ++j) cc[j] = gs[buf][r][tn + TNW * j];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < NT; ++j) acc[i][j] = fmaf(a[i], cc[j], acc[i][j]);
      }
    }
    if (do_bias) {
#pragma unroll
      for (int r = 0; r < kWgR; ++r) bacc += gs[buf][r][tid];
    }
    if (c + 1 < c_hi) stash(buf ^ 1);              // the other buffer was last read before the previous barrier
    __syncthreads();
  }
  if (owner) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + tm * 8 + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int n = n0 + (S::kStrided ? tn + TNW * j : tn * NT + j);
        if (n < Cout) atomicAdd(dw + (int64_t)m * Cout + n, acc[i][j]);
      }
    }
  }
  if (do_bias && n0 + tid < Cout) atomicAdd(db + n0 + tid, bacc);
}

template <int TMW, int TNW, int NT>
void launch_wgrad(const float* x, const float* g, float* dw, float* db, int64_t B, int Lin, int Lout, int Cin, int Cout, int K,
                  int dil, int stride, int padL, cudaStream_t st) {
  using S = WgShape<TMW, TNW, NT>;
  const int gx = ceil_div(K * Cin, S::kRows), gy = ceil_div(Cout, S::kCols);
  const int64_t total = B * (int64_t)ceil_div(Lout, kWgR);
  int64_t splits = ceil_div(4 * sm_count(), gx * gy);   // two resident CTAs per SM, two waves (2 -> 4 CTAs per SM of grid: 7.6 -> 7.4 ms)
  if (splits > total) splits = total;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  const int64_t cps = ceil_div64(total, splits);
  splits = ceil_div64(total, cps);
  wgrad_kernel<TMW, TNW, NT><<<dim3(gx, gy, (unsigned)splits), 256, 0, st>>>(x, g, dw, db, B, Lin, Lout, Cin, Cout, K, dil, stride, padL, cps);
}

// Weight gradient of the 1-channel heads (k55, Cout = 1): dW[t][ci] = sum_{b,p} x[b][ci][p + t - padL] g[b][p] is a correlation,
// not a GEMM -- on the tiled kernel its single output column pads to 32 and the two heads cost a quarter of all weight-gradient
// time.  One CTA per (input channel, batch split): the channel's row (with its zero halo) and the gradient row sit in shared
// memory, thread t owns tap t.
constexpr int kWhThreads = 64, kWhMaxL = 1024;

__global__ void __launch_bounds__(kWhThreads)
wgrad_head_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ dw, float* __restrict__ db,
                  int64_t B, int L, int Cin, int K, int padL, int frames_per_split) {
  __shared__ float xs[kWhMaxL + kWhThreads];
  __shared__ float gs[kWhMaxL];
  const int ci = blockIdx.x, t = threadIdx.x;
  const int64_t b_lo = (int64_t)blockIdx.y * frames_per_split;
  int64_t b_hi = b_lo + frames_per_split;
  if (b_hi > B) b_hi = B;
  float acc = 0.f, bacc = 0.f;
  for (int64_t b = b_lo; b < b_hi; ++b) {
    const float* xr = x + (b * Cin + ci) * (int64_t)L;
    const float* gr = g + b * (int64_t)L;
    for (int i = t; i < L + K - 1; i += kWhThreads) {
      const int u = i - padL;
      xs[i] = (u >= 0 && u < L) ? __ldg(xr + u) : 0.f;
    }
    for (int i = t; i < L; i += kWhThreads) {
      const float v = __ldg(gr + i);
      gs[i] = v;
      bacc += v;
    }
    __syncthreads();
    if (t < K) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;     // four chains: the loop is latency-, not throughput-bound per thread
      for (int p = 0; p < L; p += 4) {
        a0 = fmaf(xs[p + t], gs[p], a0);
        a1 = fmaf(xs[p + 1 + t], gs[p + 1], a1);
        a2 = fmaf(xs[p + 2 + t], gs[p + 2], a2);
        a3 = fmaf(xs[p + 3 + t], gs[p + 3], a3);
      }
      acc += (a0 + a1) + (a2 + a3);
    }
    __syncthreads();
  }
  if (t < K) atomicAdd(dw + (int64_t)t * Cin + ci, acc);
  if (db != nullptr && ci == 0) atomicAdd(db, bacc);
}

// ------------------------------------------------------------------------------------------------ quantiser backward
// dH/dh_k of entropy_coding_loss (loss_terms_and_measures.py:262-267), times coef (= tau * w_e * global batch)
__global__ void entropy_grad_kernel(const float* __restrict__ hist, int n, float coef, float* __restrict__ ge) {
  const int lane = threadIdx.x;
  float tot = 0.f;
  for (int k = lane; k < n; k += 32) tot += hist[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  const float ln2 = logf(2.0f);
  float dot = 0.f;
  for (int k = lane; k < n; k += 32) {
    const float p = hist[k] / tot;
    const float dHdp = -(logf(p + 1e-7f) + p / (p + 1e-7f)) / ln2;
    dot += p * dHdp;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int k = lane; k < n; k += 32) {
    const float p = hist[k] / tot;
    const float dHdp = -(logf(p + 1e-7f) + p / (p + 1e-7f)) / ln2;
    ge[k] = coef * (dHdp - dot) / tot;
  }
}

constexpr int kQbWarps = 8;

// one warp per code row; lane = bin (NPL bins per lane).  gout may be null (LSF codebook: value path carries no gradient)
template <int NPL>
__global__ void __launch_bounds__(kQbWarps * 32)
quantize_backward_kernel(const float* __restrict__ x, int64_t rows, const float* __restrict__ bins, int n,
                         const float* __restrict__ alpha_p, float iq, const float* __restrict__ gout, float cq,
                         const float* __restrict__ ge, float* __restrict__ gx, float* __restrict__ galpha,
                         float* __restrict__ gbins) {
  __shared__ float red_a[kQbWarps];
  __shared__ float red_b[kQbWarps][NPL * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float alpha = *alpha_p;
  float b[NPL], gev[NPL], gb_acc[NPL];
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    const int k = lane + 32 * j;
    b[j] = k < n ? bins[k] : 0.f;
    gev[j] = (k < n && ge) ? ge[k] : 0.f;
    gb_acc[j] = 0.f;
  }
  float ga_acc = 0.f;
  const int64_t row_stride = (int64_t)gridDim.x * kQbWarps;
  for (int64_t r = (int64_t)blockIdx.x * kQbWarps + warp; r < rows; r += row_stride) {
    const float xv = x[r];
    const float go = gout ? gout[r] : 0.f;
    float lg[NPL], dist[NPL], sgn[NPL];
    float best = -INFINITY;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const float d = xv - b[j];
      dist[j] = fabsf(d);
      sgn[j] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      lg[j] = alpha * dist[j];
      if (lane + 32 * j < n) best = fmaxf(best, lg[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    float s[NPL], es = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      s[j] = (lane + 32 * j) < n ? expf(lg[j] - best) : 0.f;
      es += s[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
    // gs_k = dL/ds_k ;  softmax backward gl_k = s_k (gs_k - sum_j s_j gs_j)
    float gs[NPL], dot = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      s[j] /= es;
      const bool v = (lane + 32 * j) < n;
      gs[j] = v ? (go * iq * b[j] + cq * 0.5f / sqrtf(s[j] + 1e-20f) + gev[j]) : 0.f;
      dot += s[j] * gs[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    float gxr = 0.f;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const float gl = s[j] * (gs[j] - dot);
      ga_acc += gl * dist[j];                       // d logit / d alpha = |x - b|
      gb_acc[j] += -gl * alpha * sgn[j] + go * iq * s[j];   // d|x-b|/db = -sign(x-b); plus q = sum s_k b_k directly
      gxr += gl * alpha * sgn[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gxr += __shfl_xor_sync(0xffffffffu, gxr, o);
    if (lane == 0 && gx) gx[r] = go * (1.f - iq) + gxr;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ga_acc += __shfl_xor_sync(0xffffffffu, ga_acc, o);
  if (lane == 0) red_a[warp] = ga_acc;
#pragma unroll
  for (int j = 0; j < NPL; ++j) red_b[warp][lane + 32 * j] = gb_acc[j];
  __syncthreads();
  if (threadIdx.x == 0 && galpha) {
    float t = 0.f;
    for (int w = 0; w < kQbWarps; ++w) t += red_a[w];
    atomicAdd(galpha, t);
  }
  if (gbins) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < kQbWarps; ++w) t += red_b[w][k];
      atomicAdd(gbins + k, t);
    }
  }
}

int launch_quantize_backward(const float* x, int64_t rows, const float* bins, int n, const float* alpha, float iq,
                             const float* gout, float cq, const float* ge, float* gx, float* galpha, float* gbins,
                             cudaStream_t st) {
  if (rows == 0) return NSC_OK;
  int64_t grid = ceil_div64(rows, kQbWarps * 8);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  const int npl = n <= 32 ? 1 : n <= 64 ? 2 : n <= 128 ? 4 : 8;
  ProfScope prof(st, "quantize_backward", (double)rows * n * 12.0, (double)rows * 12.0);
#define NSC_QB(NPL) quantize_backward_kernel<NPL><<<(unsigned)grid, kQbWarps * 32, 0, st>>>(x, rows, bins, n, alpha, iq, gout, cq, ge, gx, galpha, gbins)
  switch (npl) {
    case 1: NSC_QB(1); break;
    case 2: NSC_QB(2); break;
    case 4: NSC_QB(4); break;
    default: NSC_QB(8); break;
  }
#undef NSC_QB
  NSC_LAUNCH_OK();
  return NSC_OK;
}

// ------------------------------------------------------------------------------------------------ loss backward
constexpr int kN = NSC_FRAME_LENGTH, kBins = NSC_MEL_BINS, kMel = NSC_MEL_TOTAL;
__device__ __forceinline__ int mel_off(int i) { return i == 0 ? 0 : i == 1 ? 8 : i == 2 ? 24 : i == 3 ? 56 : 184; }

__device__ __forceinline__ void fft512_inplace(float2* z, int tid) {
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1);
    const int j = tid & (half - 1);
    const int i0 = ((tid >> (s - 1)) << s) + j;
    const int i1 = i0 + half;
    float sn, cs;
    sincospif(-(float)j / (float)half, &sn, &cs);
    const float2 u = z[i0], v = z[i1];
    const float2 vt = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    z[i0] = make_float2(u.x + vt.x, u.y + vt.y);
    z[i1] = make_float2(u.x - vt.x, u.y - vt.y);
    __syncthreads();
  }
}

// d( c0 * sum_b time_b + c1 * sum_b freq_b ) / d decoded   (mse_loss :77-79, mfcc_loss :151-175)
__global__ void __launch_bounds__(256)
losses_backward_kernel(const float* __restrict__ dec, const float* __restrict__ ori, const float* __restrict__ melw,
                       float c0, float c1, float* __restrict__ gdec) {
  __shared__ float2 z[kN];
  __shared__ float2 spec_d[kBins + 1];
  __shared__ float psd_d[kBins + 3], psd_o[kBins + 3];
  __shared__ float gmel[kMel];
  __shared__ float red[8];
  __shared__ float bank_r[4];
  const int tid = threadIdx.x;
  const int64_t f = blockIdx.x;
  const float* d = dec + f * kN;
  const float* o = ori + f * kN;
  float e0, e1, se = 0.f;
  {
    const float d0 = d[tid], o0 = o[tid], d1 = d[tid + 256], o1 = o[tid + 256];
    e0 = d0 - o0; e1 = d1 - o1;
    se = e0 * e0 + e1 * e1;
    z[__brev((unsigned)tid) >> 23] = make_float2(d0, o0);
    z[__brev((unsigned)(tid + 256)) >> 23] = make_float2(d1, o1);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) se += __shfl_xor_sync(0xffffffffu, se, s);
  if ((tid & 31) == 0) red[tid >> 5] = se;
  __syncthreads();
  float tl = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tl += red[w];
  tl = sqrtf(tl / (float)kN + 1e-07f);
  fft512_inplace(z, tid);
  for (int k = tid; k < kBins; k += 256) {
    const float2 a = z[k];
    const float2 c = z[(kN - k) & (kN - 1)];
    const float dr = 0.5f * (a.x + c.x), di = 0.5f * (a.y - c.y);
    const float orr = 0.5f * (a.y + c.y), oi = -0.5f * (a.x - c.x);
    spec_d[k] = make_float2(dr, di);
    psd_d[k] = (1.0f / (float)kN) * (dr * dr + di * di + 1e-7f);
    psd_o[k] = (1.0f / (float)kN) * (orr * orr + oi * oi + 1e-7f);
  }
  __syncthreads();
  const int* rng = reinterpret_cast<const int*>(melw + kBins * kMel);
  float delta = 0.f, md = 1.f;
  if (tid < kMel) {
    const int lo = rng[tid], hi = rng[kMel + tid];
    float ad = 0.f, ao = 0.f;
    for (int k = lo; k <= hi; ++k) {
      const float w = melw[k * kMel + tid];
      ad = fmaf(psd_d[k], w, ad);
      ao = fmaf(psd_o[k], w, ao);
    }
    md = ad + 1e-7f;
    delta = logf(md) - logf(ao + 1e-7f);
    gmel[tid] = delta * delta;
  }
  __syncthreads();
  if (tid < 4) {
    float s = 0.f;
    for (int m = mel_off(tid); m < mel_off(tid + 1); ++m) s += gmel[m];
    bank_r[tid] = sqrtf(s / (float)(mel_off(tid + 1) - mel_off(tid)) + 1e-07f);
  }
  __syncthreads();
  if (tid < kMel) {
    int bank = 0;
    while (tid >= mel_off(bank + 1)) ++bank;
    const float nb = (float)(mel_off(bank + 1) - mel_off(bank));
    // d freq / d Md_m = (1/4) * delta / (n_bank * r_bank) / (Md_m + 1e-7)
    gmel[tid] = c1 * 0.25f * delta / (nb * bank_r[bank]) / md;
  }
  __syncthreads();
  // g_psd[k] = sum_m gmel[m] W[k][m];  G_k = g_psd[k] * (2/512) * (re + i im);  gd = Re FFT(conj G)
  for (int k = tid; k < kN; k += 256) {
    float2 zz = make_float2(0.f, 0.f);
    if (k < kBins) {
      float gp = 0.f;
      for (int m = 0; m < kMel; ++m) {
        const float w = melw[k * kMel + m];
        gp = fmaf(gmel[m], w, gp);
      }
      const float sc = gp * (2.0f / (float)kN);
      zz = make_float2(sc * spec_d[k].x, -sc * spec_d[k].y);   // conj(G_k)
    }
    z[__brev((unsigned)k) >> 23] = zz;
  }
  __syncthreads();
  fft512_inplace(z, tid);
  const float ct = c0 / ((float)kN * tl);
  gdec[f * kN + tid] = z[tid].x + ct * e0;
  gdec[f * kN + tid + 256] = z[tid + 256].x + ct * e1;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            int64_t n, float lr_t, float b1, float b2, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

inline unsigned ew_grid(int64_t n) {
  int64_t g = ceil_div64(n, 256);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ------------------------------------------------------------------------------------------------ host program
constexpr int64_t kWflipFloats = 1 << 18;   // flipped weights of one layer (<= 90,000 floats on this path)

int64_t arena_bytes(const nsc_codec_cfg& c, int64_t B) {
  // dry-measure: run the replay walk with a null base and read how much it took
  Walker w;
  w.cfg = c;
  w.dry = false;
  w.launch = false;
  w.B = B;
  w.layers = make_layout(c).layers;
  Carver cv(nullptr, INT64_MAX);
  w.arena = &cv;
  float* cin = cv.take(B * kFrameLen);
  float* fcode = nullptr;
  float* out = nullptr;
  w.encoder(cin, &fcode);
  float* code = cv.take(B * code_length(c));
  w.decoder(code, &out);
  return cv.used;
}

struct CodecTrainBufs {
  float* cin;     // codec input  (B,512)
  float* fcode;   // floating code (B,Lc)
  float* code;    // quantised (soft) code (B,Lc)
  float* out;     // raw decoder output (B,512)
  std::vector<ConvRec> enc_tape, dec_tape;
  char* act_base;
  int64_t bytes;
};

// walks one codec into its arena; launch = forward pass, !launch = pointer replay for backward
int walk_codec(const nsc_codec_cfg& cfg, const CodecLayout& lay, const float* params, char* arena_base, int64_t arena_cap,
               int64_t B, bool launch, cudaStream_t st, void* wpack, CodecTrainBufs* tb, float iq, float* hist, float* qloss) {
  Walker w;
  w.cfg = cfg;
  w.dry = false;
  w.launch = launch;
  w.params = params;
  w.B = B;
  w.st = st;
  w.layers = lay.layers;
  w.wpack = wpack;
  Carver cv(arena_base, arena_cap);
  w.arena = &cv;
  tb->act_base = arena_base;
  tb->cin = cv.take(B * kFrameLen);
  w.tape = &tb->enc_tape;
  tb->enc_tape.clear();
  tb->dec_tape.clear();
  tb->fcode = nullptr;
  w.encoder(tb->cin, &tb->fcode);
  NSC_TRY(w.rc);
  tb->code = cv.take(B * lay.code_len);
  if (launch) {
    const float* alpha = params + lay.conv_floats;
    NSC_TRY(launch_quantize(tb->fcode, B, lay.code_len, alpha + 1, cfg.num_bins, alpha, iq, 1, tb->code, nullptr, nullptr, hist, qloss, st));
  }
  w.tape = &tb->dec_tape;
  tb->out = nullptr;
  w.decoder(tb->code, &tb->out);
  NSC_TRY(w.rc);
  tb->bytes = cv.used;
  if (cv.used > arena_cap) { set_error("training arena overflow"); return NSC_E_WORKSPACE; }
  return NSC_OK;
}

// backward through one conv record.  G(ptr) maps an activation pointer to its gradient twin.
int conv_backward(const ConvRec& r, const CodecLayout& lay, const float* params, float* grads, char* act_base, char* grad_base,
                  int64_t B, float* gpre, float* wflip, float* gtmp, int precision, void* wpack, bool need_dx, cudaStream_t st,
                  float* wg_scratch = nullptr) {
  auto G = [&](const float* p) { return reinterpret_cast<float*>(grad_base + (reinterpret_cast<const char*>(p) - act_base)); };
  if (r.op == OP_MUL) {   // y = x * res
    const int64_t n = B * (int64_t)r.Cin * r.Lin;
    ProfScope prof(st, "gate_product_backward", 4.0 * (double)n, 28.0 * (double)n);
    mul_backward_kernel<<<ew_grid(n), 256, 0, st>>>(G(r.y), r.x, r.res, G(r.x), G(r.res), n);
    NSC_LAUNCH_OK();
    return NSC_OK;
  }
  if (r.op == OP_DEPTHWISE) {   // taps (K, C) at the head of the separable layer's table entry; no bias, no activation
    const LayerInfo& li = lay.layers[r.layer];
    NSC_CHECK_ARG(r.K <= kDwMaxK && (r.K & 1) && (int64_t)r.K * r.Cin <= kWflipFloats, "training: depthwise layer with k = %d", r.K);
    const int padL = (r.K - 1) / 2;
    int splits = ceil_div(8 * sm_count(), r.Cin);
    if (splits > B) splits = (int)B;
    if (splits < 1) splits = 1;
    const int fps = (int)ceil_div64(B, splits);
    splits = (int)ceil_div64(B, fps);
    {
      ProfScope prof(st, "wgrad_depthwise", 2.0 * B * r.Lin * (double)r.K * r.Cin, 8.0 * B * (double)r.Lin * r.Cin);
      depthwise_wgrad_kernel<<<dim3(r.Cin, splits), 256, 0, st>>>(r.x, G(r.y), grads + li.off, B, r.Lin, r.Cin, r.K, padL, fps);
      NSC_LAUNCH_OK();
    }
    if (!need_dx) return NSC_OK;
    flip_taps_kernel<<<ew_grid((int64_t)r.K * r.Cin), 256, 0, st>>>(params + li.off, wflip, r.K, r.Cin);
    NSC_LAUNCH_OK();
    NSC_TRY(launch_depthwise(G(r.y), wflip, gtmp, B, r.Lin, r.Cin, r.K, 1, 1, 0, 0, st));
    const int64_t n4 = B * (int64_t)r.Cin * r.Lin / 4;
    add_inplace_kernel<<<ew_grid(n4), 256, 0, st>>>(G(r.x), gtmp, n4);
    NSC_LAUNCH_OK();
    return NSC_OK;
  }
  int Lout, padL;
  same_padding(r.Lin, r.K, r.dil, r.stride, &Lout, &padL);
  const LayerInfo& li = lay.layers[r.layer];
  // (the pointwise half of a separable layer sits behind the depthwise taps of the same table entry)
  const int64_t woff = li.off + (r.op == OP_POINTWISE ? (int64_t)r.kdw * r.Cin : 0);
  const float* w = params + woff;
  float* dw = grads + woff;
  float* db = dw + (int64_t)r.K * r.Cin * r.Cout;
  {
    ProfScope prof(st, "epilogue_backward", (double)B * Lout * r.Cout * 4.0, (double)B * Lout * r.Cout * 16.0);
    epilogue_backward_kernel<<<ew_grid(B * (int64_t)Lout), 256, 0, st>>>(G(r.y), r.y, gpre, r.res ? G(r.res) : nullptr, B, Lout, r.Cout,
                                                                      r.act, r.res_mode, r.post_act, r.shuffle);
    NSC_LAUNCH_OK();
  }
  if (r.Cout == 1 && r.stride == 1 && r.dil == 1 && r.K <= kWhThreads && r.Lin == Lout && Lout <= kWhMaxL && (Lout & 3) == 0) {
    int splits = ceil_div(16 * sm_count(), r.Cin);      // small CTAs (64 threads, 5 KB): many per SM
    if (splits > B) splits = (int)B;
    if (splits < 1) splits = 1;
    const int fps = (int)ceil_div64(B, splits);
    splits = (int)ceil_div64(B, fps);
    ProfScope prof(st, "wgrad_head", 2.0 * B * Lout * (double)r.K * r.Cin, 4.0 * B * ((double)r.Lin * r.Cin + (double)Lout));
    wgrad_head_kernel<<<dim3(r.Cin, splits), kWhThreads, 0, st>>>(r.x, gpre, dw, db, B, Lout, r.Cin, r.K, padL, fps);
    NSC_LAUNCH_OK();
  } else if (precision > 0 && wg_scratch != nullptr && wgrad_tc_supported(B, r.Lin, Lout, r.Cin, r.Cout, r.K, r.dil, r.stride, padL)) {
    // tensor cores (fp16 hi/lo split like the forward and data-gradient convs of this precision mode): positions are the K dimension
    NSC_TRY(launch_wgrad_tc(r.x, gpre, dw, db, B, Lout, r.Cin, r.Cout, r.K, r.dil, padL, wg_scratch, st, r.stride));
  } else {
    const int M = r.K * r.Cin;
    ProfScope prof(st, "wgrad", 2.0 * B * Lout * (double)M * r.Cout, 4.0 * B * ((double)r.Lin * r.Cin + (double)Lout * r.Cout));
    // tile shape by output width: 20 -> 200 x 20, 50 -> 200 x 50, 100 -> 96 x 100; anything else the power-of-two tiles
    if (r.Cout <= 20) launch_wgrad<25, 10, 2>(r.x, gpre, dw, db, B, r.Lin, Lout, r.Cin, r.Cout, r.K, r.dil, r.stride, padL, st);
    else if (r.Cout > 32 && r.Cout <= 50) launch_wgrad<25, 10, 5>(r.x, gpre, dw, db, B, r.Lin, Lout, r.Cin, r.Cout, r.K, r.dil, r.stride, padL, st);
    else if (r.Cout > 64 && r.Cout <= 100) launch_wgrad<12, 20, 5>(r.x, gpre, dw, db, B, r.Lin, Lout, r.Cin, r.Cout, r.K, r.dil, r.stride, padL, st);
    else if (r.Cout <= 32) launch_wgrad<16, 16, 2>(r.x, gpre, dw, db, B, r.Lin, Lout, r.Cin, r.Cout, r.K, r.dil, r.stride, padL, st);
    else launch_wgrad<16, 16, 4>(r.x, gpre, dw, db, B, r.Lin, Lout, r.Cin, r.Cout, r.K, r.dil, r.stride, padL, st);
    NSC_LAUNCH_OK();
  }
  if (!need_dx) return NSC_OK;
  if (r.stride == 1 && (r.K & 1)) {
    NSC_CHECK_ARG((int64_t)r.K * r.Cin * r.Cout <= kWflipFloats, "training: layer too large for the flip scratch");
    flip_weights_kernel<<<ew_grid((int64_t)r.K * r.Cin * r.Cout), 256, 0, st>>>(w, wflip, r.K, r.Cin, r.Cout);
    NSC_LAUNCH_OK();
    ConvArgs a;   // gx += conv(gpre, Wflip): forward engine, Cout -> Cin channels, accumulate through the residual input
    a.x = gpre; a.w = wflip; a.bias = nullptr; a.y = G(r.x); a.res = G(r.x); a.res_mode = RES_ADD;
    a.B = B; a.Lin = Lout; a.Cin = r.Cout; a.Cout = r.Cin; a.K = r.K; a.dil = r.dil; a.stride = 1;
    // same engine choice as the forward walk (walker.cuh): tensor cores (fp16 hi/lo split) when the codec asks for them
    if (precision > 0 && wpack != nullptr && tc_conv_supported(a)) NSC_TRY(launch_conv_tc(a, precision, wpack, st));
    else NSC_TRY(launch_conv(a, st));
  } else if (r.stride == 2 && r.dil == 1 && r.K == 9 && padL == 3 && r.Lin == 2 * Lout && (r.Lin & 3) == 0 &&
             (int64_t)5 * r.Cout * 2 * r.Cin <= kWflipFloats && precision > 0 && wpack != nullptr && [&] {
               ConvArgs t;
               t.B = B; t.Lin = Lout; t.Cin = r.Cout; t.Cout = r.Cin; t.K = 5; t.dil = 1; t.stride = 1;
               return tc_conv_supported(t);
             }()) {
    // tensor engine: one dense k5 conv per input-position parity, interleaved into the gradient
    subpixel_dgrad_weights_split_kernel<<<ew_grid((int64_t)10 * r.Cout * r.Cin), 256, 0, st>>>(w, wflip, r.K, r.Cin, r.Cout, padL);
    NSC_LAUNCH_OK();
    const int64_t half = B * (int64_t)r.Cin * Lout;
    for (int par = 0; par < 2; ++par) {
      ConvArgs a;
      a.x = gpre; a.w = wflip + (int64_t)par * 5 * r.Cout * r.Cin; a.bias = nullptr; a.y = gtmp + par * half; a.res_mode = RES_NONE;
      a.B = B; a.Lin = Lout; a.Cin = r.Cout; a.Cout = r.Cin; a.K = 5; a.dil = 1; a.stride = 1;
      NSC_TRY(launch_conv_tc(a, precision, wpack, st));
    }
    add_interleaved_kernel<<<ew_grid(2 * half), 256, 0, st>>>(G(r.x), gtmp, B * (int64_t)r.Cin, Lout);
    NSC_LAUNCH_OK();
  } else if (r.stride == 2 && r.dil == 1 && r.K == 9 && padL == 3 && r.Lin == 2 * Lout && (r.Lin & 3) == 0 &&
             (int64_t)5 * r.Cout * 2 * r.Cin <= kWflipFloats) {
    subpixel_dgrad_weights_kernel<<<ew_grid((int64_t)5 * r.Cout * 2 * r.Cin), 256, 0, st>>>(w, wflip, r.K, r.Cin, r.Cout, padL);
    NSC_LAUNCH_OK();
    ConvArgs a;   // gtmp = shuffle2(conv_k5(gpre, Wt)), then gx += gtmp (the residual operand of a shuffled conv has the pre-shuffle shape)
    a.x = gpre; a.w = wflip; a.bias = nullptr; a.y = gtmp; a.res_mode = RES_NONE; a.shuffle = 2;
    a.B = B; a.Lin = Lout; a.Cin = r.Cout; a.Cout = 2 * r.Cin; a.K = 5; a.dil = 1; a.stride = 1;
    NSC_TRY(launch_conv(a, st));
    const int64_t n4 = B * (int64_t)r.Cin * r.Lin / 4;
    add_inplace_kernel<<<ew_grid(n4), 256, 0, st>>>(G(r.x), gtmp, n4);
    NSC_LAUNCH_OK();
  } else {
    ProfScope prof(st, "dgrad_strided", 2.0 * B * Lout * (double)r.K * r.Cin * r.Cout, 4.0 * B * ((double)r.Lin * r.Cin + (double)Lout * r.Cout));
    dgrad_strided_kernel<<<ew_grid(B * (int64_t)r.Cin * r.Lin), 256, 0, st>>>(gpre, w, G(r.x), B, r.Lin, Lout, r.Cin, r.Cout, r.K, r.dil,
                                                                           r.stride, padL);
    NSC_LAUNCH_OK();
  }
  return NSC_OK;
}

struct TrainLayout {
  std::vector<int64_t> arena;   // activation arena bytes per codec
  int64_t act_off[NSC_MAX_CODECS], grad_off[NSC_MAX_CODECS];
  int64_t gpre_off, gtmp_off, wflip_off, gdec_off, acc_off, ge_off, wpack_off, wg_off, total;
};

TrainLayout train_layout(const nsc_codec_cfg* cfgs, int n, int64_t B) {
  TrainLayout t;
  int64_t off = 0;
  int max_wide = 1;
  for (int i = 0; i < n; ++i) {
    t.arena.push_back(align_up(arena_bytes(cfgs[i], B), 256));
    t.act_off[i] = off; off += t.arena[i];
    t.grad_off[i] = off; off += t.arena[i];
    if (cfgs[i].wide > max_wide) max_wide = cfgs[i].wide;
  }
  t.gpre_off = off; off += align_up(B * (int64_t)max_wide * kFrameLen * 4, 256);
  t.gtmp_off = off; off += align_up(B * (int64_t)max_wide * kFrameLen * 4, 256);   // data gradient of the strided conv before it is added
  t.wflip_off = off; off += kWflipFloats * 4;
  t.gdec_off = off; off += align_up(B * (int64_t)kFrameLen * 4, 256);
  t.acc_off = off; off += align_up(B * (int64_t)kFrameLen * 4, 256);
  t.ge_off = off; off += 256 * 4;
  t.wpack_off = off; off += kTcWpackBytes;
  t.wg_off = off; off += align_up(wgrad_tc_scratch_floats() * 4, 256);   // per-CTA accumulator slices of the tensor-core weight gradient
  t.total = off;
  return t;
}

int check_train_args(const nsc_codec_cfg* cfgs, int n, const float* const* params, int64_t B, void* ws, int64_t ws_bytes,
                     const TrainLayout** out_layout) {
  (void)out_layout;
  NSC_CHECK_ARG(cfgs && n >= 1 && n <= NSC_MAX_CODECS && params && ws, "training: bad arguments");
  for (int i = 0; i < n; ++i) {
    NSC_TRY(validate_cfg(&cfgs[i]));
    NSC_CHECK_ARG(params[i] != nullptr, "training: params[%d] is null", i);
  }
  NSC_CHECK_ARG(B >= 1, "training: empty batch");
  (void)ws_bytes;
  return NSC_OK;
}

}  // namespace
}  // namespace nsc

extern "C" {

int64_t nsc_train_workspace_bytes(const nsc_codec_cfg* cfgs, int32_t n_codecs, int64_t B) {
  if (cfgs == nullptr || n_codecs < 1 || n_codecs > NSC_MAX_CODECS || B < 1) return -1;
  for (int i = 0; i < n_codecs; ++i)
    if (nsc::validate_cfg(&cfgs[i]) != NSC_OK) return -1;
  return nsc::train_layout(cfgs, n_codecs, B).total;
}

int nsc_train_forward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                      const float* lsf_params, int32_t n_lsf_bins, const float* res_x, const float* lsf, int64_t B,
                      float res_scalar, float is_quan_on, const float* melw, float* decoded, float* time_loss,
                      float* freq_loss, float* const* qloss_ptrs_host, float* const* hist_ptrs_host, void* workspace,
                      int64_t workspace_bytes, void* stream) {
  NSC_TRY(nsc::check_train_args(cfgs, n_codecs, params_ptrs_host, B, workspace, workspace_bytes, nullptr));
  NSC_CHECK_ARG(res_x && decoded && qloss_ptrs_host && hist_ptrs_host && res_scalar != 0.f, "nsc_train_forward: null pointer");
  const nsc::TrainLayout tl = nsc::train_layout(cfgs, n_codecs, B);
  if (workspace_bytes < tl.total) { nsc::set_error("nsc_train_forward: workspace %lld < %lld", (long long)workspace_bytes, (long long)tl.total); return NSC_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = static_cast<char*>(workspace);
  const int64_t nfl = B * nsc::kFrameLen;
  // LSF codebook (index 0 of the per-quantiser arrays): soft assignment statistics only
  if (lsf != nullptr && lsf_params != nullptr)
    NSC_TRY(nsc::launch_quantize(lsf, B, NSC_LPC_ORDER, lsf_params + 1, n_lsf_bins, lsf_params, is_quan_on, 1, nullptr, nullptr, nullptr,
                                 hist_ptrs_host[0], qloss_ptrs_host[0], st));
  for (int i = 0; i < n_codecs; ++i) {
    const nsc::CodecLayout lay = nsc::make_layout(cfgs[i]);
    nsc::CodecTrainBufs tb;
    // codec input: res_scalar * (res_x - sum_{j<i} out_j)   (cmrl.py:430-433)
    float* cin = reinterpret_cast<float*>(ws + tl.act_off[i]);   // first carve of the arena
    NSC_TRY(nsc::launch_axpby(cin, res_x, res_scalar, i == 0 ? nullptr : decoded, 1.0f, nfl, st));
    NSC_TRY(nsc::walk_codec(cfgs[i], lay, params_ptrs_host[i], ws + tl.act_off[i], tl.arena[i], B, true, st, ws + tl.wpack_off, &tb,
                            is_quan_on, hist_ptrs_host[i + 1], qloss_ptrs_host[i + 1]));
    NSC_TRY(nsc::launch_accum_div(decoded, tb.out, res_scalar, i == 0 ? 1 : 0, nfl, st));
  }
  if (time_loss || freq_loss) NSC_TRY(nsc_losses_forward(decoded, res_x, B, melw, time_loss, freq_loss, stream));
  return NSC_OK;
}

int nsc_train_backward(const nsc_codec_cfg* cfgs, int32_t n_codecs, const float* const* params_ptrs_host,
                       const float* lsf_params, int32_t n_lsf_bins, const float* res_x, const float* lsf, int64_t B,
                       float res_scalar, float is_quan_on, const float* melw, const float* decoded,
                       const float* loss_coeff_host, const float* quan_w_host, const float* ent_w_host, int64_t global_B,
                       const float* const* hist_global_ptrs_host, const int32_t* trainable_host,
                       float* const* grad_ptrs_host, float* lsf_grad, void* workspace, int64_t workspace_bytes, void* stream) {
  NSC_TRY(nsc::check_train_args(cfgs, n_codecs, params_ptrs_host, B, workspace, workspace_bytes, nullptr));
  NSC_CHECK_ARG(res_x && decoded && melw && loss_coeff_host && quan_w_host && ent_w_host && hist_global_ptrs_host && trainable_host &&
                    grad_ptrs_host, "nsc_train_backward: null pointer");
  const nsc::TrainLayout tl = nsc::train_layout(cfgs, n_codecs, B);
  if (workspace_bytes < tl.total) { nsc::set_error("nsc_train_backward: workspace too small"); return NSC_E_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = static_cast<char*>(workspace);
  const float c0 = loss_coeff_host[0], c1 = loss_coeff_host[1], c2 = loss_coeff_host[2], tau = loss_coeff_host[3];
  const int64_t nfl = B * nsc::kFrameLen;
  float* gpre = reinterpret_cast<float*>(ws + tl.gpre_off);
  float* gtmp = reinterpret_cast<float*>(ws + tl.gtmp_off);
  float* wflip = reinterpret_cast<float*>(ws + tl.wflip_off);
  float* gdec = reinterpret_cast<float*>(ws + tl.gdec_off);
  float* acc = reinterpret_cast<float*>(ws + tl.acc_off);     // sum over later codecs of -rs * d/d(codec input)
  float* ge = reinterpret_cast<float*>(ws + tl.ge_off);

  // d(c0*sum time + c1*sum freq)/d decoded
  {
    nsc::ProfScope prof(st, "losses_backward", (double)B * 4.0e5, (double)B * 6148.0);
    nsc::losses_backward_kernel<<<(unsigned)B, 256, 0, st>>>(decoded, res_x, melw, c0, c1, gdec);
    NSC_LAUNCH_OK();
  }
  // LSF codebook: only quan_loss and entropy of its own soft assignment (the value path is cut by tf.py_func)
  if (lsf != nullptr && lsf_params != nullptr && lsf_grad != nullptr && trainable_host[0]) {
    NSC_CUDA_OK(cudaMemsetAsync(lsf_grad, 0, sizeof(float) * (1 + n_lsf_bins), st));
    nsc::entropy_grad_kernel<<<1, 32, 0, st>>>(hist_global_ptrs_host[0], n_lsf_bins, tau * ent_w_host[0] * (float)global_B, ge);
    NSC_LAUNCH_OK();
    NSC_TRY(nsc::launch_quantize_backward(lsf, B * NSC_LPC_ORDER, lsf_params + 1, n_lsf_bins, lsf_params, is_quan_on, nullptr,
                                          c2 * quan_w_host[0] / (float)NSC_LPC_ORDER, ge, nullptr, lsf_grad, lsf_grad + 1, st));
  }
  // codecs in reverse: d/d out'_j = gdec + sum_{m>j} (-rs) d/d in_m   (out'_j = out_j / rs, in_m = rs (x - sum_{j<m} out'_j))
  bool any_earlier_trainable = false;
  for (int i = 0; i < n_codecs; ++i) any_earlier_trainable = any_earlier_trainable || trainable_host[i + 1];
  NSC_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(float) * nfl, st));
  for (int i = n_codecs - 1; i >= 0; --i) {
    bool needed = false;    // this codec's backward is needed if it or any earlier codec is trainable
    for (int j = 0; j <= i; ++j) needed = needed || trainable_host[j + 1];
    if (!needed) break;
    const nsc::CodecLayout lay = nsc::make_layout(cfgs[i]);
    nsc::CodecTrainBufs tb;
    NSC_TRY(nsc::walk_codec(cfgs[i], lay, params_ptrs_host[i], ws + tl.act_off[i], tl.arena[i], B, false, st, nullptr, &tb, is_quan_on,
                            nullptr, nullptr));
    char* act_base = ws + tl.act_off[i];
    char* grad_base = ws + tl.grad_off[i];
    auto G = [&](const float* p) { return reinterpret_cast<float*>(grad_base + (reinterpret_cast<const char*>(p) - act_base)); };
    NSC_CUDA_OK(cudaMemsetAsync(grad_base, 0, (size_t)tl.arena[i], st));
    float* grads = grad_ptrs_host[i];
    NSC_CHECK_ARG(grads != nullptr, "nsc_train_backward: grad buffer %d is null", i);
    NSC_CUDA_OK(cudaMemsetAsync(grads, 0, sizeof(float) * (size_t)(lay.conv_floats + 1 + cfgs[i].num_bins), st));
    // seed: d/d(raw decoder output) = (gdec + acc) / rs
    NSC_TRY(nsc::launch_axpby(G(tb.out), gdec, 1.0f / res_scalar, acc, 1.0f, nfl, st));   // (gdec - acc) / rs, acc = sum_{m>i} rs * d/d in_m
    for (int k = (int)tb.dec_tape.size() - 1; k >= 0; --k)
      NSC_TRY(nsc::conv_backward(tb.dec_tape[k], lay, params_ptrs_host[i], grads, act_base, grad_base, B, gpre, wflip, gtmp, cfgs[i].precision, ws + tl.wpack_off, true, st,
                                 reinterpret_cast<float*>(ws + tl.wg_off)));
    // quantiser (soft path)
    const float* alpha = params_ptrs_host[i] + lay.conv_floats;
    nsc::entropy_grad_kernel<<<1, 32, 0, st>>>(hist_global_ptrs_host[i + 1], cfgs[i].num_bins, tau * ent_w_host[i + 1] * (float)global_B, ge);
    NSC_LAUNCH_OK();
    NSC_TRY(nsc::launch_quantize_backward(tb.fcode, B * lay.code_len, alpha + 1, cfgs[i].num_bins, alpha, is_quan_on, G(tb.code),
                                          c2 * quan_w_host[i + 1] / (float)lay.code_len, ge, G(tb.fcode), grads + lay.conv_floats,
                                          grads + lay.conv_floats + 1, st));
    bool earlier = false;
    for (int j = 0; j < i; ++j) earlier = earlier || trainable_host[j + 1];
    for (int k = (int)tb.enc_tape.size() - 1; k >= 0; --k)
      NSC_TRY(nsc::conv_backward(tb.enc_tape[k], lay, params_ptrs_host[i], grads, act_base, grad_base, B, gpre, wflip, gtmp, cfgs[i].precision, ws + tl.wpack_off, k > 0 || earlier, st,
                                 reinterpret_cast<float*>(ws + tl.wg_off)));
    if (earlier) {
      // acc += rs * d/d in_i   (applied with a minus sign in the seed of every earlier codec)
      NSC_TRY(nsc::launch_axpby(acc, G(tb.cin), res_scalar, acc, -1.0f / res_scalar, nfl, st));
    }
  }
  (void)any_earlier_trainable;
  return NSC_OK;
}

int nsc_adam_step(float* params, const float* grad, float* m, float* v, int64_t n, float lr, int64_t t, float beta1,
                  float beta2, float eps, void* stream) {
  NSC_CHECK_ARG(params && grad && m && v && n >= 0 && t >= 1, "nsc_adam_step: bad arguments");
  if (n == 0) return NSC_OK;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, (double)t)) / (1.0 - pow((double)beta1, (double)t));
  nsc::ProfScope prof((cudaStream_t)stream, "adam", 10.0 * n, 28.0 * n);
  nsc::adam_kernel<<<nsc::ew_grid(n), 256, 0, (cudaStream_t)stream>>>(params, grad, m, v, n, (float)lr_t, beta1, beta2, eps);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // extern "C"
