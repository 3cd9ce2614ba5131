// K1 / K2 / K7 -- the LPC front and back end of collaborative quantisation (lpc_utilities.py).
//
// The reference runs these in Python float64 (audiolazy generators, spectrum + numpy.roots) behind
// tf.py_func.  They are a few 10k MAC per frame -- nothing next to the 0.3 GFLOP of one codec -- so they
// keep the reference's float64 arithmetic (B200's fp64 pipe is ample) and are organised purely for
// parallelism across frames and coalesced HBM access:
//   nsc_lpc_analyze   one warp per frame: windowed signal staged in smem as double, 17 autocorrelation lags
//                     by strided partial sums + shuffle reduction, Levinson-Durbin in registers, LSFs by a
//                     Chebyshev-series sign scan (2048 intervals over [0,pi]) + bisection, one root per lane.
//   nsc_lsf2poly      one thread per frame: product of the 8+8 unit-circle quadratics.
//   nsc_lpc_residual  one thread per output sample: the (<=2) zero-state sub-frame FIRs covering it.
//   nsc_lpc_synth     one thread per frame, 16-deep history in registers, smem transpose for coalescing.
#include "common.cuh"

namespace nsc {

constexpr int kOrder = NSC_LPC_ORDER;
constexpr double kPi = 3.14159265358979323846;

// numpy.hanning(M)[j] = 0.5 + 0.5*cos(pi*(2j-(M-1))/(M-1))
__device__ __forceinline__ double np_hanning(int j, int M) {
  return 0.5 + 0.5 * cospi((double)(2 * j - (M - 1)) / (double)(M - 1));
}

// ------------------------------------------------------------------------------------------------
// shared pieces of the analysis: autocorrelation (warp), Levinson-Durbin, poly -> LSF
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// s: smem signal of length len followed by >= kOrder zeros.  Returns r[0..16] in every lane.
__device__ __forceinline__ void warp_autocorr(const double* s, int len, int lane, double r[kOrder + 1]) {
#pragma unroll
  for (int t = 0; t <= kOrder; ++t) r[t] = 0.0;
  for (int n = lane; n < len; n += 32) {
    const double v = s[n];
#pragma unroll
    for (int t = 0; t <= kOrder; ++t) r[t] = fma(v, s[n + t], r[t]);
  }
#pragma unroll
  for (int t = 0; t <= kOrder; ++t) r[t] = warp_sum_d(r[t]);
}

// Levinson-Durbin (audiolazy.lpc autocorrelation strategy).  Returns false when the frame is not analysable.
__device__ __forceinline__ bool levinson(const double r[kOrder + 1], double a[kOrder + 1]) {
#pragma unroll
  for (int i = 0; i <= kOrder; ++i) a[i] = 0.0;
  a[0] = 1.0;
  double err = r[0];
  bool ok = (err > 0.0) && isfinite(err);
#pragma unroll
  for (int m = 1; m <= kOrder; ++m) {
    double acc = r[m];
#pragma unroll
    for (int i = 1; i < m; ++i) acc = fma(a[i], r[m - i], acc);
    const double k = -acc / err;
    if (!(fabs(k) < 1.0)) ok = false;
    double prev[kOrder + 1];
#pragma unroll
    for (int i = 1; i < m; ++i) prev[i] = a[i];
#pragma unroll
    for (int i = 1; i < m; ++i) a[i] = fma(k, prev[m - i], prev[i]);
    a[m] = k;
    err *= (1.0 - k * k);
  }
  return ok;
}

// f(w) = c0 + sum_{m=1}^{8} c_m cos(m w) via Clenshaw on x = cos w
__device__ __forceinline__ double cheb_eval(const double c[9], double x) {
  double b1 = 0.0, b2 = 0.0;
  const double x2 = 2.0 * x;
#pragma unroll
  for (int m = 8; m >= 1; --m) {
    const double b0 = fma(x2, b1, c[m] - b2);
    b2 = b1;
    b1 = b0;
  }
  return fma(x, b1, c[0] - b2);
}

constexpr int kGrid = 2048;  // sign-scan intervals over [0, pi]

// spectrum.poly2lsf: roots of the sum/difference polynomials on the unit circle, ascending angles.
// All lanes hold the same a[]; scratch: 2*8 doubles + 2 ints of shared memory per warp.
__device__ __forceinline__ bool warp_poly2lsf(const double a[kOrder + 1], int lane, double* roots_s /*16*/,
                                              int* count_s /*2*/, double* out_lsf /*lane<16 valid*/) {
  // P1 = a1 - rev(a1), Q1 = a1 + rev(a1) with a1 = [a, 0]; deflate z=1 / z=-1.
  double p[kOrder + 1], q[kOrder + 1];
  {
    double pp = 0.0, qq = 0.0;
#pragma unroll
    for (int k = 0; k <= kOrder; ++k) {
      const double a1k = a[k];
      const double a2k = (k == 0) ? 0.0 : a[kOrder + 1 - k];
      pp = (a1k - a2k) + pp;   // P = P1 / (1 - z^-1)
      qq = (a1k + a2k) - qq;   // Q = Q1 / (1 + z^-1)
      p[k] = pp;
      q[k] = qq;
    }
  }
  double cp[9], cq[9];
  cp[0] = p[8];
  cq[0] = q[8];
#pragma unroll
  for (int m = 1; m <= 8; ++m) {
    cp[m] = 2.0 * p[8 - m];
    cq[m] = 2.0 * q[8 - m];
  }
  if (lane < 2) count_s[lane] = 0;
  __syncwarp();
  // scan: lane owns kGrid/32 consecutive intervals; ordered compaction keeps each list ascending.
  constexpr int PER = kGrid / 32;
  double brk_lo[2][8];  // a polynomial has 8 roots in total, so a lane can never hold more than 8 brackets
  int total_found[2];
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const double* c = which == 0 ? cq : cp;  // Q's first root precedes P's; order is fixed later by sorting
    double w0 = kPi * (double)(lane * PER) / (double)kGrid;
    double f0 = cheb_eval(c, cos(w0));
    int found = 0;
    for (int i = 1; i <= PER; ++i) {
      const double w1 = kPi * (double)(lane * PER + i) / (double)kGrid;
      const double f1 = cheb_eval(c, cos(w1));
      if ((f0 > 0.0) != (f1 > 0.0)) {
        if (found < 8) brk_lo[which][found] = w0;
        ++found;
      }
      w0 = w1;
      f0 = f1;
    }
    // exclusive prefix over lanes
    int incl = found;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int excl = incl - found;
    total_found[which] = __shfl_sync(0xffffffffu, incl, 31);
    const bool lane_overflow = found > 8;
    const unsigned any_over = __ballot_sync(0xffffffffu, lane_overflow);
    if (any_over) total_found[which] = -1;
    if (total_found[which] == 8) {
      for (int j = 0; j < found; ++j) roots_s[which * 8 + excl + j] = brk_lo[which][j];
    }
  }
  __syncwarp();
  const bool ok = (total_found[0] == 8) && (total_found[1] == 8);
  double root = 0.0;
  if (ok && lane < 16) {
    const int which = lane >> 3;
    const double* c = which == 0 ? cq : cp;
    double lo = roots_s[lane], hi = lo + kPi / (double)kGrid;
    const bool lo_pos = cheb_eval(c, cos(lo)) > 0.0;
    for (int it = 0; it < 50; ++it) {
      const double mid = 0.5 * (lo + hi);
      const bool mid_pos = cheb_eval(c, cos(mid)) > 0.0;
      if (mid_pos == lo_pos) lo = mid; else hi = mid;
    }
    root = 0.5 * (lo + hi);
  }
  __syncwarp();
  if (ok && lane < 16) roots_s[lane] = root;
  __syncwarp();
  if (ok && lane < 16) {
    // rank = own position + number of roots of the other polynomial below this one
    const int which = lane >> 3, pos = lane & 7;
    int rank = pos;
#pragma unroll
    for (int j = 0; j < 8; ++j) rank += (roots_s[(1 - which) * 8 + j] < root) ? 1 : 0;
    out_lsf[rank] = root;
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// nsc_lpc_analyze: lpc_utilities.py:112-124 (MODE 0, 1024-sample windows) and :14-25 (MODE 1, 512 frames)
// ------------------------------------------------------------------------------------------------
constexpr int kAWarps = 4;

template <int MODE>
__global__ void __launch_bounds__(kAWarps * 32)
lpc_analyze_kernel(const float* __restrict__ in, int64_t N, double* __restrict__ lsf, float* __restrict__ lsf32,
                   int* __restrict__ status) {
  constexpr int LEN = MODE == 0 ? 1024 : 512;
  __shared__ double sig_s[kAWarps][LEN + kOrder + 2];
  __shared__ double roots_s[kAWarps][16];
  __shared__ int count_s[kAWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t frame = (int64_t)blockIdx.x * kAWarps + warp;
  if (frame >= N) return;  // warp-uniform
  double* s = sig_s[warp];
  const float* x = in + frame * LEN;
  if (MODE == 0) {
    // window: [hanning(512)[:256], ones(512), hanning(512)[256:]]  (lpc_utilities.py:120-121)
    for (int n = lane; n < LEN; n += 32) {
      double w = 1.0;
      if (n < 256) w = np_hanning(n, 512);
      else if (n >= 768) w = np_hanning(n - 512, 512);
      s[n] = (double)x[n] * w;
    }
  } else {
    // highpass biquad then pre-emphasis, zero state per frame (lpc_utilities.py:8-11, :20); sequential -> lane 0
    for (int n = lane; n < LEN; n += 32) s[n] = (double)x[n];
    __syncwarp();
    if (lane == 0) {
      const double b0 = 0.989502, b1 = -1.979004, b2 = 0.989592, a1 = -1.978882, a2 = 0.979126;
      double x1 = 0, x2 = 0, y1 = 0, y2 = 0, e1 = 0;
      for (int n = 0; n < LEN; ++n) {
        const double xv = s[n];
        const double y = b0 * xv + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
        x2 = x1; x1 = xv; y2 = y1; y1 = y;
        s[n] = y + (-0.68) * e1;   // empha_filter = 1 - 0.68 z^-1 (constants.py:64)
        e1 = y;
      }
    }
  }
  for (int n = LEN + lane; n < LEN + kOrder + 2; n += 32) s[n] = 0.0;
  __syncwarp();
  double r[kOrder + 1], a[kOrder + 1];
  warp_autocorr(s, LEN, lane, r);
  bool ok = levinson(r, a);
  __shared__ double out_s[kAWarps][kOrder];
  double* out = out_s[warp];
  bool ok2 = false;
  if (ok) ok2 = warp_poly2lsf(a, lane, roots_s[warp], count_s[warp], out);
  __syncwarp();
  if (lane < kOrder) {
    const double v = (ok && ok2) ? out[lane] : nan("");
    if (lsf) lsf[frame * kOrder + lane] = v;
    if (lsf32) lsf32[frame * kOrder + lane] = (float)v;   // the cast the reference does when feeding lpc_x (float32 placeholder)
  }
  if (!(ok && ok2) && lane == 0 && status) atomicAdd(status, 1);
}

// ------------------------------------------------------------------------------------------------
// nsc_lsf2poly: spectrum.lsf2poly per row (lpc_utilities.py:28-33)
// ------------------------------------------------------------------------------------------------
__global__ void lsf2poly_kernel(const float* __restrict__ lsf, int64_t B, float* __restrict__ poly,
                                int* __restrict__ status) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  double w[kOrder];
  bool ok = true;
#pragma unroll
  for (int i = 0; i < kOrder; ++i) {
    w[i] = (double)lsf[f * kOrder + i];
    if (!(w[i] >= 0.0 && w[i] <= kPi)) ok = false;   // spectrum raises ValueError outside [0, pi]
  }
  float* out = poly + f * (kOrder + 1);
  if (!ok) {
#pragma unroll
    for (int i = 0; i <= kOrder; ++i) out[i] = nanf("");
    if (status) atomicAdd(status, 1);
    return;
  }
  // even-indexed LSFs -> Q, odd-indexed -> P ; each polynomial is a product of (1 - 2cos(w) z^-1 + z^-2)
  double P[kOrder + 1], Q[kOrder + 1];
#pragma unroll
  for (int i = 0; i <= kOrder; ++i) P[i] = Q[i] = 0.0;
  P[0] = Q[0] = 1.0;
#pragma unroll
  for (int j = 0; j < kOrder / 2; ++j) {
    const double cq = -2.0 * cos(w[2 * j]);
    const double cp = -2.0 * cos(w[2 * j + 1]);
#pragma unroll
    for (int i = 2 * j + 2; i >= 0; --i) {
      const double q1 = i >= 1 ? Q[i - 1] : 0.0, q2 = i >= 2 ? Q[i - 2] : 0.0;
      const double p1 = i >= 1 ? P[i - 1] : 0.0, p2 = i >= 2 ? P[i - 2] : 0.0;
      Q[i] = Q[i] + cq * q1 + q2;
      P[i] = P[i] + cp * p1 + p2;
    }
  }
  // P1 = P * (1 - z^-1), Q1 = Q * (1 + z^-1); a = (P1 + Q1)/2 without the last coefficient
#pragma unroll
  for (int i = 0; i <= kOrder; ++i) {
    const double p1 = P[i] - (i >= 1 ? P[i - 1] : 0.0);
    const double q1 = Q[i] + (i >= 1 ? Q[i - 1] : 0.0);
    out[i] = (float)(0.5 * (p1 + q1));
  }
}

// ------------------------------------------------------------------------------------------------
// nsc_lpc_residual: lpc_utilities.py:37-77
// ------------------------------------------------------------------------------------------------
constexpr int kFrame = NSC_FRAME_LENGTH;

__global__ void __launch_bounds__(kFrame)
lpc_residual_kernel(const float* __restrict__ x, const float* __restrict__ poly, int64_t B,
                    float* __restrict__ res) {
  __shared__ float xs[kFrame];
  __shared__ double as[kOrder + 1];
  const int64_t f = blockIdx.x;
  const int n = threadIdx.x;
  xs[n] = x[f * kFrame + n];
  if (n <= kOrder) as[n] = (double)poly[f * (kOrder + 1) + n];
  __syncthreads();
  constexpr int SUB = kFrame / 4, HALF = SUB / 2;  // 128, 64
  double total = 0.0;
  const int s_hi = n / HALF;          // sub-frame starting at 64*s_hi covers n with j < 64
  const int s_lo = s_hi - 1;          // previous sub-frame covers n with j >= 64
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int s = pass == 0 ? s_lo : s_hi;   // the reference accumulates sub-frames in ascending order
    if (s < 0 || s > 6) continue;
    const int j = n - s * HALF;              // 0..127
    const int kmax = j < kOrder ? j : kOrder;
    double acc = 0.0;
    for (int k = 0; k <= kmax; ++k) acc = fma(as[k], (double)xs[n - k], acc);  // product exact in fp64
    double w;
    if (s == 0) w = j < HALF ? 1.0 : np_hanning(j, SUB);
    else if (s == 6) w = j < HALF ? np_hanning(j, SUB) : 1.0;
    else w = np_hanning(j, SUB);
    total = __dadd_rn(total, __dmul_rn(acc, w));
  }
  res[f * kFrame + n] = (float)total;
}

// ------------------------------------------------------------------------------------------------
// nsc_lpc_synth: lpc_utilities.py:137-156
// ------------------------------------------------------------------------------------------------
constexpr int kSynFrames = 64;   // frames (= threads) per CTA
constexpr int kSynChunk = 32;    // samples staged per step

__global__ void __launch_bounds__(kSynFrames)
lpc_synth_kernel(const float* __restrict__ poly, const float* __restrict__ res, int64_t B,
                 float* __restrict__ y) {
  __shared__ float tile[kSynFrames][kSynChunk + 1];
  const int t = threadIdx.x;
  const int64_t f0 = (int64_t)blockIdx.x * kSynFrames;
  const int64_t f = f0 + t;
  const bool live = f < B;
  double a[kOrder + 1];
#pragma unroll
  for (int k = 0; k <= kOrder; ++k) a[k] = live ? (double)poly[f * (kOrder + 1) + k] : (k == 0 ? 1.0 : 0.0);
  const double a0 = a[0];
  double h[kOrder];
#pragma unroll
  for (int k = 0; k < kOrder; ++k) h[k] = 0.0;   // h[(n-1-k) mod 16] layout: h[i] holds y[n'] with n' mod 16 == i

  const int lane = t & 31, wrp = t >> 5;
  constexpr int NW = kSynFrames / 32;
  for (int c0 = 0; c0 < kFrame; c0 += kSynChunk) {
    // coalesced load: each warp reads 32 consecutive samples of one frame per step
    for (int fr = wrp; fr < kSynFrames; fr += NW) {
      const int64_t ff = f0 + fr;
      tile[fr][lane] = ff < B ? res[ff * kFrame + c0 + lane] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSynChunk; ++j) {
      // y[n] = (x[n] - sum_{k=1..16} a_k y[n-k]) / a0 ; y[n-k] sits in h[(j-k) mod 16] (chunk is a multiple of 16)
      double acc = (double)tile[t][j];
#pragma unroll
      for (int k = kOrder; k >= 1; --k) acc = fma(-a[k], h[(j - k + 2 * kOrder) % kOrder], acc);
      acc = acc / a0;
      h[j % kOrder] = acc;
      tile[t][j] = (float)acc;
    }
    __syncthreads();
    for (int fr = wrp; fr < kSynFrames; fr += NW) {
      const int64_t ff = f0 + fr;
      if (ff < B) y[ff * kFrame + c0 + lane] = tile[fr][lane];
    }
    __syncthreads();
  }
}

}  // namespace nsc

extern "C" {

int nsc_lpc_analyze(const float* windows, int64_t N, double* lsf_out, float* lsf_out_f32, int32_t* status,
                    void* stream) {
  if (N == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(windows && (lsf_out || lsf_out_f32), "nsc_lpc_analyze: null pointer");
  nsc::ProfScope prof((cudaStream_t)stream, "lpc_analyze", (double)N * 2.0 * 17.0 * 1024.0, (double)N * (4096.0 + 128.0));
  nsc::lpc_analyze_kernel<0><<<(unsigned)nsc::ceil_div64(N, nsc::kAWarps), nsc::kAWarps * 32, 0,
                               (cudaStream_t)stream>>>(windows, N, lsf_out, lsf_out_f32, status);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_lpc_analyze_train(const float* frames, int64_t B, double* lsf_out, float* lsf_out_f32, int32_t* status,
                          void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(frames && (lsf_out || lsf_out_f32), "nsc_lpc_analyze_train: null pointer");
  nsc::lpc_analyze_kernel<1><<<(unsigned)nsc::ceil_div64(B, nsc::kAWarps), nsc::kAWarps * 32, 0,
                               (cudaStream_t)stream>>>(frames, B, lsf_out, lsf_out_f32, status);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_lsf2poly(const float* lsf, int64_t B, float* poly, int32_t* status, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(lsf && poly, "nsc_lsf2poly: null pointer");
  nsc::ProfScope prof((cudaStream_t)stream, "lsf2poly", (double)B * 600.0, (double)B * (64.0 + 68.0));
  nsc::lsf2poly_kernel<<<(unsigned)nsc::ceil_div64(B, 128), 128, 0, (cudaStream_t)stream>>>(lsf, B, poly, status);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_lpc_residual(const float* x, const float* poly, int64_t B, float* res, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(x && poly && res, "nsc_lpc_residual: null pointer");
  nsc::ProfScope prof((cudaStream_t)stream, "lpc_residual", (double)B * 2.0 * 15232.0, (double)B * 4164.0);
  nsc::lpc_residual_kernel<<<(unsigned)B, nsc::kFrame, 0, (cudaStream_t)stream>>>(x, poly, B, res);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_lpc_synth(const float* poly, const float* res, int64_t B, float* y, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(poly && res && y, "nsc_lpc_synth: null pointer");
  nsc::ProfScope prof((cudaStream_t)stream, "lpc_synth", (double)B * 2.0 * 8192.0, (double)B * 4164.0);
  nsc::lpc_synth_kernel<<<(unsigned)nsc::ceil_div64(B, nsc::kSynFrames), nsc::kSynFrames, 0,
                          (cudaStream_t)stream>>>(poly, res, B, y);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // extern "C"
