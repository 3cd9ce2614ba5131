// K1 / K2 / K7 -- the LPC front and back end of collaborative quantisation (lpc_utilities.py).
//
// The reference runs these in Python float64 (audiolazy generators, spectrum + numpy.roots) behind
// tf.py_func.  They are a few 10k MAC per frame -- nothing next to the 0.3 GFLOP of one codec -- so they
// keep the reference's float64 arithmetic (B200's fp64 pipe is ample) and are organised purely for
// parallelism across frames and coalesced HBM access:
//   nsc_lpc_analyze   persistent CTAs, batches of frames in three phases: warp per frame for the 17 autocorrelation lags (lane =
//                     contiguous chunk, sums out of registers), THREAD per frame for Levinson-Durbin, warp per frame for the LSFs
//                     (Chebyshev-series sign scan over 2049 tabulated grid points, bisection on cos w, one root per lane).
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
// shared pieces of the analysis: Levinson-Durbin, Chebyshev-series evaluation
// ------------------------------------------------------------------------------------------------
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

// ------------------------------------------------------------------------------------------------
// nsc_lpc_analyze: lpc_utilities.py:112-124 (MODE 0, 1024-sample windows) and :14-25 (MODE 1, 512 frames)
//
// The roof of this kernel is the fp64 pipe, not HBM: a frame costs 1024 x 17 autocorrelation DFMAs plus 2 x 2049 x 9 for the
// sign scan of the two LSF polynomials (~55 k DFMA per 4 KB window).  Round 1's warp-per-frame kernel spent most of its time
// elsewhere: 18 shared-memory loads per 17 DFMAs in the autocorrelation, cos() per scan point and per bisection step, cospi() per
// windowed sample, and the (sequential) Levinson recursion executed redundantly by all 32 lanes.  This one is a persistent CTA
// working in batches of frames, three phases per batch:
//   A  warp per frame: window from a shared table -> padded shared buffer; each lane takes a CONTIGUOUS chunk of samples (+16 of the
//      next chunk) into registers and runs its 17 lag sums out of registers; partials are summed in a fixed order through the buffer;
//   B  THREAD per frame: Levinson-Durbin and the Chebyshev coefficients of the deflated sum / difference polynomials;
//   C  warp per frame: sign scan over the 2049 grid points with cos(pi k / 2048) from a shared table (lane = point, ballots give the
//      brackets in ascending order), bisection on x = cos w (no cos in the loop), w = acos(x), rank-merge of the two root lists.
// ------------------------------------------------------------------------------------------------
constexpr int kAWarps = 8;

template <int MODE> struct AnaShape {
  static constexpr int LEN = MODE == 0 ? 1024 : 512;
  static constexpr int CH = LEN / 32;        // samples per lane
  static constexpr int STRIDE = CH + 1;      // padded chunk stride in doubles (odd: lanes a chunk apart hit different banks)
  static constexpr int BUF = 33 * STRIDE;    // 32 chunks + a zero chunk behind the last one
  __host__ __device__ static constexpr size_t smem(int fpw) {
    return sizeof(double) * ((size_t)kAWarps * BUF + (kGrid + 1) + 512 + (size_t)kAWarps * fpw * (kOrder + 1) +
                             (size_t)kAWarps * fpw * 19 + kAWarps * 16) +
           sizeof(int) * ((size_t)kAWarps * fpw + kAWarps * 16);
  }
};

template <int MODE>
__global__ void __launch_bounds__(kAWarps * 32, 2)
lpc_analyze_kernel(const float* __restrict__ in, int64_t N, int fpw, double* __restrict__ lsf, float* __restrict__ lsf32,
                   int* __restrict__ status) {
  using S = AnaShape<MODE>;
  constexpr int LEN = S::LEN, CH = S::CH, STRIDE = S::STRIDE;
  extern __shared__ double ana_smem[];
  const int batch = kAWarps * fpw;
  double* sig_all = ana_smem;                                  // [kAWarps][BUF]
  double* cosg = sig_all + kAWarps * S::BUF;                   // [kGrid + 1]  cos(pi k / kGrid)
  double* hann = cosg + (kGrid + 1);                           // [512]        numpy.hanning(512)
  double* r_s = hann + 512;                                    // [batch][17]
  double* coef_s = r_s + batch * (kOrder + 1);                 // [batch][19]  cq[9] | cp[9]
  double* roots_all = coef_s + batch * 19;                     // [kAWarps][16]
  int* ok_s = reinterpret_cast<int*>(roots_all + kAWarps * 16);   // [batch]
  int* brk_all = ok_s + batch;                                 // [kAWarps][16]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* s = sig_all + warp * S::BUF;
  double* roots_s = roots_all + warp * 16;
  int* brk_s = brk_all + warp * 16;

  for (int k = threadIdx.x; k <= kGrid; k += blockDim.x) cosg[k] = cos(kPi * (double)k / (double)kGrid);
  if (MODE == 0)
    for (int k = threadIdx.x; k < 512; k += blockDim.x) hann[k] = np_hanning(k, 512);
  __syncthreads();

  const int64_t n_batches = (N + batch - 1) / batch;
  for (int64_t bt = blockIdx.x; bt < n_batches; bt += gridDim.x) {
    const int64_t frame0 = bt * batch;
    // ---- phase A: autocorrelation, warp per frame ------------------------------------------------------------------
    for (int q = 0; q < fpw; ++q) {
      const int slot = warp * fpw + q;
      const int64_t frame = frame0 + slot;
      if (frame >= N) break;   // warp-uniform
      const float* x = in + frame * LEN;
      if (MODE == 0) {
        // window: [hanning(512)[:256], ones(512), hanning(512)[256:]]  (lpc_utilities.py:120-121)
#pragma unroll 8
        for (int n = lane; n < LEN; n += 32) {
          const double w = n < 256 ? hann[n] : (n >= 768 ? hann[n - 512] : 1.0);
          s[(n / CH) * STRIDE + n % CH] = (double)x[n] * w;
        }
      } else {
        // highpass biquad then pre-emphasis, zero state per frame (lpc_utilities.py:8-11, :20); sequential -> lane 0
        for (int n = lane; n < LEN; n += 32) s[(n / CH) * STRIDE + n % CH] = (double)x[n];
        __syncwarp();
        if (lane == 0) {
          const double b0 = 0.989502, b1 = -1.979004, b2 = 0.989592, a1 = -1.978882, a2 = 0.979126;
          double x1 = 0, x2 = 0, y1 = 0, y2 = 0, e1 = 0;
          for (int n = 0; n < LEN; ++n) {
            double* p = s + (n / CH) * STRIDE + n % CH;
            const double xv = *p;
            const double y = b0 * xv + b1 * x1 + b2 * x2 - a1 * y1 - a2 * y2;
            x2 = x1; x1 = xv; y2 = y1; y1 = y;
            *p = y + (-0.68) * e1;   // empha_filter = 1 - 0.68 z^-1 (constants.py:64)
            e1 = y;
          }
        }
      }
      if (lane < kOrder) s[32 * STRIDE + lane] = 0.0;   // the chunk behind the last one
      __syncwarp();
      double r[kOrder + 1];
#pragma unroll
      for (int t = 0; t <= kOrder; ++t) r[t] = 0.0;
      // the lane's chunk in steps of 16 samples: 32 values in registers (16 + the 16 behind them), 16 x 17 DFMAs out of registers
#pragma unroll
      for (int h = 0; h < CH / 16; ++h) {
        double c[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int m = 16 * h + j;                     // offset in the lane's chunk; beyond CH: the next chunk's head
          c[j] = m < CH ? s[lane * STRIDE + m] : s[(lane + 1) * STRIDE + (m - CH)];
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
          for (int t = 0; t <= kOrder; ++t) r[t] = fma(c[j], c[j + t], r[t]);
        }
      }
      __syncwarp();
      // fixed-order sum of the 32 partials per lag: lag-major rows of 33 doubles, lane t adds row t
#pragma unroll
      for (int t = 0; t <= kOrder; ++t) s[t * 33 + lane] = r[t];
      __syncwarp();
      if (lane <= kOrder) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int l = 0; l < 32; l += 4) {
          a0 += s[lane * 33 + l];
          a1 += s[lane * 33 + l + 1];
          a2 += s[lane * 33 + l + 2];
          a3 += s[lane * 33 + l + 3];
        }
        r_s[slot * (kOrder + 1) + lane] = (a0 + a1) + (a2 + a3);
      }
      __syncwarp();
    }
    __syncthreads();
    // ---- phase B: Levinson-Durbin + LSF polynomials, thread per frame ---------------------------------------------
    if ((int)threadIdx.x < batch && frame0 + threadIdx.x < N) {
      const int slot = threadIdx.x;
      double r[kOrder + 1], a[kOrder + 1];
#pragma unroll
      for (int t = 0; t <= kOrder; ++t) r[t] = r_s[slot * (kOrder + 1) + t];
      const bool ok = levinson(r, a);
      // spectrum.poly2lsf: P1 = a1 - rev(a1), Q1 = a1 + rev(a1) with a1 = [a, 0]; deflate z = 1 / z = -1
      double p[kOrder + 1], q[kOrder + 1];
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
      double* cf = coef_s + slot * 19;
      cf[0] = q[8];
      cf[9] = p[8];
#pragma unroll
      for (int m = 1; m <= 8; ++m) {
        cf[m] = 2.0 * q[8 - m];
        cf[9 + m] = 2.0 * p[8 - m];
      }
      ok_s[slot] = ok ? 1 : 0;
    }
    __syncthreads();
    // ---- phase C: roots on the unit circle, warp per frame ---------------------------------------------------------
    for (int q = 0; q < fpw; ++q) {
      const int slot = warp * fpw + q;
      const int64_t frame = frame0 + slot;
      if (frame >= N) break;   // warp-uniform
      bool ok = ok_s[slot] != 0;
      double cq[9], cp[9];
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        cq[m] = coef_s[slot * 19 + m];
        cp[m] = coef_s[slot * 19 + 9 + m];
      }
      if (ok) {
        // interval k = [k, k + 1] pi / kGrid holds a root when the sign of f differs at its ends; which = 0: Q (its first root
        // precedes P's), 1: P
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const double* c = which == 0 ? cq : cp;
          int found = 0;
          unsigned prev = 0;
          for (int g = 0; g < kGrid / 32; g += 4) {
            double f[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) f[u] = cheb_eval(c, cosg[(g + u) * 32 + lane]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const unsigned mask = __ballot_sync(0xffffffffu, f[u] > 0.0);
              unsigned chg = (mask ^ (mask >> 1)) & 0x7fffffffu;      // bit b: points 32 (g+u) + b and + b + 1 differ
              if ((g + u) > 0 && ((prev >> 31) != (mask & 1u))) {     // last point of the previous group vs this group's first
                if (found < 8 && lane == 0) brk_s[which * 8 + found] = (g + u) * 32 - 1;
                ++found;
              }
              while (chg) {
                const int b = __ffs(chg) - 1;
                chg &= chg - 1;
                if (found < 8 && lane == 0) brk_s[which * 8 + found] = (g + u) * 32 + b;
                ++found;
              }
              prev = mask;
            }
          }
          const bool last_pos = cheb_eval(c, cosg[kGrid]) > 0.0;      // point kGrid closes interval kGrid - 1
          if (((prev >> 31) != 0u) != last_pos) {
            if (found < 8 && lane == 0) brk_s[which * 8 + found] = kGrid - 1;
            ++found;
          }
          if (found != 8) ok = false;
        }
      }
      __syncwarp();
      double root = 0.0;
      if (ok && lane < 16) {
        const int which = lane >> 3;
        double c[9];
#pragma unroll
        for (int m = 0; m < 9; ++m) c[m] = which == 0 ? cq[m] : cp[m];
        const int k = brk_s[lane];
        double xl = cosg[k], xr = cosg[k + 1];          // bisection on x = cos w: the bracket in w maps to [xr, xl]
        const bool l_pos = cheb_eval(c, xl) > 0.0;
        for (int it = 0; it < 52; ++it) {
          const double mid = 0.5 * (xl + xr);
          const bool mid_pos = cheb_eval(c, mid) > 0.0;
          if (mid_pos == l_pos) xl = mid; else xr = mid;
        }
        root = acos(0.5 * (xl + xr));
        roots_s[lane] = root;
      }
      __syncwarp();
      if (lane < kOrder) {
        double v = nan("");
        int rank = lane;
        if (ok) {
          // rank = own position + number of roots of the other polynomial below this one
          const int which = lane >> 3;
          rank = lane & 7;
#pragma unroll
          for (int j = 0; j < 8; ++j) rank += (roots_s[(1 - which) * 8 + j] < root) ? 1 : 0;
          v = root;
        }
        if (lsf) lsf[frame * kOrder + rank] = v;
        if (lsf32) lsf32[frame * kOrder + rank] = (float)v;   // the cast the reference does when feeding lpc_x (float32 placeholder)
      }
      if (!ok && lane == 0 && status) atomicAdd(status, 1);
      __syncwarp();
    }
    __syncthreads();   // the next batch overwrites r_s / coef_s / ok_s
  }
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
  constexpr int SUB = kFrame / 4, HALF = SUB / 2;  // 128, 64
  __shared__ float xs[kFrame];
  __shared__ double as[kOrder + 1];
  __shared__ double hw[SUB];             // hanning(128): one cospi per CTA thread instead of two per sample
  const int64_t f = blockIdx.x;
  const int n = threadIdx.x;
  xs[n] = x[f * kFrame + n];
  if (n <= kOrder) as[n] = (double)poly[f * (kOrder + 1) + n];
  if (n >= kFrame - SUB) hw[n - (kFrame - SUB)] = np_hanning(n - (kFrame - SUB), SUB);
  __syncthreads();
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
    if (s == 0) w = j < HALF ? 1.0 : hw[j];
    else if (s == 6) w = j < HALF ? hw[j] : 1.0;
    else w = hw[j];
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

namespace nsc {
template <int MODE>
static int launch_lpc_analyze(const float* in, int64_t N, double* lsf, float* lsf32, int32_t* status, cudaStream_t st) {
  // two frames per warp and batch once every SM has work for both of its CTAs; one below that (latency, not throughput)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int fpw = N >= (int64_t)sms * 2 * kAWarps * 2 ? 2 : 1;
  const size_t smem = AnaShape<MODE>::smem(fpw);
  static bool attr_set[2] = {false, false};
  if (!attr_set[MODE]) {
    NSC_CUDA_OK(cudaFuncSetAttribute(lpc_analyze_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AnaShape<MODE>::smem(2)));
    attr_set[MODE] = true;
  }
  int64_t grid = ceil_div64(N, (int64_t)kAWarps * fpw);
  if (grid > (int64_t)sms * 2) grid = (int64_t)sms * 2;
  lpc_analyze_kernel<MODE><<<(unsigned)grid, kAWarps * 32, smem, st>>>(in, N, fpw, lsf, lsf32, status);
  NSC_LAUNCH_OK();
  return NSC_OK;
}
}  // namespace nsc

extern "C" {

int nsc_lpc_analyze(const float* windows, int64_t N, double* lsf_out, float* lsf_out_f32, int32_t* status,
                    void* stream) {
  if (N == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(windows && (lsf_out || lsf_out_f32), "nsc_lpc_analyze: null pointer");
  nsc::ProfScope prof((cudaStream_t)stream, "lpc_analyze", (double)N * 2.0 * (17.0 * 1024.0 + 2.0 * 9.0 * 2049.0), (double)N * (4096.0 + 128.0));
  return nsc::launch_lpc_analyze<0>(windows, N, lsf_out, lsf_out_f32, status, (cudaStream_t)stream);
}

int nsc_lpc_analyze_train(const float* frames, int64_t B, double* lsf_out, float* lsf_out_f32, int32_t* status,
                          void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(frames && (lsf_out || lsf_out_f32), "nsc_lpc_analyze_train: null pointer");
  return nsc::launch_lpc_analyze<1>(frames, B, lsf_out, lsf_out_f32, status, (cudaStream_t)stream);
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
