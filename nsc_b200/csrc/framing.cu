// The steps either side of the codec pass, all plain HBM-bound gathers / scans:
//   * utterance -> 512-sample frames at hop 480 (utilities.py:25-39) and the 1024-sample LPC windows that
//     lpc_analysis_at_test cuts out of the FLATTENED frame matrix (lpc_utilities.py:98-104 -- frames overlap by 32
//     samples, so the flattened signal repeats 32 samples at every 512-sample boundary; reproduced literally);
//   * trapezoid-Hann overlap-add of the decoded frames (utilities.py:7-22; cmrl.py:595-597, :710-716);
//   * utterance-level filters: high-pass biquad, pre-emphasis, de-emphasis (lpc_utilities.py:8-11, cmrl.py:671, :735)
//     as a chunked parallel scan in float64 (the reference runs them in Python floats through audiolazy);
//   * fixed-width bit packing of the hard codes (the reference has no bitstream; bitrate is estimated from entropy).
#include <math.h>

#include "common.cuh"

namespace nsc {
namespace {

constexpr int kFrame = NSC_FRAME_LENGTH, kOverlap = 32, kHop = kFrame - kOverlap;

// np.hanning(M)[n] = 0.5 - 0.5 cos(2 pi n / (M - 1))
__device__ __forceinline__ double hanning(int n, int M) { return 0.5 - 0.5 * cos(2.0 * 3.14159265358979323846 * n / (M - 1)); }

// utilities.py:10-12 `the_window`: hanning(63)[:32], ones(448), hanning(63)[31:]
__device__ __forceinline__ double the_window(int i) {
  if (i < kOverlap) return hanning(i, 2 * kOverlap - 1);
  if (i >= kFrame - kOverlap) return hanning(i - (kFrame - kOverlap) + kOverlap - 1, 2 * kOverlap - 1);
  return 1.0;
}
// utilities.py:14 `first_window`: ones(480), hanning(64)[32:]   /  :15 `last_window`: hanning(64)[:32], ones(480)
__device__ __forceinline__ double first_window(int i) { return i >= kFrame - kOverlap ? hanning(i - (kFrame - kOverlap) + kOverlap, 2 * kOverlap) : 1.0; }
__device__ __forceinline__ double last_window(int i) { return i < kOverlap ? hanning(i, 2 * kOverlap) : 1.0; }
__device__ __forceinline__ double ola_window(int i, int64_t j, int64_t seg_amount) {
  return j == 0 ? first_window(i) : (j == seg_amount - 1 ? last_window(i) : the_window(i));
}

// (blockIdx.y = signal of a batch of equal-length utterances: utt_stride / seg_stride floats apart)
__global__ void segment_kernel(const float* __restrict__ utt, int64_t offset, int post_window, float* __restrict__ seg, int64_t N,
                               int64_t utt_stride = 0, int64_t seg_stride = 0) {
  utt += blockIdx.y * utt_stride;
  seg += blockIdx.y * seg_stride;
  const int64_t total = N * kFrame;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = k / kFrame;
    const int i = (int)(k - j * kFrame);
    const float v = utt[offset + j * kHop + i];
    seg[k] = post_window ? v : (float)((double)v * the_window(i));
  }
}

__global__ void lpc_windows_kernel(const float* __restrict__ utt, float* __restrict__ win, int64_t Nw, int64_t utt_stride = 0,
                                   int64_t win_stride = 0) {
  utt += blockIdx.y * utt_stride;
  win += blockIdx.y * win_stride;
  const int64_t total = Nw * 2 * kFrame;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = k / (2 * kFrame);
    const int64_t m = w * kFrame + (k - w * 2 * kFrame);   // index into the flattened (N, 512) hop-480 frame matrix
    win[k] = utt[(m / kFrame) * kHop + (m % kFrame)];
  }
}

// out[t] = sum_j w_j[t - 480 j] * frames[j][t - 480 j] over the (at most two) frames covering t: a gather, no atomics
__global__ void overlap_add_kernel(const float* __restrict__ frames, int64_t n_used, int64_t seg_amount, float* __restrict__ out,
                                   int64_t out_len, int64_t frames_stride = 0, int64_t out_stride = 0) {
  frames += blockIdx.y * frames_stride;
  out += blockIdx.y * out_stride;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < out_len; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j1 = t / kHop;
    double acc = 0.0;
    for (int64_t j = j1 - 1; j <= j1; ++j) {
      if (j < 0 || j >= n_used) continue;
      const int64_t i = t - j * kHop;
      if (i < 0 || i >= kFrame) continue;
      acc += (double)frames[j * kFrame + i] * ola_window((int)i, j, seg_amount);
    }
    out[t] = (float)acc;
  }
}

// ---- second-order recursive filter over a long signal, float64, chunked scan ---------------------------------------
//   y[n] = b0 x[n] + b1 x[n-1] + b2 x[n-2] - a1 y[n-1] - a2 y[n-2]        zero initial state
// pass 1: every chunk's zero-state response and final state;  pass 2: chunk-to-chunk state carry (one thread per signal);
// pass 3: add the homogeneous response of the carried-in state.  State s = (y[n], y[n-1]),  s <- A s + (u, 0).
constexpr int kScanChunk = 256;

struct Biquad { double b0, b1, b2, a1, a2; };

__global__ void iir_pass1_kernel(const float* __restrict__ x, int64_t T, int64_t n_sig, Biquad f, double* __restrict__ yz,
                                 double* __restrict__ state, int64_t n_chunks) {
  const int64_t total = n_sig * n_chunks;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t sig = k / n_chunks, c = k - sig * n_chunks;
    const float* xs = x + sig * T;
    double* ys = yz + sig * T;
    const int64_t n0 = c * kScanChunk, n1 = n0 + kScanChunk < T ? n0 + kScanChunk : T;
    double y1 = 0.0, y2 = 0.0;
    double x1 = n0 >= 1 ? (double)xs[n0 - 1] : 0.0, x2 = n0 >= 2 ? (double)xs[n0 - 2] : 0.0;   // the FIR part sees the real past input
    for (int64_t n = n0; n < n1; ++n) {
      const double xn = (double)xs[n];
      const double y = f.b0 * xn + f.b1 * x1 + f.b2 * x2 - f.a1 * y1 - f.a2 * y2;
      ys[n] = y;
      x2 = x1; x1 = xn; y2 = y1; y1 = y;
    }
    state[2 * k] = y1;
    state[2 * k + 1] = y2;
  }
}

// carry[c] = state entering chunk c:  carry[c + 1] = M carry[c] + z[c]  with  M = A^Lc  (2 x 2, built once per thread by running the
// homogeneous recursion on the two basis vectors), so the sequential part is 4 FMAs per chunk, not Lc steps
__global__ void iir_pass2_kernel(int64_t T, int64_t n_sig, Biquad f, double* __restrict__ state, int64_t n_chunks) {
  const int64_t sig = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (sig >= n_sig) return;
  (void)T;
  double m11 = 1.0, m21 = 0.0, m12 = 0.0, m22 = 1.0;      // columns = images of (1,0) and (0,1)
  for (int n = 0; n < kScanChunk; ++n) {
    const double h = -f.a1 * m11 - f.a2 * m21; m21 = m11; m11 = h;
    const double g = -f.a1 * m12 - f.a2 * m22; m22 = m12; m12 = g;
  }
  double* st = state + sig * n_chunks * 2;
  double c1 = 0.0, c2 = 0.0;      // state entering the current chunk
  for (int64_t c = 0; c < n_chunks; ++c) {
    const double z1 = st[2 * c], z2 = st[2 * c + 1];
    st[2 * c] = c1;
    st[2 * c + 1] = c2;
    const double n1 = m11 * c1 + m12 * c2 + z1, n2 = m21 * c1 + m22 * c2 + z2;   // (only full chunks are ever carried out of)
    c1 = n1;
    c2 = n2;
  }
}

__global__ void iir_pass3_kernel(int64_t T, int64_t n_sig, Biquad f, const double* __restrict__ yz, const double* __restrict__ state,
                                 int64_t n_chunks, float* __restrict__ y32, double* __restrict__ y64) {
  const int64_t total = n_sig * n_chunks;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t sig = k / n_chunks, c = k - sig * n_chunks;
    const int64_t n0 = c * kScanChunk, n1 = n0 + kScanChunk < T ? n0 + kScanChunk : T;
    double h1 = state[2 * k], h2 = state[2 * k + 1];
    for (int64_t n = n0; n < n1; ++n) {
      const double h = -f.a1 * h1 - f.a2 * h2;
      h2 = h1; h1 = h;
      const double y = yz[sig * T + n] + h;
      if (y32) y32[sig * T + n] = (float)y;
      if (y64) y64[sig * T + n] = y;
    }
  }
}

// ---- fixed-width bit packing (little-endian bit order inside the stream of one row) --------------------------------
__global__ void pack_kernel(const uint8_t* __restrict__ idx, int64_t rows, int L, int bits, uint8_t* __restrict__ out, int row_bytes) {
  const int64_t total = rows * row_bytes;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k / row_bytes;
    const int b = (int)(k - r * row_bytes);
    const uint8_t* src = idx + r * L;
    uint32_t byte = 0;
    for (int bit = 0; bit < 8; ++bit) {
      const int pos = 8 * b + bit;
      const int c = pos / bits;
      if (c < L) byte |= ((uint32_t)(src[c] >> (pos - c * bits)) & 1u) << bit;
    }
    out[k] = (uint8_t)byte;
  }
}

__global__ void unpack_kernel(const uint8_t* __restrict__ in, int64_t rows, int L, int bits, uint8_t* __restrict__ idx, int row_bytes) {
  const int64_t total = rows * L;
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k / L;
    const int c = (int)(k - r * L);
    const uint8_t* src = in + r * row_bytes;
    uint32_t v = 0;
    for (int bit = 0; bit < bits; ++bit) {
      const int pos = c * bits + bit;
      v |= ((uint32_t)(src[pos >> 3] >> (pos & 7)) & 1u) << bit;
    }
    idx[k] = (uint8_t)v;
  }
}

inline int grid_for(int64_t total, int threads) {
  const int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148LL * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace
}  // namespace nsc

using namespace nsc;

extern "C" {

int64_t nsc_segment_count(int64_t T) { return T > kFrame ? (T - kFrame + kHop - 1) / kHop : 0; }

int64_t nsc_lpc_window_count(int64_t n_segments) { return n_segments >= 3 ? n_segments - 2 : 0; }

int nsc_utterance_to_segment(const float* utterance, int64_t T, int64_t offset, int32_t post_window, float* segments, void* stream) {
  NSC_CHECK_ARG(offset >= 0 && offset <= T, "nsc_utterance_to_segment: offset %lld outside the signal", (long long)offset);
  const int64_t N = nsc_segment_count(T - offset);
  if (N == 0) return NSC_OK;
  NSC_CHECK_ARG(utterance && segments, "nsc_utterance_to_segment: null pointer");
  ProfScope prof((cudaStream_t)stream, "utterance_to_segment", 0.0, 8.0 * N * kFrame);
  segment_kernel<<<grid_for(N * kFrame, 256), 256, 0, (cudaStream_t)stream>>>(utterance, offset, post_window, segments, N);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_lpc_windows(const float* utterance, int64_t T, float* windows, void* stream) {
  const int64_t Nw = nsc_lpc_window_count(nsc_segment_count(T));
  if (Nw == 0) return NSC_OK;
  NSC_CHECK_ARG(utterance && windows, "nsc_lpc_windows: null pointer");
  ProfScope prof((cudaStream_t)stream, "lpc_windows", 0.0, 8.0 * Nw * 2 * kFrame);
  lpc_windows_kernel<<<grid_for(Nw * 2 * kFrame, 256), 256, 0, (cudaStream_t)stream>>>(utterance, windows, Nw);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_overlap_add(const float* frames, int64_t n_used, int64_t seg_amount, float* out, int64_t out_len, void* stream) {
  if (out_len == 0) return NSC_OK;
  NSC_CHECK_ARG(out != nullptr && (frames != nullptr || n_used == 0), "nsc_overlap_add: null pointer");
  NSC_CHECK_ARG(n_used >= 0 && seg_amount >= n_used, "nsc_overlap_add: n_used=%lld seg_amount=%lld", (long long)n_used, (long long)seg_amount);
  ProfScope prof((cudaStream_t)stream, "overlap_add", 0.0, 4.0 * (n_used * kFrame + out_len));
  overlap_add_kernel<<<grid_for(out_len, 256), 256, 0, (cudaStream_t)stream>>>(frames, n_used, seg_amount, out, out_len);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

// ---- batches of equal-length utterances (a corpus is coded utterance by utterance in the reference, cmrl.py:666-737; at 10 s per
// utterance the per-utterance launches, not the work, were a third of the corpus workload's time) --------------------------------
static int grid_y_ok(int64_t n, const char* what) {
  NSC_CHECK_ARG(n >= 1 && n <= 65535, "%s: n_signals=%lld (1..65535)", what, (long long)n);
  return NSC_OK;
}

int nsc_utterances_to_segments(const float* utterances, int64_t T, int64_t n_signals, int64_t offset, int32_t post_window,
                               int64_t n_take, float* segments, void* stream) {
  NSC_CHECK_ARG(offset >= 0 && offset <= T, "nsc_utterances_to_segments: offset %lld outside the signal", (long long)offset);
  NSC_CHECK_ARG(n_take >= 0 && n_take <= nsc_segment_count(T - offset), "nsc_utterances_to_segments: n_take=%lld", (long long)n_take);
  if (n_take == 0 || n_signals == 0) return NSC_OK;
  NSC_TRY(grid_y_ok(n_signals, "nsc_utterances_to_segments"));
  NSC_CHECK_ARG(utterances && segments, "nsc_utterances_to_segments: null pointer");
  ProfScope prof((cudaStream_t)stream, "utterance_to_segment", 0.0, 8.0 * n_signals * n_take * kFrame);
  int64_t gx = grid_for(n_take * kFrame, 256);
  if (gx > 64) gx = 64;
  segment_kernel<<<dim3((unsigned)gx, (unsigned)n_signals), 256, 0, (cudaStream_t)stream>>>(utterances, offset, post_window, segments, n_take, T,
                                                                                             n_take * kFrame);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_lpc_windows_batch(const float* utterances, int64_t T, int64_t n_signals, int64_t n_take, float* windows, void* stream) {
  NSC_CHECK_ARG(n_take >= 0 && n_take <= nsc_lpc_window_count(nsc_segment_count(T)), "nsc_lpc_windows_batch: n_take=%lld", (long long)n_take);
  if (n_take == 0 || n_signals == 0) return NSC_OK;
  NSC_TRY(grid_y_ok(n_signals, "nsc_lpc_windows_batch"));
  NSC_CHECK_ARG(utterances && windows, "nsc_lpc_windows_batch: null pointer");
  ProfScope prof((cudaStream_t)stream, "lpc_windows", 0.0, 8.0 * n_signals * n_take * 2 * kFrame);
  int64_t gx = grid_for(n_take * 2 * kFrame, 256);
  if (gx > 64) gx = 64;
  lpc_windows_kernel<<<dim3((unsigned)gx, (unsigned)n_signals), 256, 0, (cudaStream_t)stream>>>(utterances, windows, n_take, T, n_take * 2 * kFrame);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_overlap_add_batch(const float* frames, int64_t n_signals, int64_t n_used, int64_t seg_amount, float* out, int64_t out_len,
                          void* stream) {
  if (out_len == 0 || n_signals == 0) return NSC_OK;
  NSC_TRY(grid_y_ok(n_signals, "nsc_overlap_add_batch"));
  NSC_CHECK_ARG(out != nullptr && (frames != nullptr || n_used == 0), "nsc_overlap_add_batch: null pointer");
  NSC_CHECK_ARG(n_used >= 0 && seg_amount >= n_used, "nsc_overlap_add_batch: n_used=%lld seg_amount=%lld", (long long)n_used, (long long)seg_amount);
  ProfScope prof((cudaStream_t)stream, "overlap_add", 0.0, 4.0 * n_signals * (n_used * kFrame + out_len));
  int64_t gx = grid_for(out_len, 256);
  if (gx > 64) gx = 64;
  overlap_add_kernel<<<dim3((unsigned)gx, (unsigned)n_signals), 256, 0, (cudaStream_t)stream>>>(frames, n_used, seg_amount, out, out_len,
                                                                                                 n_used * kFrame, out_len);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int64_t nsc_iir_workspace_bytes(int64_t T, int64_t n_signals) {
  const int64_t n_chunks = (T + kScanChunk - 1) / kScanChunk;
  return (n_signals * T + 2 * n_signals * n_chunks) * (int64_t)sizeof(double) + 512;
}

int nsc_iir_biquad(const float* x, int64_t T, int64_t n_signals, const double* b_host, const double* a_host, float* y_f32,
                   double* y_f64, void* workspace, int64_t workspace_bytes, void* stream) {
  if (T == 0 || n_signals == 0) return NSC_OK;
  NSC_CHECK_ARG(x && b_host && a_host && workspace && (y_f32 || y_f64), "nsc_iir_biquad: null pointer");
  NSC_CHECK_ARG(a_host[0] == 1.0, "nsc_iir_biquad: a[0] must be 1");
  if (workspace_bytes < nsc_iir_workspace_bytes(T, n_signals)) {
    set_error("nsc_iir_biquad: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)nsc_iir_workspace_bytes(T, n_signals));
    return NSC_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_chunks = (T + kScanChunk - 1) / kScanChunk;
  double* yz = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  double* state = yz + n_signals * T;
  const Biquad f{b_host[0], b_host[1], b_host[2], a_host[1], a_host[2]};
  ProfScope prof(st, "iir_biquad", 0.0, (double)n_signals * T * (4.0 + 8.0 + 8.0 + 4.0));
  iir_pass1_kernel<<<grid_for(n_signals * n_chunks, 128), 128, 0, st>>>(x, T, n_signals, f, yz, state, n_chunks);
  NSC_LAUNCH_OK();
  iir_pass2_kernel<<<(unsigned)((n_signals + 63) / 64), 64, 0, st>>>(T, n_signals, f, state, n_chunks);
  NSC_LAUNCH_OK();
  iir_pass3_kernel<<<grid_for(n_signals * n_chunks, 128), 128, 0, st>>>(T, n_signals, f, yz, state, n_chunks, y_f32, y_f64);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int32_t nsc_packed_row_bytes(int32_t L, int32_t bits) { return (L * bits + 7) / 8; }

int nsc_pack_codes(const uint8_t* idx, int64_t rows, int32_t L, int32_t bits, uint8_t* packed, void* stream) {
  if (rows == 0) return NSC_OK;
  NSC_CHECK_ARG(idx && packed && L > 0 && bits >= 1 && bits <= 8, "nsc_pack_codes: bad argument");
  const int rb = nsc_packed_row_bytes(L, bits);
  ProfScope prof((cudaStream_t)stream, "pack_codes", 0.0, (double)rows * (L + rb));
  pack_kernel<<<grid_for(rows * rb, 256), 256, 0, (cudaStream_t)stream>>>(idx, rows, L, bits, packed, rb);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_unpack_codes(const uint8_t* packed, int64_t rows, int32_t L, int32_t bits, uint8_t* idx, void* stream) {
  if (rows == 0) return NSC_OK;
  NSC_CHECK_ARG(idx && packed && L > 0 && bits >= 1 && bits <= 8, "nsc_unpack_codes: bad argument");
  const int rb = nsc_packed_row_bytes(L, bits);
  ProfScope prof((cudaStream_t)stream, "unpack_codes", 0.0, (double)rows * (L + rb));
  unpack_kernel<<<grid_for(rows * L, 256), 256, 0, (cudaStream_t)stream>>>(packed, rows, L, bits, idx, rb);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // extern "C"
