// K8 -- time-domain RMSE (mse_loss, loss_terms_and_measures.py:77-79) and the 4-resolution mel loss
// (tf_stft :178-183, mfcc_transform :130-148, mfcc_loss :151-175).
//
// One CTA per frame.  The decoded and the original frame are packed as the real and imaginary parts of ONE
// length-512 complex FFT (radix-2 in shared memory); both 257-bin spectra are recovered from Z[k] and
// conj(Z[512-k]).  The four HTK filterbanks (8/16/32/128 bands = 184 columns) are triangular, so each of the
// 184 threads walks only the support [lo, hi] of its column.  HBM traffic per frame is the 4 KB of the two
// signals plus 8 B of results; the 189 KB filterbank stays in L2.
#include "common.cuh"

namespace nsc {

constexpr int kN = NSC_FRAME_LENGTH;      // 512
constexpr int kBins = NSC_MEL_BINS;       // 257
constexpr int kMel = NSC_MEL_TOTAL;       // 184
// column offsets of the 8/16/32/128-band banks inside the 184-column matrix
__host__ __device__ __forceinline__ int mel_off(int i) { return i == 0 ? 0 : i == 1 ? 8 : i == 2 ? 24 : i == 3 ? 56 : 184; }

// melw layout: [257][184] weights, then int32 lo[184], hi[184] (support of each column; lo > hi when empty)
__global__ void mel_weights_kernel(float* __restrict__ melw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kBins * kMel) return;
  const int k = i / kMel, col = i % kMel;
  int bank = 0;
  while (col >= mel_off(bank + 1)) ++bank;
  const int nmel = mel_off(bank + 1) - mel_off(bank);
  const int m = col - mel_off(bank);
  float wv = 0.f;
  if (k >= 1) {  // bands_to_zero = 1: the DC row is zero
    // tf.signal.linear_to_mel_weight_matrix, HTK scale, float64 then cast
    const double nyq = 8000.0;
    const double f = nyq * (double)k / (double)(kBins - 1);
    const double mel = 1127.0 * log(1.0 + f / 700.0);
    const double mel_hi = 1127.0 * log(1.0 + nyq / 700.0);
    const double step = mel_hi / (double)(nmel + 1);
    const double lower = step * (double)m;
    const double center = step * (double)(m + 1);
    const double upper = (m + 2 == nmel + 1) ? mel_hi : step * (double)(m + 2);
    const double ls = (mel - lower) / (center - lower);
    const double us = (upper - mel) / (upper - center);
    const double v = fmax(0.0, fmin(ls, us));
    wv = (float)v;
  }
  melw[i] = wv;
}

__global__ void mel_ranges_kernel(float* __restrict__ melw) {
  const int col = threadIdx.x;
  if (col >= kMel) return;
  int lo = kBins, hi = -1;
  for (int k = 0; k < kBins; ++k) {
    if (melw[k * kMel + col] != 0.f) {
      if (k < lo) lo = k;
      hi = k;
    }
  }
  int* rng = reinterpret_cast<int*>(melw + kBins * kMel);
  rng[col] = lo;
  rng[kMel + col] = hi;
}

__global__ void __launch_bounds__(256)
losses_kernel(const float* __restrict__ dec, const float* __restrict__ ori, int64_t B,
              const float* __restrict__ melw, float* __restrict__ time_loss, float* __restrict__ freq_loss) {
  __shared__ float2 z[kN];
  __shared__ float psd_d[kBins + 3], psd_o[kBins + 3];
  __shared__ float red[8];
  __shared__ float dsq[kMel];
  const int tid = threadIdx.x;
  const int64_t f = blockIdx.x;
  const float* d = dec + f * kN;
  const float* o = ori + f * kN;

  float se = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = tid + h * 256;
    const float dv = d[n], ov = o[n];
    const float e = dv - ov;
    se = fmaf(e, e, se);
    z[__brev((unsigned)n) >> 23] = make_float2(dv, ov);
  }
  // ---- RMSE (mse_loss)
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) se += __shfl_xor_sync(0xffffffffu, se, s);
  if ((tid & 31) == 0) red[tid >> 5] = se;
  __syncthreads();
  if (tid == 0 && time_loss) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    time_loss[f] = sqrtf(t / (float)kN + 1e-07f);
  }
  if (freq_loss == nullptr) return;

  // ---- 512-point complex FFT, decimation in time (input already bit-reversed)
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1);
    const int j = tid & (half - 1);
    const int i0 = ((tid >> (s - 1)) << s) + j;
    const int i1 = i0 + half;
    float sn, cs;
    sincospif(-(float)j / (float)half, &sn, &cs);   // exp(-2*pi*i*j / 2^s)
    const float2 u = z[i0], v = z[i1];
    const float2 vt = make_float2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    z[i0] = make_float2(u.x + vt.x, u.y + vt.y);
    z[i1] = make_float2(u.x - vt.x, u.y - vt.y);
    __syncthreads();
  }
  // ---- split the two real spectra, PSD = (1/512) * (sqrt(re^2 + im^2 + 1e-7))^2
  for (int k = tid; k < kBins; k += 256) {
    const float2 a = z[k];
    const float2 c = z[(kN - k) & (kN - 1)];
    const float dr = 0.5f * (a.x + c.x), di = 0.5f * (a.y - c.y);     // D = (Z[k] + conj(Z[N-k])) / 2
    const float orr = 0.5f * (a.y + c.y), oi = -0.5f * (a.x - c.x);   // O = (Z[k] - conj(Z[N-k])) / (2i)
    const float md = sqrtf(dr * dr + di * di + 1e-7f);
    const float mo = sqrtf(orr * orr + oi * oi + 1e-7f);
    psd_d[k] = (1.0f / (float)kN) * (md * md);
    psd_o[k] = (1.0f / (float)kN) * (mo * mo);
  }
  __syncthreads();
  // ---- mel banks + log + squared difference
  if (tid < kMel) {
    const int* rng = reinterpret_cast<const int*>(melw + kBins * kMel);
    const int lo = rng[tid], hi = rng[kMel + tid];
    float ad = 0.f, ao = 0.f;
    for (int k = lo; k <= hi; ++k) {
      const float w = melw[k * kMel + tid];
      ad = fmaf(psd_d[k], w, ad);
      ao = fmaf(psd_o[k], w, ao);
    }
    const float e = logf(ad + 1e-7f) - logf(ao + 1e-7f);
    dsq[tid] = e * e;
  }
  __syncthreads();
  if (tid < 4) {
    float s = 0.f;
    for (int m = mel_off(tid); m < mel_off(tid + 1); ++m) s += dsq[m];
    red[tid] = sqrtf(s / (float)(mel_off(tid + 1) - mel_off(tid)) + 1e-07f);
  }
  __syncthreads();
  if (tid == 0) freq_loss[f] = (red[0] + red[1] + red[2] + red[3]) / 4.0f;
}

}  // namespace nsc

extern "C" {

int nsc_mel_filterbank(float* melw, void* stream) {
  NSC_CHECK_ARG(melw != nullptr, "nsc_mel_filterbank: null pointer");
  const int total = nsc::kBins * nsc::kMel;
  nsc::mel_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(melw);
  NSC_LAUNCH_OK();
  nsc::mel_ranges_kernel<<<1, 192, 0, (cudaStream_t)stream>>>(melw);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_losses_forward(const float* decoded, const float* original, int64_t B, const float* melw,
                       float* time_loss, float* freq_loss, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(decoded && original, "nsc_losses_forward: null signal");
  NSC_CHECK_ARG(freq_loss == nullptr || melw != nullptr, "nsc_losses_forward: mel loss requested without filterbank");
  nsc::ProfScope prof((cudaStream_t)stream, "losses", (double)B * 2.0e5, (double)B * 4104.0);
  nsc::losses_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(decoded, original, B, melw, time_loss, freq_loss);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // extern "C"
