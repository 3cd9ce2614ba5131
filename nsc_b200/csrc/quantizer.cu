// K5 -- soft-to-hard scalar quantiser (nn_core_operator.py:140-164), quan_loss
// (loss_terms_and_measures.py:257-259) and the soft histogram behind entropy_coding_loss (:262-267).
//
// Mapping: one warp per code row, lane = bin (NPL bins per lane: 1 for 32 bins, 8 for the 256-entry LSF
// codebook).  Distances, fp32 logits, first-index arg-max (tf.nn.top_k tie rule) and the softmax are
// reduced with warp shuffles; the per-bin histogram lives in registers across the rows a warp owns and is
// merged through shared memory -> one global atomic per bin per CTA.  Nothing of the (rows x n) soft tensor
// touches HBM unless the caller asks for it.
#include "common.cuh"

namespace nsc {

constexpr int kQWarps = 8;

template <int NPL>
__global__ void __launch_bounds__(kQWarps * 32)
quantize_kernel(const float* __restrict__ x, int64_t n_frames, int L, int frames_per_cta,
                const float* __restrict__ bins, int n, const float* __restrict__ alpha_p, float iq, int use_soft,
                float* __restrict__ out, uint8_t* __restrict__ idx, float* __restrict__ soft,
                float* __restrict__ hist, float* __restrict__ qloss) {
  extern __shared__ float smem[];
  float* bins_s = smem;                       // n
  float* hist_s = bins_s + NPL * 32;          // kQWarps * NPL*32
  float* rowterm_s = hist_s + kQWarps * NPL * 32;  // frames_per_cta * L

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < NPL * 32; k += blockDim.x) bins_s[k] = k < n ? bins[k] : 0.f;
  __syncthreads();

  const float alpha = *alpha_p;
  const int64_t frame0 = (int64_t)blockIdx.x * frames_per_cta;
  int frames_here = frames_per_cta;
  if (frame0 + frames_here > n_frames) frames_here = (int)(n_frames - frame0);
  const int rows_here = frames_here * L;
  const int64_t row0 = frame0 * L;
  const bool need_soft = use_soft || soft != nullptr || hist != nullptr || qloss != nullptr;

  float b[NPL], hacc[NPL];
#pragma unroll
  for (int j = 0; j < NPL; ++j) {
    b[j] = bins_s[lane + 32 * j];
    hacc[j] = 0.f;
  }

  for (int r = warp; r < rows_here; r += kQWarps) {
    const float xv = x[row0 + r];
    float lg[NPL];
    float best = -INFINITY;
    int besti = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      const int k = lane + 32 * j;
      // fp32 distance, then a separate fp32 multiply (no FMA contraction) -- the bit-exact contract.
      const float d = fabsf(__fsub_rn(xv, b[j]));
      lg[j] = __fmul_rn(alpha, d);
      if (k < n && (besti == 0x7fffffff || lg[j] > best)) {  // strictly greater keeps the lowest index
        best = lg[j];
        besti = k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (oi != 0x7fffffff && (besti == 0x7fffffff || ov > best || (ov == best && oi < besti))) {
        best = ov;
        besti = oi;
      }
    }
    if (besti == 0x7fffffff) besti = 0;  // all-NaN row: top_k would still return an index
    float q = bins_s[besti];
    if (need_soft) {
      float e[NPL], es = 0.f;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        e[j] = (lane + 32 * j) < n ? expf(lg[j] - best) : 0.f;
        es += e[j];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(0xffffffffu, es, o);
      float qs = 0.f, rt = 0.f;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        const int k = lane + 32 * j;
        const float s = e[j] / es;
        if (k < n) {
          qs = fmaf(s, b[j], qs);
          rt += sqrtf(s + 1e-20f);
          hacc[j] += s;
          if (soft) soft[(row0 + r) * n + k] = s;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        qs += __shfl_xor_sync(0xffffffffu, qs, o);
        rt += __shfl_xor_sync(0xffffffffu, rt, o);
      }
      if (use_soft) q = qs;
      if (lane == 0) rowterm_s[r] = rt;
    }
    if (lane == 0) {
      if (out) out[row0 + r] = __fadd_rn(__fmul_rn(1.f - iq, xv), __fmul_rn(iq, q));
      if (idx) idx[row0 + r] = (uint8_t)besti;
    }
  }

  if (hist != nullptr) {
#pragma unroll
    for (int j = 0; j < NPL; ++j) hist_s[warp * NPL * 32 + lane + 32 * j] = hacc[j];
  }
  __syncthreads();
  if (hist != nullptr) {
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kQWarps; ++w) s += hist_s[w * NPL * 32 + k];
      atomicAdd(hist + k, s);
    }
  }
  if (qloss != nullptr) {
    for (int f = threadIdx.x; f < frames_here; f += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < L; ++l) s += rowterm_s[f * L + l];
      qloss[frame0 + f] = s / (float)L;
    }
  }
}

__global__ void dequantize_kernel(const uint8_t* __restrict__ idx, int64_t rows, const float* __restrict__ bins,
                                  int n, float* __restrict__ out) {
  __shared__ float bins_s[256];
  for (int k = threadIdx.x; k < 256; k += blockDim.x) bins_s[k] = k < n ? bins[k] : 0.f;
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
    out[r] = bins_s[idx[r]];
}

__global__ void entropy_kernel(const float* __restrict__ hist, int n, float* __restrict__ ent) {
  // single warp; loss_terms_and_measures.py:262-267
  const int lane = threadIdx.x;
  float tot = 0.f;
  for (int k = lane; k < n; k += 32) tot += hist[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  const float ln2 = logf(2.0f);
  float acc = 0.f;
  for (int k = lane; k < n; k += 32) {
    const float p = hist[k] / tot;
    acc += p * logf(p + 1e-7f) / ln2;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) *ent = -acc;
}

// quan_loss on a materialised soft assignment: one CTA per frame (loss_terms_and_measures.py:257-259)
__global__ void __launch_bounds__(256)
quan_loss_kernel(const float* __restrict__ soft, int L, int n, float* __restrict__ out) {
  __shared__ float red[8];
  const float* s = soft + (int64_t)blockIdx.x * L * n;
  float acc = 0.f;
  for (int i = threadIdx.x; i < L * n; i += 256) acc += sqrtf(s[i] + 1e-20f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    out[blockIdx.x] = t / (float)L;
  }
}

// column sums of a (rows, n) soft assignment -- the histogram of entropy_coding_loss (:262-267)
__global__ void __launch_bounds__(256)
soft_hist_kernel(const float* __restrict__ soft, int64_t rows, int n, int rows_per_cta, float* __restrict__ hist) {
  __shared__ float h[256];
  for (int k = threadIdx.x; k < 256; k += 256) h[k] = 0.f;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > rows) r1 = rows;
  const int64_t e0 = r0 * n, e1 = r1 * n;
  for (int64_t i = e0 + threadIdx.x; i < e1; i += 256) atomicAdd(&h[(int)(i % n)], soft[i]);
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += 256) atomicAdd(hist + k, h[k]);
}


// Hard path only (no soft tensor, histogram or quantisation loss wanted): one THREAD per code, bins in shared memory.
// Same arithmetic as quantize_kernel -- fp32 distance, separate fp32 multiply by alpha, first maximum wins -- so the
// indices are bit-identical; it just does not spend a warp and five shuffle rounds per code.
__global__ void __launch_bounds__(256) quantize_hard_kernel(const float* __restrict__ x, int64_t rows, const float* __restrict__ bins,
                                                            int n, const float* __restrict__ alpha_p, float iq,
                                                            float* __restrict__ out, uint8_t* __restrict__ idx) {
  __shared__ float bins_s[256];
  for (int k = threadIdx.x; k < n; k += blockDim.x) bins_s[k] = bins[k];
  __syncthreads();
  const float alpha = *alpha_p;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[r];
    float best = 0.f;
    int besti = -1;
    for (int k = 0; k < n; ++k) {
      const float lg = __fmul_rn(alpha, fabsf(__fsub_rn(xv, bins_s[k])));
      if (besti < 0 || lg > best) { best = lg; besti = k; }   // strictly greater keeps the lowest index; NaN never wins
    }
    if (besti < 0) besti = 0;
    if (out) out[r] = __fadd_rn(__fmul_rn(1.f - iq, xv), __fmul_rn(iq, bins_s[besti]));
    if (idx) idx[r] = (uint8_t)besti;
  }
}

int launch_quantize(const float* x, int64_t B, int L, const float* bins, int n, const float* alpha, float iq,
                    int use_soft, float* out, uint8_t* idx, float* soft, float* hist, float* qloss,
                    cudaStream_t st) {
  if (B == 0) return NSC_OK;   // empty batch
  NSC_CHECK_ARG(n >= 1 && n <= 256, "nsc_quantize_scalar: num bins %d not in [1,256]", n);
  NSC_CHECK_ARG(L >= 1 && L <= 4096, "nsc_quantize_scalar: code length %d not in [1,4096]", L);
  NSC_CHECK_ARG(x && bins && alpha, "nsc_quantize_scalar: null input");
  if (B == 0) return NSC_OK;
  if (!use_soft && soft == nullptr && hist == nullptr && qloss == nullptr) {
    char hname[32];
    snprintf(hname, sizeof(hname), "quantize_hard_n%d_L%d", n, L);
    ProfScope hprof(st, hname, (double)B * L * n * 4.0, (double)B * L * 9.0);
    const int64_t rows = B * (int64_t)L;
    int64_t g = ceil_div64(rows, 256);
    if (g > 148 * 8) g = 148 * 8;
    quantize_hard_kernel<<<(unsigned)g, 256, 0, st>>>(x, rows, bins, n, alpha, iq, out, idx);
    NSC_LAUNCH_OK();
    return NSC_OK;
  }
  const int frames_per_cta = L >= 256 ? 1 : 256 / L;
  const int64_t grid = ceil_div64(B, frames_per_cta);
  const int npl = n <= 32 ? 1 : n <= 64 ? 2 : n <= 128 ? 4 : 8;
  const size_t smem = sizeof(float) * (npl * 32 + kQWarps * npl * 32 + (size_t)frames_per_cta * L);
#define NSC_Q_LAUNCH(NPL)                                                                                   \
  quantize_kernel<NPL><<<(unsigned)grid, kQWarps * 32, smem, st>>>(x, B, L, frames_per_cta, bins, n, alpha, iq, \
                                                                   use_soft, out, idx, soft, hist, qloss)
  char name[32];
  snprintf(name, sizeof(name), "quantize_n%d_L%d", n, L);
  // algorithmic traffic: code in, value out, uint8 index (2,304 B / frame / codec at L = 256)
  ProfScope prof(st, name, (double)B * L * n * 4.0, (double)B * L * 9.0 + (soft ? (double)B * L * n * 4.0 : 0.0));
  switch (npl) {
    case 1: NSC_Q_LAUNCH(1); break;
    case 2: NSC_Q_LAUNCH(2); break;
    case 4: NSC_Q_LAUNCH(4); break;
    default: NSC_Q_LAUNCH(8); break;
  }
#undef NSC_Q_LAUNCH
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // namespace nsc

extern "C" {

int nsc_quantize_scalar(const float* x, int64_t B, int32_t L, const float* bins, int32_t n, const float* alpha,
                        float is_quan_on, int32_t use_soft, float* out, uint8_t* idx, float* soft, float* hist,
                        float* qloss, void* stream) {
  return nsc::launch_quantize(x, B, L, bins, n, alpha, is_quan_on, use_soft, out, idx, soft, hist, qloss,
                              (cudaStream_t)stream);
}

int nsc_dequantize_scalar(const uint8_t* idx, int64_t rows, const float* bins, int32_t n, float* out, void* stream) {
  if (rows == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(n >= 1 && n <= 256, "nsc_dequantize_scalar: num bins %d not in [1,256]", n);
  int64_t grid = nsc::ceil_div64(rows, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  nsc::dequantize_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(idx, rows, bins, n, out);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_quan_loss(const float* soft, int64_t B, int32_t L, int32_t n, float* qloss, void* stream) {
  if (B == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(soft && qloss && L >= 1 && n >= 1, "nsc_quan_loss: bad argument");
  nsc::quan_loss_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(soft, L, n, qloss);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_soft_histogram(const float* soft, int64_t rows, int32_t n, float* hist, void* stream) {
  if (rows == 0) return NSC_OK;   // empty batch: nothing to validate or launch
  NSC_CHECK_ARG(soft && hist && n >= 1 && n <= 256, "nsc_soft_histogram: bad argument (n=%d)", n);
  const int rows_per_cta = 1024;
  nsc::soft_hist_kernel<<<(unsigned)nsc::ceil_div64(rows, rows_per_cta), 256, 0, (cudaStream_t)stream>>>(
      soft, rows, n, rows_per_cta, hist);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int nsc_entropy_from_hist(const float* hist, int32_t n, float* entropy, void* stream) {
  NSC_CHECK_ARG(n >= 1, "nsc_entropy_from_hist: n=%d", n);
  nsc::entropy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(hist, n, entropy);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // extern "C"
