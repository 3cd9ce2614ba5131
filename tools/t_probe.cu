// t_probe -- bring-up probe for the "taps in N" formulation of a narrow conv layer on tcgen05 (sm_100a).
//
//   P[row, (tap, co)] = sum_ci X[row, ci] * W[tap, ci, co]       one MMA chain, N = 9 * 20 = 180 (padded to 192)
//   y[row, co]        = sum_tap P[row + (tap - 4) * dil, tap, co]   shifted sum across TMEM lanes, done with warp shuffles
//
// Checks the result against a CPU reference and times (a) the MMA chain at N = 192 and (b) the shuffle epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o t_probe tools/t_probe.cu && ./t_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__host__ __device__ inline uint32_t sw128_off(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
}

constexpr int kN = 192, kTaps = 9, kCo = 20, kK = 64;

__device__ __forceinline__ void tmem_ld20(uint32_t taddr, float (&v)[20]) {
  uint32_t r[20];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]) : "r"(taddr + 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 20; ++i) v[i] = __uint_as_float(r[i]);
}

template <int DIL>
__device__ __forceinline__ void shifted_sum(uint32_t tmem_row, int lane, float (&acc)[20], float (&up)[20], float (&down)[20]) {
#pragma unroll
  for (int c = 0; c < 20; ++c) { acc[c] = 0.f; up[c] = 0.f; down[c] = 0.f; }
#pragma unroll
  for (int t = 0; t < kTaps; ++t) {
    constexpr int dummy = 0; (void)dummy;
    const int s = (t - 4) * DIL;
    float v[20];
    tmem_ld20(tmem_row + (uint32_t)(t * kCo), v);
    const int src = (lane + s) & 31;
    const bool inr = (unsigned)(lane + s) < 32u;
#pragma unroll
    for (int c = 0; c < 20; ++c) {
      const float x = s == 0 ? v[c] : __shfl_sync(0xffffffffu, v[c], src);
      if (s == 0) acc[c] += x;
      else if (s > 0) { if (inr) acc[c] += x; else down[c] += x; }
      else { if (inr) acc[c] += x; else up[c] += x; }
    }
  }
}

struct Args {
  const __half* a;   // [128][64]
  const __half* b;   // [192][64]
  float* acc; float* up; float* down;   // [128][20]
  int dil, reps_mma, reps_epi, epi_warps;
  long long* cyc;    // [0] mma chain, [1] epilogue
};

__global__ void __launch_bounds__(128) t_probe_kernel(Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 128 rows
  uint8_t* sb = smem + 128 * 128;     // 192 rows
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * kK; i += 128) *(__half*)(sa + sw128_off(i >> 6, i & 63)) = p.a[i];
  for (int i = tid; i < kN * kK; i += 128) *(__half*)(sb + sw128_off(i >> 6, i & 63)) = p.b[i];
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, kN);
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps_mma; ++rep)
      for (int kk = 0; kk < 4; ++kk)
        mma_f16_ss(tmem, make_desc_sw128(smem_u32(sa) + kk * 32), make_desc_sw128(smem_u32(sb) + kk * 32), idesc, (rep | kk) ? 1u : 0u);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    p.cyc[0] = clock64() - t0;
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  float acc[20], up[20], down[20];
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
  const long long t0 = clock64();
  if (warp < p.epi_warps) {
    for (int rep = 0; rep < p.reps_epi; ++rep) {
      if (p.dil == 1) shifted_sum<1>(trow, lane, acc, up, down);
      else shifted_sum<2>(trow, lane, acc, up, down);
      // keep the work alive across repetitions
      if (rep + 1 < p.reps_epi && acc[0] == 1234.5f) p.acc[tid] = acc[1] + up[2] + down[3];
    }
  }
  __syncthreads();
  if (tid == 0) p.cyc[1] = clock64() - t0;
  if (warp < p.epi_warps)
    for (int c = 0; c < 20; ++c) { p.acc[tid * 20 + c] = acc[c]; p.up[tid * 20 + c] = up[c]; p.down[tid * 20 + c] = down[c]; }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main() {
  srand(2);
  std::vector<float> X(128 * kK), W(kTaps * kK * kCo);
  for (auto& v : X) v = frand();
  for (auto& v : W) v = frand() * 0.1f;
  std::vector<__half> a(128 * kK), b(kN * kK);
  for (size_t i = 0; i < a.size(); ++i) a[i] = __float2half(X[i]);
  for (int n = 0; n < kN; ++n)
    for (int k = 0; k < kK; ++k) {
      const int t = n / kCo, co = n % kCo;
      b[n * kK + k] = __float2half(n < kTaps * kCo ? W[(t * kK + k) * kCo + co] : 0.f);
    }
  __half *d_a, *d_b; float *d_acc, *d_up, *d_down; long long* d_cyc;
  CK(cudaMalloc(&d_a, a.size() * 2)); CK(cudaMalloc(&d_b, b.size() * 2));
  CK(cudaMalloc(&d_acc, 128 * 20 * 4)); CK(cudaMalloc(&d_up, 128 * 20 * 4)); CK(cudaMalloc(&d_down, 128 * 20 * 4)); CK(cudaMalloc(&d_cyc, 16));
  CK(cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = (128 + kN) * 128 + 1024;
  CK(cudaFuncSetAttribute(t_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int fails = 0;
  for (int dil : {1, 2}) {
    Args p{d_a, d_b, d_acc, d_up, d_down, dil, 1, 1, 4, d_cyc};
    t_probe_kernel<<<1, 128, smem>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> acc(128 * 20), up(128 * 20), down(128 * 20);
    CK(cudaMemcpy(acc.data(), d_acc, acc.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(up.data(), d_up, up.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(down.data(), d_down, down.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    for (int r = 0; r < 128; ++r)
      for (int co = 0; co < kCo; ++co) {
        double ref = 0;
        for (int t = 0; t < kTaps; ++t) {
          const int src = r + (t - 4) * dil;
          if (src < 0 || src >= 128) continue;
          for (int k = 0; k < kK; ++k) ref += (double)__half2float(a[src * kK + k]) * (double)__half2float(b[(t * kCo + co) * kK + k]);
        }
        double got = acc[r * 20 + co];
        if (r >= 32) got += up[(r - 32) * 20 + co];      // up-spill of the previous quarter's same lane
        if (r < 96) got += down[(r + 32) * 20 + co];     // down-spill of the next quarter's same lane
        max_err = fmax(max_err, fabs(ref - got));
        max_ref = fmax(max_ref, fabs(ref));
      }
    const bool ok = max_err <= 1e-5 * max_ref + 1e-6;
    if (!ok) ++fails;
    printf("taps-in-N dil=%d: max_err=%.3e (max_ref %.3f) %s\n", dil, max_err, max_ref, ok ? "ok" : "FAIL");
  }
  for (int dil : {1, 2})
    for (int ew : {1, 2, 4}) {
      Args p{d_a, d_b, d_acc, d_up, d_down, dil, 256, 64, ew, d_cyc};
      t_probe_kernel<<<1, 128, smem>>>(p);
      CK(cudaDeviceSynchronize());
      long long cyc[2];
      CK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
      printf("dil=%d epi_warps=%d: MMA N=192 %.1f cycles/MMA; shuffle epilogue %.0f cycles per 32-row quarter pass (%d warps concurrently)\n",
             dil, ew, (double)cyc[0] / (256 * 4), (double)cyc[1] / 64, ew);
    }
  printf(fails ? "T-PROBE FAILED\n" : "T-PROBE PASSED\n");
  return fails ? 1 : 0;
}
