// tc_probe -- bring-up probe for the tcgen05 path of the conv engine (sm_100a).  Not part of the library.
//
// Validates, against a CPU reference, the three hardware facts the implicit-GEMM conv kernel relies on:
//   1. K-major SWIZZLE_128B shared-memory descriptors whose start address is shifted by an arbitrary number of
//      128-byte rows (conv taps = row shifts of ONE staged activation tile; swizzle is a function of the absolute
//      shared-memory address, so a shifted start must read consistently);
//   2. N = 32 and N = 112 instruction shapes at M = 128, fp16 inputs / fp32 accumulation in TMEM;
//   3. the 3-MMA hi/lo split (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi) reproducing an fp32 product to ~1e-6.
// It also times a long chain of MMAs to measure the issue cost per instruction shape.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tools/tc_probe.cu && ./tc_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: fp16 A/B (K-major both), fp32 accumulate, M x N
__host__ __device__ inline uint32_t make_idesc_f16(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                  // c_format = F32
  d |= 0u << 7;                  // a_format = F16
  d |= 0u << 10;                 // b_format = F16
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
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
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

// byte offset of element (row, k) (k in halfs, < 64) inside a SW128 K-major slab whose base is 1024-aligned
__host__ __device__ inline uint32_t sw128_off(int row, int k) {
  const uint32_t chunk = (uint32_t)(k >> 3), within = (uint32_t)(k & 7);
  return (uint32_t)row * 128u + ((chunk ^ ((uint32_t)row & 7u)) << 4) + within * 2u;
}

constexpr int kRowsA = 160;   // staged activation rows (128 + halo)
constexpr int kMaxN = 128;

struct ProbeArgs {
  const __half* a_hi;   // [kRowsA][64]
  const __half* a_lo;
  const __half* b_hi;   // [N][64]
  const __half* b_lo;
  float* d;             // [128][N]
  int N, shift, split, reps;
  long long* cycles;
};

__global__ void __launch_bounds__(128) probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // manual 1024-byte alignment of the dynamic segment
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa_hi = smem;                         // kRowsA * 128 B
  uint8_t* sa_lo = sa_hi + kRowsA * 128;
  uint8_t* sb_hi = sa_lo + kRowsA * 128;         // kMaxN * 128 B
  uint8_t* sb_lo = sb_hi + kMaxN * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < kRowsA * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    *(__half*)(sa_hi + sw128_off(r, k)) = p.a_hi[i];
    *(__half*)(sa_lo + sw128_off(r, k)) = p.a_lo[i];
  }
  for (int i = tid; i < kMaxN * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    const bool v = r < p.N;
    *(__half*)(sb_hi + sw128_off(r, k)) = v ? p.b_hi[i] : __float2half(0.f);
    *(__half*)(sb_lo + sw128_off(r, k)) = v ? p.b_lo[i] : __float2half(0.f);
  }
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, p.N);
    const uint32_t a_hi = smem_u32(sa_hi) + p.shift * 128, a_lo = smem_u32(sa_lo) + p.shift * 128;
    const uint32_t b_hi = smem_u32(sb_hi), b_lo = smem_u32(sb_lo);
    t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) {
      uint32_t acc = rep > 0 ? 1u : 0u;
      for (int kk = 0; kk < 4; ++kk) {         // 4 x K16 = 64
        const uint32_t ko = kk * 32;           // 16 halfs = 32 bytes inside the 128-byte swizzle row
        mma_f16_ss(tmem, make_desc_sw128(a_hi + ko), make_desc_sw128(b_hi + ko), idesc, acc);
        acc = 1u;
        if (p.split) {
          mma_f16_ss(tmem, make_desc_sw128(a_hi + ko), make_desc_sw128(b_lo + ko), idesc, 1u);
          mma_f16_ss(tmem, make_desc_sw128(a_lo + ko), make_desc_sw128(b_hi + ko), idesc, 1u);
        }
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) { t1 = clock64(); if (p.cycles) *p.cycles = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes [32w, 32w+32); thread = one row, 32 columns per load
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = warp * 32 + (tid & 31);
    for (int j = 0; j < 32; ++j)
      if (c0 + j < p.N) p.d[row * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128) : "memory");
}

// Issue-rate probe: one thread issues `reps` x 8 MMAs with PRECOMPUTED descriptors, rotating over `nacc`
// accumulators (column offsets) -- separates the tensor-pipe floor from descriptor arithmetic / dependency stalls.
template <int NACC>
__global__ void __launch_bounds__(128) issue_kernel(int N, int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kRowsA + 256) * 32; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;  // fp16 1.0 pairs
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    uint64_t ad[4], bd[4];
    for (int k = 0; k < 4; ++k) {
      ad[k] = make_desc_sw128(smem_u32(smem) + k * 32);
      bd[k] = make_desc_sw128(smem_u32(smem) + kRowsA * 128 + k * 32);
    }
    const int ncols = (N + 31) & ~31;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t d = tmem + (uint32_t)((u % NACC) * ncols);
        mma_f16_ss(d, ad[u & 3], bd[u & 3], idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    *cycles = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main() {
  srand(1);
  std::vector<float> A(kRowsA * 64), Bm(kMaxN * 64);
  for (auto& v : A) v = frand();
  for (auto& v : Bm) v = frand() * 0.1f;
  std::vector<__half> a_hi(A.size()), a_lo(A.size()), b_hi(Bm.size()), b_lo(Bm.size());
  for (size_t i = 0; i < A.size(); ++i) { a_hi[i] = __float2half(A[i]); a_lo[i] = __float2half(A[i] - __half2float(a_hi[i])); }
  for (size_t i = 0; i < Bm.size(); ++i) { b_hi[i] = __float2half(Bm[i]); b_lo[i] = __float2half(Bm[i] - __half2float(b_hi[i])); }
  __half *d_ah, *d_al, *d_bh, *d_bl; float* d_d; long long* d_cyc;
  CK(cudaMalloc(&d_ah, a_hi.size() * 2)); CK(cudaMalloc(&d_al, a_lo.size() * 2));
  CK(cudaMalloc(&d_bh, b_hi.size() * 2)); CK(cudaMalloc(&d_bl, b_lo.size() * 2));
  CK(cudaMalloc(&d_d, 128 * kMaxN * 4)); CK(cudaMalloc(&d_cyc, 8));
  CK(cudaMemcpy(d_ah, a_hi.data(), a_hi.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_al, a_lo.data(), a_lo.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bh, b_hi.data(), b_hi.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_bl, b_lo.data(), b_lo.size() * 2, cudaMemcpyHostToDevice));
  const size_t smem = 2 * kRowsA * 128 + 2 * kMaxN * 128 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  std::vector<float> D(128 * kMaxN);
  int fails = 0;
  for (int N : {32, 112, 16, 64}) {
    for (int shift : {0, 1, 3, 8, 11, 27}) {
      for (int split : {0, 1}) {
        ProbeArgs p{d_ah, d_al, d_bh, d_bl, d_d, N, shift, split, 1, d_cyc};
        CK(cudaMemset(d_d, 0, 128 * kMaxN * 4));
        probe_kernel<<<1, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), d_d, 128 * N * 4, cudaMemcpyDeviceToHost));
        double max_err = 0, max_ref = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) {
              const double a = split ? (double)A[(m + shift) * 64 + k] : (double)__half2float(a_hi[(m + shift) * 64 + k]);
              const double b = split ? (double)Bm[n * 64 + k] : (double)__half2float(b_hi[n * 64 + k]);
              ref += a * b;
            }
            max_err = fmax(max_err, fabs(ref - (double)D[m * N + n]));
            max_ref = fmax(max_ref, fabs(ref));
          }
        const double tol = split ? 3e-6 : 1e-5;
        const bool ok = max_err <= tol * max_ref + 1e-7;
        if (!ok) ++fails;
        printf("N=%3d shift=%2d split=%d  max_err=%.3e (max_ref %.3f)  %s\n", N, shift, split, max_err, max_ref, ok ? "ok" : "FAIL");
      }
    }
  }
  // issue-cost measurement: reps * 4 (x3 if split) MMAs back to back from one thread, one commit at the end
  for (int N : {16, 32, 64, 112, 128}) {
    for (int split : {0, 1}) {
      const int reps = 256;
      ProbeArgs p{d_ah, d_al, d_bh, d_bl, d_d, N, 0, split, reps, d_cyc};
      probe_kernel<<<1, 128, smem>>>(p);
      CK(cudaDeviceSynchronize());
      long long cyc = 0;
      CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
      const int n_mma = reps * 4 * (split ? 3 : 1);
      printf("timing N=%3d split=%d: %d MMAs (M128 K16) in %lld cycles -> %.1f cycles/MMA, %.0f MAC/cycle/SM\n", N, split,
             n_mma, cyc, (double)cyc / n_mma, 128.0 * N * 16 * n_mma / (double)cyc);
    }
  }
  {
    const size_t smem2 = (kRowsA + 256) * 128 + 1024;
    CK(cudaFuncSetAttribute(issue_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    CK(cudaFuncSetAttribute(issue_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    CK(cudaFuncSetAttribute(issue_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    for (int N : {16, 32, 48, 64, 112, 128, 256}) {
      for (int nacc : {1, 2, 4}) {
        if (((N + 31) & ~31) * nacc > 512) continue;
        const int reps = 512;
        if (nacc == 1) issue_kernel<1><<<1, 128, smem2>>>(N, reps, d_cyc);
        else if (nacc == 2) issue_kernel<2><<<1, 128, smem2>>>(N, reps, d_cyc);
        else issue_kernel<4><<<1, 128, smem2>>>(N, reps, d_cyc);
        CK(cudaDeviceSynchronize());
        long long cyc = 0;
        CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("issue N=%3d nacc=%d: %.1f cycles/MMA (floor %d), %.0f MAC/cycle/SM\n", N, nacc, (double)cyc / (reps * 8), N / 2,
               128.0 * N * 16 * reps * 8 / (double)cyc);
      }
    }
  }
  printf(fails ? "PROBE FAILED (%d cases)\n" : "PROBE PASSED\n", fails);
  return fails ? 1 : 0;
}
