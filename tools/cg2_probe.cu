// cg2_probe -- does a 2-CTA tcgen05.mma (cta_group::2, M = 256) work with this repo's operand layout?  (round-2 enabler: a CTA pair
// shares every weight tile -- each CTA stages only HALF of the B operand -- which halves the L2 weight stream of the 100 -> 100
// convs and the resident-weight footprint that keeps the 100->20 / 20->20 fusion out of shared memory.)
//
// Layout under test: each CTA holds its own 128 rows of A (K-major SWIZZLE_128B, 64 fp16 per row) and N/2 rows of B
// (CTA r holds output columns [r N/2, (r+1) N/2)); the leader CTA issues ONE instruction per K step; D rows [128 r, 128 r + 128)
// land in CTA r's TMEM, all N columns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cg2_probe tools/cg2_probe.cu && ./cg2_probe
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
__host__ __device__ inline uint32_t sw128_off(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
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
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int kK = 64;

struct Args {
  const __half* a;   // [256][64]
  const __half* b;   // [N][64]
  float* d;          // [256][N]
  int N;
  long long* cyc;
  int reps;
};

__global__ void __launch_bounds__(128) cg2_kernel(Args p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;                 // 128 rows of A
  uint8_t* sb = smem + 128 * 128;     // N/2 rows of B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int nh = p.N / 2;
  for (int i = tid; i < 128 * kK; i += 128) *(__half*)(sa + sw128_off(i >> 6, i & 63)) = p.a[(rank * 128) * kK + i];
  for (int i = tid; i < nh * kK; i += 128) *(__half*)(sb + sw128_off(i >> 6, i & 63)) = p.b[(rank * nh) * kK + i];
  if (tid == 0) mbar_init(&bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                 // both CTAs' operands and barriers are in place
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (rank == 0 && tid == 0) {
    const uint32_t idesc = make_idesc_f16(256, p.N);
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep)
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ad = make_desc_sw128(smem_u32(sa) + kk * 32), bd = make_desc_sw128(smem_u32(sb) + kk * 32);
        const uint32_t acc = (rep | kk) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    // completion arrives on the barrier at the same offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    mbar_wait(&bar, 0);
    if (p.cyc) *p.cyc = clock64() - t0;
  }
  mbar_wait(&bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = rank * 128 + warp * 32 + (tid & 31);
    for (int j = 0; j < 16; ++j) p.d[row * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();                 // nobody frees TMEM while the peer's instruction stream can still touch it
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
}

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

int main() {
  srand(3);
  int fails = 0;
  for (int N : {64, 128, 224, 256}) {
    std::vector<__half> a(256 * kK), b(N * kK);
    for (auto& v : a) v = __float2half(frand());
    for (auto& v : b) v = __float2half(frand() * 0.1f);
    __half *d_a, *d_b; float* d_d; long long* d_cyc;
    CK(cudaMalloc(&d_a, a.size() * 2)); CK(cudaMalloc(&d_b, b.size() * 2)); CK(cudaMalloc(&d_d, 256 * N * 4)); CK(cudaMalloc(&d_cyc, 8));
    CK(cudaMemcpy(d_a, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_d, 0, 256 * N * 4));
    const size_t smem = 128 * 128 + 128 * 128 + 1024;
    CK(cudaFuncSetAttribute(cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int reps : {1, 256}) {
      Args p{d_a, d_b, d_d, N, d_cyc, reps};
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, cg2_kernel, p));
      CK(cudaDeviceSynchronize());
      if (reps == 1) {
        std::vector<float> D(256 * N);
        CK(cudaMemcpy(D.data(), d_d, D.size() * 4, cudaMemcpyDeviceToHost));
        double max_err = 0, max_ref = 0;
        for (int m = 0; m < 256; ++m)
          for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < kK; ++k) ref += (double)__half2float(a[m * kK + k]) * (double)__half2float(b[n * kK + k]);
            max_err = fmax(max_err, fabs(ref - (double)D[m * N + n]));
            max_ref = fmax(max_ref, fabs(ref));
          }
        const bool ok = max_err <= 1e-5 * max_ref + 1e-6;
        if (!ok) ++fails;
        printf("cta_group::2 M=256 N=%3d: max_err=%.3e (max_ref %.3f) %s\n", N, max_err, max_ref, ok ? "ok" : "FAIL");
      } else {
        long long cyc = 0;
        CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("   %d instructions (M256 x N%d x K16) in %lld cycles -> %.1f cycles each, %.0f MAC/cycle/SM\n", reps * 4, N, cyc,
               (double)cyc / (reps * 4), 256.0 * N * 16 * reps * 4 / (double)cyc / 2);
      }
    }
    cudaFree(d_a); cudaFree(d_b); cudaFree(d_d); cudaFree(d_cyc);
  }
  printf(fails ? "CG2 PROBE FAILED (%d)\n" : "CG2 PROBE PASSED\n", fails);
  return fails ? 1 : 0;
}
