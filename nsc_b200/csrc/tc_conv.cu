// K3/K4 (tensor engine) -- 1-D SAME convolution as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), same fused epilogue and same NCL fp32 tensors as the FFMA engine (conv.cu).
//
//   D[position, cout] = sum_taps A_tap[position, cin] * W_tap[cout, cin]^T        M = 128 positions per MMA
//
// * One CTA owns one whole frame: its Lout/128 accumulators (N_pad columns each) live in TMEM side by side, so a
//   layer's weights are streamed ONCE per frame (per pass) and every weight stage feeds all position tiles.
// * The activation frame is staged in shared memory in the K-major SWIZZLE_128B layout (row = position + left
//   padding, 64 channels = 128 bytes per row, slabs of 64 channels).  A conv tap is nothing but a ROW SHIFT of the
//   A descriptor's start address -- no im2col is ever materialised (verified by tools/tc_probe.cu: the swizzle is a
//   function of the absolute shared-memory address, so shifted starts read consistently).  Zero rows on both ends
//   give SAME padding; a stride-2 conv de-interleaves positions into STRIDE row buffers by (pos+padL) % STRIDE.
// * fp32 parity: activations and weights are split x = hi + lo into two fp16 planes and three MMAs
//   (hi*hi + hi*lo + lo*hi) accumulate in fp32 -- ~2^-21 relative error per product, i.e. fp32-class results at a
//   third of the fp16 tensor rate ("precision 1").  "precision 2" issues only hi*hi (plain fp16 inputs) and is
//   reported separately as reduced precision.
// * Measured on B200 (tools/tc_probe.cu): an M128 x K16 MMA costs max(N/2, 32 + N/4, 44) cycles -- the A tile
//   (4 KB) is re-read from shared memory by every instruction, so narrow layers (N_pad = 32) are bound by
//   shared-memory bandwidth, not by the tensor pipe.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "conv.cuh"

namespace nsc {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B descriptor: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// kind::f16 instruction descriptor: fp16 A and B (both K-major), fp32 accumulator, M = 128
__host__ __device__ inline uint32_t make_idesc_f16(int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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

// byte offset of (row, channel c < 64) inside one SWIZZLE_128B slab whose base is 1024-byte aligned
__host__ __device__ inline uint32_t sw128_off(int row, int c) {
  return (uint32_t)row * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)c & 7u) * 2u;
}

struct TcLaunch {
  ConvArgs a;
  int Lout, padL;
  int n_pad;        // Cout rounded up to 16 (MMA N)
  int ksteps;       // ceil(Cin / 16)
  int slabs;        // ceil(Cin / 64)
  int mtiles;       // Lout / 128
  int rows;         // rows per A buffer (multiple of 8)
  int nbuf;         // = stride (row buffers)
  int passes;       // 3 = hi*hi + hi*lo + lo*hi, 1 = hi*hi only
  int tmem_cols;    // power of two >= mtiles * n_pad
  const uint4* wpack;  // [plane][tap][slab][n_pad * 8] uint4
};

// Weights (K, Cin, Cout) fp32 -> two fp16 planes in the exact shared-memory image of each (tap, slab) stage.
__global__ void tc_pack_weights_kernel(const float* __restrict__ w, int K, int Cin, int Cout, int n_pad, int slabs,
                                       __half* __restrict__ out) {
  const int per_plane = K * slabs * n_pad * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per_plane; i += gridDim.x * blockDim.x) {
    const int plane = i / per_plane;
    int r = i - plane * per_plane;
    const int c = r & 63; r >>= 6;
    const int n = r % n_pad; r /= n_pad;
    const int slab = r % slabs;
    const int t = r / slabs;
    const int ci = slab * 64 + c;
    float v = 0.f;
    if (ci < Cin && n < Cout) v = w[((int64_t)t * Cin + ci) * Cout + n];
    const __half hi = __float2half_rn(v);
    const __half val = plane == 0 ? hi : __float2half_rn(v - __half2float(hi));
    const size_t stage = ((size_t)(plane * K + t) * slabs + slab) * n_pad * 128;
    *reinterpret_cast<__half*>(reinterpret_cast<char*>(out) + stage + sw128_off(n, c)) = val;
  }
}

constexpr int kTcThreads = 256;

__global__ void __launch_bounds__(kTcThreads, 1) tc_conv_kernel(const __grid_constant__ TcLaunch p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_stage[2];
  __shared__ uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;

  const ConvArgs& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b = blockIdx.x;
  const uint32_t slab_bytes = (uint32_t)p.rows * 128u;                 // one slab of one row buffer
  const uint32_t buf_bytes = slab_bytes * (uint32_t)p.slabs;
  const uint32_t a_bytes = buf_bytes * (uint32_t)p.nbuf;
  const uint32_t bslab_bytes = (uint32_t)p.n_pad * 128u;
  const uint32_t bstage_bytes = bslab_bytes * (uint32_t)p.slabs;
  uint8_t* sA = smem;
  uint8_t* sB = smem + a_bytes;                                        // 2 stages

  // ---- one-time setup: zero the A region (halo rows / channel padding stay zero), barriers, TMEM
  for (uint32_t i = tid; i < a_bytes / 16; i += kTcThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(&bar_stage[0], 1);
    mbar_init(&bar_stage[1], 1);
    mbar_init(&bar_done, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = make_idesc_f16(p.n_pad);

  const float* xb = a.x + b * (int64_t)a.Cin * a.Lin;   // NCL input frame
  const int groups = (a.Cin + 7) >> 3;                   // 8-channel (16-byte) groups
  uint32_t stage_phase[2] = {0, 0};                      // parity each stage barrier will complete next
  int stage_used[2] = {0, 0};                            // outstanding commits per stage
  uint32_t n_issued = 0;                                 // MMAs issued so far (first one overwrites the accumulator)
  int it = 0;

  for (int pass = 0; pass < p.passes; ++pass) {
    const int a_plane = pass == 2 ? 1 : 0;               // hi, hi, lo
    const int b_plane = pass == 1 ? 1 : 0;               // hi, lo, hi
    if (pass == 0 || pass == 2) {
      // all MMAs that read the previous A plane must have completed before it is overwritten
      if (pass == 2) {
        for (int s = 0; s < 2; ++s)
          if (stage_used[s]) { mbar_wait(&bar_stage[s], stage_phase[s]); stage_phase[s] ^= 1; stage_used[s] = 0; }
      }
      // ---- stage the frame: NCL fp32 -> fp16 plane, K-major SW128 rows (row = (pos + padL) / stride)
      for (int i = tid; i < groups * a.Lin; i += kTcThreads) {
        const int g = i / a.Lin, pos = i - g * a.Lin;
        const int c0 = g << 3;
        __half h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = (c0 + j < a.Cin) ? xb[(int64_t)(c0 + j) * a.Lin + pos] : 0.f;
          const __half hi = __float2half_rn(v);
          h[j] = a_plane == 0 ? hi : __float2half_rn(v - __half2float(hi));
        }
        const int u = pos + p.padL;
        const int buf = u % p.nbuf, row = u / p.nbuf;
        const int slab = g >> 3, chunk = g & 7;
        uint8_t* dst = sA + (uint32_t)buf * buf_bytes + (uint32_t)slab * slab_bytes + (uint32_t)row * 128u +
                       (uint32_t)((chunk ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
      }
    }
    for (int t = 0; t < a.K; ++t, ++it) {
      const int s = it & 1;
      // stage s is free once the MMAs committed two iterations ago have completed
      if (stage_used[s]) { mbar_wait(&bar_stage[s], stage_phase[s]); stage_phase[s] ^= 1; stage_used[s] = 0; }
      const uint4* src = p.wpack + ((size_t)(b_plane * a.K + t) * p.slabs) * (size_t)(p.n_pad * 8);
      uint4* dstB = reinterpret_cast<uint4*>(sB + (uint32_t)s * bstage_bytes);
      for (uint32_t i = tid; i < bstage_bytes / 16; i += kTcThreads) dstB[i] = src[i];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int shift = t * a.dil;                                   // in input positions
        const int buf = shift % p.nbuf, rowoff = shift / p.nbuf;
        const uint32_t a_base = smem_u32(sA) + (uint32_t)buf * buf_bytes;
        const uint32_t b_base = smem_u32(sB) + (uint32_t)s * bstage_bytes;
        for (int mt = 0; mt < p.mtiles; ++mt) {
          const uint32_t d = tmem + (uint32_t)(mt * p.n_pad);
          const uint32_t a_row = a_base + (uint32_t)(mt * 128 + rowoff) * 128u;
          for (int ks = 0; ks < p.ksteps; ++ks) {
            const uint32_t slab = (uint32_t)ks >> 2, ko = ((uint32_t)ks & 3u) * 32u;
            mma_f16_ss(d, make_desc_sw128(a_row + slab * slab_bytes + ko), make_desc_sw128(b_base + slab * bslab_bytes + ko),
                       idesc, (pass | t | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&bar_stage[s]);
      }
      stage_used[s] = 1;
      (void)n_issued;
    }
  }
  if (tid == 0) umma_commit(&bar_done);
  mbar_wait(&bar_done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: TMEM -> registers -> bias / activation / residual -> NCL fp32 (thread = one position)
  const int lane_base = (warp & 3) * 32;
  const int mt_lo = (warp < 4) ? 0 : (p.mtiles + 1) / 2;
  const int mt_hi = (warp < 4) ? (p.mtiles + 1) / 2 : p.mtiles;
  const int Cres = a.res_mode == RES_ADD_BCAST ? 1 : a.Cout;
  const int r = a.shuffle;
  const int Lout_y = p.Lout * r, Cout_y = a.Cout / r;
  float* yb = a.y + b * (int64_t)Lout_y * Cout_y;
  const float* rb = a.res ? a.res + b * (int64_t)p.Lout * Cres : nullptr;
  for (int mt = mt_lo; mt < mt_hi; ++mt) {
    const int pos = mt * 128 + lane_base + lane;
    for (int c0 = 0; c0 < a.Cout; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + ((uint32_t)lane_base << 16) + (uint32_t)(mt * p.n_pad + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int co = c0 + j;
        if (co >= a.Cout) break;
        float o = apply_act(__uint_as_float(v[j]) + (a.bias ? a.bias[co] : 0.f), a.act);
        if (a.res_mode != RES_NONE) {
          const float rv = rb[(int64_t)(a.res_mode == RES_ADD_BCAST ? 0 : co) * p.Lout + pos];
          o = a.res_mode == RES_MUL ? o * rv : o + rv;
        }
        o = apply_act(o, a.post_act);
        if (r == 1) yb[(int64_t)co * Lout_y + pos] = o;
        else yb[(int64_t)(co / r) * Lout_y + (int64_t)pos * r + (co % r)] = o;   // sub-pixel (nscm.py:158-167)
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(p.tmem_cols) : "memory");
}

// ================================================================================================
// Persistent, warp-specialised pipeline (the default).  Work unit = TILE output positions of one frame
// (TILE = 256 = two M-tiles, or 128 when Lout = 128); CTAs loop over tiles round-robin.
//
//   warps 0-7   epilogue      wait acc_full[k]  -> TMEM -> bias/act/residual -> NCL fp32 stores -> arrive acc_empty[k]
//                             (two warps per TMEM lane quarter; 32 residual loads in flight per thread)
//   warps 8-9   MMA issuers   one elected thread each, one per M-tile of the unit: wait a_full / b_full, issue tcgen05.mma,
//                             tcgen05.commit -> b_empty, a_empty, acc_full (barrier counts = number of issuers)
//   warp  10    B producer    one elected thread: cp.async.bulk of pre-packed weight stages (tap, slab) into a ring
//   warps 11-18 A producers   NCL fp32 -> fp16 plane -> K-major SW128 rows of A buffer (job & 1); arrive a_full
//
// Two A buffers (hi / lo plane of a tile in split mode, consecutive tiles otherwise), a ring of weight stages and two
// TMEM accumulator sets decouple the four roles: staging of tile i+1 and the epilogue of tile i-1 overlap the MMAs
// of tile i.  All hand-offs are mbarriers; tensor-core completions arrive through tcgen05.commit.
// ================================================================================================
constexpr int kEpiWarps = 8, kAProdWarps = 8;
constexpr int kMmaWarps = 2;       // one issuing thread per M-tile of the work unit (issue overhead, not the tensor pipe, bounds narrow layers)
constexpr int kPipeThreads = (kEpiWarps + kMmaWarps + 1 + kAProdWarps) * 32;   // 608
constexpr int kATasks = 2;        // A-producer tasks (8 x 128-bit loads each) in flight per thread
constexpr int kMaxBStages = 32;   // weight ring depth is chosen per layer from the shared memory left over (TcPipe::bstages)

struct TcPipe {
  ConvArgs a;
  int Lout, padL;
  int n_pad, ksteps, slabs;
  int tile;          // output positions per work unit (128 or 256)
  int mt;            // tile / 128
  int tiles_per_frame;
  int64_t n_tiles;
  int rows;          // rows per (sub-buffer, slab), multiple of 8
  int nsub;          // = stride
  int passes;
  int tmem_cols;
  int bstages;       // weight ring depth (2..kMaxBStages)
  int group;         // (tap, slab) weight units per ring stage
  int stages_per_pass;
  const uint4* wpack;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(kPipeThreads, 1) tc_conv_pipe_kernel(const __grid_constant__ TcPipe p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full[2], a_empty[2], b_full[kMaxBStages], b_empty[kMaxBStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const ConvArgs& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t slab_bytes = (uint32_t)p.rows * 128u;
  const uint32_t sub_bytes = slab_bytes * (uint32_t)p.slabs;
  const uint32_t abuf_bytes = sub_bytes * (uint32_t)p.nsub;
  const uint32_t unit_bytes = (uint32_t)p.n_pad * 128u;            // one (tap, slab) weight unit
  const uint32_t bstage_bytes = unit_bytes * (uint32_t)p.group;    // one ring stage = `group` consecutive units
  uint8_t* sA = smem;                           // 2 A buffers
  uint8_t* sB = smem + 2u * abuf_bytes;         // p.bstages weight stages

  for (uint32_t i = tid; i < 2u * abuf_bytes / 16; i += kPipeThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
  const uint32_t n_issuers = (uint32_t)(p.mt < kMmaWarps ? p.mt : kMmaWarps);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kAProdWarps / 2); // one arrival per warp of the producer group that owns buffer i
      mbar_init(&a_empty[i], n_issuers);      // tcgen05.commit of every issuing thread
      mbar_init(&acc_full[i], n_issuers);     // tcgen05.commit of every issuing thread
      mbar_init(&acc_empty[i], kEpiWarps);    // one arrival per epilogue warp
    }
    for (int i = 0; i < p.bstages; ++i) {
      mbar_init(&b_full[i], 1);               // expect_tx arrival + bulk-copy bytes
      mbar_init(&b_empty[i], n_issuers);      // tcgen05.commit of every issuing thread
    }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == kEpiWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int acc_cols = p.mt * p.n_pad;
  const int jobs_per_tile = p.passes == 3 ? 2 : 1;          // A planes needed per tile

  if (warp < kEpiWarps) {
    // =========================== epilogue ===========================
    const int Cres = a.res_mode == RES_ADD_BCAST ? 1 : a.Cout;
    const int r = a.shuffle;
    const int Lout_y = p.Lout * r, Cout_y = a.Cout / r;
    const int quarter = warp & 3, half = warp >> 2;        // TMEM lanes [32*quarter, +32); two warps share a quarter
    const int nbatch = (a.Cout + 15) >> 4;                 // 16-column units per M-tile
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
      const int64_t b = tile / p.tiles_per_frame;
      const int q0 = (int)(tile - b * p.tiles_per_frame) * p.tile;
      const uint32_t acc = it & 1u;
      mbar_wait(&acc_full[acc], (it >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* __restrict__ yb = a.y + b * (int64_t)Lout_y * Cout_y;
      const float* __restrict__ rb = a.res ? a.res + b * (int64_t)p.Lout * Cres : nullptr;
      const int n_units = p.mt * nbatch;
      auto load_res = [&](int u, float (&rv)[16]) {
        const int mt = u / nbatch, c0 = (u - mt * nbatch) << 4;
        const int pos = q0 + mt * 128 + quarter * 32 + lane;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = c0 + j;
          rv[j] = (a.res_mode != RES_NONE && u < n_units && co < a.Cout)
                      ? __ldg(rb + (int64_t)(a.res_mode == RES_ADD_BCAST ? 0 : co) * p.Lout + pos) : 0.f;
        }
      };
      float rv[16], rvn[16];
      load_res(half, rv);
      for (int u = half; u < n_units; u += 2) {
        const int mt = u / nbatch, c0 = (u - mt * nbatch) << 4;
        const int pos = q0 + mt * 128 + quarter * 32 + lane;
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * acc_cols + mt * p.n_pad + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        load_res(u + 2, rvn);        // residual of this warp's NEXT unit is in flight while this one is written
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int co = c0 + j;
          if (co < a.Cout) {
            float o = apply_act(__uint_as_float(v[j]) + (a.bias ? __ldg(a.bias + co) : 0.f), a.act);
            if (a.res_mode == RES_MUL) o *= rv[j];
            else if (a.res_mode != RES_NONE) o += rv[j];
            o = apply_act(o, a.post_act);
            if (r == 1) yb[(int64_t)co * Lout_y + pos] = o;
            else yb[(int64_t)(co / r) * Lout_y + (int64_t)pos * r + (co % r)] = o;
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) rv[j] = rvn[j];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  } else if (warp < kEpiWarps + kMmaWarps) {
    // =========================== MMA issuer ===========================
    // One thread feeds the tensor core.  Descriptors are NOT rebuilt per instruction: the upper 32 bits (stride,
    // version, swizzle mode) are constant and the lower word is (smem address >> 4) | LBO, so advancing along K or
    // to the next position tile is an integer add (the first build spent ~140 cycles of address arithmetic per MMA,
    // three times the 44-cycle cost of the instruction itself).
    const int issuer = warp - kEpiWarps;                // issuer i owns M-tiles i, i + n_issuers, ...
    if (issuer < (int)n_issuers && elect_one()) {
      const uint32_t idesc = make_idesc_f16(p.n_pad);
      const uint32_t desc_hi = (uint32_t)(make_desc_sw128(0) >> 32);
      auto mk = [desc_hi](uint32_t lo) { return ((uint64_t)desc_hi << 32) | (uint64_t)lo; };
      auto lo_of = [](uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | 0x10000u; };
      const uint32_t mt_step = (128u * 128u) >> 4;           // next 128-position tile of the same buffer
      uint32_t it = 0, job = 0, bit = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        mbar_wait(&acc_empty[acc], ((it >> 1) & 1u) ^ 1u);
        const uint32_t d0 = tmem + acc * (uint32_t)acc_cols;
        const uint32_t job_hi = job, job_lo = job + 1;     // job_lo only meaningful in split mode
        for (int pass = 0; pass < p.passes; ++pass) {
          const uint32_t jb = pass == 2 ? job_lo : job_hi;
          if (pass == 0 || pass == 2) mbar_wait(&a_full[jb & 1u], (jb >> 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_buf = smem_u32(sA) + (jb & 1u) * abuf_bytes;
          int t = 0, sl = 0;
          for (int st = 0; st < p.stages_per_pass; ++st, ++bit) {
            const uint32_t s = bit % (uint32_t)p.bstages;
            mbar_wait(&b_full[s], (bit / (uint32_t)p.bstages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t b_lo = lo_of(smem_u32(sB) + s * bstage_bytes);
            for (int g = 0; g < p.group; ++g) {               // (tap, slab) units of this stage
              const int shift = t * a.dil;
              const uint32_t sub = (uint32_t)(shift % p.nsub), rowoff = (uint32_t)(shift / p.nsub);
              const uint32_t a_lo0 = lo_of(a_buf + sub * sub_bytes + (uint32_t)sl * slab_bytes + rowoff * 128u);
              const int ks_n = min(4, p.ksteps - sl * 4);
              const uint32_t first = (uint32_t)(pass | t | sl);
              for (int mt = issuer; mt < p.mt; mt += (int)n_issuers) {
                const uint32_t d = d0 + (uint32_t)(mt * p.n_pad);
                const uint32_t a_lo = a_lo0 + (uint32_t)mt * mt_step;
#pragma unroll 4
                for (int ks = 0; ks < ks_n; ++ks)
                  mma_f16_ss(d, mk(a_lo + 2u * (uint32_t)ks), mk(b_lo + 2u * (uint32_t)ks), idesc, (first | (uint32_t)ks) != 0 ? 1u : 0u);
              }
              b_lo += unit_bytes >> 4;
              if (++sl == p.slabs) { sl = 0; ++t; }
            }
            umma_commit(&b_empty[s]);            // weight stage reusable once these MMAs retire
          }
          // A buffers: hi is last read in pass 1 (split) or pass 0 (plain); lo in pass 2
          if (p.passes == 1) umma_commit(&a_empty[job_hi & 1u]);
          else if (pass == 1) umma_commit(&a_empty[job_hi & 1u]);
          else if (pass == 2) umma_commit(&a_empty[job_lo & 1u]);
        }
        umma_commit(&acc_full[acc]);
        job += (uint32_t)jobs_per_tile;
      }
    }
  } else if (warp == kEpiWarps + kMmaWarps) {
    // =========================== B producer (bulk copies) ===========================
    if (elect_one()) {
      uint32_t bit = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int pass = 0; pass < p.passes; ++pass) {
          const int b_plane = pass == 1 ? 1 : 0;
          const uint8_t* src0 = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)b_plane * p.stages_per_pass * (size_t)bstage_bytes;
          for (int st = 0; st < p.stages_per_pass; ++st, ++bit) {
            const uint32_t s = bit % (uint32_t)p.bstages;
            mbar_wait(&b_empty[s], ((bit / (uint32_t)p.bstages) & 1u) ^ 1u);
            mbar_expect_tx(&b_full[s], bstage_bytes);
            bulk_g2s(sB + s * bstage_bytes, src0 + (size_t)st * bstage_bytes, bstage_bytes, &b_full[s]);
          }
        }
      }
    }
  } else {
    // =========================== A producers ===========================
    // Two producer groups of kAProdWarps/2 warps: group g fills buffer g (jobs with job & 1 == g), so two staging
    // jobs -- the hi and lo plane of a tile in split mode, consecutive tiles otherwise -- are in flight at once and a
    // job's global-load latency no longer bounds the tile rate.
    const int pgroup = (warp - (kEpiWarps + kMmaWarps + 1)) / (kAProdWarps / 2);
    const int nprod = (kAProdWarps / 2) * 32;
    const int ptid = tid - (kEpiWarps + kMmaWarps + 1) * 32 - pgroup * nprod;
    const int groups = (a.Cin + 7) >> 3;
    const int span = p.rows * p.nsub;               // u-coordinates covered by one buffer
    uint32_t job = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const int64_t b = tile / p.tiles_per_frame;
      const int q0 = (int)(tile - b * p.tiles_per_frame) * p.tile;
      const float* xb = a.x + b * (int64_t)a.Cin * a.Lin;
      const int u0 = q0 * p.nsub;                    // first u = pos + padL of this tile
      for (int plane = 0; plane < jobs_per_tile; ++plane, ++job) {
        const uint32_t buf = job & 1u;
        if ((int)buf != pgroup) continue;             // the other group's job (the `continue` still advances job)
        mbar_wait(&a_empty[buf], ((job >> 1) & 1u) ^ 1u);
        uint8_t* dstA = sA + buf * abuf_bytes;
        // A task = 8 channels x 4 consecutive positions: eight 128-bit loads (lanes walk position quads, so a warp
        // request is 512 contiguous bytes), four 16-byte swizzled row stores.  kATasks tasks are in flight per thread
        // (1 KB of loads): the staging rate is set by bytes in flight, not by instruction count.
        const int pos_lo = u0 - p.padL;                       // position of buffer coordinate j = 0
        const int pos_al = pos_lo & ~3;                       // floor to a multiple of 4 (two's complement: works for < 0)
        const int quads = (pos_lo - pos_al + span + 3) >> 2;  // position quads covering [pos_lo, pos_lo + span)
        const int total = groups * quads;
        for (int i0 = ptid; i0 < total; i0 += kATasks * nprod) {
          float4 v[kATasks][8];
          int gg[kATasks], pp[kATasks];
#pragma unroll
          for (int q = 0; q < kATasks; ++q) {
            const int i = i0 + q * nprod;
            const int g = i < total ? i / quads : 0, k = i < total ? i - g * quads : 0;
            const int pos = pos_al + 4 * k;
            gg[q] = g; pp[q] = pos;
            const bool inside = i < total && pos >= 0 && pos < a.Lin;    // Lin % 4 == 0: a quad is all in or all out
            const int c0 = g << 3;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              v[q][c] = (inside && c0 + c < a.Cin) ? __ldg(reinterpret_cast<const float4*>(xb + (int64_t)(c0 + c) * a.Lin + pos))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int q = 0; q < kATasks; ++q) {
            if (i0 + q * nprod >= total) break;
            const int g = gg[q];
            const int slab = g >> 3, chunk = g & 7;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = pp[q] + e - pos_lo;
              if (j < 0 || j >= span) continue;
              __half2 h[4];     // packed conversions: one F2FP per channel pair
#pragma unroll
              for (int c = 0; c < 8; c += 2) {
                const float fa = e == 0 ? v[q][c].x : e == 1 ? v[q][c].y : e == 2 ? v[q][c].z : v[q][c].w;
                const float fb = e == 0 ? v[q][c + 1].x : e == 1 ? v[q][c + 1].y : e == 2 ? v[q][c + 1].z : v[q][c + 1].w;
                const __half2 hi = __floats2half2_rn(fa, fb);
                if (plane == 0) {
                  h[c >> 1] = hi;
                } else {
                  const float2 back = __half22float2(hi);
                  h[c >> 1] = __floats2half2_rn(fa - back.x, fb - back.y);
                }
              }
              const int sub = j % p.nsub, row = j / p.nsub;
              uint8_t* dst = dstA + (uint32_t)sub * sub_bytes + (uint32_t)slab * slab_bytes + (uint32_t)row * 128u +
                             (uint32_t)((chunk ^ (row & 7)) << 4);
              *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
            }
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[buf]);
      }
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == kEpiWarps) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(p.tmem_cols) : "memory");
  }
}

}  // namespace

// bytes of packed weights the tensor engine needs for one layer
int64_t tc_wpack_bytes(int K, int Cin, int Cout) {
  const int n_pad = (Cout + 15) & ~15, slabs = (Cin + 63) / 64;
  return 2LL * K * slabs * n_pad * 128;
}

bool tc_conv_supported(const ConvArgs& a) {
  if (a.x_cl || a.y_cl || a.res_cl) return false;
  if (a.stride != 1 && a.stride != 2) return false;
  if (a.stride > 1 && a.dil != 1) return false;
  if (a.Cin < 16 || a.Cin > 128 || a.Cout < 16 || a.Cout > 128) return false;
  int Lout, padL;
  same_padding(a.Lin, a.K, a.dil, a.stride, &Lout, &padL);
  if (Lout % 128 != 0 || a.Lin % a.stride != 0) return false;
  const int n_pad = (a.Cout + 15) & ~15, mt = Lout / 128;
  if (mt * n_pad > 512) return false;
  if (a.shuffle != 1 && a.Cout % a.shuffle != 0) return false;
  return true;
}


static int launch_conv_tc_pipe(const ConvArgs& a, int precision, void* wpack, cudaStream_t st) {
  TcPipe p;
  p.a = a;
  same_padding(a.Lin, a.K, a.dil, a.stride, &p.Lout, &p.padL);
  p.n_pad = (a.Cout + 15) & ~15;
  p.ksteps = (a.Cin + 15) / 16;
  p.slabs = (a.Cin + 63) / 64;
  p.tile = p.Lout % 256 == 0 ? 256 : 128;
  p.mt = p.tile / 128;
  p.tiles_per_frame = p.Lout / p.tile;
  p.n_tiles = a.B * p.tiles_per_frame;
  p.nsub = a.stride;
  p.rows = ((p.tile + ((a.K - 1) * a.dil) / a.stride + 1) + 7) & ~7;
  p.passes = precision == 1 ? 3 : 1;
  int cols = 32;
  while (cols < 2 * p.mt * p.n_pad) cols *= 2;
  p.tmem_cols = cols;
  p.wpack = reinterpret_cast<const uint4*>(wpack);
  const size_t a_bytes = 2ull * p.nsub * p.slabs * p.rows * 128, unit_bytes = (size_t)p.n_pad * 128;
  const size_t budget = 224 * 1024 - 1024;
  if (cols > 512 || a_bytes + 2 * unit_bytes > budget || (a.Lin & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(a.x) & 15) != 0) return 1;   // caller falls back to the simple kernel
  // group several (tap, slab) units into one ring stage so that the issuing thread pays one barrier round trip per
  // ~24 KB of weights, while keeping at least 3 stages in flight
  const int units = a.K * p.slabs;
  int group = 1;
  for (int g = 1; g <= units; ++g)
    if (units % g == 0 && (size_t)g * unit_bytes <= 24 * 1024 && a_bytes + 3 * (size_t)g * unit_bytes <= budget) group = g;
  p.group = group;
  p.stages_per_pass = units / group;
  const size_t stage_bytes = (size_t)group * unit_bytes;
  size_t nst = (budget - a_bytes) / stage_bytes;
  const size_t per_tile = (size_t)p.passes * p.stages_per_pass;    // no point in a ring deeper than two tiles' worth
  if (nst > 2 * per_tile) nst = 2 * per_tile;
  if (nst < 2) nst = 2;
  if (nst > (size_t)kMaxBStages) nst = kMaxBStages;
  p.bstages = (int)nst;
  const size_t smem = 1024 + a_bytes + nst * stage_bytes;
  NSC_CUDA_OK(cudaFuncSetAttribute(tc_conv_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  char name[32];
  snprintf(name, sizeof(name), "tc%d_k%dd%ds%d_c%dto%d", precision, a.K, a.dil, a.stride, a.Cin, a.Cout);
  const double macs = (double)a.B * p.Lout * a.K * a.Cin * a.Cout;
  const double bytes = 4.0 * ((double)a.B * ((double)a.Lin * a.Cin + (double)p.Lout * a.Cout) + (double)a.K * a.Cin * a.Cout);
  ProfScope prof(st, name, 2.0 * macs, bytes);
  const int64_t grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  tc_conv_pipe_kernel<<<(unsigned)grid, kPipeThreads, smem, st>>>(p);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

// precision: 1 = fp16 hi/lo split (3 MMAs, fp32-class), 2 = fp16 inputs only.  wpack: tc_wpack_bytes() scratch.
int launch_conv_tc(const ConvArgs& a, int precision, void* wpack, cudaStream_t st) {
  if (a.B == 0) return NSC_OK;
  NSC_CHECK_ARG(tc_conv_supported(a), "tensor conv: unsupported shape");
  NSC_CHECK_ARG(wpack != nullptr && (reinterpret_cast<uintptr_t>(wpack) & 15) == 0, "tensor conv: bad weight scratch");
  NSC_CHECK_ARG(tc_wpack_bytes(a.K, a.Cin, a.Cout) <= kTcWpackBytes, "tensor conv: packed weights exceed the scratch");
  TcLaunch p;
  p.a = a;
  same_padding(a.Lin, a.K, a.dil, a.stride, &p.Lout, &p.padL);
  p.n_pad = (a.Cout + 15) & ~15;
  p.ksteps = (a.Cin + 15) / 16;
  p.slabs = (a.Cin + 63) / 64;
  p.mtiles = p.Lout / 128;
  p.nbuf = a.stride;
  const int max_shift_rows = ((a.K - 1) * a.dil) / a.stride;
  p.rows = ((p.Lout + max_shift_rows + 1) + 7) & ~7;
  // every input position must land inside a buffer: row = (pos + padL) / stride <= (Lin - 1 + padL) / stride
  const int need_rows = (a.Lin - 1 + p.padL) / a.stride + 1;
  if (p.rows < ((need_rows + 7) & ~7)) p.rows = (need_rows + 7) & ~7;
  p.passes = precision == 1 ? 3 : 1;
  int cols = 32;
  while (cols < p.mtiles * p.n_pad) cols *= 2;
  p.tmem_cols = cols;
  p.wpack = reinterpret_cast<const uint4*>(wpack);
  const size_t smem = 1024 + (size_t)p.nbuf * p.slabs * p.rows * 128 + 2ull * p.slabs * p.n_pad * 128;
  NSC_CHECK_ARG(smem + 256 <= 227 * 1024, "tensor conv: needs %zu bytes of shared memory", smem);
  {
    const int total = 2 * a.K * p.slabs * p.n_pad * 64;
    ProfScope prof(st, "tc_pack_weights", 0.0, 4.0 * a.K * a.Cin * a.Cout + 2.0 * total);
    tc_pack_weights_kernel<<<ceil_div(total, 256) < 592 ? ceil_div(total, 256) : 592, 256, 0, st>>>(
        a.w, a.K, a.Cin, a.Cout, p.n_pad, p.slabs, reinterpret_cast<__half*>(wpack));
    NSC_LAUNCH_OK();
  }
  static const bool force_simple = getenv("NSC_TC_SIMPLE") != nullptr;   // debugging aid: one-CTA-per-frame kernel
  if (!force_simple) {
    const int rc = launch_conv_tc_pipe(a, precision, wpack, st);
    if (rc <= 0) return rc;
  }
  NSC_CUDA_OK(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  char name[32];
  snprintf(name, sizeof(name), "tc%d_k%dd%ds%d_c%dto%d", precision, a.K, a.dil, a.stride, a.Cin, a.Cout);
  const double macs = (double)a.B * p.Lout * a.K * a.Cin * a.Cout;
  const double bytes = 4.0 * ((double)a.B * ((double)a.Lin * a.Cin + (double)p.Lout * a.Cout) + (double)a.K * a.Cin * a.Cout);
  ProfScope prof(st, name, 2.0 * macs, bytes);
  tc_conv_kernel<<<(unsigned)a.B, kTcThreads, smem, st>>>(p);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // namespace nsc
