// Weight gradient of a stride-1 conv on the tensor cores (training step).
//
//   dW[t][ci][co] = sum over frames b and positions p of  x[b][ci][p + s_t] * g[b][co][p],   s_t = t * dil - padL  (zero outside the frame)
//
// is a GEMM whose K dimension is the POSITION axis -- 65,536 positions for 128 frames -- and whose other two dimensions are tiny
// (100 x 180 for the codec's 100 -> 20 conv).  The fp32 training tensors are (B, C, L): positions are contiguous per channel, i.e.
// both operands are K-major as they lie.  Mapping:
//   M rows   = channels of the WIDER of the two tensors, unshifted               ("W operand": x if Cin >= Cout, else g)
//   N rows   = (tap, channel) of the narrower one, shifted by the tap's offset   ("S operand": taps in N; 9 x 20 = 180 columns)
//   K        = 64 positions per block; a launch splits the B * L / 64 blocks over the CTAs, each CTA accumulates its share in ONE
//              TMEM accumulator (fp32) and writes it out once.
// Producer warps read the fp32 rows (coalesced along positions), split every value into fp16 hi + lo (the same x = hi + lo,
// three-MMA arithmetic as the forward engines: hi*hi + lo*hi + hi*lo) and write the two operand tiles in the SWIZZLE_128B K-major
// image the MMA reads; the shift of a tap is just an offset of the source pointer, so no shifted copy ever exists in memory.  One
// elected thread issues the tcgen05.mma chain; four epilogue warps copy the accumulator to a per-CTA slice of a scratch buffer and a
// second small kernel sums the slices in a fixed order (deterministic, unlike the CUDA-core kernel's atomics) into dW.  The bias
// gradient rides along as a row (or column) of ones.  Layers whose tap count times channel count exceeds 255 columns run as
// several launches over tap ranges (k15 gates: 12 + 3 taps; 100 -> 100: two taps per launch).
#include <stdlib.h>

#include "conv.cuh"
#include "tc_common.cuh"

namespace nsc {

using namespace tc;

namespace {

constexpr int kWgEpiWarps = 4, kWgProdWarps = 16;
constexpr int kWgATasks = 128 * 8 / (kWgProdWarps * 32), kWgSTasks = 256 * 8 / (kWgProdWarps * 32);   // operand chunks per producer thread
constexpr int kWgThreads = (kWgEpiWarps + 1 + kWgProdWarps) * 32;
constexpr int kWgMaxGrid = 148;

struct WgTc {
  const float* w_src;   // (B, Cw, L) unshifted operand (rows of the accumulator)
  const float* s_src;   // (B, Cs, L) shifted operand (columns: (tap, channel))
  int Cw, Cs, L;        // L: positions of the W operand (= output positions of the conv)
  int Ls, s_stride;     // the S operand has Ls positions and is read at s_stride * position + shift (2 for a stride-2 conv: S = x)
  int shift_sign;       // column (tt, ch) reads s_src[.., pos + shift_sign * s_t]:  -1 when S = g (mode 0), +1 when S = x (mode 1)
  int dil, padL, t0, taps;
  int n_used, npad;     // columns in use (taps * Cs [+ 1 ones column]) and the MMA's N
  int ones_row_w;       // row of ones in the W tile (bias gradient when S = g), or -1
  int ones_row_s;       // column of ones in the S tile (bias gradient when W = g), or -1
  int64_t nblk;         // B * L / 64 position blocks
  float* scratch;       // [gridDim.x][128][npad]
  // staged producers: the S operand's source rows of a block sit in shared memory as fp32 (window floats per row, covering the
  // block's positions plus `halo` on either side, zeros outside the frame), so the tap shifts are shared-memory offsets
  int halo, window, n_raw;
};

__device__ __forceinline__ uint32_t wg_sw128(int row, int chunk) { return (uint32_t)row * 128u + ((((uint32_t)chunk) ^ ((uint32_t)row & 7u)) << 4); }

constexpr int kWgRawMax = 8;    // fp32 float4 chunks of the raw S rows per producer thread

template <bool kStaged>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgTc p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[2], empty[2], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s_tile = (uint32_t)p.npad * 128u;
  const uint32_t stage_bytes = 2u * 16384u + 2u * s_tile;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], kWgProdWarps); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == kWgEpiWarps) tmem_alloc(&tmem_base_s, 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int bpf = p.L >> 6;                       // 64-position blocks per frame

  if (warp < kWgEpiWarps) {
    // =========================== epilogue: accumulator -> this CTA's scratch slice ===========================
    mbar_wait_relaxed(&acc_full, 0u);
    tc_fence_after();
    float* dst = p.scratch + ((size_t)blockIdx.x * 128 + (size_t)(warp * 32 + lane)) * (size_t)p.npad;
    for (int c0 = 0; c0 < p.npad; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; e += 4)
        *reinterpret_cast<float4*>(dst + c0 + e) = make_float4(__uint_as_float(r[e]), __uint_as_float(r[e + 1]), __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
    }
    tc_fence_before();
  } else if (warp == kWgEpiWarps) {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(p.npad);
      uint32_t j = 0, accum = 0;
      for (int64_t blk = blockIdx.x; blk < p.nblk; blk += gridDim.x, ++j) {
        const uint32_t st = j & 1u;
        mbar_wait(&full[st], (j >> 1) & 1u);
        tc_fence_after();
        const uint32_t base = smem_u32(smem + st * stage_bytes);
        const uint32_t a_hi = desc_lo(base), a_lo = desc_lo(base + 16384u), s_hi = desc_lo(base + 32768u), s_lo = desc_lo(base + 32768u + s_tile);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) { mma_f16_ss(tmem, desc_from_lo(a_hi + 2u * ks), desc_from_lo(s_hi + 2u * ks), idesc, accum); accum = 1; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_f16_ss(tmem, desc_from_lo(a_lo + 2u * ks), desc_from_lo(s_hi + 2u * ks), idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_f16_ss(tmem, desc_from_lo(a_hi + 2u * ks), desc_from_lo(s_lo + 2u * ks), idesc, 1u);
        umma_commit(&empty[st]);
      }
      umma_commit(&acc_full);
    }
  } else if constexpr (kStaged) {
    // =========================== producers, staged form ===========================
    // Global loads only fetch what is unique -- the W operand's chunks and the S operand's source rows (one aligned window per
    // channel) -- and they are issued ONE BLOCK AHEAD into registers, so their latency hides behind the conversion of the current
    // block.  The source rows go to shared memory as fp32; the (tap, channel) rows of the S tile are then shifted reads of them.
    const int ptid = tid - (kWgEpiWarps + 1) * 32;
    constexpr int kProd = kWgProdWarps * 32;
    const uint4 ones = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u), zero4 = make_uint4(0, 0, 0, 0);
    float* raw0 = reinterpret_cast<float*>(smem + 2 * stage_bytes);
    const int win = p.window, w4 = win >> 2, n_raw4 = p.Cs * w4;
    int a_kind[kWgATasks], a_src[kWgATasks];
    uint32_t a_off[kWgATasks];
#pragma unroll
    for (int i = 0; i < kWgATasks; ++i) {
      const int a = ptid + i * kProd, row = a >> 3, c = a & 7;
      a_kind[i] = row < p.Cw ? 1 : (row == p.ones_row_w ? 2 : 0);
      a_src[i] = row * p.L + 8 * c;
      a_off[i] = wg_sw128(row, c);
    }
    int s_kind[kWgSTasks], s_rel[kWgSTasks];
    uint32_t s_off[kWgSTasks];
#pragma unroll
    for (int i = 0; i < kWgSTasks; ++i) {
      const int t = ptid + i * kProd, row = t >> 3, c = t & 7;
      s_kind[i] = t >= p.npad * 8 ? -1 : (row < p.taps * p.Cs ? 1 : (row == p.ones_row_s ? 2 : 0));
      const int tt = row / p.Cs, ch = row - tt * p.Cs;
      s_rel[i] = ch * win + p.halo + p.s_stride * 8 * c + p.shift_sign * ((p.t0 + tt) * p.dil - p.padL);
      s_off[i] = wg_sw128(row, c);
    }
    int r_dst[kWgRawMax], r_src[kWgRawMax], r_pos[kWgRawMax];      // shared-memory float index, source row offset, position relative to the window start
#pragma unroll
    for (int k = 0; k < kWgRawMax; ++k) {
      const int idx = ptid + k * kProd;
      const int ch = idx / w4, f = idx - ch * w4;
      r_dst[k] = idx < n_raw4 ? idx * 4 : -1;
      r_src[k] = ch * p.Ls;
      r_pos[k] = 4 * f - p.halo;
    }
    float4 fa[kWgATasks][2], fr[kWgRawMax];
    auto issue_loads = [&](int64_t blk) {
      const int64_t b = blk / bpf;
      const int q0 = (int)(blk - b * bpf) * 64;
      const float* wb = p.w_src + (size_t)b * p.Cw * p.L + q0;
      const float* sb = p.s_src + (size_t)b * p.Cs * p.Ls;
#pragma unroll
      for (int i = 0; i < kWgATasks; ++i)
        if (a_kind[i] == 1) {
          const float4* src = reinterpret_cast<const float4*>(wb + a_src[i]);
          fa[i][0] = __ldg(src);
          fa[i][1] = __ldg(src + 1);
        }
#pragma unroll
      for (int k = 0; k < kWgRawMax; ++k)
        if (r_dst[k] >= 0) {
          const int pos = p.s_stride * q0 + r_pos[k];                 // a multiple of 4: the float4 is wholly inside or outside the frame
          fr[k] = (pos >= 0 && pos < p.Ls) ? __ldg(reinterpret_cast<const float4*>(sb + r_src[k] + pos)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    uint32_t j = 0;
    if ((int64_t)blockIdx.x < p.nblk) issue_loads(blockIdx.x);
    for (int64_t blk = blockIdx.x; blk < p.nblk; blk += gridDim.x, ++j) {
      const uint32_t st = j & 1u;
      float* raw = raw0 + (p.n_raw == 2 ? (size_t)(j & 1u) * (size_t)p.Cs * win : 0);
      uint8_t* A_hi = smem + st * stage_bytes;
      uint8_t* A_lo = A_hi + 16384;
      uint8_t* S_hi = A_hi + 32768;
      uint8_t* S_lo = S_hi + s_tile;
      mbar_wait_relaxed(&empty[st], ((j >> 1) & 1u) ^ 1u);
#pragma unroll
      for (int k = 0; k < kWgRawMax; ++k)
        if (r_dst[k] >= 0) *reinterpret_cast<float4*>(raw + r_dst[k]) = fr[k];
#pragma unroll
      for (int i = 0; i < kWgATasks; ++i) {
        uint4 hi = zero4, lo = zero4;
        if (a_kind[i] == 1) {
          const float v[8] = {fa[i][0].x, fa[i][0].y, fa[i][0].z, fa[i][0].w, fa[i][1].x, fa[i][1].y, fa[i][1].z, fa[i][1].w};
          split8(v, hi, lo);
        } else if (a_kind[i] == 2) {
          hi = ones;
        }
        *reinterpret_cast<uint4*>(A_hi + a_off[i]) = hi;
        *reinterpret_cast<uint4*>(A_lo + a_off[i]) = lo;
      }
      asm volatile("bar.sync 1, %0;" :: "n"(kProd) : "memory");        // the source rows are in shared memory
      if (blk + gridDim.x < p.nblk) issue_loads(blk + gridDim.x);       // next block's loads fly while this one is converted
#pragma unroll
      for (int i = 0; i < kWgSTasks; ++i) {
        if (s_kind[i] < 0) continue;
        uint4 hi = zero4, lo = zero4;
        if (s_kind[i] == 1) {
          const float* src = raw + s_rel[i];
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = src[p.s_stride * e];
          split8(v, hi, lo);
        } else if (s_kind[i] == 2) {
          hi = ones;
        }
        *reinterpret_cast<uint4*>(S_hi + s_off[i]) = hi;
        *reinterpret_cast<uint4*>(S_lo + s_off[i]) = lo;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
      if (p.n_raw == 1) asm volatile("bar.sync 1, %0;" :: "n"(kProd) : "memory");   // single source buffer: everybody has read it
    }
  } else {
    // =========================== producers: fp32 rows -> fp16 hi / lo operand tiles ===========================
    // Every thread owns the same operand chunks (row, 8 positions) in every block: what it reads (source row, tap shift) and where it
    // writes (swizzled offset) is worked out once; per block it issues ALL of its loads before converting anything, so that the
    // latencies overlap.
    const int ptid = tid - (kWgEpiWarps + 1) * 32;
    const uint4 ones = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u), zero4 = make_uint4(0, 0, 0, 0);
    int a_kind[kWgATasks], a_src[kWgATasks];          // kind: 0 zeros, 1 data, 2 ones, -1 no task
    uint32_t a_off[kWgATasks];
#pragma unroll
    for (int i = 0; i < kWgATasks; ++i) {
      const int a = ptid + i * kWgProdWarps * 32, row = a >> 3, c = a & 7;
      a_kind[i] = row < p.Cw ? 1 : (row == p.ones_row_w ? 2 : 0);
      a_src[i] = row * p.L + 8 * c;
      a_off[i] = wg_sw128(row, c);
    }
    int s_kind[kWgSTasks], s_src[kWgSTasks], s_pos[kWgSTasks];
    uint32_t s_off[kWgSTasks];
#pragma unroll
    for (int i = 0; i < kWgSTasks; ++i) {
      const int t = ptid + i * kWgProdWarps * 32, row = t >> 3, c = t & 7;
      s_kind[i] = t >= p.npad * 8 ? -1 : (row < p.taps * p.Cs ? 1 : (row == p.ones_row_s ? 2 : 0));
      const int tt = row / p.Cs, ch = row - tt * p.Cs;
      s_src[i] = ch * p.Ls;
      s_pos[i] = p.s_stride * 8 * c + p.shift_sign * ((p.t0 + tt) * p.dil - p.padL);
      s_off[i] = wg_sw128(row, c);
    }
    uint32_t j = 0;
    for (int64_t blk = blockIdx.x; blk < p.nblk; blk += gridDim.x, ++j) {
      const uint32_t st = j & 1u;
      const int64_t b = blk / bpf;
      const int q0 = (int)(blk - b * bpf) * 64;
      uint8_t* A_hi = smem + st * stage_bytes;
      uint8_t* A_lo = A_hi + 16384;
      uint8_t* S_hi = A_hi + 32768;
      uint8_t* S_lo = S_hi + s_tile;
      const float* wb = p.w_src + (size_t)b * p.Cw * p.L + q0;
      const float* sb = p.s_src + (size_t)b * p.Cs * p.Ls;
      float4 fa[kWgATasks][2];
      float vs[kWgSTasks][8];
#pragma unroll
      for (int i = 0; i < kWgATasks; ++i)
        if (a_kind[i] == 1) {
          const float4* src = reinterpret_cast<const float4*>(wb + a_src[i]);
          fa[i][0] = __ldg(src);
          fa[i][1] = __ldg(src + 1);
        }
#pragma unroll
      for (int i = 0; i < kWgSTasks; ++i)
        if (s_kind[i] == 1) {
          const float* src = sb + s_src[i];
          const int pos0 = p.s_stride * q0 + s_pos[i];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int pos = pos0 + p.s_stride * e;
            vs[i][e] = (pos >= 0 && pos < p.Ls) ? __ldg(src + pos) : 0.f;
          }
        }
      mbar_wait_relaxed(&empty[st], ((j >> 1) & 1u) ^ 1u);      // (the loads above do not touch the stage: they overlap the wait)
#pragma unroll
      for (int i = 0; i < kWgATasks; ++i) {
        uint4 hi = zero4, lo = zero4;
        if (a_kind[i] == 1) {
          const float v[8] = {fa[i][0].x, fa[i][0].y, fa[i][0].z, fa[i][0].w, fa[i][1].x, fa[i][1].y, fa[i][1].z, fa[i][1].w};
          split8(v, hi, lo);
        } else if (a_kind[i] == 2) {
          hi = ones;
        }
        *reinterpret_cast<uint4*>(A_hi + a_off[i]) = hi;
        *reinterpret_cast<uint4*>(A_lo + a_off[i]) = lo;
      }
#pragma unroll
      for (int i = 0; i < kWgSTasks; ++i) {
        if (s_kind[i] < 0) continue;
        uint4 hi = zero4, lo = zero4;
        if (s_kind[i] == 1) split8(vs[i], hi, lo);
        else if (s_kind[i] == 2) hi = ones;
        *reinterpret_cast<uint4*>(S_hi + s_off[i]) = hi;
        *reinterpret_cast<uint4*>(S_lo + s_off[i]) = lo;
      }
      fence_async_smem();                     // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[st]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kWgEpiWarps) tmem_dealloc(tmem, 256u);
}

// sums the CTAs' slices in a fixed order and adds the result into dW (and the bias gradient from the row / column of ones)
struct WgRed {
  const float* scratch;
  int grid, npad, Cw, Cs, taps, t0;
  int mode;             // 0: rows = ci, columns = (tap, co);  1: rows = co, columns = (tap, ci)
  int Cin, Cout, dil, padL;
  int ones_row_w, ones_row_s;
  float* dw;
  float* db;
};

// One block = one accumulator row x 32 columns; its 8 warps take every 8th slice each (all of a thread's loads are in flight at once:
// a serial walk over 148 slices costs 148 L2 latencies -- 31 us measured -- whatever the layer's size), then the eight partial sums
// are added in a fixed order through shared memory.
constexpr int kWgRedGroups = 8, kWgRedMaxPer = (kWgMaxGrid + kWgRedGroups - 1) / kWgRedGroups;
__global__ void __launch_bounds__(32 * kWgRedGroups) wgrad_tc_reduce_kernel(const WgRed p) {
  __shared__ float part[kWgRedGroups][32];
  const int ncol = p.taps * p.Cs;
  const int cols = ncol + (p.ones_row_s >= 0 ? 1 : 0);
  const int cgroups = (cols + 31) >> 5;
  const int ri = blockIdx.x / cgroups, j = (blockIdx.x - ri * cgroups) * 32 + (threadIdx.x & 31);
  const int sg = threadIdx.x >> 5;
  const int w = ri < p.Cw ? ri : p.ones_row_w;
  const int col = j < ncol ? j : p.ones_row_s;
  float v[kWgRedMaxPer];
  if (j < cols) {
    const float* src = p.scratch + (size_t)w * p.npad + col;
    const size_t slice = (size_t)128 * p.npad;
#pragma unroll
    for (int k = 0; k < kWgRedMaxPer; ++k) {
      const int c = sg + k * kWgRedGroups;
      v[k] = c < p.grid ? src[(size_t)c * slice] : 0.f;
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kWgRedMaxPer; ++k) s += v[k];
    part[sg][threadIdx.x & 31] = s;
  }
  __syncthreads();
  if (sg != 0 || j >= cols) return;
  float s = 0.f;
#pragma unroll
  for (int g8 = 0; g8 < kWgRedGroups; ++g8) s += part[g8][threadIdx.x];
  if (ri < p.Cw && j < ncol) {
    const int tt = j / p.Cs, ch = j - tt * p.Cs, t = p.t0 + tt;
    const size_t idx = p.mode == 0 ? ((size_t)t * p.Cin + w) * p.Cout + ch : ((size_t)t * p.Cin + ch) * p.Cout + w;
    p.dw[idx] += s;
  } else if (ri >= p.Cw && j < ncol) {          // row of ones x shifted g: the tap with zero offset sums g over the whole frame
    const int tt = j / p.Cs, ch = j - tt * p.Cs;
    if (p.db != nullptr && (p.t0 + tt) * p.dil - p.padL == 0) p.db[ch] += s;
  } else if (ri < p.Cw && j >= ncol) {          // g x column of ones
    if (p.db != nullptr) p.db[w] += s;
  }
}

bool wgrad_tc_on() {
  static const bool on = [] { const char* e = getenv("NSC_WGRAD_TC"); return !(e && e[0] == '0'); }();
  return on;
}

}  // namespace

int64_t wgrad_tc_scratch_floats() { return (int64_t)kWgMaxGrid * 128 * 256; }

bool wgrad_tc_supported(int64_t B, int Lin, int Lout, int Cin, int Cout, int K, int dil, int stride, int padL) {
  if (!wgrad_tc_on()) return false;
  if ((stride != 1 && stride != 2) || Lin != Lout * stride || (Lout & 63) || B < 1) return false;
  if (Cin < 1 || Cout < 1 || Cin > 128 || Cout > 128) return false;
  if (K < 1 || dil < 1 || (stride == 2 && dil != 1)) return false;
  // the bias gradient needs a tap with zero offset when the rows are input channels (SAME padding with an odd kernel has one)
  if (stride == 1 && Cin >= Cout && (padL % dil != 0 || padL / dil >= K)) return false;
  return true;
}

int launch_wgrad_tc(const float* x, const float* g, float* dw, float* db, int64_t B, int L, int Cin, int Cout, int K, int dil, int padL,
                    float* scratch, cudaStream_t st, int stride) {
  // L = OUTPUT positions.  A stride-2 conv reads x at 2 p + s_t: x must be the shifted (S) operand
  const int mode = (Cin >= Cout && stride == 1) ? 0 : 1;
  const int Cw = mode == 0 ? Cin : Cout, Cs = mode == 0 ? Cout : Cin;
  const int64_t nblk = B * (L >> 6);
  int grid = sm_count() < kWgMaxGrid ? sm_count() : kWgMaxGrid;
  if (nblk < grid) grid = (int)nblk;
  const int tpp_max = 255 / Cs;                   // leaves room for the column of ones
  ProfScope prof(st, "wgrad_tc", 2.0 * B * L * (double)K * Cin * Cout, 4.0 * B * ((double)L * Cin + (double)L * Cout));
  bool bias_done = false;
  for (int t0 = 0; t0 < K; t0 += tpp_max) {
    const int taps = K - t0 < tpp_max ? K - t0 : tpp_max;
    WgTc p;
    p.w_src = mode == 0 ? x : g; p.s_src = mode == 0 ? g : x;
    p.Cw = Cw; p.Cs = Cs; p.L = L; p.shift_sign = mode == 0 ? -1 : 1;
    p.Ls = mode == 0 ? L : L * stride; p.s_stride = mode == 0 ? 1 : stride;
    p.dil = dil; p.padL = padL; p.t0 = t0; p.taps = taps;
    p.ones_row_w = -1; p.ones_row_s = -1;
    if (db != nullptr) {
      if (mode == 0) p.ones_row_w = Cw;                                             // (Cw <= 100 < 128 on this path; checked below)
      else if (!bias_done) { p.ones_row_s = taps * Cs; bias_done = true; }
    }
    NSC_CHECK_ARG(p.ones_row_w < 128, "wgrad_tc: no spare row for the bias gradient (%d channels)", Cw);
    p.n_used = taps * Cs + (p.ones_row_s >= 0 ? 1 : 0);
    p.npad = (p.n_used + 15) & ~15;
    p.nblk = nblk;
    p.scratch = scratch;
    const size_t stages = 2 * (2 * 16384 + 2 * (size_t)p.npad * 128);
    // staged producers when the fp32 source rows of a block fit beside the operand stages (twice, or once)
    int max_shift = 0;
    for (int tt = 0; tt < taps; ++tt) {
      const int sh = (t0 + tt) * dil - padL;
      if (sh > max_shift) max_shift = sh;
      if (-sh > max_shift) max_shift = -sh;
    }
    p.halo = (max_shift + 3) & ~3;
    p.window = 64 * p.s_stride + 2 * p.halo;
    const size_t raw_bytes = (size_t)Cs * p.window * 4;
    const size_t budget = 227 * 1024 - 2048;
    p.n_raw = stages + 2 * raw_bytes <= budget ? 2 : (stages + raw_bytes <= budget ? 1 : 0);
    // Measured (profiles/r02_wgrad_tc.log, 128 frames, all 54 layers of a step): direct producers 2.43 ms, staged 2.63 ms -- the kernel
    // is bound by the producers' instruction issue (fp16 hi/lo conversion of nine shifted copies of the narrow operand), not by load
    // latency, so hiding the loads buys nothing.  Default: direct; NSC_WGRAD_TC_STAGED=1 selects the staged form (parity-tested).
    static const bool staged = [] { const char* e = getenv("NSC_WGRAD_TC_STAGED"); return e && e[0] == '1'; }();
    if (!staged || Cs * (p.window / 4) > kWgRawMax * kWgProdWarps * 32) p.n_raw = 0;
    const size_t smem = 1024 + stages + (size_t)p.n_raw * raw_bytes;
    if (p.n_raw > 0) {
      NSC_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      wgrad_tc_kernel<true><<<grid, kWgThreads, smem, st>>>(p);
    } else {
      NSC_CUDA_OK(cudaFuncSetAttribute(wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      wgrad_tc_kernel<false><<<grid, kWgThreads, smem, st>>>(p);
    }
    NSC_LAUNCH_OK();
    WgRed r;
    r.scratch = scratch; r.grid = grid; r.npad = p.npad; r.Cw = Cw; r.Cs = Cs; r.taps = taps; r.t0 = t0; r.mode = mode;
    r.Cin = Cin; r.Cout = Cout; r.dil = dil; r.padL = padL; r.ones_row_w = p.ones_row_w; r.ones_row_s = p.ones_row_s;
    r.dw = dw; r.db = db;
    const int rows = Cw + (p.ones_row_w >= 0 ? 1 : 0), cols = taps * Cs + (p.ones_row_s >= 0 ? 1 : 0);
    wgrad_tc_reduce_kernel<<<rows * ((cols + 31) / 32), 32 * kWgRedGroups, 0, st>>>(r);
    NSC_LAUNCH_OK();
  }
  return NSC_OK;
}

}  // namespace nsc
