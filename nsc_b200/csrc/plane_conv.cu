// Plane engine: the codec's conv layers on tcgen05 with fp16 plane-image activations (plane.cuh).
//
// Three kernels, all persistent (one CTA per SM) and warp-specialised -- bulk-copy loader threads, one or two MMA-issuing
// threads, epilogue warps; every hand-off is an mbarrier and tensor-core completions arrive through tcgen05.commit:
//
//   plane_t_kernel   PK_T  "taps in N".  A CTA walks whole frames, tile by tile (128 positions).  Per tile ONE MMA chain
//                    computes P[row, (tap, co)] for all taps at once (N = 9 * 20 -> 192: 96 cycles per K step instead of
//                    9 x 44 for nine N = 32 instructions) with the layer's packed weights RESIDENT in shared memory and
//                    the input streamed slab by slab through a ring.  The tap sum y[row] = sum_t P[row + s_t, t] crosses
//                    TMEM lanes: each epilogue warp owns 32 rows (12 warps: 4 lane quarters x three 8-channel chunks), pulls
//                    P[row + s_t] from lane (lane + s_t) mod 32 with a shuffle, and the wrapped lanes accumulate the
//                    contribution that belongs to the SAME lane of the neighbouring quarter ("up"/"down" spill).  Quarters
//                    are finished in row order: a quarter's first rows take the previous quarter's up-spill, its last rows
//                    are parked in shared memory and finished by the next quarter, which holds their down-spill.  No halo
//                    is ever loaded or recomputed.  Finished rows go to a shared-memory ring that mirrors the output image
//                    and leave by one bulk store per tile.
//                    Narrow inputs (20 -> 20) use the GROUPED form: the nine taps are three groups of three, a group is a row shift
//                    of the A descriptor (-2d, 0, +2d) and only the shifts INSIDE a group cross TMEM lanes -- five tap slots
//                    (shifts -2d .. +2d) instead of nine, i.e. half the lane-crossing volume that bounds this layer.  The outer
//                    groups shift one way only ({-2d,-d,0} and {0,+d,+2d}), which keeps frame borders exact: a row that would
//                    take a contribution from outside the frame only ever needed zeros from there.
//   plane_x_kernel   PK_X  one MMA per (tap, K step): a tap is a row shift of the A descriptor inside the staged tile
//                    (+8 halo rows each side, which are the zero rows of the image at frame borders).  Weights stay
//                    resident when they fit, else stream through a ring; with hi/lo planes every W_hi unit is used for
//                    both the A_hi and A_lo products while it is resident.  The 100 -> 100 convs (stride 2 / sub-pixel).
//                    <true, .>: PK_GEN, the same pipeline fed from a Toeplitz tile that producer warps build from a 1-channel
//                    fp32 signal (decoder's k9 1 -> 20).
//   plane_xs_kernel  PK_X / PK_GEN with a STAGED epilogue (narrow-input 20 -> 100 / 20 -> 50 + residual, and the k55 stem): the
//                    residual tile comes in and the result goes out by bulk copies through shared-memory units; no thread
//                    touches global memory.
//
// CTA PAIRS (plane_x_kernel<false, true>, plane_xs_kernel<true>; cta_group::2): with an even number of work units two CTAs on one
// TPC run as a pair.  Each keeps its own tile(s), loaders, epilogue and TMEM half, but only HALF of every weight unit (output columns
// [rank N/2, (rank + 1) N/2)); the leader's issuing thread multiplies both tiles with one M = 256 instruction per K step.  The peer's
// issuing thread forwards "my operand landed" to twin barriers in the leader (plain remote mbarrier.arrive -- a cluster-scope
// release costs ~900 cycles per arrive), tcgen05.commit multicasts "slot free" / "accumulator full" back to both CTAs, and both
// CTAs' epilogue warps arrive on the leader's "accumulator empty".  What it buys: half the weight stream into and half the B-operand
// reads out of each shared memory, and room for a fourth staging unit in plane_xs_kernel.
//
// Epilogues write the next layer's plane image directly (bias, activation, residual add from the residual's planes,
// hi/lo split, sub-pixel shuffle, stride-2 de-interleave are all index arithmetic on the way out).
#include <stdlib.h>

#include <type_traits>

#include "plane.cuh"
#include "tc_common.cuh"

namespace nsc {

using namespace tc;

namespace {

// ------------------------------------------------------------------------------------------------
// plane image addressing
// ------------------------------------------------------------------------------------------------
// channels [8g, 8g + 8) of position `pos` of the frame image `img`
__device__ __forceinline__ void pt_store8(const PlaneTensor& t, uint8_t* img, int pos, int g, const float (&v)[8]) {
  uint4 hi, lo;
  split8(v, hi, lo);
  const int sub = t.deint ? (pos & 1) : 0;
  const int row = (t.deint ? (pos >> 1) : pos) + 8;
  const uint32_t sw = (uint32_t)row & 7u;
  const int64_t sb = pt_slab_bytes(t);
  if (t.packed) {
    uint8_t* r = img + sub * sb + (int64_t)row * 128;
    if (t.planes == 2) {   // K-concatenated row (plane.cuh)
      if (g < 2) {
        *reinterpret_cast<uint4*>(r + (((uint32_t)g ^ sw) << 4)) = hi;
        *reinterpret_cast<uint4*>(r + (((uint32_t)(g + 2) ^ sw) << 4)) = lo;
        *reinterpret_cast<uint4*>(r + (((uint32_t)(g + 4) ^ sw) << 4)) = hi;
      } else if (g == 2) {
        *reinterpret_cast<uint4*>(r + ((6u ^ sw) << 4)) = make_uint4(hi.x, hi.y, lo.x, lo.y);
        *reinterpret_cast<uint4*>(r + ((7u ^ sw) << 4)) = make_uint4(hi.x, hi.y, 0u, 0u);
      }                    // g == 3: channels 24-31 are not represented (C <= 20)
    } else {
      *reinterpret_cast<uint4*>(r + (((uint32_t)g ^ sw) << 4)) = hi;
    }
  } else {
    uint8_t* r = img + (int64_t)pt_slab_index(t, sub, 0, g >> 3) * sb + (int64_t)row * 128 + ((((uint32_t)g & 7u) ^ sw) << 4);
    *reinterpret_cast<uint4*>(r) = hi;
    if (t.planes == 2) *reinterpret_cast<uint4*>(r + (int64_t)t.spp * sb) = lo;
  }
}

__device__ __forceinline__ void pt_load8(const PlaneTensor& t, const uint8_t* img, int pos, int g, float (&v)[8]) {
  const int sub = t.deint ? (pos & 1) : 0;
  const int row = (t.deint ? (pos >> 1) : pos) + 8;
  const uint32_t sw = (uint32_t)row & 7u;
  const int64_t sb = pt_slab_bytes(t);
  uint4 hi, lo = make_uint4(0, 0, 0, 0);
  if (t.packed) {
    const uint8_t* r = img + sub * sb + (int64_t)row * 128;
    if (t.planes == 2) {
      if (g < 2) {
        hi = __ldg(reinterpret_cast<const uint4*>(r + (((uint32_t)g ^ sw) << 4)));
        lo = __ldg(reinterpret_cast<const uint4*>(r + (((uint32_t)(g + 2) ^ sw) << 4)));
      } else if (g == 2) {
        const uint4 c6 = __ldg(reinterpret_cast<const uint4*>(r + ((6u ^ sw) << 4)));
        hi = make_uint4(c6.x, c6.y, 0u, 0u);
        lo = make_uint4(c6.z, c6.w, 0u, 0u);
      } else {
        hi = make_uint4(0, 0, 0, 0);
      }
    } else {
      hi = __ldg(reinterpret_cast<const uint4*>(r + (((uint32_t)g ^ sw) << 4)));
    }
  } else {
    const uint8_t* r = img + (int64_t)pt_slab_index(t, sub, 0, g >> 3) * sb + (int64_t)row * 128 + ((((uint32_t)g & 7u) ^ sw) << 4);
    hi = __ldg(reinterpret_cast<const uint4*>(r));
    if (t.planes == 2) lo = __ldg(reinterpret_cast<const uint4*>(r + (int64_t)t.spp * sb));
  }
  unpack8(hi, v);
  if (t.planes == 2) {
    float l[8];
    unpack8(lo, l);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += l[i];
  }
}

__host__ __device__ inline int pt_chunks_per_row(const PlaneTensor& t) { return t.packed ? 4 : t.spp * 8; }
// ALGORITHMIC bytes of an activation tensor of C real channels over `positions` positions: 2 bytes per channel and plane (the
// padding channels and zero rows of the image are not payload)
inline double pt_real_bytes(int positions, int C, int planes) { return (double)positions * C * 2.0 * planes; }

// ------------------------------------------------------------------------------------------------
// fp32 <-> planes
// ------------------------------------------------------------------------------------------------
__global__ void plane_from_f32_kernel(const float* __restrict__ x, int x_cl, int64_t B, int L, int C, PlaneTensor t) {
  const int nch = pt_chunks_per_row(t);
  const int64_t total = B * (int64_t)nch * L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pos = (int)(i % L);
    const int g = (int)((i / L) % nch);
    const int64_t b = i / ((int64_t)L * nch);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = 8 * g + e;
      v[e] = c < C ? (x_cl ? x[(b * L + pos) * C + c] : x[(b * C + c) * L + pos]) : 0.f;
    }
    pt_store8(t, t.base + b * t.frame_bytes, pos, g, v);
  }
}

__global__ void plane_to_f32_kernel(PlaneTensor t, float* __restrict__ y, int y_cl, int64_t B, int L, int C) {
  const int nch = (C + 7) / 8;
  const int64_t total = B * (int64_t)nch * L;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int pos = (int)(i % L);
    const int g = (int)((i / L) % nch);
    const int64_t b = i / ((int64_t)L * nch);
    float v[8];
    pt_load8(t, t.base + b * t.frame_bytes, pos, g, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = 8 * g + e;
      if (c < C) {
        if (y_cl) y[(b * L + pos) * C + c] = v[e];
        else y[(b * C + c) * L + pos] = v[e];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// depthwise FIR over the rows of a wide image (Keras SeparableConv1D's depthwise half, nscm.py:175-177): one thread per
// (frame, block of 8 positions, 8-channel chunk) slides over 8 + K - 1 rows, so every row is read twice instead of K times.
// SAME padding comes from the image's zero rows (K <= 17).  fp32 arithmetic on hi + lo, result split back into planes.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plane_depthwise_kernel(PlaneTensor in, PlaneTensor out, const float* __restrict__ w, int K, int C, int L, int64_t B) {
  __shared__ float s_w[17 * 128];
  const int nch = in.spp * 8;
  for (int i = threadIdx.x; i < K * nch * 8; i += blockDim.x) {
    const int t = i / (nch * 8), c = i - t * (nch * 8);
    s_w[i] = c < C ? __ldg(w + (int64_t)t * C + c) : 0.f;
  }
  __syncthreads();
  const int padL = (K - 1) / 2, nblk = L / 8;
  const int64_t total = B * (int64_t)nblk * nch;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % nch);
    const int pos0 = (int)((i / nch) % nblk) * 8;
    const int64_t b = i / ((int64_t)nch * nblk);
    const uint8_t* img = in.base + b * in.frame_bytes;
    float acc[8][8];
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[o][e] = 0.f;
    for (int r = 0; r < 8 + K - 1; ++r) {
      float v[8];
      pt_load8(in, img, pos0 - padL + r, g, v);        // rows -8 .. L + 7 exist (zero rows)
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int t = r - o;
        if (t >= 0 && t < K) {
          const float* wt = s_w + (t * nch + g) * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[o][e] = fmaf(v[e], wt[e], acc[o][e]);
        }
      }
    }
    uint8_t* oimg = out.base + b * out.frame_bytes;
#pragma unroll
    for (int o = 0; o < 8; ++o) pt_store8(out, oimg, pos0 + o, g, acc[o]);
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 (K, Cin, Cout) -> fp16 hi/lo operand slabs in the exact shared-memory image
// ------------------------------------------------------------------------------------------------
// Grouped taps-in-N (narrow inputs, k = 9): tap groups are row shifts of the A descriptor, the shifts inside a group cross TMEM lanes
// as "tap slots".  Outer groups shift one way only, which keeps frame borders exact (see the file header).
//   3 groups x 3 taps, 5 slots (-2d .. +2d):  shifts {-2d, 0, +2d};            group g, slot s in [g, g + 2]  -> tap s + 2 g
//   5 groups {0,1} {2,3} {4} {5,6} {7,8}, 3 slots (-d, 0, +d):  shifts {-3d, -d, 0, +d, +3d}
__host__ __device__ constexpr int tgroup_shift(int groups, int g) {      // in units of the dilation
  return groups == 3 ? 2 * (g - 1) : (g == 0 ? -3 : g == 1 ? -1 : g == 2 ? 0 : g == 3 ? 1 : 3);
}
__host__ __device__ constexpr int tgroup_tap(int groups, int g, int slot) {   // -1: the slot is unused by this group
  if (groups == 3) return (slot >= g && slot <= g + 2) ? slot + 2 * g : -1;
  if (g < 2) return slot <= 1 ? 2 * g + slot : -1;
  if (g == 2) return slot == 1 ? 4 : -1;
  return slot >= 1 ? 2 * g - 2 + slot : -1;      // g = 3: slots 1, 2 -> taps 5, 6;  g = 4: -> taps 7, 8
}

// weights of the folded narrow conv (plane.cuh): w (9, 20, 20), bias (20) -> out (5, 48, 48) followed by the 48 biases
// (Kf = 5: dilation 1, tap 2 s + ph' - ph; Kf = 9: dilation 2 on the same image -- a position only meets its own phase, tap s)
__global__ void fold2_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int Kf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nW = Kf * kFoldC * kFoldC;
  if (i < nW) {
    const int s = i / (kFoldC * kFoldC), rem = i - s * kFoldC * kFoldC;
    const int a = rem / kFoldC, b = rem - a * kFoldC;          // input channel 24 ph' + ci, output channel 24 ph + co
    const int phi = a / 24, ci = a - 24 * phi, pho = b / 24, co = b - 24 * pho;
    const int t = Kf == 9 ? (phi == pho ? s : -1) : 2 * s + phi - pho;
    out[i] = (ci < 20 && co < 20 && t >= 0 && t <= 8) ? w[(t * 20 + ci) * 20 + co] : 0.f;
  } else if (i < nW + kFoldC) {
    const int b = i - nW, co = b % 24;
    out[i] = (co < 20 && bias != nullptr) ? bias[co] : 0.f;
  }
}

struct PackArgs {
  const float* w;
  __half* out;
  int kind, K, Cin, Cout, planes, in_packed, in_spp;
  int rows;     // rows per unit (MMA N)
  int C;        // PK_T: output channels per tap
  int n_units;
  int groups;   // PK_T grouped form (3 or 5 groups): unit = tap group g, row = (slot, co)
  const float* w2;   // gated linear unit: output columns [csplit, Cout) come from this second (K, Cin, Cout - csplit) kernel
  int csplit;
};

__device__ __forceinline__ uint32_t sw128_off(int row, int k) {
  return (uint32_t)row * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
}

// K slot k (0..63) of a weight row that meets a packed input row: which input channel and which weight plane it holds
// (planes = 2: the K-concatenated layout of plane.cuh; planes = 1: channels 0..31 of the hi plane)
__device__ __forceinline__ void kcat_slot(int planes, int k, int* ci, int* lo_plane) {
  if (planes != 2) { *ci = k & 31; *lo_plane = k >> 5; return; }
  const int c = k >> 3, e = k & 7;
  if (c < 6) { *ci = (c & 1) * 8 + e; *lo_plane = c >= 4 ? 1 : 0; }
  else if (c == 6) { *ci = 16 + (e & 3); *lo_plane = 0; }
  else { *ci = e < 4 ? 16 + e : -1; *lo_plane = 1; }
}

__global__ void plane_pack_kernel(PackArgs a) {
  const int64_t total = (int64_t)a.n_units * a.rows * 64;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i & 63);
    const int n = (int)((i >> 6) % a.rows);
    const int u = (int)(i / (64LL * a.rows));
    int t = 0, ci = -1, co = -1, lo_plane = 0;
    if (a.kind == PK_T) {
      // units = weight slabs: unpacked input -> (plane, slab) ; packed input -> one slab [hi | lo]
      t = n / a.C;
      co = n - t * a.C;
      if (a.groups > 1) {                        // n = (slot, co); unit = group
        const int tt = t < (a.groups == 3 ? 5 : 3) ? tgroup_tap(a.groups, u, t) : -1;
        t = tt < 0 ? a.K : tt;
      }
      if (t >= a.K) co = -1;
      if (a.in_packed) kcat_slot(a.planes, k, &ci, &lo_plane);
      else { ci = (u % a.in_spp) * 64 + k; lo_plane = u / a.in_spp; }
    } else if (a.kind == PK_X) {
      co = n;
      if (a.in_packed) { t = u; kcat_slot(a.planes, k, &ci, &lo_plane); }
      else {
        // u = ((slab * K) + tap) * planes + plane
        lo_plane = u % a.planes;
        t = (u / a.planes) % a.K;
        ci = (u / a.planes / a.K) * 64 + k;
      }
    } else {   // PK_GEN: taps are the K dimension, unit = plane
      co = n; t = k; ci = 0; lo_plane = u;
    }
    float v = 0.f;
    if (co >= 0 && co < a.Cout && ci >= 0 && ci < a.Cin && t < a.K && lo_plane < a.planes) {
      if (a.w2 == nullptr) v = a.w[((int64_t)t * a.Cin + ci) * a.Cout + co];
      else if (co < a.csplit) v = a.w[((int64_t)t * a.Cin + ci) * a.csplit + co];
      else v = a.w2[((int64_t)t * a.Cin + ci) * (a.Cout - a.csplit) + (co - a.csplit)];
    }
    const __half hi = __float2half_rn(v);
    const __half val = lo_plane == 0 ? hi : __float2half_rn(v - __half2float(hi));
    *reinterpret_cast<__half*>(reinterpret_cast<char*>(a.out) + (size_t)u * a.rows * 128 + sw128_off(n, k)) = val;
  }
}

// ================================================================================================
// shared pieces of the two kernels
// ================================================================================================
// activation as a slope: leaky_relu(v) = max(v, 0.2 v), identity = max(v, 1.0 v); tanh takes the slow uniform branch
__device__ __forceinline__ float act_slope_of(int act) { return act == NSC_ACT_LRELU ? kLeakySlope : 1.0f; }
__device__ __forceinline__ float act_fast(float v, float slope) { return fmaxf(v, v * slope); }
__device__ __noinline__ float act_slow(float v, int act) { return apply_act(v, act); }

// K steps of one (A tile, B unit) pair from the issuing thread; descriptors advance 32 bytes (2 units of 16 B) per step
// MODE: 0 = one CTA; 1 = leader of a CTA pair (cta_group::2 instructions, M = 256); 2 = the leader's peer, which issues nothing
// (its issuing thread only forwards "operand landed" to the leader, see XIssue)
template <int NKS, int MODE = 0>
__device__ __forceinline__ void issue_ks(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum0) {
  if constexpr (MODE == 2) return;
#pragma unroll
  for (int ks = 0; ks < NKS; ++ks) {
    if constexpr (MODE == 1)
      mma_f16_ss_cg2(d, desc_from_lo(a_lo + 2u * (uint32_t)ks), desc_from_lo(b_lo + 2u * (uint32_t)ks), idesc, ks == 0 ? accum0 : 1u);
    else
      mma_f16_ss(d, desc_from_lo(a_lo + 2u * (uint32_t)ks), desc_from_lo(b_lo + 2u * (uint32_t)ks), idesc, ks == 0 ? accum0 : 1u);
  }
}
template <int MODE>
__device__ __forceinline__ void commit_mode(uint64_t* bar) {
  if constexpr (MODE == 0) umma_commit(bar);
  else if constexpr (MODE == 1) umma_commit_cg2(bar);
}
__device__ __forceinline__ void issue_n(int nks, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum0) {
  switch (nks) {
    case 4: issue_ks<4>(d, a_lo, b_lo, idesc, accum0); break;
    case 3: issue_ks<3>(d, a_lo, b_lo, idesc, accum0); break;
    case 2: issue_ks<2>(d, a_lo, b_lo, idesc, accum0); break;
    default: issue_ks<1>(d, a_lo, b_lo, idesc, accum0); break;
  }
}

__device__ __forceinline__ void issue_n_cg2(int nks, uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum0) {
  switch (nks) {
    case 4: issue_ks<4, 1>(d, a_lo, b_lo, idesc, accum0); break;
    case 3: issue_ks<3, 1>(d, a_lo, b_lo, idesc, accum0); break;
    case 2: issue_ks<2, 1>(d, a_lo, b_lo, idesc, accum0); break;
    default: issue_ks<1, 1>(d, a_lo, b_lo, idesc, accum0); break;
  }
}

// ------------------------------------------------------------------------------------------------
// frame flags of the fused block kernel: CTAs of different roles hand frames to each other through global-memory rings.
// A producer's bulk stores are complete (cp.async.bulk.wait_group) before it publishes; a consumer orders its bulk loads after
// the acquire with a proxy fence.  A flag that stays unset for seconds means a broken launch: trap instead of hanging the GPU.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long flag_wait(const uint32_t* p, uint32_t target) {
  long long waited = 0;
  if (ld_acquire_gpu(p) < target) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(p) < target) {
      __nanosleep(64);
      if (clock64() - t0 > (1ll << 33)) __trap();
    }
    waited = clock64() - t0;
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  return waited;
}
// optional per-CTA counters of the fused block kernel (NSC_BLOCK_STATS=1, nsc_debug_block_stats): 8 words per CTA --
// role, epilogue loop cycles, cycles the storer waited for a free ring slot, cycles the loader waited for a ready frame,
// work units, cycles the storer waited for its own stores, cycles the MMA issuer waited for its input stage, ... for a free accumulator
constexpr int kStatWords = 8, kStatCtas = 160;
__device__ unsigned long long g_block_stats[kStatCtas * kStatWords];
__device__ __forceinline__ void stat_add(unsigned long long* stats, int word, long long v) {
  if (stats != nullptr && blockIdx.x < kStatCtas) atomicAdd(stats + blockIdx.x * kStatWords + word, (unsigned long long)v);
}
__device__ __forceinline__ void flag_signal(uint32_t* p) {
  asm volatile("fence.proxy.async;" ::: "memory");
  __threadfence();
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(1u) : "memory");
}
// "I have read my part of the frame out of the ring": nothing of this thread's has to become visible, the bump only has to stay
// behind the mbarrier wait (an acquire) that observed the bulk loads' completion -- no fence on the issuing thread's path
__device__ __forceinline__ void flag_bump(uint32_t* p) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(1u) : "memory");
}

// ================================================================================================
// PK_T
// ================================================================================================
struct TParams {
  PlaneTensor in, out;
  const uint8_t* wpack;
  const float* bias;
  float* yvec;
  int N, n_wslab, ksteps;
  int stage_bytes;           // one input slab of one tile: 128 rows, + 8 halo rows each side in the grouped form
  int L, dil, act, na;
  int64_t B;
  // role-relative CTA numbering (a plain launch: cta0 = 0, ncta = gridDim.x) and the frame flags of ring tensors (fused block)
  int cta0, ncta;
  uint32_t *in_ready, *in_free, *out_ready, *out_free;
  int out_free_target;       // arrivals that free a frame slot of `out`: its consumer's CTAs per frame
  unsigned long long* stats; // optional counters (fused block kernel)
  HeadFold fold;             // k55 heads: hard quantiser / cascade accumulation in the epilogue (plane.cuh)
  int fold_out;              // 20-channel output written as the FOLDED image (plane.cuh): 0 no, 1 pairs of positions, 2 pairs within each parity
};

// spill slots: the quarters of the current and the previous tile; the k55 head has no per-tile barrier after its reads (no staged
// output), so it keeps a third tile's worth.  (Eight slots instead of twelve give the 100 -> 20 layer a fifth input stage: the
// layer is bound by bytes in flight -- 4 x 16 KB per SM against ~2 us of loaded HBM latency is 32 GB/s per SM, 73 % of peak.)
constexpr int t_slots(int C) { return C == 1 ? 12 : 8; }
constexpr int kORing = 256;         // rows of the output staging ring (two tiles)

// Epilogue organisation: C = 20 -> three warps per TMEM lane quarter, one 8-channel chunk of the output row each
// (channels 0-7, 8-15, 16-19 + zero padding); C = 1 (k55 head) -> one warp per quarter.
// TAPS = tap slots the epilogue sums across lanes (9; 5 in the grouped form; 55 for the head), GROUPS = MMA groups by row shift
template <int C, int TAPS>
struct TShape {
  static constexpr int kGroups = (C == 1) ? 1 : 3;
  static constexpr int kEpiWarps = 4 * kGroups;
  static constexpr int kThreads = (kEpiWarps + 2) * 32;   // + MMA issuer, loader
  static constexpr int kMaxM = (C == 1) ? 32 : 8;         // largest row shift handled (lanes that can wrap)
  static constexpr int kRowFloats = (C == 1) ? 1 : 24;    // floats per spilled row
  static constexpr int kSlots = t_slots(C);
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}

// ---- phase 1: shifted sums of NCH channels of one 32-row quarter ------------------------------------
// tcol: TMEM address of this warp's first column of tap 0; taps are `tap_stride` columns apart.
template <int NCH, int TAPS>
__device__ __forceinline__ void t_phase1(uint32_t tcol, const int (&srcl)[TAPS], uint32_t inrm, float (&acc)[NCH], float (&up)[NCH], float (&down)[NCH]) {
  constexpr int kMid = TAPS / 2;
  // all tap slots are fetched before the single wait: the independent shuffles then pipeline back to back
  uint32_t r[TAPS][8];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    if constexpr (NCH == 8) tmem_ld8(tcol + (uint32_t)(t * 20), r[t]);
    else {
      uint32_t b[4];
      tmem_ld4(tcol + (uint32_t)(t * 20), b);
#pragma unroll
      for (int i = 0; i < 4; ++i) r[t][i] = b[i];
    }
  }
  tmem_ld_wait();
  // packed fp32x2 adds (sm_100 FADD2): the epilogue is bound by ALU issue (3-register fp32 ops issue every other cycle), and two
  // channels per instruction halve the add count
  float2 a2[NCH / 2], u2[NCH / 2], d2[NCH / 2];
#pragma unroll
  for (int c = 0; c < NCH / 2; ++c) {
    a2[c] = make_float2(__uint_as_float(r[kMid][2 * c]), __uint_as_float(r[kMid][2 * c + 1]));
    u2[c] = make_float2(0.f, 0.f);
    d2[c] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int t = 0; t < TAPS; ++t) {
    if (t == kMid) continue;
    const int src = srcl[t];
    const bool inr = (inrm >> t) & 1u;
#pragma unroll
    for (int c = 0; c < NCH / 2; ++c) {
      float2 x;
      x.x = __shfl_sync(0xffffffffu, __uint_as_float(r[t][2 * c]), src);
      x.y = __shfl_sync(0xffffffffu, __uint_as_float(r[t][2 * c + 1]), src);
      if (inr) a2[c] = __fadd2_rn(a2[c], x);
      else if (t > kMid) d2[c] = __fadd2_rn(d2[c], x);   // source row is in this quarter, target row in the previous one
      else u2[c] = __fadd2_rn(u2[c], x);                 // ... in the next one
    }
  }
#pragma unroll
  for (int c = 0; c < NCH / 2; ++c) {
    acc[2 * c] = a2[c].x; acc[2 * c + 1] = a2[c].y;
    up[2 * c] = u2[c].x; up[2 * c + 1] = u2[c].y;
    down[2 * c] = d2[c].x; down[2 * c + 1] = d2[c].y;
  }
}

__device__ __forceinline__ void t_phase1_k55(uint32_t tcol, int lane, float& acc, float& up, float& down) {
  uint32_t r0[32], r1[32];
  tmem_ld32(tcol, r0);
  tmem_ld32(tcol + 32, r1);
  tmem_ld_wait();
  float a = 0.f, u = 0.f, d = 0.f;
#pragma unroll
  for (int t = 0; t < 55; ++t) {
    const float v = __uint_as_float(t < 32 ? r0[t & 31] : r1[t & 31]);
    const int s = t - 27;
    if (s == 0) { a += v; continue; }
    const float x = __shfl_sync(0xffffffffu, v, (lane + s) & 31);
    const bool inr = (unsigned)(lane + s) < 32u;
    if (inr) a += x;
    else if (s > 0) d += x;
    else u += x;
  }
  acc = a; up = u; down = d;
}

// kPair (wide-input C = 20 layers, an even number of frames): two CTAs of a cluster run as a PAIR (cta_group::2).  Each walks its own
// frames in lockstep with the other and keeps its own tiles, epilogue and TMEM half, but only HALF of every weight slab (output
// columns [rank N/2, (rank + 1) N/2)); the leader's issuing thread multiplies both CTAs' tiles with one M = 256 instruction, the
// peer's forwards "my slab has landed".  What it buys: 49 KB of shared memory, i.e. eight input stages in flight instead of five
// for the layer that is bound by exactly that (100 -> 20).
template <int C, int TAPS, int GROUPS, bool kPair = false>
__device__ __forceinline__ void t_body(const TParams& p) {
  using S = TShape<C, TAPS>;
  const int rank = (int)blockIdx.x - p.cta0, nranks = p.ncta;
  constexpr int kTE = (C == 1) ? 1 : TAPS;              // tap slots of the shuffle tap sum (the head has its own phase 1)
  const uint32_t kAStage = (uint32_t)p.stage_bytes;
  constexpr int kMaxM = S::kMaxM, kRF = S::kRowFloats, kEpi = S::kEpiWarps;
  constexpr int NCHMAX = (C == 1) ? 1 : 8;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t w_full, a_full[8], a_empty[8], acc_full[2], acc_empty[2];
  __shared__ uint64_t w_full2, a_full2[8];           // pair: the peer's operands have landed (live in the leader)
  __shared__ uint64_t so_ready[2], so_free[2];      // ring mode: output windows handed to / returned by the storer warp (by tile parity)
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_qbins[C == 1 ? 256 : 1];       // code head: the quantiser's bins

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if constexpr (C == 1) {
    if (p.fold.q_bins != nullptr)
      for (int k = tid; k < p.fold.q_n; k += blockDim.x) s_qbins[k] = __ldg(p.fold.q_bins + k);
  }
  const bool ring_out = (C != 1) && p.out_ready != nullptr;   // fused block: a dedicated warp stores and publishes (the CTA has spare warps)
  const uint32_t crank = kPair ? cluster_ctarank() : 0u;
  const uint32_t wslab_bytes = (uint32_t)p.N * (kPair ? 64u : 128u);      // bytes of one weight slab in THIS CTA
  uint8_t* sW = smem;
  uint8_t* sA = sW + (uint32_t)p.n_wslab * wslab_bytes;
  constexpr int kTSlots = S::kSlots;
  float* sU = reinterpret_cast<float*>(sA + (uint32_t)p.na * kAStage);     // [kTSlots][kMaxM][kRF]
  float* sP = sU + kTSlots * kMaxM * kRF;                                   // [kTSlots][kMaxM][kRF]
  // output staging ring (C = 20): kORing image rows of 128 B, same swizzle as the global image, bulk-stored one window per tile
  uint8_t* sO = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sP + kTSlots * kMaxM * kRF) + 1023) & ~(uintptr_t)1023);
  const int T = p.L / 128;
  const int nst = pt_n_slabs(p.in);
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2u * (uint32_t)p.N) tmem_cols *= 2;      // two accumulator slots: the next tile's MMAs run under this tile's epilogue

  if (tid == 0) {
    mbar_init(&w_full, 1);
    mbar_init(&w_full2, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_full2[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kPair ? 2 * kEpi : kEpi); }
    for (int i = 0; i < 2; ++i) { mbar_init(&so_ready[i], kEpi); mbar_init(&so_free[i], 1); }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == kEpi) {
    if constexpr (kPair) tmem_alloc_cg2(&tmem_base_s, tmem_cols);
    else tmem_alloc(&tmem_base_s, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers (and TMEM) exist before anything arrives remotely
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();      // barriers, TMEM and the bias table are set up under the previous kernel's tail; its output is only touched from here

  if (warp < kEpi) {
    // =========================== epilogue ===========================
    const int q = warp & 3, grp = warp >> 2;          // TMEM lane quarter; 8-channel chunk of the output row
    const int nch = (C == 1) ? 1 : (grp == 2 ? 4 : 8);
    float bias[NCHMAX];
#pragma unroll
    for (int c = 0; c < NCHMAX; ++c) bias[c] = (p.bias && c < nch) ? __ldg(p.bias + grp * 8 + c) : 0.f;
    const float slope = act_slope_of(p.act);
    const int m = ((TAPS - 1) / 2) * p.dil;
    const int nq = 4 * T;
    const bool low = lane < m, high = lane >= 32 - m;
    // everything that depends only on the lane is computed once: shuffle sources and the in-range mask of the nine taps,
    // this thread's rows inside a spill slot, its swizzle phase
    int srcl[kTE];
    uint32_t inrm = 0;
#pragma unroll
    for (int t = 0; t < kTE; ++t) {
      const int s = (t - kTE / 2) * p.dil;
      srcl[t] = (lane + s) & 31;
      if ((unsigned)(lane + s) < 32u) inrm |= 1u << t;
    }
    constexpr int kSlotF = kMaxM * kRF;
    const int coff = (C == 1) ? 0 : grp * 8;
    float* const sU_l = sU + lane * kRF + coff;
    float* const sP_l = sP + (lane - (32 - m)) * kRF + coff;
    const int out_planes = p.out.planes;
    const int fold_out = (C == 1) ? 0 : p.fold_out;
    const int fsh = fold_out == 2 ? 2 : 1;             // positions per folded row of one slab, log2
    const int fring = kORing >> fsh;                    // ring rows per (parity, plane) region of the folded staging buffer
    uint32_t it = 0, tb = 0, tbp = kTSlots - 4;       // tb: first spill slot of this tile's quarters (0, 4[, 8] in turn)
    const long long t_begin = clock64();
    for (int64_t f = rank; f < p.B; f += nranks) {
      uint8_t* const orow = (C == 1) ? nullptr : p.out.base + pt_frame_off(p.out, f) + 8 * 128;   // position 0 of the packed image
      const int ring0 = (int)((it * 128u) & (kORing - 1));      // ring row of this frame's position 0 (frames are whole tiles)
      int done_rows = 0;                                        // positions of this frame already handed to the copy engine
      float* const yrow = (C == 1 && p.yvec != nullptr) ? p.yvec + f * p.L : nullptr;
      float pre_own = 0.f, pre_prev = 0.f;                      // output head: running sums of the rows this thread may finish
      int pre_row = -1;
      auto store_row = [&](int row, const float (&r)[NCHMAX]) {
        if constexpr (C == 1) {
          const float v = apply_act(r[0] + bias[0], p.act);
          if (yrow) yrow[row] = v;
          const int64_t gi = f * p.L + row;
          if (p.fold.q_bins != nullptr) {
            // hard scalar quantiser (nn_core_operator.py:140-164), the arithmetic of quantize_kernel to the bit: fp32 distance, a
            // separate fp32 multiply by alpha, arg-max with the lowest index on ties (ascending scan, strictly greater)
            const float alpha = __ldg(p.fold.q_alpha);
            float best = -INFINITY;
            int besti = -1;
            int k = 0;
            for (; k + 4 <= p.fold.q_n; k += 4) {       // four independent logits per step; the selects keep the lowest index
              const float l0 = __fmul_rn(alpha, fabsf(__fsub_rn(v, s_qbins[k])));
              const float l1 = __fmul_rn(alpha, fabsf(__fsub_rn(v, s_qbins[k + 1])));
              const float l2 = __fmul_rn(alpha, fabsf(__fsub_rn(v, s_qbins[k + 2])));
              const float l3 = __fmul_rn(alpha, fabsf(__fsub_rn(v, s_qbins[k + 3])));
              float m = l0;
              int mi = k;
              if (l1 > m) { m = l1; mi = k + 1; }
              if (l2 > m) { m = l2; mi = k + 2; }
              if (l3 > m) { m = l3; mi = k + 3; }
              if (besti < 0 || m > best) { best = m; besti = mi; }
            }
            for (; k < p.fold.q_n; ++k) {
              const float lg = __fmul_rn(alpha, fabsf(__fsub_rn(v, s_qbins[k])));
              if (besti < 0 || lg > best) { best = lg; besti = k; }
            }
            if (besti < 0) besti = 0;
            if (p.fold.q_idx) p.fold.q_idx[gi] = (uint8_t)besti;
            if (p.fold.q_code) p.fold.q_code[gi] = __fadd_rn(__fmul_rn(1.f - p.fold.q_iq, v), __fmul_rn(p.fold.q_iq, s_qbins[besti]));
          }
          if (p.fold.acc != nullptr) {       // decoded (+)= out / res_scalar (cmrl.py:522-531, :822-830), the arithmetic of accum_div_kernel
            const float q = v / p.fold.div;
            // (the running sum of this thread's two candidate rows was fetched before the accumulator wait)
            p.fold.acc[gi] = p.fold.acc_first ? q : (row == pre_row ? pre_own : pre_prev) + q;
            if (p.fold.quot) p.fold.quot[gi] = q;
          }
        } else {
          float v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) v[c] = c < nch ? act_fast(r[c] + bias[c], slope) : 0.f;
          uint4 hi, lo;
          split8(v, hi, lo);
          // A thread owns a ROW, so direct global stores would touch 32 different 128-byte lines per warp instruction (measured:
          // 1,600 of the 2,900 cycles of a tile).  The row goes into a shared-memory ring that mirrors the image; one thread
          // bulk-stores a whole window of finished rows per tile.
          if (fold_out) {
            // folded image (plane.cuh): `fsh` consecutive positions (of one parity for fold_out = 2) share a row of 24-channel groups;
            // hi and lo planes (and the parities) are separate slabs, each with its own ring of kORing >> fsh rows
            const int sp = ring0 + row;
            const int frow = (sp >> fsh) & (fring - 1);
            const int par = fold_out == 2 ? (row & 1) : 0;
            const int ph = (row >> (fsh - 1)) & 1;
            const uint32_t fsw = (uint32_t)(row >> fsh) & 7u;
            uint8_t* rp = sO + (uint32_t)((par * 2) * fring + frow) * 128u + ((((uint32_t)(3 * ph + grp)) ^ fsw) << 4);
            *reinterpret_cast<uint4*>(rp) = hi;
            if (out_planes == 2) *reinterpret_cast<uint4*>(rp + (uint32_t)fring * 128u) = lo;
            return;
          }
          uint8_t* rp = sO + (uint32_t)((ring0 + row) & (kORing - 1)) * 128u;
          const uint32_t sw = (uint32_t)row & 7u;       // (row + 8) & 7
          if (out_planes == 2) {                        // K-concatenated row (plane.cuh)
            if (grp < 2) {
              *reinterpret_cast<uint4*>(rp + (((uint32_t)grp ^ sw) << 4)) = hi;
              *reinterpret_cast<uint4*>(rp + (((uint32_t)(grp + 2) ^ sw) << 4)) = lo;
              *reinterpret_cast<uint4*>(rp + (((uint32_t)(grp + 4) ^ sw) << 4)) = hi;
            } else {
              *reinterpret_cast<uint4*>(rp + ((6u ^ sw) << 4)) = make_uint4(hi.x, hi.y, lo.x, lo.y);
              *reinterpret_cast<uint4*>(rp + ((7u ^ sw) << 4)) = make_uint4(hi.x, hi.y, 0u, 0u);
            }
          } else {
            *reinterpret_cast<uint4*>(rp + (((uint32_t)grp ^ sw) << 4)) = hi;
            if (grp == 2) *reinterpret_cast<uint4*>(rp + ((3u ^ sw) << 4)) = make_uint4(0, 0, 0, 0);   // channels 24-31: K padding the consumer reads
          }
        }
      };
      for (int j = 0; j < T; ++j, ++it) {
        const uint32_t acc_i = it & 1u;
        if constexpr (C == 1) {
          if (p.fold.acc != nullptr && !p.fold.acc_first) {      // a thread finishes its own row or the one 32 above it
            pre_row = (j * 4 + q) * 32 + lane;
            pre_own = p.fold.acc[f * p.L + pre_row];
            pre_prev = pre_row >= 32 ? p.fold.acc[f * p.L + pre_row - 32] : 0.f;
          }
        }
        mbar_wait_relaxed(&acc_full[acc_i], (it >> 1) & 1u);
        tc_fence_after();
        float acc[NCHMAX], up[NCHMAX], down[NCHMAX];
        const uint32_t tcol = tmem + ((uint32_t)(q * 32) << 16) + acc_i * (uint32_t)p.N + (uint32_t)(grp * 8);
        if constexpr (C == 1) {
          t_phase1_k55(tcol, lane, acc[0], up[0], down[0]);
        } else {
          if (grp == 2) {
            float a4[4], u4[4], d4[4];
            t_phase1<4, kTE>(tcol, srcl, inrm, a4, u4, d4);
#pragma unroll
            for (int c = 0; c < 8; ++c) { acc[c] = c < 4 ? a4[c & 3] : 0.f; up[c] = c < 4 ? u4[c & 3] : 0.f; down[c] = c < 4 ? d4[c & 3] : 0.f; }
          } else {
            t_phase1<8, kTE>(tcol, srcl, inrm, acc, up, down);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                 // TMEM slot free: the next tile's MMAs run under phase 2
          if constexpr (kPair) mbar_arrive_remote(&acc_empty[acc_i], 0u);   // the leader's barrier counts both CTAs' epilogue warps
          else mbar_arrive(&acc_empty[acc_i]);
        }
        // ring mode: this tile's rows overwrite the staging window of two tiles ago -- the storer warp has seen it leave shared memory
        if (ring_out) mbar_wait_relaxed(&so_free[it & 1u], ((it >> 1) & 1u) ^ 1u);

        const int fqi = j * 4 + q;
        const bool lastq = fqi == nq - 1;
        const uint32_t slot = tb + (uint32_t)q;
        const uint32_t slot1 = q ? slot - 1u : tbp + 3u;                       // previous quarter (possibly of the previous tile)
        const uint32_t slot2 = q >= 2 ? slot - 2u : tbp + 2u + (uint32_t)q;    // the one before it
        if (low) {
#pragma unroll
          for (int c = 0; c < NCHMAX; ++c) sU_l[slot * kSlotF + c] = up[c];
        }
        if (high && !lastq) {
#pragma unroll
          for (int c = 0; c < NCHMAX; ++c) sP_l[slot * kSlotF + c] = acc[c];
        }
        asm volatile("bar.sync 1, %0;" :: "n"(kEpi * 32) : "memory");
        // Every thread finishes ONE row: its own if nothing is missing from the next quarter, else the parked row of the
        // previous quarter whose down-spill it holds.  The frame's last quarter also finishes its own parked rows.
        // (one code path for both kinds of lane: the row's own part comes from registers, the neighbour's part from shared memory)
        float r[NCHMAX];
        {
          const bool take = high ? (fqi >= 1) : (low && fqi >= 1);
          const float* src = (high ? sP_l : sU_l) + slot1 * kSlotF;
#pragma unroll
          for (int c = 0; c < NCHMAX; ++c) r[c] = (high ? down[c] : acc[c]) + (take ? src[c] : 0.f);
          if (C == 1 && high && low && fqi >= 2) {     // shifts beyond 16 rows (k55): the parked row also took an up-spill
#pragma unroll
            for (int c = 0; c < NCHMAX; ++c) r[c] += sU_l[slot2 * kSlotF + c];
          }
          if (!high || fqi >= 1) store_row(fqi * 32 + lane - (high ? 32 : 0), r);
        }
        if (lastq && high) {
#pragma unroll
          for (int c = 0; c < NCHMAX; ++c) r[c] = acc[c];
          if (C == 1 && low && fqi >= 1) {
#pragma unroll
            for (int c = 0; c < NCHMAX; ++c) r[c] += sU_l[slot1 * kSlotF + c];
          }
          store_row(fqi * 32 + lane, r);
        }
        if constexpr (C != 1) {
          fence_async_smem();                                      // generic-proxy row writes -> visible to the bulk store
          if (ring_out) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&so_ready[it & 1u]);         // the storer warp takes it from here
          } else {
            asm volatile("bar.sync 2, %0;" :: "n"(kEpi * 32) : "memory");
            // rows final after this tile: everything up to 8 rows before its end (those wait for the next tile's down-spill)
            const int upto = (j == T - 1) ? p.L : (j + 1) * 128 - 8;
            if (fold_out) {
              // every (parity, plane) slab of the folded image takes its own window of rows [done >> fsh, upto >> fsh); one issuing
              // thread PER SLAB (lane 0 of warps 0-3): issuing a bulk copy costs the thread ~150 cycles and the issuing thread is on
              // the epilogue's critical path -- eight copies from one thread made the 50 -> 20 layer 30 % slower
              const int nreg = (fold_out == 2 ? 2 : 1) * out_planes;
              if (lane == 0 && warp < nreg) {
                const int64_t sb = pt_slab_bytes(p.out);
                const int par = warp / out_planes, pl_i = warp - par * out_planes;
                uint8_t* const gbase = orow + (int64_t)pt_slab_index(p.out, par, pl_i, 0) * sb;
                const uint8_t* const sbase = sO + (uint32_t)((par * 2 + pl_i) * fring) * 128u;
                int r0 = done_rows >> fsh;
                const int r1 = upto >> fsh;
                while (r0 < r1) {                                    // at most two pieces (ring wrap)
                  const int ring_r = ((ring0 >> fsh) + r0) & (fring - 1);
                  int n = r1 - r0;
                  if (ring_r + n > fring) n = fring - ring_r;
                  bulk_s2g(gbase + (int64_t)r0 * 128, sbase + (uint32_t)ring_r * 128u, (uint32_t)n * 128u);
                  r0 += n;
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // this slab's window before this one has left shared memory
              }
            } else if (warp == 0 && lane == 0) {
              int r0 = done_rows;
              while (r0 < upto) {                                    // at most two pieces (ring wrap)
                const int ring_r = (ring0 + r0) & (kORing - 1);
                int n = upto - r0;
                if (ring_r + n > kORing) n = kORing - ring_r;
                bulk_s2g(orow + (int64_t)r0 * 128, sO + (uint32_t)ring_r * 128u, (uint32_t)n * 128u);
                r0 += n;
              }
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the window before this one has left shared memory
            }
            done_rows = (j == T - 1) ? p.L : (j + 1) * 128 - 8;
          }
        }
        tbp = tb;
        tb = tb == (uint32_t)(kTSlots - 4) ? 0u : tb + 4u;
      }
    }
    if constexpr (C != 1) {
      if (warp < 4 && lane == 0 && !ring_out) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // (folded output: one issuing thread per slab)
    }
    if (warp == 0 && lane == 0 && p.stats != nullptr) {
      stat_add(p.stats, 0, p.in_ready != nullptr ? 2 : 1);
      stat_add(p.stats, 1, clock64() - t_begin);
      stat_add(p.stats, 4, (long long)it);
    }
  } else if (warp == kEpi) {
    // =========================== MMA issuer ===========================
    if (kPair && crank != 0) {
      // the peer of a pair issues nothing: it forwards "landed" of its own weights and input slabs to the leader's twin barriers
      if (elect_one()) {
        mbar_wait(&w_full, 0);
        mbar_arrive_remote(&w_full2, 0u);
        uint32_t slot = 0, sph = 0;
        for (int64_t f = rank; f < p.B; f += nranks)
          for (int j = 0; j < T; ++j)
            for (int s = 0; s < nst; ++s) {
              mbar_wait(&a_full[slot], sph);
              mbar_arrive_remote(&a_full2[slot], 0u);
              if (++slot == (uint32_t)p.na) { slot = 0; sph ^= 1u; }
            }
      }
    } else if (elect_one()) {
      mbar_wait(&w_full, 0);
      if constexpr (kPair) mbar_wait(&w_full2, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16(p.N, kPair ? 256 : 128);
      const uint32_t w_lo0 = desc_lo(smem_u32(sW));
      const uint32_t a_lo0 = desc_lo(smem_u32(sA));
      const uint32_t wslab_lo = wslab_bytes >> 4;
      const int spp = p.in.spp, planes = p.in.planes;
      const bool packed = p.in.packed != 0;
      uint32_t it = 0, slot = 0, sph = 0;
      for (int64_t f = rank; f < p.B; f += nranks) {
        for (int j = 0; j < T; ++j, ++it) {
          const uint32_t acc_i = it & 1u;
          long long tw = p.stats != nullptr ? clock64() : 0;
          mbar_wait(&acc_empty[acc_i], ((it >> 1) & 1u) ^ 1u);
          if (p.stats != nullptr) stat_add(p.stats, 7, clock64() - tw);
          tc_fence_after();
          const uint32_t d = tmem + acc_i * (uint32_t)p.N;
          uint32_t accum = 0;
          for (int s = 0; s < nst; ++s) {
            tw = p.stats != nullptr ? clock64() : 0;
            mbar_wait(&a_full[slot], sph);
            if constexpr (kPair) mbar_wait(&a_full2[slot], sph);
            if (p.stats != nullptr) stat_add(p.stats, 6, clock64() - tw);
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + slot * (uint32_t)(kAStage >> 4);
            if constexpr (GROUPS > 1) {
              // tap groups = row shifts of the staged tile (8 halo rows above)
#pragma unroll
              for (int g = 0; g < GROUPS; ++g) {
                const uint32_t ag = a_lo + (uint32_t)(8 + tgroup_shift(GROUPS, g) * p.dil) * 8u;   // one row = 128 bytes = 8 descriptor units
                const uint32_t bg = w_lo0 + (uint32_t)g * wslab_lo;
                issue_n(p.ksteps, d, ag, bg, idesc, g == 0 ? accum : 1u);   // planes = 2: hi*wh + lo*wh + hi*wl along ONE K axis (4 K steps)
              }
            } else if (packed) {
              issue_n(p.ksteps, d, a_lo, w_lo0, idesc, accum);            // planes = 2: the K-concatenated row, 4 K steps
            } else {
              const int plane = s >= spp ? 1 : 0, sl = s - plane * spp;
              const int nks = min(4, p.ksteps - 4 * sl);
              if constexpr (kPair) {
                issue_n_cg2(nks, d, a_lo, w_lo0 + (uint32_t)sl * wslab_lo, idesc, accum);
                if (plane == 0 && planes == 2) issue_n_cg2(nks, d, a_lo, w_lo0 + (uint32_t)(spp + sl) * wslab_lo, idesc, 1u);
              } else {
              issue_n(nks, d, a_lo, w_lo0 + (uint32_t)sl * wslab_lo, idesc, accum);                                   // (hi | lo) * W_hi
              if (plane == 0 && planes == 2) issue_n(nks, d, a_lo, w_lo0 + (uint32_t)(spp + sl) * wslab_lo, idesc, 1u);   // hi * W_lo
              }
            }
            accum = 1;
            if constexpr (kPair) umma_commit_cg2(&a_empty[slot]);      // "slot free" in both CTAs
            else umma_commit(&a_empty[slot]);
            if (++slot == (uint32_t)p.na) { slot = 0; sph ^= 1u; }
          }
          if (j == T - 1 && p.in_free != nullptr) flag_bump(p.in_free + f);   // every stage of the frame has been read out of the ring
          if constexpr (kPair) umma_commit_cg2(&acc_full[acc_i]);
          else umma_commit(&acc_full[acc_i]);
        }
      }
    }
  } else if (warp == kEpi + 1) {
    // =========================== loader ===========================   (the fused block kernel has more warps than this role uses)
    if (elect_one()) {
      mbar_expect_tx(&w_full, (uint32_t)p.n_wslab * wslab_bytes);
      // (pair: rows [rank N/2, (rank + 1) N/2) of every slab)
      for (int i = 0; i < p.n_wslab; ++i)
        bulk_g2s(sW + (uint32_t)i * wslab_bytes, p.wpack + (size_t)i * ((size_t)p.N * 128u) + (size_t)crank * wslab_bytes, wslab_bytes, &w_full);
      const int64_t sb = pt_slab_bytes(p.in);
      uint32_t slot = 0, sph = 1;
      for (int64_t f = rank; f < p.B; f += nranks) {
        const uint8_t* img = p.in.base + pt_frame_off(p.in, f);
        if (p.in_ready != nullptr) stat_add(p.stats, 3, flag_wait(p.in_ready + f, 1u));
        for (int j = 0; j < T; ++j) {
          for (int s = 0; s < nst; ++s) {
            mbar_wait_relaxed(&a_empty[slot], sph);
            mbar_expect_tx(&a_full[slot], kAStage);
            bulk_g2s(sA + slot * kAStage, img + s * sb + (int64_t)((GROUPS > 1 ? 0 : 8) + 128 * j) * 128, kAStage, &a_full[slot]);
            if (++slot == (uint32_t)p.na) { slot = 0; sph ^= 1u; }
          }
        }
      }
    }
  }
  else if (warp == kEpi + 2) {
    // =========================== storer (ring mode: the fused block kernel) ===========================
    // One thread sends every finished window to the ring in global memory and publishes a frame once its stores have COMPLETED.
    // The completion wait is deferred by one frame: frame f - 1 is published after the last window of frame f has been handed to the
    // copy engine (cp.async.bulk.wait_group T leaves this frame's T groups in flight), so the write latency to L2 (~2-4 us under
    // load, half of this role's loop time when it was waited for in line -- profiles/r02_block_stats3.log) hides behind a whole
    // frame of work.  A CTA never holds an unpublished frame while it BLOCKS on a free slot (that wait would close a cycle through
    // the other roles): the pending frame is published first.  The epilogue warps never wait for a store.
    if constexpr (C != 1) {
      if (ring_out && elect_one()) {
        uint32_t it = 0;
        int64_t pending = -1;                                      // frame whose stores are issued but not yet known complete
        for (int64_t f = rank; f < p.B; f += nranks) {
          uint8_t* const orow = p.out.base + pt_frame_off(p.out, f) + 8 * 128;
          const int ring0 = (int)((it * 128u) & (kORing - 1));
          int done_rows = 0;
          // the slot's previous frame has been consumed before this frame's first rows go out
          if (p.out_free != nullptr && f >= p.out.ring) {
            const uint32_t* fr = p.out_free + (f - p.out.ring);
            if (pending >= 0 && ld_acquire_gpu(fr) < (uint32_t)p.out_free_target) {
              asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
              flag_signal(p.out_ready + pending);
              pending = -1;
            }
            stat_add(p.stats, 2, flag_wait(fr, (uint32_t)p.out_free_target));
          }
          for (int j = 0; j < T; ++j, ++it) {
            mbar_wait_relaxed(&so_ready[it & 1u], (it >> 1) & 1u);
            const int upto = (j == T - 1) ? p.L : (j + 1) * 128 - 8;
            int r0 = done_rows;
            while (r0 < upto) {                                    // at most two pieces (ring wrap)
              const int ring_r = (ring0 + r0) & (kORing - 1);
              int n = upto - r0;
              if (ring_r + n > kORing) n = kORing - ring_r;
              bulk_s2g(orow + (int64_t)r0 * 128, sO + (uint32_t)ring_r * 128u, (uint32_t)n * 128u);
              r0 += n;
            }
            done_rows = upto;
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            const long long t0 = p.stats != nullptr ? clock64() : 0;
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the window has left shared memory
            mbar_arrive(&so_free[it & 1u]);
            if (j == T - 1) {
              if (pending >= 0) {
                // everything but this frame's T groups is complete in global memory -> the previous frame can be published
                if (T == 4) asm volatile("cp.async.bulk.wait_group 4;" ::: "memory");
                else if (T == 2) asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
                else if (T == 1) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                flag_signal(p.out_ready + pending);
              }
              pending = f;
            }
            if (p.stats != nullptr) stat_add(p.stats, 5, clock64() - t0);
          }
        }
        if (pending >= 0) {
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          flag_signal(p.out_ready + pending);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) cluster_sync_all();   // nobody frees TMEM (or retires) while the leader's instruction stream can still touch it
  tc_fence_after();
  if (warp == kEpi) {
    if constexpr (kPair) tmem_dealloc_cg2(tmem, tmem_cols);
    else tmem_dealloc(tmem, tmem_cols);
  }
}

template <int C, int TAPS, int GROUPS, bool kPair = false>
__global__ void __launch_bounds__(TShape<C, TAPS>::kThreads, 1) plane_t_kernel(const __grid_constant__ TParams p) {
  pdl_trigger();
  t_body<C, TAPS, GROUPS, kPair>(p);
}

// ================================================================================================
// PK_X / PK_GEN
// ================================================================================================
struct XParams {
  int kind;
  PlaneTensor in, out, res;
  const float* xvec;
  const float* xsub;
  float xscale;
  const float* resvec;
  const uint8_t* wpack;
  const float* bias;
  const float* bias2;        // gated linear unit: bias of the tanh gate (columns [20, 40))
  int glu, ileave;           // see PlaneConv
  int fold_out, unfold;      // folded narrow images (plane.cuh): the output is written folded / the 48-channel result is unfolded
  int psplit;                // two MMA-issuing threads BY WEIGHT PLANE (0: W_hi units = hi and lo activation planes, 1: W_lo units), one accumulator
                             // each, summed in the epilogue: the stride-2 conv's 189 MMAs per tile are bound by what ONE thread can issue
  int pairtiles;             // unfold + ileave: a CTA takes its tiles in PAIRS (2u, 2u + 1) = the two position parities of codec frame u,
                             // so that the unfolded frame is complete in its staging buffer after the second one
  int Lin, Lout, Cin, Cout, K, dil, stride, padL;
  int act, post_act, res_mode, shuffle, planes;
  int Npad, ksteps, mt, tile, tiles_per_frame;
  int n_stage, kbuf, stage_bytes;
  int halo;                  // rows staged above and below the tile: 8 (the image's own zero rows), or 16 for packed-input layers whose
                             // taps reach further (k15 dilation 2: +-14) -- the 8 rows beyond an image's zero rows are the zero rows of the
                             // neighbouring frame's image (or of the zeroed guard in front of / behind the buffer), see plane_codec.cuh
  int n_units, unit_bytes, wslots, resident;
  int tmem_cols;
  int zero_from, zero_to;    // output chunks the epilogue must clear (K padding the consumer will read)
  int staged;                // plane_xs_kernel: epilogue through shared-memory units and bulk copies
  int n_iss;                 // MMA-issuing threads (1, or one per M tile)
  int pair;                  // CTA pairs (cta_group::2): each CTA stages its own tile and HALF of every weight unit; the leader's
                             // M = 256 instructions cover both tiles (half the weight stream, half the instructions per tile)
  int slot_bytes;            // shared-memory bytes of one weight unit in this CTA (unit_bytes, or half of it in a pair)
  int s_units;               // staged epilogue: residual / output units in the ring (3, or 4 in a pair)
  int64_t B, n_tiles;
  // role-relative CTA numbering (a plain launch: cta0 = 0, ncta = gridDim.x) and the frame flags of a ring input (fused block)
  int cta0, ncta;
  uint32_t *in_ready, *in_free;
  unsigned long long* stats; // optional counters (fused block kernel)
};

// tile of this CTA's i-th iteration (every role of plane_x_kernel walks the same sequence)
__device__ __forceinline__ int64_t x_tile_at(const XParams& p, int64_t i) {
  const int64_t c = (int64_t)blockIdx.x - p.cta0;
  return p.pairtiles ? 2 * (c + (i >> 1) * p.ncta) + (i & 1) : c + i * p.ncta;
}
// unfold: rows of the output staging window -- one 256-position tile (dilation 1) or one frame of two parity tiles (dilation 2)
__host__ __device__ constexpr int x_out_rows(int ileave) { return ileave ? 512 : 256; }

constexpr int kXEpiGroups = 3;                   // epilogue warps per TMEM lane quarter (each takes every third 16-column batch)
constexpr int kXEpiWarps = 4 * kXEpiGroups, kXGenWarps = 4;
constexpr int kXThreadsX = (kXEpiWarps + 4) * 32;                 // + MMA issuer, A loader, W loader, second MMA issuer
constexpr int kXThreadsGen = (kXEpiWarps + 4 + kXGenWarps) * 32;   // + Toeplitz producers (the second issuer is the last warp)
constexpr int kGenSeg = 336;                                       // staged input samples per tile: 256 + 63 taps, rounded up
constexpr int kXMaxStage = 16, kXMaxW = 40;

struct ResRaw { uint4 h0, h1, l0, l1; float s; };

__device__ __forceinline__ void pt_load_raw(const PlaneTensor& t, const uint8_t* img, int pos, int g, uint4& hi, uint4& lo) {
  const int row = pos + 8;                       // residual tensors are never de-interleaved or packed
  const uint32_t sw = (uint32_t)row & 7u;
  const int64_t sb = pt_slab_bytes(t);
  const uint8_t* r = img + (int64_t)(g >> 3) * sb + (int64_t)row * 128 + ((((uint32_t)g & 7u) ^ sw) << 4);
  hi = __ldg(reinterpret_cast<const uint4*>(r));
  lo = t.planes == 2 ? __ldg(reinterpret_cast<const uint4*>(r + (int64_t)t.spp * sb)) : make_uint4(0, 0, 0, 0);
}


// channels [8g, 8g + 8) of output row `pos` of a plain-epilogue layer; omul / opar: `ileave`.  Only the Toeplitz layer (decoder's
// k9 1 -> 20) ever writes a FOLDED image (plane.cuh: `fold_out` 1 = pairs of positions, 2 = pairs within each position parity), so
// the other instantiations keep the bare store (the index maps cost the HBM-bound k1 layers of 'gln' 30 % when every store took them).
template <bool kFoldable>
__device__ __forceinline__ void x_store8(const XParams& p, uint8_t* oimg, int pos, int g, const float (&v)[8], int omul, int opar) {
  pos = omul * pos + opar;
  if constexpr (kFoldable) {
    if (p.fold_out) {
      if (g >= 3) return;                 // channels 24-31 of a 20-channel result: no place in a 24-channel group
      if (p.fold_out == 1) pt_store8(p.out, oimg, pos >> 1, 3 * (pos & 1) + g, v);
      else pt_store8(p.out, oimg, ((pos >> 2) << 1) | (pos & 1), 3 * ((pos >> 1) & 1) + g, v);
      return;
    }
  }
  pt_store8(p.out, oimg, pos, g, v);
}

// ---- Toeplitz producers (PK_GEN), shared by plane_x_kernel<true> and plane_xs_kernel ---------------------------------
__device__ __forceinline__ void gen_producer(const XParams& p, int ptid, int lane, uint8_t* sA, uint64_t* a_full, uint64_t* a_empty,
                                             __half* s_xh, __half* s_xl) {
      const int nch = p.ksteps * 2;     // 16-byte chunks per row that the MMAs read
      const int nch_sh = nch == 8 ? 3 : (nch == 4 ? 2 : (nch == 2 ? 1 : 0));
      const uint32_t* wh = reinterpret_cast<const uint32_t*>(s_xh);
      const uint32_t* wl = reinterpret_cast<const uint32_t*>(s_xl);
      uint32_t kb = 0, ph = 1;
      for (int64_t tile = (int64_t)blockIdx.x - p.cta0; tile < p.n_tiles; tile += p.ncta) {
        const int64_t f = tile / p.tiles_per_frame;
        const int q0 = (int)(tile - f * p.tiles_per_frame) * p.tile;
        const float* xv = p.xvec + f * p.Lin;
        const float* xs = p.xsub ? p.xsub + f * p.Lin : nullptr;
        for (int j = ptid; j < kGenSeg; j += kXGenWarps * 32) {
          const int pos = q0 - p.padL + j;
          float x = 0.f;
          if (pos >= 0 && pos < p.Lin) x = p.xscale * (__ldg(xv + pos) - (xs ? __ldg(xs + pos) : 0.f));
          const __half h = __float2half_rn(x);
          s_xh[j] = h;
          s_xl[j] = __float2half_rn(x - __half2float(h));
        }
        for (int ap = 0; ap < p.planes; ++ap) mbar_wait_relaxed(&a_empty[kb + ap], ph);
        asm volatile("bar.sync 2, %0;" :: "n"(kXGenWarps * 32) : "memory");
        uint8_t* dhi = sA + kb * (uint32_t)p.stage_bytes;
        uint8_t* dlo = dhi + p.stage_bytes;
        for (int item = ptid; item < (p.tile << nch_sh); item += kXGenWarps * 32) {
          const int i = item >> nch_sh, g = item & (nch - 1);
          const int j0 = i + 8 * g;
          const int w = j0 >> 1;
          const uint32_t off = (uint32_t)i * 128u + (((uint32_t)g ^ ((uint32_t)i & 7u)) << 4);
          uint32_t a0 = wh[w], a1 = wh[w + 1], a2 = wh[w + 2], a3 = wh[w + 3], a4 = wh[w + 4];
          uint4 o;
          if (j0 & 1) o = make_uint4(__funnelshift_r(a0, a1, 16), __funnelshift_r(a1, a2, 16), __funnelshift_r(a2, a3, 16), __funnelshift_r(a3, a4, 16));
          else o = make_uint4(a0, a1, a2, a3);
          *reinterpret_cast<uint4*>(dhi + off) = o;
          if (p.planes == 2) {
            a0 = wl[w]; a1 = wl[w + 1]; a2 = wl[w + 2]; a3 = wl[w + 3]; a4 = wl[w + 4];
            if (j0 & 1) o = make_uint4(__funnelshift_r(a0, a1, 16), __funnelshift_r(a1, a2, 16), __funnelshift_r(a2, a3, 16), __funnelshift_r(a3, a4, 16));
            else o = make_uint4(a0, a1, a2, a3);
            *reinterpret_cast<uint4*>(dlo + off) = o;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0)
          for (int ap = 0; ap < p.planes; ++ap) mbar_arrive(&a_full[kb + ap]);
        asm volatile("bar.sync 2, %0;" :: "n"(kXGenWarps * 32) : "memory");   // staging array is reused by the next tile
        if (p.kbuf == 2) { kb = kb ? 0u : (uint32_t)p.n_stage; if (kb == 0) ph ^= 1u; }
        else ph ^= 1u;
      }
}

// ---- MMA issue (shared by plane_x_kernel and plane_xs_kernel) ------------------------------------------------------
// One elected thread per issuing warp.  With two M tiles per work unit there are TWO issuing threads, one per M tile
// (independent accumulators): descriptor arithmetic and the tcgen05.mma issue itself cost ~50-90 cycles per instruction
// from a single thread, more than the 44-60 cycles the narrow MMAs of this codec take on the tensor pipe.
struct XBars {
  uint64_t *a_full, *a_empty, *w_full, *w_empty, *acc_full, *acc_empty;
  uint64_t *a_full2, *w_full2;      // CTA pairs: the peer's "landed" reports, in the leader's shared memory
};

// "operand landed" of a CTA pair: every CTA's copies complete on its own barrier; the peer's issuing thread forwards each
// completion to the leader's twin barrier, the leader waits for both.
template <int MODE>
__device__ __forceinline__ void wait_full(uint64_t* mine, uint64_t* twin, uint32_t parity) {
  mbar_wait(mine, parity);
  if constexpr (MODE == 1) mbar_wait(twin, parity);
  if constexpr (MODE == 2) mbar_arrive_remote(twin, 0u);
  tc_fence_after();
}

struct XIssue {
  uint32_t d, idesc;                // this issuer's first accumulator
  int nmt;                          // M tiles this thread issues for (accumulators Npad columns apart, A rows 128 apart)
  uint32_t npad;
  uint32_t w_lo_base, unit_lo;
  uint32_t ws, wph, u;              // weight ring slot / phase, unit counter inside the tile
  bool resident, w_ready;
  int wslots;
  int issuer, psplit;
  const XBars* b;
  // a unit that belongs to the other issuing thread: keep the ring position, touch nothing
  __device__ __forceinline__ void skip_w() {
    if (!resident) { if (++ws == (uint32_t)wslots) { ws = 0; wph ^= 1u; } }
    ++u;
  }
  template <int MODE>
  __device__ __forceinline__ uint32_t wait_w() {
    if (resident) {
      if (!w_ready) wait_full<MODE>(&b->w_full[u], &b->w_full2[u], 0u);
      return w_lo_base + u * unit_lo;
    }
    wait_full<MODE>(&b->w_full[ws], &b->w_full2[ws], wph);
    return w_lo_base + ws * unit_lo;
  }
  template <int MODE>
  __device__ __forceinline__ void done_w() {
    if (!resident) {
      commit_mode<MODE>(&b->w_empty[ws]);
      if (++ws == (uint32_t)wslots) { ws = 0; wph ^= 1u; }
    }
    ++u;
  }
};

template <int NKS, int MODE>
__device__ __forceinline__ void x_issue_mt(const XIssue& x, uint32_t a_lo, uint32_t b_lo, uint32_t accum) {
  issue_ks<NKS, MODE>(x.d, a_lo, b_lo, x.idesc, accum);
  if (x.nmt == 2) issue_ks<NKS, MODE>(x.d + x.npad, a_lo + ((128u * 128u) >> 4), b_lo, x.idesc, accum);
}

// narrow (packed) input: per tap ONE chain over the row's K axis (planes = 2: K-concatenated hi*wh + lo*wh + hi*wl, 4 K steps)
template <int NKS, int MODE>
__device__ __forceinline__ void x_issue_packed(XIssue& x, const XParams& p, uint32_t a_lo0) {
  uint32_t accum = 0;
  for (int t = 0; t < p.K; ++t) {
    const uint32_t b_lo = x.wait_w<MODE>();
    const uint32_t a_lo = a_lo0 + (uint32_t)(t * p.dil) * 8u;      // one row = 128 bytes = 8 descriptor units
    x_issue_mt<NKS, MODE>(x, a_lo, b_lo, accum);
    accum = 1;
    x.done_w<MODE>();
  }
}

// one 64-channel slab of a wide input: per tap the W_hi unit meets the hi and the lo plane, the W_lo unit the hi plane
// (kSplit: the experimental two-thread form, XParams::psplit -- a template parameter because three extra predicates per unit in the
// ONE-thread loop cost the stride-2 conv 15 %)
template <int NKS, int MODE, bool kSplit = false>
__device__ __forceinline__ void x_issue_slab(XIssue& x, const XParams& p, int s, uint32_t a_stage0, uint32_t stage_lo, uint32_t kb,
                                             uint32_t a_phase, uint32_t& waited, uint32_t& accum) {
  const int spp = p.in.spp;
  auto wait_stage = [&](int stage) {
    if (!((waited >> stage) & 1u)) {
      const long long tw = p.stats != nullptr ? clock64() : 0;
      wait_full<MODE>(&x.b->a_full[kb + stage], &x.b->a_full2[kb + stage], a_phase);
      if (p.stats != nullptr) stat_add(p.stats, 6, clock64() - tw);
      waited |= 1u << stage;
    }
  };
  for (int t = 0; t < p.K; ++t) {
    int sub = 0, rowoff;
    if (p.stride == 1) rowoff = 8 + t * p.dil - p.padL;
    else { const int tp = t - p.padL; sub = tp & 1; rowoff = 8 + (tp - sub) / 2; }
    const int st_hi = pt_slab_index(p.in, sub, 0, s);
    const uint32_t a_hi = a_stage0 + (uint32_t)st_hi * stage_lo + (uint32_t)rowoff * 8u;
    const uint32_t a_lo_pl = a_hi + (uint32_t)spp * stage_lo;     // the lo plane's stage is spp slabs further
    if constexpr (kSplit) {
      if (x.issuer == 1) x.skip_w();
      else {
        const uint32_t b_lo = x.wait_w<MODE>();
        wait_stage(st_hi);
        x_issue_mt<NKS, MODE>(x, a_hi, b_lo, accum);
        accum = 1;
        if (p.planes == 2) { wait_stage(st_hi + spp); x_issue_mt<NKS, MODE>(x, a_lo_pl, b_lo, 1u); }
        x.done_w<MODE>();
      }
      if (p.planes == 2) {
        if (x.issuer == 0) x.skip_w();
        else {
          const uint32_t b_lo = x.wait_w<MODE>();
          wait_stage(st_hi);
          x_issue_mt<NKS, MODE>(x, a_hi, b_lo, accum);
          accum = 1;
          x.done_w<MODE>();
        }
      }
    } else {
      {
        const uint32_t b_lo = x.wait_w<MODE>();
        wait_stage(st_hi);
        x_issue_mt<NKS, MODE>(x, a_hi, b_lo, accum);
        accum = 1;
        if (p.planes == 2) { wait_stage(st_hi + spp); x_issue_mt<NKS, MODE>(x, a_lo_pl, b_lo, 1u); }
        x.done_w<MODE>();
      }
      if (p.planes == 2) {
        const uint32_t b_lo = x.wait_w<MODE>();
        x_issue_mt<NKS, MODE>(x, a_hi, b_lo, 1u);
        x.done_w<MODE>();
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void x_issuer(const XParams& p, bool gen, int issuer, uint8_t* sA, uint8_t* sW, uint32_t tmem, const XBars& bars) {
  const uint32_t mt_step = (128u * 128u) >> 4;
  const int m0 = (p.n_iss == 2 && !p.psplit) ? issuer : 0;        // first M tile of this thread
  const uint32_t a_lo_base = desc_lo(smem_u32(sA)) + (uint32_t)m0 * mt_step;
  const uint32_t stage_lo = (uint32_t)p.stage_bytes >> 4;
  const int acc_cols = (p.n_iss == 3 ? 3 : p.psplit ? 2 : p.mt) * p.Npad;
  XIssue x;
  x.issuer = issuer; x.psplit = p.psplit;
  x.idesc = make_idesc_f16(p.Npad, MODE == 0 ? 128 : 256);
  x.w_lo_base = desc_lo(smem_u32(sW));
  x.unit_lo = (uint32_t)p.slot_bytes >> 4;
  x.ws = 0; x.wph = 0; x.u = 0;
  x.resident = p.resident != 0;
  x.w_ready = false;
  x.wslots = p.wslots;
  x.b = &bars;
  x.nmt = (p.n_iss == 2 && !p.psplit) ? 1 : p.mt;
  x.npad = (uint32_t)p.Npad;
  uint32_t it = 0, kb = 0, a_phase = 0;
  for (int64_t ti = 0, tile; (tile = x_tile_at(p, ti)) < p.n_tiles; ++ti, ++it) {
    const uint32_t acc_i = it & 1u;
    long long tw = p.stats != nullptr ? clock64() : 0;
    if constexpr (MODE != 2) mbar_wait(&bars.acc_empty[acc_i], ((it >> 1) & 1u) ^ 1u);   // (pair: both CTAs' epilogues arrive on the leader's)
    if (p.stats != nullptr) stat_add(p.stats, 7, clock64() - tw);
    tc_fence_after();
    x.d = tmem + acc_i * (uint32_t)acc_cols + (uint32_t)(((p.n_iss == 3 || p.psplit) ? issuer : m0) * p.Npad);
    x.u = 0;
    if (gen) {
      if constexpr (MODE == 0) {
        uint32_t accum = 0;
        for (int wp = 0; wp < p.planes; ++wp) {
          const uint32_t b_lo = x.wait_w<0>();
          for (int ap = 0; ap < (wp == 0 ? p.planes : 1); ++ap) {
            if (wp == 0) { mbar_wait(&bars.a_full[kb + ap], a_phase); tc_fence_after(); }
            for (int m = 0; m < x.nmt; ++m)
              issue_n(p.ksteps, x.d + (uint32_t)m * x.npad, a_lo_base + (kb + (uint32_t)ap) * stage_lo + (uint32_t)m * mt_step, b_lo, x.idesc, accum);
            accum = 1;
          }
          x.done_w<0>();
        }
        for (int ap = 0; ap < p.planes; ++ap) umma_commit(&bars.a_empty[kb + ap]);
      }
    } else if (p.n_iss == 3) {
      // folded narrow conv (plane.cuh): 45 narrow MMAs per tile cost one issuing thread ~85 cycles each (measured: 4.0 k cycles per
      // tile in this loop, 0.2 k of them waiting), twice what the tensor pipe needs.  THREE issuing threads, one per product --
      // 0: hi x W_hi, 1: lo x W_hi, 2: hi x W_lo -- each into its own accumulator; the epilogue adds the three.
      if constexpr (MODE == 0) {
        const int st = issuer == 1 ? 1 : 0;                        // this product's activation plane (spp = 1: stage = plane)
        const long long t_iss = p.stats != nullptr ? clock64() : 0;
        mbar_wait(&bars.a_full[kb + (uint32_t)st], a_phase);
        if (p.stats != nullptr) stat_add(p.stats, 6, clock64() - t_iss);
        tc_fence_after();
        const uint32_t a0 = a_lo_base + (kb + (uint32_t)st) * stage_lo + (uint32_t)(8 - p.padL) * 8u;
        for (int t = 0; t < p.K; ++t) {
          const uint32_t unit = (uint32_t)(t * 2 + (issuer == 2 ? 1 : 0));   // ((slab K) + tap) planes + plane
          if (!x.w_ready) { mbar_wait(&bars.w_full[unit], 0u); tc_fence_after(); }
          issue_ks<3, 0>(x.d, a0 + (uint32_t)t * 8u, x.w_lo_base + unit * x.unit_lo, x.idesc, t == 0 ? 0u : 1u);
        }
        umma_commit(&bars.a_empty[kb]);                            // (every issuing thread reports on both planes' stages)
        umma_commit(&bars.a_empty[kb + 1u]);
        if (p.stats != nullptr && issuer == 0) { stat_add(p.stats, 5, clock64() - t_iss); stat_add(p.stats, 4, 1); }
      }
    } else if (p.in.packed) {
      tw = p.stats != nullptr ? clock64() : 0;
      wait_full<MODE>(&bars.a_full[kb], &bars.a_full2[kb], a_phase);
      if (p.stats != nullptr) stat_add(p.stats, 6, clock64() - tw);
      if (p.in_free != nullptr && issuer == 0) flag_bump(p.in_free + tile / p.tiles_per_frame);   // this CTA's tile has been read out of the ring
      const uint32_t a_lo0 = a_lo_base + kb * stage_lo + (uint32_t)(p.halo - p.padL) * 8u;
      switch (p.ksteps) {
        case 2: x_issue_packed<2, MODE>(x, p, a_lo0); break;
        case 1: x_issue_packed<1, MODE>(x, p, a_lo0); break;
        case 3: x_issue_packed<3, MODE>(x, p, a_lo0); break;
        default: x_issue_packed<4, MODE>(x, p, a_lo0); break;
      }
      commit_mode<MODE>(&bars.a_empty[kb]);
    } else {
      const int nsub = p.in.deint ? 2 : 1;
      uint32_t waited = 0, accum = 0;
      const long long t_iss = p.stats != nullptr ? clock64() : 0;
      for (int s = 0; s < p.in.spp; ++s) {
        const int nks = min(4, p.ksteps - 4 * s);
        const uint32_t a0 = a_lo_base + kb * stage_lo;
        if (p.psplit) {
          if constexpr (MODE == 0) {
            switch (nks) {
              case 4: x_issue_slab<4, 0, true>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
              case 3: x_issue_slab<3, 0, true>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
              case 2: x_issue_slab<2, 0, true>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
              default: x_issue_slab<1, 0, true>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
            }
          }
        } else
        switch (nks) {
          case 4: x_issue_slab<4, MODE>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
          case 3: x_issue_slab<3, MODE>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
          case 2: x_issue_slab<2, MODE>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
          default: x_issue_slab<1, MODE>(x, p, s, a0, stage_lo, kb, a_phase, waited, accum); break;
        }
        for (int sub = 0; sub < nsub; ++sub)
          for (int ap = 0; ap < p.planes; ++ap) commit_mode<MODE>(&bars.a_empty[kb + pt_slab_index(p.in, sub, ap, s)]);
      }
      if (p.stats != nullptr) { stat_add(p.stats, 5, clock64() - t_iss); stat_add(p.stats, 4, 1); }   // issue section incl. its waits; tiles
    }
    commit_mode<MODE>(&bars.acc_full[acc_i]);
    x.w_ready = true;
    kb += (uint32_t)p.n_stage;                       // next of the kbuf tile buffers
    if (kb == (uint32_t)(p.n_stage * p.kbuf)) { kb = 0; a_phase ^= 1u; }
  }
}

// kFold: the folded narrow conv (plane.cuh) -- its own instantiation, so that its unfolding epilogue does not raise the register pressure
// of the other layers' (with one kernel for both, every instantiation spilled)
template <bool kGen, bool kPair, bool kFold = false>
__global__ void __launch_bounds__(kGen ? kXThreadsGen : kXThreadsX, 1) plane_x_kernel(const __grid_constant__ XParams p) {
  static_assert(!(kGen && kPair), "Toeplitz layers do not run as CTA pairs");
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full[kXMaxStage], a_empty[kXMaxStage], w_full[kXMaxW], w_empty[kXMaxW], acc_full[2], acc_empty[2];
  __shared__ uint64_t a_full2[kXMaxStage], w_full2[kXMaxW];   // CTA pairs: the peer's operands have landed (live in the leader)
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[128];
  __shared__ __align__(16) __half s_xh[kGen ? kGenSeg : 8], s_xl[kGen ? kGenSeg : 8];   // hi / lo halves of the tile's input samples

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbuf = p.n_stage * p.kbuf;
  uint8_t* sA = smem;
  uint8_t* sW = smem + (uint32_t)nbuf * (uint32_t)p.stage_bytes;
  const int acc_cols = (p.n_iss == 3 ? 3 : p.psplit ? 2 : p.mt) * p.Npad;
  const int n_iss = p.n_iss;           // issuing threads: one, or one per M tile
  constexpr int kIssuer1 = kGen ? kXEpiWarps + 3 + kXGenWarps : kXEpiWarps + 3;

  if (tid < 128) {
    if (p.glu) s_bias[tid] = tid < 20 ? (p.bias ? __ldg(p.bias + tid) : 0.f) : (tid < 40 && p.bias2) ? __ldg(p.bias2 + tid - 20) : 0.f;
    else s_bias[tid] = (p.bias && tid < p.Cout) ? __ldg(p.bias + tid) : 0.f;
  }
  constexpr bool pair = kPair;          // (a kernel that contains cta_group::2 instructions can only be launched as a cluster)
  const uint32_t crank = pair ? cluster_ctarank() : 0u;
  if (tid == 0) {
    for (int i = 0; i < nbuf; ++i) { mbar_init(&a_full[i], kGen ? kXGenWarps : 1); mbar_init(&a_empty[i], n_iss); mbar_init(&a_full2[i], 1); }
    // (psplit: a weight unit is read by exactly one of the two issuing threads)
    for (int i = 0; i < p.wslots; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], p.psplit ? 1 : n_iss); mbar_init(&w_full2[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], n_iss); mbar_init(&acc_empty[i], pair ? 2 * kXEpiWarps : kXEpiWarps); }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == kXEpiWarps) {
    if constexpr (pair) tmem_alloc_cg2(&tmem_base_s, (uint32_t)p.tmem_cols);
    else tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (pair) cluster_sync_all();   // the peer's barriers (and TMEM) exist before anything arrives remotely
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();      // barriers, TMEM and the bias table are set up under the previous kernel's tail; its output is only touched from here

  if (warp < kXEpiWarps) {
    // =========================== epilogue ===========================
    const int quarter = warp & 3, grp = warp >> 2;
    const int nb = p.Npad >> 4;
    const int n_e = p.mt * nb;
    const float slope = act_slope_of(p.act), pslope = act_slope_of(p.post_act);
    const bool slow_act = p.act == NSC_ACT_TANH || p.post_act == NSC_ACT_TANH;
    uint32_t it = 0;
    for (int64_t ti = 0, tile; (tile = x_tile_at(p, ti)) < p.n_tiles; ++ti, ++it) {
      const int64_t f = tile / p.tiles_per_frame;
      const int q0 = (int)(tile - f * p.tiles_per_frame) * p.tile;
      const uint32_t acc_i = it & 1u;
      // `ileave`: this layer's frames are the two position-parity sub-images of the output's frames
      uint8_t* oimg = p.out.base + (p.ileave ? (f >> 1) : f) * p.out.frame_bytes;
      const int opar = p.ileave ? (int)(f & 1) : 0, omul = p.ileave ? 2 : 1;
      const uint8_t* rimg = p.res_mode == RES_ADD ? p.res.base + f * p.res.frame_bytes : nullptr;
      const float* rvec = p.res_mode == RES_ADD_BCAST ? p.resvec + f * p.Lout : nullptr;
      const int row0 = q0 + quarter * 32 + lane;
      if constexpr (kFold) {
        // folded narrow conv (plane.cuh): row q of this tile holds output positions 2 q and 2 q + 1 (24 columns each).  A thread owns
        // a row, so direct global stores would touch 32 lines per warp instruction (the load/store unit, not the MMAs, then bounds
        // the layer: measured 67 us per launch against 35 of MMAs).  The rows go to a staging ring that mirrors the packed output
        // image and leave by ONE bulk store per 256 positions (dilation 1) or per frame (dilation 2: the two parities of a frame are
        // consecutive tiles of this CTA, `pairtiles`).  One 16-column batch per epilogue group.
        uint8_t* const sO = sW + (uint32_t)p.wslots * (uint32_t)p.slot_bytes;
        mbar_wait_relaxed(&acc_full[acc_i], (it >> 1) & 1u);
        tc_fence_after();
        uint32_t r[16], r1[16], r2[16];
        const uint32_t tcol = tmem + ((uint32_t)(quarter * 32) << 16) + acc_i * (uint32_t)acc_cols + (uint32_t)(grp * 16);
        tmem_ld16(tcol, r);
        if (n_iss == 3) { tmem_ld16(tcol + (uint32_t)p.Npad, r1); tmem_ld16(tcol + 2u * (uint32_t)p.Npad, r2); }   // one accumulator per product
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                          // TMEM slot free: the next tile's MMAs run under the rest
          if (pair) mbar_arrive_remote(&acc_empty[acc_i], 0u);
          else mbar_arrive(&acc_empty[acc_i]);
        }
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float a = __uint_as_float(r[e]);
          if (n_iss == 3) a += __uint_as_float(r1[e]) + __uint_as_float(r2[e]);
          v[e] = act_fast(a + s_bias[grp * 16 + e], slope);
        }
        const int q = quarter * 32 + lane;
        // window rows: dilation 1 -> tile parity picks a 256-row half, row 2 q + ph; dilation 2 -> the whole ring, row 4 q + 2 ph + parity
        const int wbase = p.ileave ? (int)(f & 1) : 0;
        const int wmul = p.ileave ? 4 : 2, wph = p.ileave ? 2 : 1;
        // the window's previous contents must have left shared memory: its bulk store was issued a whole tile ago (dilation 1: one
        // 256-row window, stored per tile) or when the previous frame was complete (dilation 2), so this wait is normally free --
        // waiting right after the store instead stalled the epilogue for the duration of the read
        if (warp == 0 && lane == 0 && (!p.ileave || !(f & 1))) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 3, %0;" :: "n"(kXEpiWarps * 32) : "memory");
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g = 2 * grp + h, ph = g >= 3 ? 1 : 0, gg = g - 3 * ph;
          float a8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) a8[e] = v[8 * h + e];
          uint4 hi, lo;
          split8(a8, hi, lo);
          const int R = wbase + wmul * q + wph * ph;
          uint8_t* rp = sO + (uint32_t)R * 128u;
          const uint32_t sw = (uint32_t)R & 7u;
          if (gg < 2) {                                           // K-concatenated row (plane.cuh)
            *reinterpret_cast<uint4*>(rp + (((uint32_t)gg ^ sw) << 4)) = hi;
            *reinterpret_cast<uint4*>(rp + (((uint32_t)(gg + 2) ^ sw) << 4)) = lo;
            *reinterpret_cast<uint4*>(rp + (((uint32_t)(gg + 4) ^ sw) << 4)) = hi;
          } else {
            *reinterpret_cast<uint4*>(rp + ((6u ^ sw) << 4)) = make_uint4(hi.x, hi.y, lo.x, lo.y);
            *reinterpret_cast<uint4*>(rp + ((7u ^ sw) << 4)) = make_uint4(hi.x, hi.y, 0u, 0u);
          }
        }
        fence_async_smem();
        asm volatile("bar.sync 3, %0;" :: "n"(kXEpiWarps * 32) : "memory");
        if (warp == 0 && lane == 0) {
          if (!p.ileave) {
            bulk_s2g(p.out.base + f * p.out.frame_bytes + (int64_t)(8 + 2 * q0) * 128, sO, 256u * 128u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          } else if (f & 1) {
            bulk_s2g(p.out.base + (f >> 1) * p.out.frame_bytes + 8 * 128, sO, 512u * 128u);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        continue;
      }
      if constexpr (!kFold) {
      if (p.glu) {
        // gated linear unit: columns [0, 20) linear gate, [20, 40) tanh gate -> 20-channel packed row (nn_core_operator.py:91-102)
        mbar_wait_relaxed(&acc_full[acc_i], (it >> 1) & 1u);
        tc_fence_after();
        for (int mt_i = grp; mt_i < p.mt; mt_i += kXEpiGroups) {
          const int pos = row0 + mt_i * 128;
          uint32_t r0[16], r1[16], r2[16];
          const uint32_t tb = tmem + ((uint32_t)(quarter * 32) << 16) + acc_i * (uint32_t)acc_cols + (uint32_t)(mt_i * p.Npad);
          tmem_ld16(tb, r0);
          tmem_ld16(tb + 16u, r1);
          tmem_ld16(tb + 32u, r2);
          tmem_ld_wait();
          float v[24];
#pragma unroll
          for (int c = 0; c < 20; ++c) {
            const float a = __uint_as_float(c < 16 ? r0[c & 15] : r1[c & 3]) + s_bias[c];
            const float g = __uint_as_float(c < 12 ? r1[(4 + c) & 15] : r2[(c - 12) & 15]) + s_bias[20 + c];
            v[c] = a * tanhf(g);
          }
          v[20] = v[21] = v[22] = v[23] = 0.f;
          float a8[8];
#pragma unroll
          for (int g3 = 0; g3 < 3; ++g3) {
#pragma unroll
            for (int e = 0; e < 8; ++e) a8[e] = v[8 * g3 + e];
            pt_store8(p.out, oimg, omul * pos + opar, g3, a8);
          }
          if (p.out.planes == 1) {   // channels 24-31: K padding the consumer reads
#pragma unroll
            for (int e = 0; e < 8; ++e) a8[e] = 0.f;
            pt_store8(p.out, oimg, omul * pos + opar, 3, a8);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (pair) mbar_arrive_remote(&acc_empty[acc_i], 0u);
          else mbar_arrive(&acc_empty[acc_i]);
        }
        continue;
      }
      auto load_res = [&](int u, ResRaw& rr) {
        if (kGen || u >= n_e) return;      // 1-channel-input layers carry no residual
        const int mt_i = u / nb, c0 = (u - mt_i * nb) << 4;
        const int pos = row0 + mt_i * 128;
        if (rimg) {
          pt_load_raw(p.res, rimg, pos, c0 >> 3, rr.h0, rr.l0);
          pt_load_raw(p.res, rimg, pos, (c0 >> 3) + 1, rr.h1, rr.l1);
        } else if (rvec) {
          rr.s = __ldg(rvec + pos);
        }
      };
      ResRaw rn, rn2;
      rn.s = 0.f; rn2.s = 0.f;
      load_res(grp, rn);                 // the residual does not depend on the accumulator: two batches are fetched
      load_res(grp + kXEpiGroups, rn2);  // before waiting for it, and the pipeline stays two batches deep
      mbar_wait_relaxed(&acc_full[acc_i], (it >> 1) & 1u);
      tc_fence_after();
      for (int u = grp; u < n_e; u += kXEpiGroups) {
        const int mt_i = u / nb, c0 = (u - mt_i * nb) << 4;
        const int pos = row0 + mt_i * 128;
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + acc_i * (uint32_t)acc_cols + (uint32_t)(mt_i * p.Npad + c0), r);
        if (p.psplit) {      // the W_lo products sit in their own accumulator (second issuing thread)
          uint32_t r2[16];
          tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + acc_i * (uint32_t)acc_cols + (uint32_t)(p.Npad + c0), r2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
        }
        const ResRaw rc = rn;
        rn = rn2;
        load_res(u + 2 * kXEpiGroups, rn2);
        float rs[16];
        if (rimg) {
          float a[8], b[8];
          unpack8(rc.h0, a); unpack8(rc.l0, b);
#pragma unroll
          for (int e = 0; e < 8; ++e) rs[e] = a[e] + b[e];
          unpack8(rc.h1, a); unpack8(rc.l1, b);
#pragma unroll
          for (int e = 0; e < 8; ++e) rs[8 + e] = a[e] + b[e];
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) rs[e] = (rvec && c0 + e < p.Cout) ? rc.s : 0.f;
        }
        float bs[16];
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + e]);
          bs[e] = b4.x; bs[e + 1] = b4.y; bs[e + 2] = b4.z; bs[e + 3] = b4.w;
        }
        tmem_ld_wait();
        float v[16];
        if (!slow_act) {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = act_fast(act_fast(__uint_as_float(r[e]) + bs[e], slope) + rs[e], pslope);
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = act_slow(act_slow(__uint_as_float(r[e]) + bs[e], p.act) + rs[e], p.post_act);
        }
        if (p.shuffle == 1) {
          float a[8], b[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { a[e] = v[e]; b[e] = v[8 + e]; }
          x_store8<kGen>(p, oimg, pos, c0 >> 3, a, omul, opar);
          x_store8<kGen>(p, oimg, pos, (c0 >> 3) + 1, b, omul, opar);
        } else {   // sub-pixel: out[2 pos + r, c] = y[pos, 2 c + r]   (nscm.py:158-167)
          float a[8], b[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { a[e] = v[2 * e]; b[e] = v[2 * e + 1]; }
          pt_store8(p.out, oimg, 2 * pos, c0 >> 4, a);
          pt_store8(p.out, oimg, 2 * pos + 1, c0 >> 4, b);
        }
      }
      if (grp == 0 && p.zero_from < p.zero_to) {
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int mt_i = 0; mt_i < p.mt; ++mt_i) {
          const int pos = row0 + mt_i * 128;
          for (int g = p.zero_from; g < p.zero_to; ++g) {
            if (p.shuffle == 1) pt_store8(p.out, oimg, omul * pos + opar, g, z);
            else { pt_store8(p.out, oimg, 2 * pos, g, z); pt_store8(p.out, oimg, 2 * pos + 1, g, z); }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pair) mbar_arrive_remote(&acc_empty[acc_i], 0u);   // the leader's barrier counts both CTAs' epilogue warps
        else mbar_arrive(&acc_empty[acc_i]);
      }
      }   // !kFold
    }
    if (kFold && warp == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp == kXEpiWarps || warp == kIssuer1) {
    // =========================== MMA issuers ===========================
    const int issuer = warp == kXEpiWarps ? 0 : 1;
    if (issuer < n_iss && elect_one()) {
      const XBars bars{a_full, a_empty, w_full, w_empty, acc_full, acc_empty, a_full2, w_full2};
      if constexpr (kGen) x_issuer<0>(p, true, issuer, sA, sW, tmem, bars);
      else {
        if constexpr (!pair) x_issuer<0>(p, false, issuer, sA, sW, tmem, bars);
        else {
          if (crank == 0) x_issuer<1>(p, false, issuer, sA, sW, tmem, bars);
          else if (issuer == 0) x_issuer<2>(p, false, issuer, sA, sW, tmem, bars);   // the peer only reports what has landed
        }
      }
    }
  } else if (warp == kXEpiWarps + 1) {
    // =========================== A loader (bulk copies of plane tiles) ===========================
    if (!kGen && elect_one()) {
      const int64_t sb = pt_slab_bytes(p.in);
      const int nsub = p.in.deint ? 2 : 1;
      const int npl = p.in.packed ? 1 : p.in.planes;
      uint32_t kb = 0, ph = 1;
      for (int64_t ti = 0, tile; (tile = x_tile_at(p, ti)) < p.n_tiles; ++ti) {
        const int64_t f = tile / p.tiles_per_frame;
        const int q0 = (int)(tile - f * p.tiles_per_frame) * p.tile;
        const uint8_t* img = p.in.base + f * p.in.frame_bytes;
        for (int s = 0; s < p.in.spp; ++s)          // consumption order of the issuer: slab-major
          for (int sub = 0; sub < nsub; ++sub)
            for (int ap = 0; ap < npl; ++ap) {
              const int stage = pt_slab_index(p.in, sub, ap, s);
              mbar_wait_relaxed(&a_empty[kb + stage], ph);
              mbar_expect_tx(&a_full[kb + stage], (uint32_t)p.stage_bytes);
              bulk_g2s(sA + (uint32_t)(kb + stage) * (uint32_t)p.stage_bytes, img + stage * sb + (int64_t)(q0 - (p.halo - 8)) * 128,
                       (uint32_t)p.stage_bytes, &a_full[kb + stage]);
            }
        kb += (uint32_t)p.n_stage;
        if (kb == (uint32_t)(p.n_stage * p.kbuf)) { kb = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kXEpiWarps + 2) {
    // =========================== W loader ===========================
    if (elect_one()) {
      // a CTA of a pair stages its half of every unit (output columns [rank Npad/2, (rank + 1) Npad/2))
      const uint8_t* wsrc = p.wpack + (pair ? (size_t)crank * (size_t)p.slot_bytes : 0);
      const uint32_t sbytes = (uint32_t)p.slot_bytes;
      if (p.resident) {
        for (int u = 0; u < p.n_units; ++u) {
          mbar_expect_tx(&w_full[u], sbytes);
          bulk_g2s(sW + (uint32_t)u * sbytes, wsrc + (size_t)u * p.unit_bytes, sbytes, &w_full[u]);
        }
        if constexpr (!kGen && !pair) {
          if (n_iss == 3) {        // resident weights leave this thread idle: it is the third MMA-issuing thread of the folded conv
            const XBars bars{a_full, a_empty, w_full, w_empty, acc_full, acc_empty, a_full2, w_full2};
            x_issuer<0>(p, false, 2, sA, sW, tmem, bars);
          }
        }
      } else {
        uint32_t ws = 0, wph = 1;
        for (int64_t ti = 0; x_tile_at(p, ti) < p.n_tiles; ++ti) {
          for (int u = 0; u < p.n_units; ++u) {
            mbar_wait_relaxed(&w_empty[ws], wph);
            mbar_expect_tx(&w_full[ws], sbytes);
            bulk_g2s(sW + ws * sbytes, wsrc + (size_t)u * p.unit_bytes, sbytes, &w_full[ws]);
            if (++ws == (uint32_t)p.wslots) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else {
    // =========================== Toeplitz producers (PK_GEN) ===========================
    // The tile's input samples are split into hi / lo halves ONCE into a small staging array; row i of the Toeplitz
    // tile is then the 2 * ksteps * 8 consecutive halves starting at sample i, i.e. per 16-byte chunk five 32-bit
    // shared loads and (for odd i) four funnel shifts.  Entries past the K taps meet zero weights.
    if constexpr (kGen) gen_producer(p, tid - (kXEpiWarps + 3) * 32, lane, sA, a_full, a_empty, s_xh, s_xl);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (pair) cluster_sync_all();   // nobody frees TMEM (or retires) while the leader's instruction stream can still touch it
  tc_fence_after();
  if (warp == kXEpiWarps) {
    if constexpr (pair) tmem_dealloc_cg2(tmem, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

// ================================================================================================
// PK_X with a staged epilogue (narrow-input layers: 20 -> 100 / 20 -> 50 + residual, the HBM-bound third conv of a block)
//
// A thread owns one ROW of the accumulator, so direct global loads / stores of its 16-byte chunks touch 32 different
// 128-byte lines per warp instruction -- the load/store unit then moves one line per cycle and caps the kernel far below
// the HBM rate.  Here the residual tile arrives by bulk copy into a shared-memory unit laid out exactly like the image
// (128 rows x 128 B per plane), the epilogue adds it and writes the result over it IN PLACE (16-byte shared accesses,
// conflict-free thanks to the swizzle), and a dedicated thread sends the finished unit back with one bulk store per
// plane.  No thread touches global memory; latency hiding is the copy engine's job.
//   unit = (M tile, 64-channel output slab) x planes; ring of kSUnits units:  residual loader -> epilogue -> storer.
// ================================================================================================
constexpr int kSEpiGroups = 4;                     // epilogue warps per TMEM lane quarter of the stand-alone kernel (one 16-column batch each)
constexpr int kSThreads = (4 * kSEpiGroups + 6 + kXGenWarps) * 32;   // + MMA issuer, A loader, W loader, residual loader, storer, second MMA issuer, Toeplitz producers
constexpr int kSUnits = 3, kSMaxUnits = 4;
constexpr int kSPlane = 128 * 128;                 // one plane of one unit

// kEG: epilogue warps per lane quarter.  4 = one 16-column batch of a 64-channel unit per warp; the fused block kernel runs 3 (the
// register budget of its taps-in-N roles caps the CTA at 18 warps), where the first group takes two batches of a full unit.
template <bool kPair, int kEG>
__device__ __forceinline__ void xs_body(const XParams& p) {
  constexpr int kSEpiWarps = 4 * kEG;
  constexpr int kNB = (4 + kEG - 1) / kEG;          // batches a warp may own in one unit
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_full[4], a_empty[4], w_full[kXMaxW], w_empty[kXMaxW], acc_full[2], acc_empty[2];
  __shared__ uint64_t a_full2[4], w_full2[kXMaxW];     // CTA pairs: the peer's operands have landed (live in the leader)
  __shared__ uint64_t st_full[kSMaxUnits], st_done[kSMaxUnits], st_empty[kSMaxUnits];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[128];
  __shared__ __align__(16) __half s_xh[kPair ? 8 : kGenSeg], s_xl[kPair ? 8 : kGenSeg];   // PK_GEN: hi / lo halves of the tile's input samples
  const bool gen = !kPair && p.kind == PK_GEN;
  constexpr bool pair = kPair;
  const uint32_t crank = pair ? cluster_ctarank() : 0u;
  const uint32_t n_su = (uint32_t)p.s_units;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbuf = p.n_stage * p.kbuf;                 // n_stage == 1 (packed input)
  const uint32_t stg_bytes = (uint32_t)p.planes * kSPlane;
  uint8_t* sA = smem;
  uint8_t* sW = sA + (uint32_t)nbuf * (uint32_t)p.stage_bytes;
  uint8_t* sS = sW + (uint32_t)p.wslots * (uint32_t)p.slot_bytes;
  const int acc_cols = p.mt * p.Npad;
  const int ospp = p.out.spp;
  const int nb = p.Npad >> 4;

  if (tid < 128) s_bias[tid] = (p.bias && tid < p.Cout) ? __ldg(p.bias + tid) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < nbuf; ++i) { mbar_init(&a_full[i], gen ? kXGenWarps : 1); mbar_init(&a_empty[i], p.n_iss); mbar_init(&a_full2[i], 1); }
    for (int i = 0; i < p.wslots; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], p.n_iss); mbar_init(&w_full2[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], p.n_iss); mbar_init(&acc_empty[i], pair ? 2 * kSEpiWarps : kSEpiWarps); }
    for (int i = 0; i < kSMaxUnits; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_done[i], kSEpiWarps); mbar_init(&st_empty[i], 1); }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == kSEpiWarps) {
    if constexpr (pair) tmem_alloc_cg2(&tmem_base_s, (uint32_t)p.tmem_cols);
    else tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (pair) cluster_sync_all();   // the peer's barriers (and TMEM) exist before anything arrives remotely
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();      // barriers, TMEM and the bias table are set up under the previous kernel's tail; its output is only touched from here

  if (warp < kSEpiWarps) {
    // =========================== epilogue ===========================
    const int quarter = warp & 3, grp = warp >> 2;
    const int lr = quarter * 32 + lane;                  // row inside the M tile
    const float slope = act_slope_of(p.act), pslope = act_slope_of(p.post_act);
    const bool has_res = p.res_mode == RES_ADD;
    // what the per-value arithmetic contains: bit 0 activation, bit 1 residual add, bit 2 post-activation (slope 1 = identity)
    const int emode = (slope != 1.0f ? 1 : 0) | ((has_res || p.res_mode == RES_ADD_BCAST) ? 2 : 0) | (pslope != 1.0f ? 4 : 0);
    const bool barrier_rw = has_res && p.out.deint;      // residual and output use different unit layouts: read all, then write all
    // shared-memory offsets of this thread's two chunks inside a unit plane (layout of the OUTPUT image)
    const int srow_o = p.out.deint ? (lr >> 1) : lr;
    const uint32_t obase = (p.out.deint ? (uint32_t)(lr & 1) * (kSPlane / 2) : 0u) + (uint32_t)srow_o * 128u;
    const uint32_t osw = (uint32_t)srow_o & 7u;
    const uint32_t rbase = (uint32_t)lr * 128u, rsw = (uint32_t)lr & 7u;   // residual units are never de-interleaved
    uint32_t it = 0, slot = 0, sph = 0;
    const long long t_begin = clock64();
    for (int64_t tile = (int64_t)blockIdx.x - p.cta0; tile < p.n_tiles; tile += p.ncta, ++it) {
      const int64_t f = tile / p.tiles_per_frame;
      const int q0 = (int)(tile - f * p.tiles_per_frame) * p.tile;
      const uint32_t acc_i = it & 1u;
      const float* rvec = p.res_mode == RES_ADD_BCAST ? p.resvec + f * p.Lout : nullptr;
      mbar_wait_relaxed(&acc_full[acc_i], (it >> 1) & 1u);
      tc_fence_after();
      for (int mt_i = 0; mt_i < p.mt; ++mt_i) {
        const float rv = rvec ? __ldg(rvec + q0 + mt_i * 128 + lr) : 0.f;
        for (int s = 0; s < ospp; ++s) {
          const int nbs = min(4, nb - 4 * s);
          mbar_wait_relaxed(&st_full[slot], sph);
          uint8_t* unit = sS + slot * stg_bytes;
          float v[kNB][16];
#pragma unroll
          for (int bi = 0; bi < kNB; ++bi) {
            const int bt = grp + bi * kEG;                 // 16-column batch of this unit
            if (bt < nbs) {
              const int c0 = 64 * s + 16 * bt;             // first column / channel of the batch
              const uint32_t cg = 2u * (uint32_t)bt;       // first 16-byte chunk inside the slab
              uint32_t r[16];
              tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + acc_i * (uint32_t)acc_cols + (uint32_t)(mt_i * p.Npad + c0), r);
              float rs[16];
              if (has_res) {
                const uint4 h0 = *reinterpret_cast<const uint4*>(unit + rbase + ((cg ^ rsw) << 4));
                const uint4 h1 = *reinterpret_cast<const uint4*>(unit + rbase + (((cg + 1u) ^ rsw) << 4));
                float a[8], b[8];
                unpack8(h0, a); unpack8(h1, b);
#pragma unroll
                for (int e = 0; e < 8; ++e) { rs[e] = a[e]; rs[8 + e] = b[e]; }
                if (p.planes == 2) {
                  const uint4 l0 = *reinterpret_cast<const uint4*>(unit + kSPlane + rbase + ((cg ^ rsw) << 4));
                  const uint4 l1 = *reinterpret_cast<const uint4*>(unit + kSPlane + rbase + (((cg + 1u) ^ rsw) << 4));
                  unpack8(l0, a); unpack8(l1, b);
#pragma unroll
                  for (int e = 0; e < 8; ++e) { rs[e] += a[e]; rs[8 + e] += b[e]; }
                }
              } else if (rvec) {
#pragma unroll
                for (int e = 0; e < 16; ++e) rs[e] = (c0 + e < p.Cout) ? rv : 0.f;
              }
              float bs[16];
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c0 + e]);
                bs[e] = b4.x; bs[e + 1] = b4.y; bs[e + 2] = b4.z; bs[e + 3] = b4.w;
              }
              tmem_ld_wait();
              // The epilogue is co-limited by instruction issue (ncu: ~60 % issue-slot utilisation next to 70 % of HBM), so the
              // per-value arithmetic only contains what the layer has: an identity activation (slope 1) or a missing residual are
              // not computed (the stem ran 7 instructions per value where 3 do).  The choice is uniform: one branch per batch.
              auto finish = [&](auto kAct, auto kAdd, auto kPost) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  float t = __uint_as_float(r[e]) + bs[e];
                  if constexpr (decltype(kAct)::value) t = act_fast(t, slope);
                  if constexpr (decltype(kAdd)::value) t += rs[e];
                  if constexpr (decltype(kPost)::value) t = act_fast(t, pslope);
                  v[bi][e] = t;
                }
              };
              using T1 = std::true_type;
              using T0 = std::false_type;
              switch (emode) {
                case 0: finish(T0{}, T0{}, T0{}); break;
                case 1: finish(T1{}, T0{}, T0{}); break;
                case 2: finish(T0{}, T1{}, T0{}); break;
                case 3: finish(T1{}, T1{}, T0{}); break;
                case 4: finish(T0{}, T0{}, T1{}); break;
                case 5: finish(T1{}, T0{}, T1{}); break;
                case 6: finish(T0{}, T1{}, T1{}); break;
                default: finish(T1{}, T1{}, T1{}); break;
              }
            }
          }
          if (barrier_rw) asm volatile("bar.sync 3, %0;" :: "n"(kSEpiWarps * 32) : "memory");
#pragma unroll
          for (int bi = 0; bi < kNB; ++bi) {
            const int bt = grp + bi * kEG;
            if (bt < nbs) {
              const uint32_t cg = 2u * (uint32_t)bt;
              float a[8], b[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) { a[e] = v[bi][e]; b[e] = v[bi][8 + e]; }
              uint4 hi, lo;
              split8(a, hi, lo);
              *reinterpret_cast<uint4*>(unit + obase + ((cg ^ osw) << 4)) = hi;
              if (p.planes == 2) *reinterpret_cast<uint4*>(unit + kSPlane + obase + ((cg ^ osw) << 4)) = lo;
              split8(b, hi, lo);
              *reinterpret_cast<uint4*>(unit + obase + (((cg + 1u) ^ osw) << 4)) = hi;
              if (p.planes == 2) *reinterpret_cast<uint4*>(unit + kSPlane + obase + (((cg + 1u) ^ osw) << 4)) = lo;
            }
          }
          fence_async_smem();                            // generic-proxy writes -> visible to the bulk store
          __syncwarp();
          if (lane == 0) mbar_arrive(&st_done[slot]);
          if (++slot == n_su) { slot = 0; sph ^= 1u; }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (pair) mbar_arrive_remote(&acc_empty[acc_i], 0u);   // the leader's barrier counts both CTAs' epilogue warps
        else mbar_arrive(&acc_empty[acc_i]);
      }
    }
    if (warp == 0 && lane == 0 && p.stats != nullptr) {
      stat_add(p.stats, 0, 3);
      stat_add(p.stats, 1, clock64() - t_begin);
      stat_add(p.stats, 4, (long long)it);
    }
  } else if (warp == kSEpiWarps || warp == kSEpiWarps + 5) {
    // =========================== MMA issuers (one per M tile) ===========================
    const int issuer = warp == kSEpiWarps ? 0 : 1;
    if (issuer < p.n_iss && elect_one()) {
      const XBars bars{a_full, a_empty, w_full, w_empty, acc_full, acc_empty, a_full2, w_full2};
      if constexpr (!pair) x_issuer<0>(p, gen, issuer, sA, sW, tmem, bars);
      else {
        if (crank == 0) x_issuer<1>(p, false, issuer, sA, sW, tmem, bars);
        else if (issuer == 0) x_issuer<2>(p, false, issuer, sA, sW, tmem, bars);   // the peer only reports what has landed
      }
    }
  } else if (warp == kSEpiWarps + 1) {
    // =========================== A loader ===========================
    if (!gen && elect_one()) {
      uint32_t kb = 0, ph = 1;
      for (int64_t tile = (int64_t)blockIdx.x - p.cta0; tile < p.n_tiles; tile += p.ncta) {
        const int64_t f = tile / p.tiles_per_frame;
        const int q0 = (int)(tile - f * p.tiles_per_frame) * p.tile;
        mbar_wait_relaxed(&a_empty[kb], ph);
        if (p.in_ready != nullptr) stat_add(p.stats, 3, flag_wait(p.in_ready + f, 1u));
        mbar_expect_tx(&a_full[kb], (uint32_t)p.stage_bytes);
        bulk_g2s(sA + kb * (uint32_t)p.stage_bytes, p.in.base + pt_frame_off(p.in, f) + (int64_t)q0 * 128, (uint32_t)p.stage_bytes, &a_full[kb]);
        if (p.kbuf == 2) { kb ^= 1u; if (kb == 0) ph ^= 1u; }
        else ph ^= 1u;
      }
    }
  } else if (warp == kSEpiWarps + 2) {
    // =========================== W loader ===========================
    if (elect_one()) {
      // a CTA of a pair stages its half of every unit (output columns [rank Npad/2, (rank + 1) Npad/2))
      const uint8_t* wsrc = p.wpack + (pair ? (size_t)crank * (size_t)p.slot_bytes : 0);
      const uint32_t sbytes = (uint32_t)p.slot_bytes;
      if (p.resident) {
        for (int u = 0; u < p.n_units; ++u) {
          mbar_expect_tx(&w_full[u], sbytes);
          bulk_g2s(sW + (uint32_t)u * sbytes, wsrc + (size_t)u * p.unit_bytes, sbytes, &w_full[u]);
        }
      } else {
        uint32_t ws = 0, wph = 1;
        for (int64_t tile = (int64_t)blockIdx.x - p.cta0; tile < p.n_tiles; tile += p.ncta) {
          for (int u = 0; u < p.n_units; ++u) {
            mbar_wait_relaxed(&w_empty[ws], wph);
            mbar_expect_tx(&w_full[ws], sbytes);
            bulk_g2s(sW + ws * sbytes, wsrc + (size_t)u * p.unit_bytes, sbytes, &w_full[ws]);
            if (++ws == (uint32_t)p.wslots) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == kSEpiWarps + 3) {
    // =========================== residual loader ===========================
    if (elect_one()) {
      const bool has_res = p.res_mode == RES_ADD;
      const int64_t rsb = pt_slab_bytes(p.res);
      uint32_t slot = 0, ph = 1;
      for (int64_t tile = (int64_t)blockIdx.x - p.cta0; tile < p.n_tiles; tile += p.ncta) {
        const int64_t f = tile / p.tiles_per_frame;
        const int q0 = (int)(tile - f * p.tiles_per_frame) * p.tile;
        for (int mt_i = 0; mt_i < p.mt; ++mt_i)
          for (int s = 0; s < ospp; ++s) {
            mbar_wait_relaxed(&st_empty[slot], ph);
            if (has_res) {
              mbar_expect_tx(&st_full[slot], stg_bytes);
              for (int pl = 0; pl < p.planes; ++pl)
                bulk_g2s(sS + slot * stg_bytes + (uint32_t)pl * kSPlane,
                         p.res.base + f * p.res.frame_bytes + (int64_t)(pl * p.res.spp + s) * rsb + (int64_t)(8 + q0 + mt_i * 128) * 128, kSPlane,
                         &st_full[slot]);
            } else {
              mbar_arrive(&st_full[slot]);
            }
            if (++slot == n_su) { slot = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == kSEpiWarps + 4) {
    // =========================== storer ===========================
    if (elect_one()) {
      const int64_t osb = pt_slab_bytes(p.out);
      uint32_t slot = 0, ph = 0;
      for (int64_t tile = (int64_t)blockIdx.x - p.cta0; tile < p.n_tiles; tile += p.ncta) {
        const int64_t f = tile / p.tiles_per_frame;
        const int q0 = (int)(tile - f * p.tiles_per_frame) * p.tile;
        uint8_t* oimg = p.out.base + f * p.out.frame_bytes;
        for (int mt_i = 0; mt_i < p.mt; ++mt_i)
          for (int s = 0; s < ospp; ++s) {
            mbar_wait_relaxed(&st_done[slot], ph);
            const int p0 = q0 + mt_i * 128;
            for (int pl = 0; pl < p.planes; ++pl) {
              const uint8_t* src = sS + slot * stg_bytes + (uint32_t)pl * kSPlane;
              if (!p.out.deint) {
                bulk_s2g(oimg + (int64_t)pt_slab_index(p.out, 0, pl, s) * osb + (int64_t)(8 + p0) * 128, src, kSPlane);
              } else {
                for (int sub = 0; sub < 2; ++sub)
                  bulk_s2g(oimg + (int64_t)pt_slab_index(p.out, sub, pl, s) * osb + (int64_t)(8 + p0 / 2) * 128, src + sub * (kSPlane / 2), kSPlane / 2);
              }
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the unit may be refilled once the copy engine has read it
            mbar_arrive(&st_empty[slot]);
            if (++slot == n_su) { slot = 0; ph ^= 1u; }
          }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");             // all stores complete before the CTA retires
    }
  } else if (warp >= kSEpiWarps + 6) {
    // =========================== Toeplitz producers (PK_GEN) ===========================
    if (gen) gen_producer(p, tid - (kSEpiWarps + 6) * 32, lane, sA, a_full, a_empty, s_xh, s_xl);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (pair) cluster_sync_all();   // nobody frees TMEM (or retires) while the leader's instruction stream can still touch it
  tc_fence_after();
  if (warp == kSEpiWarps) {
    if constexpr (pair) tmem_dealloc_cg2(tmem, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
  }
}

template <bool kPair>
__global__ void __launch_bounds__(kSThreads, 1) plane_xs_kernel(const __grid_constant__ XParams p) {
  pdl_trigger();
  xs_body<kPair, kSEpiGroups>(p);
}

// ================================================================================================
// Fused bottleneck block (nn_core_operator.py:57-79): ONE persistent launch, three CTA roles.
//   role 1  conv 1 (wide -> narrow, taps-in-N)        t_body<20, 9, 1>       reads the block's input x from HBM, writes h1 into ring 1
//   role 2  conv 2 (narrow -> narrow, grouped)        t_body<20, 5, 3>       ring 1 -> ring 2
//   role 3  conv 3 (narrow -> wide) + residual + act  xs_body<pair, 3>       ring 2 + x (second read: L2) -> the block's output y
// The rings hold a few dozen frames (plane_block_ring_frames(); a few MB, far below the 126 MB L2), so h1 and h2 are produced and
// consumed out of L2 and never written back before they are overwritten; frame hand-off is one flag per frame and direction
// (flag_wait / flag_signal / flag_bump above).  All CTAs of the launch are co-resident (grid <= SM count, one CTA per SM), roles are
// contiguous CTA ranges of even size so that the CTA pairs of role 3 stay pairs.  Why not shared memory: in the fp16 hi/lo
// representation the three convs' packed weights alone are 270 KB.
// ================================================================================================
constexpr int kBlockThreads = (12 + 6) * 32;      // 12 epilogue warps in every role + the 6 service warps of role 3

struct BlockParams {
  TParams p1, p2;
  XParams p3;
};

template <int TAPS2, int GROUPS2>
__global__ void __launch_bounds__(kBlockThreads, 1) plane_block_kernel(const __grid_constant__ BlockParams bp) {
  const int b = (int)blockIdx.x;
  if (b < bp.p2.cta0) t_body<20, 9, 1>(bp.p1);
  else if (b < bp.p3.cta0) t_body<20, TAPS2, GROUPS2>(bp.p2);
  else xs_body<true, 3>(bp.p3);
}

constexpr size_t kSmemBudget = 227 * 1024 - 2048;   // dynamic bytes we allow ourselves (alignment slack + static barriers)

struct TPlan {
  int N, n_wslab, ksteps, na;
  int pair;          // CTA pairs (cta_group::2): half of every weight slab per CTA, more input stages in flight
  int groups;        // 3: grouped form (narrow input), 1: all taps in N
  int stage_bytes;
  size_t smem;
};

bool plan_t(const PlaneConv& c, TPlan* pl, bool allow_pair = true) {
  const bool k9 = (c.Cout == 20 && c.K == 9 && c.dil >= 1 && c.dil <= 2);
  const bool k55 = (c.Cout == 1 && c.K == 55 && c.dil == 1);
  if (!k9 && !k55) return false;
  if (c.stride != 1 || c.shuffle != 1 || c.res_mode != RES_NONE || c.post_act != NSC_ACT_NONE) return false;
  if (k9 && c.act == NSC_ACT_TANH) return false;      // the 20-channel epilogue applies slope-form activations only
  if (c.Cin < 2 || c.in.deint || c.Lin % 128 != 0 || c.in.rows != c.Lin) return false;
  if (k9 && !c.fold_out && (!c.out.packed || c.out.deint || c.out.rows != c.Lin)) return false;
  if (c.fold_out && (!k9 || c.out.packed || c.out.spp != 1 || (c.fold_out == 2) != (c.out.deint != 0) || c.out.rows != (c.Lin >> c.fold_out) ||
                     (c.fold_out != 1 && c.fold_out != 2)))
    return false;
  // narrow -> narrow: tap groups by row shift, 5 (or 3) tap slots across lanes instead of 9 -- the lane-crossing volume bounds the layer.
  // Measured per 33k frames (both dilations): nine taps in N 21.1 ms, 3 groups / 5 slots 18.4 ms, 5 groups / 3 slots 21.0 ms on a box
  // where 3 groups took 19.3 (30 MMAs per tile from ONE issuing thread cost more than the lanes save).
  // NSC_PLANE_TGROUPS=3 (default) / 5 / 0 (all nine taps in N).  Wide inputs stay ungrouped: three times the MMAs would exceed their HBM time.
  static const int groups_knob = [] { const char* e = getenv("NSC_PLANE_TGROUPS"); const int v = e ? atoi(e) : 3; return (v == 3 || v == 5) ? v : 1; }();
  pl->groups = (k9 && c.in.packed) ? groups_knob : 1;
  pl->stage_bytes = pl->groups > 1 ? (128 + 16) * 128 : 128 * 128;
  const size_t kAStage = (size_t)pl->stage_bytes;
  pl->N = k9 ? (pl->groups == 5 ? 64 : pl->groups == 3 ? 112 : 192) : 64;
  pl->ksteps = (c.in.packed && c.planes == 2) ? 4 : (c.Cin + 15) / 16;   // packed hi/lo rows: one K axis of 64 halves
  pl->n_wslab = pl->groups > 1 ? pl->groups : (c.in.packed ? 1 : c.in.planes * c.in.spp);
  const int maxm = k9 ? 8 : 32, rowf = k9 ? 24 : 1;
  // CTA pairs for the wide-input 20-channel layers whose stages are short of the cap (100 -> 20: five stages -> eight).  Measured
  // NEUTRAL (17.3 vs 16.8 ms per 33k frames: the layer is not bound by bytes in flight after all, its tap-sum epilogue and MMAs each
  // fill about half of a tile's time), so only NSC_PLANE_PAIR=2 selects it.  Needs an even number of frames on an even grid: the two
  // CTAs walk their frames in lockstep.
  {
    static const int pair_knob = [] { const char* e = getenv("NSC_PLANE_PAIR"); return e ? atoi(e) : 1; }();
    const int64_t grid = c.B < sm_count() ? c.B : sm_count();
    const size_t full = (size_t)pl->n_wslab * pl->N * 128 + 2ull * t_slots(20) * maxm * rowf * sizeof(float) + (size_t)kORing * 128 + 1024;
    const bool short_of_stages = k9 && (kSmemBudget - full) / kAStage < 8;
    pl->pair = (allow_pair && pair_knob >= 2 && k9 && pl->groups == 1 && !c.in.packed && short_of_stages && c.B >= 2 && c.B % 2 == 0 && grid % 2 == 0) ? 1 : 0;
  }
  const size_t fixed = (size_t)pl->n_wslab * pl->N * (pl->pair ? 64 : 128) + 2ull * t_slots(k9 ? 20 : 1) * maxm * rowf * sizeof(float) + (k9 ? (size_t)kORing * 128 + 1024 : 0);
  if (fixed + 2ull * kAStage > kSmemBudget) return false;
  // (Two CTAs per SM for the narrow-input layers measured neutral -- 4.8 vs 4.8 ms per step on 20 -> 20 -- and were removed.)
  const size_t budget = kSmemBudget;
  size_t na = (budget - fixed) / kAStage;
  if (na > 8) na = 8;
  pl->na = (int)na;
  pl->smem = 1024 + fixed + na * kAStage;
  return true;
}

TParams make_tparams(const PlaneConv& c, const TPlan& pl, int cta0, int ncta) {
  TParams p;
  p.in = c.in; p.out = c.out;
  p.wpack = static_cast<const uint8_t*>(c.wpack); p.bias = c.bias; p.yvec = c.yvec;
  p.N = pl.N; p.n_wslab = pl.n_wslab; p.ksteps = pl.ksteps; p.stage_bytes = pl.stage_bytes; p.L = c.Lin; p.dil = c.dil; p.act = c.act; p.na = pl.na; p.B = c.B;
  p.cta0 = cta0; p.ncta = ncta;
  p.in_ready = p.in_free = p.out_ready = p.out_free = nullptr;
  p.out_free_target = 1;
  p.stats = nullptr;
  p.fold = c.fold;
  p.fold_out = c.fold_out;
  return p;
}

bool plan_x(const PlaneConv& c, XParams* p, size_t smem_budget = kSmemBudget) {
  const bool gen = c.kind == PK_GEN;
  int Lout, padL;
  same_padding(c.Lin, c.K, c.dil, c.stride, &Lout, &padL);
  if (Lout % 128 != 0) return false;
  if (c.shuffle != 1 && c.shuffle != 2) return false;
  if (c.Cout < 2 || c.Cout > 128 || c.Cout % c.shuffle != 0) return false;
  p->halo = 8;
  if (gen) {
    if (c.Cin != 1 || c.stride != 1 || c.dil != 1 || c.K > 64 || c.xvec == nullptr) return false;
  } else {
    if (c.Cin < 2 || c.Cin > 128) return false;
    if (c.stride == 1) { if (c.in.deint || c.in.rows != c.Lin) return false; }
    else if (c.stride == 2) { if (!c.in.deint || c.in.packed || c.dil != 1 || c.Lin % 2 != 0 || c.in.rows != c.Lin / 2) return false; }
    else return false;
    const int span = (c.K - 1) * c.dil;           // rows touched beyond the tile: [8 - padL, 8 - padL + span] must stay in [0, 16]
    if (c.stride == 1 && (padL > 8 || span - padL > 8)) {
      // packed input, plain epilogue: 16 halo rows (the neighbouring images' zero rows) -- the k15 dilation-2 gates at 128 positions
      if (!(c.in.packed && c.out.packed && padL <= 16 && span - padL <= 16)) return false;
      p->halo = 16;
    }
    if (c.stride == 2 && (padL > 16 || span - padL > 16)) return false;
  }
  if (c.res_mode == RES_ADD && (c.res.deint || c.res.rows != Lout)) return false;
  if (c.res_mode == RES_ADD_BCAST && c.resvec == nullptr) return false;
  if (c.res_mode == RES_MUL || (gen && c.res_mode != RES_NONE)) return false;
  if (c.glu && (gen || !c.in.packed || !c.out.packed || c.Cout != 40 || c.shuffle != 1 || c.res_mode != RES_NONE || c.stride != 1)) return false;
  // folded narrow images (plane.cuh)
  if (c.fold2 && (gen || c.Cin != kFoldC || c.Cout != kFoldC || c.K != (c.fold2 == 2 ? 9 : kFoldK) || c.dil != 1 || c.stride != 1 || c.in.packed || !c.out.packed ||
                  (c.fold2 == 2 && c.ileave) ||
                  c.planes != 2 || c.res_mode != RES_NONE || c.shuffle != 1 || c.glu || c.fold_out))
    return false;
  if (c.fold_out && (!gen || c.Cout != 20 || c.shuffle != 1 || c.glu || c.ileave || c.res_mode != RES_NONE || c.out.packed || c.out.spp != 1 ||
                     (c.fold_out != 1 && c.fold_out != 2) || (c.fold_out == 2) != (c.out.deint != 0) || c.out.rows != (Lout >> c.fold_out)))
    return false;
  if (c.fold2 && c.ileave && Lout != 128) return false;   // (the unfolding epilogue stages one frame = two 128-row parity tiles)
  const int unf = c.fold2 ? 2 : 1;                 // output positions per row of this layer
  if (c.ileave && (c.out.deint || c.shuffle != 1 || c.out.rows != 2 * unf * Lout)) return false;
  if (!c.ileave && !c.fold_out && c.out.rows != unf * (c.out.deint ? Lout * c.shuffle / 2 : Lout * c.shuffle)) return false;
  p->kind = c.kind;
  p->glu = c.glu; p->ileave = c.ileave; p->bias2 = c.bias2;
  p->fold_out = c.fold_out; p->unfold = c.fold2 ? 1 : 0;
  p->cta0 = 0; p->ncta = 1; p->in_ready = nullptr; p->in_free = nullptr; p->stats = nullptr;
  p->in = c.in; p->out = c.out; p->res = c.res;
  p->xvec = c.xvec; p->xsub = c.xsub; p->xscale = c.xscale; p->resvec = c.resvec;
  p->wpack = static_cast<const uint8_t*>(c.wpack); p->bias = c.bias;
  p->Lin = c.Lin; p->Lout = Lout; p->Cin = c.Cin; p->Cout = c.Cout; p->K = c.K; p->dil = c.dil; p->stride = c.stride; p->padL = padL;
  p->act = c.act; p->post_act = c.post_act; p->res_mode = c.res_mode; p->shuffle = c.shuffle; p->planes = c.planes;
  p->Npad = (c.Cout + 15) & ~15;
  p->ksteps = gen ? (c.K + 15) / 16 : ((c.in.packed && c.planes == 2) ? 4 : (c.Cin + 15) / 16);   // packed hi/lo rows: one K axis of 64 halves
  p->n_stage = gen ? c.planes : pt_n_slabs(c.in);
  p->unit_bytes = p->Npad * 128;
  p->n_units = gen ? c.planes : (c.in.packed ? c.K : c.in.spp * c.K * c.planes);
  if (p->n_stage > kXMaxStage) return false;
  // narrow-input, wide-output layers (the HBM-bound third conv of a block) take the staged epilogue
  static const bool no_stage = getenv("NSC_PLANE_NOSTAGE") != nullptr;
  p->staged = ((gen || c.in.packed) && c.shuffle == 1 && !c.out.packed && !c.fold_out && c.stride == 1 && !no_stage && c.act != NSC_ACT_TANH &&
               c.post_act != NSC_ACT_TANH) ? 1 : 0;   // (the staged epilogue applies slope-form activations only)
  // tile: two M-tiles per work unit when everything fits (halves the weight re-streaming of ring layers)
  // CTA pairs (cta_group::2) for the layers with a plain epilogue: each CTA of a pair keeps its own tile and half of every
  // weight unit (half the weight stream into each shared memory, half the instructions).  Needs an even number of work units,
  // otherwise the one-CTA kernel runs.  Measured (profiles/r01_pair_probe.log, cycles per tile): up-sampling conv (two M tiles,
  // two issuing threads) 25.5 k -> 23.0 k; stride-2 conv (one M tile) 19.1 k -> 20.5 k, its 11-slot half-unit ring runs dry more
  // often than the 5-slot ring did; 20 -> 20 is unchanged (an M = 256 instruction occupies BOTH tensor pipes for as long as an
  // M = 128 one occupies one, so the ~44-cycle floor of narrow instructions is not halved).  Default: pairs where two M tiles
  // share a weight pass; NSC_PLANE_PAIR=0 never, =2 wherever possible.
  static const int pair_knob = [] { const char* e = getenv("NSC_PLANE_PAIR"); return e ? atoi(e) : 1; }();
  p->pairtiles = (c.fold2 && c.ileave) ? 1 : 0;
  const size_t out_ring = c.fold2 ? (size_t)x_out_rows(c.ileave) * 128 : 0;      // staging window of the unfolding epilogue
  for (int mt = (Lout % 256 == 0 && c.stride == 1 && !c.fold2) ? 2 : 1; mt >= 1; --mt) {
    p->mt = mt;
    p->tile = 128 * mt;
    p->stage_bytes = (p->tile + 2 * (gen ? 8 : p->halo)) * 128;
    p->tiles_per_frame = Lout / p->tile;
    p->n_tiles = c.B * p->tiles_per_frame;
    // (CTA pairs do not help the folded narrow conv: it is bound by the issue rate of ONE thread, and a pair still has one -- measured)
    p->pair = (pair_knob >= (mt == 2 ? 1 : 2) && !gen && !c.glu && !c.fold2 && p->n_tiles >= 2 && p->n_tiles % 2 == 0) ? 1 : 0;
    p->slot_bytes = p->pair ? p->unit_bytes / 2 : p->unit_bytes;
    const size_t slot = (size_t)p->slot_bytes;
    const size_t a1 = (size_t)p->n_stage * p->stage_bytes;
    const size_t wall = (size_t)p->n_units * slot;
    if (2 * mt * p->Npad > 512) continue;
    // staged epilogue: ring of residual / output units.  A pair's half-size weight slots leave room for a fourth unit -- more
    // bytes in flight per SM, which is what bounds the HBM-bound layers
    p->s_units = p->staged ? ((p->pair && 2 * a1 + 3 * slot + (size_t)(kSUnits + 1) * c.planes * kSPlane <= smem_budget) ? kSUnits + 1 : kSUnits) : 0;
    const size_t budget = smem_budget - (size_t)p->s_units * c.planes * kSPlane - out_ring;
    // (the folded narrow conv is bound by bytes in flight per SM -- ncu: 20 % tensor activity, 43 % of HBM with two tile buffers --
    // and its pair form has room for a third)
    if (c.fold2 && p->n_units <= kXMaxW && 3 * a1 + wall <= budget) { p->kbuf = 3; p->resident = 1; p->wslots = p->n_units; }
    else if (p->n_units <= kXMaxW && 2 * a1 + wall <= budget) { p->kbuf = 2; p->resident = 1; p->wslots = p->n_units; }
    else if (2 * a1 + (p->pair ? 3 * slot : 4ull * p->unit_bytes) <= budget) { p->kbuf = 2; p->resident = 0; }
    else if (p->n_units <= kXMaxW && a1 + wall <= budget) { p->kbuf = 1; p->resident = 1; p->wslots = p->n_units; }
    else if (a1 + 3ull * p->unit_bytes <= budget) { p->kbuf = 1; p->resident = 0; }
    else continue;
    if (p->n_stage * p->kbuf > kXMaxStage) { if (p->kbuf == 2 && !p->resident && a1 + 3ull * p->unit_bytes <= budget) p->kbuf = 1; else continue; }
    if (!p->resident) {
      size_t ws = (budget - (size_t)p->kbuf * a1) / slot;
      if (ws > (size_t)kXMaxW) ws = kXMaxW;
      if (ws > (size_t)p->n_units) ws = p->n_units;
      p->wslots = (int)ws;
    }
    int cols = 32;
    while (cols < 2 * (c.fold2 ? 3 : mt) * p->Npad) cols *= 2;      // (folded conv: one accumulator per product, see x_issuer)
    p->tmem_cols = cols;
    p->B = c.B;
    // chunks of the output row the consumer's K steps read but this layer does not write
    const int out_c = c.Cout / c.shuffle;
    const int written = c.shuffle == 1 ? p->Npad / 8 : p->Npad / 16;
    int needed = ((out_c + 15) & ~15) / 8;
    if (c.out.packed) needed = 4;
    p->zero_from = written;
    p->zero_to = needed > written ? needed : written;
    if (c.glu) p->zero_from = p->zero_to = 0;          // the gate epilogue writes the whole packed row itself
    if (p->staged && (p->zero_from != p->zero_to || p->n_stage * p->kbuf > 4)) p->staged = 0;   // (never for the codec's shapes)
    if (!p->staged) p->s_units = 0;
    {
      // one issuing thread per M tile where the issue rate, not HBM, bounds the layer (measured); the staged layers are HBM-bound
      static const int knob = [] { const char* e = getenv("NSC_PLANE_ISSUERS"); return e ? atoi(e) : 0; }();
      p->n_iss = (p->mt == 2 && !p->staged) ? 2 : 1;
      if (knob == 1) p->n_iss = 1;
      if (knob == 2 && p->mt == 2) p->n_iss = 2;
      if (c.fold2 && p->resident && p->kbuf * p->n_stage <= kXMaxStage) p->n_iss = 3;
      // two issuing threads by weight plane for one-tile layers with many MMAs per tile (the stride-2 conv: 189).  EXPERIMENT, off unless
      // NSC_PLANE_PSPLIT=1: full bench runs hung intermittently while this and the barrier waits' suspend-time hint were both on
      // (DESIGN.md finding 20); with both off 16 of 16 ran through.
      static const bool psplit_knob = [] { const char* e = getenv("NSC_PLANE_PSPLIT"); return e && e[0] == '1'; }();
      p->psplit = (psplit_knob && !gen && !c.in.packed && c.planes == 2 && p->mt == 1 && !p->pair && !p->staged && !c.glu && !c.fold2 && c.K >= 3 &&
                   p->n_iss == 1 && 4 * p->Npad <= 512) ? 1 : 0;
      if (p->psplit) {
        p->n_iss = 2;
        int cols2 = 32;
        while (cols2 < 4 * p->Npad) cols2 *= 2;
        p->tmem_cols = cols2;
      }
    }
    return true;
  }
  return false;
}

size_t x_smem_bytes(const XParams& p) {
  return 1024 + (size_t)p.n_stage * p.kbuf * p.stage_bytes + (size_t)p.wslots * p.slot_bytes + (size_t)p.s_units * p.planes * kSPlane +
         (p.unfold ? (size_t)x_out_rows(p.ileave) * 128 : 0);
}

}  // namespace

int plane_narrow_kind() {
  static const int kind = [] {
    const char* e = getenv("NSC_PLANE_NARROW");
    if (e && (e[0] == 'X' || e[0] == 'x')) return (int)PK_X;
    if (e && (e[0] == 'T' || e[0] == 't')) return (int)PK_T;
    return (int)PK_T;
  }();
  return kind;
}

bool plane_conv_supported(const PlaneConv& c) {
  if (c.planes != 1 && c.planes != 2) return false;
  if (c.kind == PK_DW)
    return c.Cin == c.Cout && c.Cin > 32 && c.Cin <= 128 && c.K >= 1 && c.K <= 17 && (c.Lin & 7) == 0 && c.dil == 1 && c.stride == 1 && c.shuffle == 1 &&
           c.res_mode == RES_NONE && !c.in.packed && !c.in.deint && !c.out.packed && !c.out.deint && c.in.rows == c.Lin && c.out.rows == c.Lin;
  if (c.kind == PK_T) { TPlan pl; return plan_t(c, &pl); }
  XParams p;
  PlaneConv cc = c;
  cc.B = 1;
  return plan_x(cc, &p);
}

int64_t plane_wpack_bytes(const PlaneConv& c) {
  if (c.kind == PK_DW) return plane_conv_supported(c) ? 0 : -1;
  if (c.kind == PK_T) {
    TPlan pl;
    if (!plan_t(c, &pl)) return -1;
    return (int64_t)pl.n_wslab * pl.N * 128;
  }
  XParams p;
  PlaneConv cc = c;
  cc.B = 1;
  if (!plan_x(cc, &p)) return -1;
  return (int64_t)p.n_units * p.unit_bytes + (c.fold2 ? plane_fold2_scratch_bytes(c.K) : 0);
}

bool plane_plan_info(const PlaneConv& c, int64_t* o) {
  for (int i = 0; i < 12; ++i) o[i] = 0;
  o[0] = c.kind;
  if (c.kind == PK_DW) return plane_conv_supported(c);
  if (c.kind == PK_T) {
    TPlan pl;
    if (!plan_t(c, &pl)) return false;
    o[1] = pl.groups > 1 ? pl.groups : 0;    // (taps-in-N: the "staged" field reports the tap groups of the grouped form)
    o[2] = pl.pair; o[3] = 1; o[4] = 1; o[5] = 1; o[6] = pl.n_wslab; o[7] = pl.na; o[8] = (int64_t)pl.smem;
    int cols = 32;
    while (cols < 2 * pl.N) cols *= 2;
    o[9] = cols;
    o[10] = c.B < sm_count() ? c.B : sm_count();
    o[11] = c.B;
    return true;
  }
  XParams p;
  if (!plan_x(c, &p)) return false;
  const int64_t n_work = p.pairtiles ? p.n_tiles / 2 : p.n_tiles;
  int64_t grid = n_work < sm_count() ? n_work : sm_count();
  if (p.pair) grid &= ~(int64_t)1;
  o[1] = p.staged; o[2] = p.pair; o[3] = p.mt; o[4] = p.n_iss; o[5] = p.resident; o[6] = p.wslots; o[7] = (int64_t)p.n_stage * p.kbuf;
  o[8] = (int64_t)x_smem_bytes(p); o[9] = p.tmem_cols; o[10] = grid; o[11] = p.n_tiles;
  return true;
}

int plane_pack_weights(const PlaneConv& c, cudaStream_t st) {
  if (c.kind == PK_DW) return NSC_OK;      // the depthwise FIR reads its fp32 taps directly
  NSC_CHECK_ARG(c.w != nullptr && c.wpack != nullptr, "plane engine: null weights");
  NSC_CHECK_ARG(!c.glu || c.w2 != nullptr, "plane engine: gated layer without its second kernel");
  PackArgs a;
  a.w = c.w; a.out = static_cast<__half*>(c.wpack);
  if (c.fold2) {
    // (9, 20, 20) -> (5, 48, 48) + bias (48) behind the packed units, then packed like any 48 -> 48 k5 layer
    XParams p;
    PlaneConv cc = c;
    cc.B = 1;
    NSC_CHECK_ARG(plan_x(cc, &p), "plane engine: unsupported folded layer");
    float* scratch = reinterpret_cast<float*>(static_cast<uint8_t*>(c.wpack) + (size_t)p.n_units * p.unit_bytes);
    fold2_weights_kernel<<<(c.K * kFoldC * kFoldC + kFoldC + 255) / 256, 256, 0, st>>>(c.w, c.bias, scratch, c.K);
    NSC_LAUNCH_OK();
    a.w = scratch;
  }
  a.w2 = c.glu ? c.w2 : nullptr; a.csplit = 20;
  a.kind = c.kind; a.K = c.K; a.Cin = c.Cin; a.Cout = c.Cout; a.planes = c.planes;
  a.in_packed = c.kind == PK_GEN ? 0 : c.in.packed;
  a.in_spp = c.kind == PK_GEN ? 1 : c.in.spp;
  a.C = c.Cout;
  a.groups = 1;
  if (c.kind == PK_T) {
    TPlan pl;
    NSC_CHECK_ARG(plan_t(c, &pl), "plane engine: unsupported taps-in-N layer (k%d d%d %d->%d)", c.K, c.dil, c.Cin, c.Cout);
    a.rows = pl.N; a.n_units = pl.n_wslab; a.groups = pl.groups;
  } else {
    XParams p;
    PlaneConv cc = c;
    cc.B = 1;
    NSC_CHECK_ARG(plan_x(cc, &p), "plane engine: unsupported layer (k%d d%d s%d %d->%d)", c.K, c.dil, c.stride, c.Cin, c.Cout);
    a.rows = p.Npad; a.n_units = p.n_units;
  }
  const int64_t total = (int64_t)a.n_units * a.rows * 64;
  const int blocks = (int)((total + 255) / 256 < 592 ? (total + 255) / 256 : 592);
  ProfScope prof(st, "plane_pack_weights", 0.0, 4.0 * c.K * c.Cin * c.Cout + 2.0 * total);
  plane_pack_kernel<<<blocks, 256, 0, st>>>(a);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

// Every conv kernel of the plane engine is launched with programmatic stream serialization (NSC_PLANE_PDL=0: plain launches): its
// prologue runs under the tail of the kernel before it, the rest after pdl_wait().  Worth 1.2 % of a large step and 15 % of a
// 128-frame call.  (It was switched off for a few hours at the end of round 2 as the suspect of a bit-identity failure that turned
// out to be the opt-in fused block kernel's own -- DESIGN.md findings 11 and 19.)
template <typename P>
static cudaError_t launch_plane(void (*kernel)(P), int64_t grid, int threads, size_t smem, cudaStream_t st, int cluster, const P& p) {
  static const bool pdl = [] { const char* e = getenv("NSC_PLANE_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

int plane_launch(const PlaneConv& c, cudaStream_t st) {
  if (c.B == 0) return NSC_OK;
  if (c.kind == PK_DW) {
    NSC_CHECK_ARG(plane_conv_supported(c) && c.w != nullptr, "plane engine: unsupported depthwise layer (k%d %d channels)", c.K, c.Cin);
    char dname[32];
    snprintf(dname, sizeof(dname), "pD%d_k%d_c%d", c.planes, c.K, c.Cin);
    ProfScope prof(st, dname, 2.0 * (double)c.B * c.Lin * c.K * c.Cin, 2.0 * (double)c.B * pt_real_bytes(c.Lin, c.Cin, c.planes));
    const int64_t total = c.B * (int64_t)(c.Lin / 8) * (c.in.spp * 8);
    const int blocks = (int)((total + 255) / 256 < (int64_t)sm_count() * 8 ? (total + 255) / 256 : (int64_t)sm_count() * 8);
    plane_depthwise_kernel<<<blocks, 256, 0, st>>>(c.in, c.out, c.w, c.K, c.Cin, c.Lin, c.B);
    NSC_LAUNCH_OK();
    return NSC_OK;
  }
  NSC_CHECK_ARG(c.wpack != nullptr, "plane engine: weights not packed");
  int Lout, padL;
  same_padding(c.Lin, c.K, c.dil, c.stride, &Lout, &padL);
  const double macs = (double)c.B * Lout * c.K * c.Cin * c.Cout;
  char name[32];
  if (c.kind == PK_T) {
    TPlan pl;
    NSC_CHECK_ARG(plan_t(c, &pl), "plane engine: unsupported taps-in-N layer (k%d d%d %d->%d)", c.K, c.dil, c.Cin, c.Cout);
    NSC_CHECK_ARG(c.Cout > 1 || c.yvec != nullptr || c.fold.acc != nullptr || c.fold.q_code != nullptr || c.fold.q_idx != nullptr,
                  "plane engine: head without an output vector");
    NSC_CHECK_ARG(c.fold.q_bins == nullptr || (c.fold.q_n >= 1 && c.fold.q_n <= 256 && c.fold.q_alpha != nullptr), "plane engine: bad quantiser fold");
    const int64_t grid = c.B < sm_count() ? c.B : sm_count();
    const TParams p = make_tparams(c, pl, 0, (int)grid);
    snprintf(name, sizeof(name), "pT%d_k%dd%d_c%dto%d", c.planes, c.K, c.dil, c.Cin, c.Cout);
    // algorithmic bytes: every input and output plane image moved exactly once (zero rows excluded)
    const double bytes = (double)c.B * (pt_real_bytes(c.Lin, c.Cin, c.planes) + (c.Cout > 1 ? pt_real_bytes(c.Lin, c.Cout, c.planes) : 4.0 * c.Lin));
    ProfScope prof(st, name, 2.0 * macs, bytes);
    if (c.Cout == 20 && pl.groups == 5) {
      NSC_CUDA_OK(cudaFuncSetAttribute(plane_t_kernel<20, 3, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
      NSC_CUDA_OK(launch_plane(plane_t_kernel<20, 3, 5>, grid, TShape<20, 3>::kThreads, pl.smem, st, 1, p));
    } else if (c.Cout == 20 && pl.groups == 3) {
      NSC_CUDA_OK(cudaFuncSetAttribute(plane_t_kernel<20, 5, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
      NSC_CUDA_OK(launch_plane(plane_t_kernel<20, 5, 3>, grid, TShape<20, 5>::kThreads, pl.smem, st, 1, p));
    } else if (c.Cout == 20 && pl.pair) {
      NSC_CUDA_OK(cudaFuncSetAttribute(plane_t_kernel<20, 9, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
      NSC_CUDA_OK(launch_plane(plane_t_kernel<20, 9, 1, true>, grid, TShape<20, 9>::kThreads, pl.smem, st, 2, p));
    } else if (c.Cout == 20) {
      NSC_CUDA_OK(cudaFuncSetAttribute(plane_t_kernel<20, 9, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
      NSC_CUDA_OK(launch_plane(plane_t_kernel<20, 9, 1>, grid, TShape<20, 9>::kThreads, pl.smem, st, 1, p));
    } else {
      NSC_CUDA_OK(cudaFuncSetAttribute(plane_t_kernel<1, 55, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
      NSC_CUDA_OK(launch_plane(plane_t_kernel<1, 55, 1>, grid, TShape<1, 55>::kThreads, pl.smem, st, 1, p));
    }
    NSC_LAUNCH_OK();
    return NSC_OK;
  }
  XParams p;
  NSC_CHECK_ARG(plan_x(c, &p), "plane engine: unsupported layer (k%d d%d s%d %d->%d)", c.K, c.dil, c.stride, c.Cin, c.Cout);
  if (c.fold2 && getenv("NSC_FOLD_STATS") != nullptr) {   // experiment: issue / wait counters of the folded conv (nsc_debug_block_stats)
    void* sp = nullptr;
    NSC_CUDA_OK(cudaGetSymbolAddress(&sp, g_block_stats));
    NSC_CUDA_OK(cudaMemsetAsync(sp, 0, sizeof(unsigned long long) * kStatCtas * kStatWords, st));
    p.stats = static_cast<unsigned long long*>(sp);
  }
  if (c.fold2)   // the folded biases sit behind the folded kernel, behind the packed units (plane_pack_weights)
    p.bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(c.wpack) + (size_t)p.n_units * p.unit_bytes) + c.K * kFoldC * kFoldC;
  const size_t smem = x_smem_bytes(p);
  const int64_t n_work = p.pairtiles ? p.n_tiles / 2 : p.n_tiles;      // (pairtiles: a CTA's unit of work is a pair of tiles)
  {
    int64_t g = n_work < sm_count() ? n_work : sm_count();
    if (p.pair) g &= ~(int64_t)1;
    p.ncta = (int)g;
  }
  if (p.staged && p.pair) NSC_CUDA_OK(cudaFuncSetAttribute(plane_xs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (p.staged) NSC_CUDA_OK(cudaFuncSetAttribute(plane_xs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (c.kind == PK_GEN) NSC_CUDA_OK(cudaFuncSetAttribute(plane_x_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (p.pair) NSC_CUDA_OK(cudaFuncSetAttribute(plane_x_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else if (p.unfold) NSC_CUDA_OK(cudaFuncSetAttribute(plane_x_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else NSC_CUDA_OK(cudaFuncSetAttribute(plane_x_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  snprintf(name, sizeof(name), "p%s%d_k%dd%ds%d_c%dto%d", c.kind == PK_GEN ? "G" : (c.glu ? (c.ileave ? "U2x" : "U") : "X"), c.planes, c.K, c.dil, c.stride,
           c.Cin, c.Cout);
  double bytes = (double)c.B * (c.kind == PK_GEN ? 4.0 * c.Lin : pt_real_bytes(c.Lin, c.Cin, c.planes));
  bytes += (double)c.B * pt_real_bytes(Lout, c.Cout, c.planes);
  double flops = 2.0 * macs;
  if (c.fold2) {   // the k9 20 -> 20 conv it stands for: its real work and payload, under its own name
    snprintf(name, sizeof(name), "pF%d_k9d%d_c20to20", c.planes, (c.ileave || c.fold2 == 2) ? 2 : 1);
    flops = 2.0 * (double)c.B * (2.0 * Lout) * 9.0 * 20.0 * 20.0;
    bytes = 2.0 * (double)c.B * pt_real_bytes(2 * Lout, 20, c.planes);
  }
  if (c.res_mode == RES_ADD) bytes += (double)c.B * pt_real_bytes(Lout, c.Cout, c.planes);
  if (c.res_mode == RES_ADD_BCAST) bytes += (double)c.B * 4.0 * Lout;
  ProfScope prof(st, name, flops, bytes);
  int64_t grid = n_work < sm_count() ? n_work : sm_count();
  if (p.pair) grid &= ~(int64_t)1;
  if (p.pair && p.staged) NSC_CUDA_OK(launch_plane(plane_xs_kernel<true>, grid, kSThreads, smem, st, 2, p));
  else if (p.pair) NSC_CUDA_OK(launch_plane(plane_x_kernel<false, true>, grid, kXThreadsX, smem, st, 2, p));
  else if (p.staged) NSC_CUDA_OK(launch_plane(plane_xs_kernel<false>, grid, kSThreads, smem, st, 1, p));
  else if (c.kind == PK_GEN) NSC_CUDA_OK(launch_plane(plane_x_kernel<true, false>, grid, kXThreadsGen, smem, st, 1, p));
  else if (p.unfold) NSC_CUDA_OK(launch_plane(plane_x_kernel<false, false, true>, grid, kXThreadsX, smem, st, 1, p));
  else NSC_CUDA_OK(launch_plane(plane_x_kernel<false, false>, grid, kXThreadsX, smem, st, 1, p));
  NSC_LAUNCH_OK();
  return NSC_OK;
}

// ---- fused bottleneck block ---------------------------------------------------------------------------------------------
int plane_block_ring_frames() {
  static const int v = [] {
    const char* e = getenv("NSC_BLOCK_RING");
    int n = e ? atoi(e) : 128;
    if (n < 2) n = 2;
    int r = 2;
    while (r * 2 <= n && r < 4096) r *= 2;
    return r;
  }();
  return v;
}

int64_t plane_block_flag_words(int64_t B) { return 4 * align_up(B < 1 ? 1 : B, 32); }

namespace {

// CTAs per role.  Defaults are measured (tools/block_stats.py, profiles/r02_block_stats*.log): per 128-position tile role 1 costs
// ~3.6 k cycles (shared-memory bandwidth of its N = 192 instructions), role 2 ~2.6 k (tap-sum epilogue), role 3 ~4.4 k (staging
// traffic of the residual / output units) for a 100-channel block; a 50-channel block is cheaper on role 3.
// NSC_BLOCK_SPLIT="n1,n2" overrides (even numbers; role 3 takes the rest).
void block_split(const PlaneBlock& b, int* n1, int* n2, int* n3) {
  const int sms = sm_count() & ~1;
  static const char* knob = getenv("NSC_BLOCK_SPLIT");
  int a = 0, c = 0;
  if (knob && sscanf(knob, "%d,%d", &a, &c) == 2 && a >= 2 && c >= 2 && a + c + 2 <= sms) { a &= ~1; c &= ~1; }
  else if (b.c1.Cin >= 64) { a = (int)(sms * 50 / 148) & ~1; c = (int)(sms * 36 / 148) & ~1; }
  else { a = (int)(sms * 52 / 148) & ~1; c = (int)(sms * 44 / 148) & ~1; }
  if (a < 2) a = 2;
  if (c < 2) c = 2;
  *n1 = a; *n2 = c; *n3 = sms - a - c;
}

struct BlockPlan {
  TPlan t1, t2;
  XParams x3;
  size_t smem;
};

bool plan_block(const PlaneBlock& b, BlockPlan* bp) {
  const PlaneConv &c1 = b.c1, &c2 = b.c2, &c3 = b.c3;
  if (c1.kind != PK_T || c2.kind != PK_T || c3.kind != PK_X) return false;
  if (c1.Cout != 20 || c2.Cin != 20 || c2.Cout != 20 || c3.Cin != 20 || c1.Cin != c3.Cout) return false;
  if (c3.res_mode != RES_ADD || c1.planes != c2.planes || c1.planes != c3.planes) return false;
  if (c1.B != c2.B || c1.B != c3.B || c1.B < 1) return false;
  if (b.ring < 2 || (b.ring & (b.ring - 1))) return false;
  if (sm_count() < 8) return false;
  if (!plan_t(c1, &bp->t1, false) || !plan_t(c2, &bp->t2, false) || bp->t2.groups != 3) return false;   // (the roles of the fused kernel are single CTAs)
  if (!plan_x(c3, &bp->x3, kSmemBudget - 2048) || !bp->x3.staged || !bp->x3.pair) return false;   // (the fused kernel's static shared memory is the sum of its roles')   // (pairs need an even number of tiles)
  bp->smem = bp->t1.smem;
  if (bp->t2.smem > bp->smem) bp->smem = bp->t2.smem;
  if (x_smem_bytes(bp->x3) > bp->smem) bp->smem = x_smem_bytes(bp->x3);
  return true;
}

}  // namespace

bool plane_block_supported(const PlaneBlock& b) {
  BlockPlan bp;
  return plan_block(b, &bp);
}

// Whether the codec program launches its bottleneck blocks fused.  Measured on the headline workload (profiles/r02_block_*.log,
// DESIGN.md section 4): the fused launch keeps the intermediates out of HBM and is bit-identical, but its three roles are bound by
// shared-memory bandwidth / the tap-sum epilogue rather than by HBM, and under the 1 kW power cap the step is 8 % SLOWER with it
// (124 vs 113 ms) -- so it is opt-in (NSC_BLOCK_FUSED=1); the operator surface (nsc_bottleneck_block_tc) uses it.
bool plane_block_default_on() {
  static const bool on = [] { const char* e = getenv("NSC_BLOCK_FUSED"); return e && e[0] == '1'; }();
  return on;
}

int plane_block_launch(const PlaneBlock& b, cudaStream_t st) {
  BlockPlan pl;
  NSC_CHECK_ARG(plan_block(b, &pl), "plane engine: block not covered by the fused kernel");
  NSC_CHECK_ARG(b.flags != nullptr && b.c1.wpack && b.c2.wpack && b.c3.wpack, "plane engine: fused block without flags / packed weights");
  const int64_t B = b.c1.B;
  const int64_t fw = align_up(B, 32);
  uint32_t* h1_ready = b.flags;
  uint32_t* h1_free = b.flags + fw;
  uint32_t* h2_ready = b.flags + 2 * fw;
  uint32_t* h2_free = b.flags + 3 * fw;
  int n1, n2, n3;
  block_split(b, &n1, &n2, &n3);
  BlockParams bp;
  PlaneConv c1 = b.c1, c2 = b.c2, c3 = b.c3;
  c1.out.ring = b.ring; c2.in.ring = b.ring; c2.out.ring = b.ring; c3.in.ring = b.ring;
  bp.p1 = make_tparams(c1, pl.t1, 0, n1);
  bp.p1.out_ready = h1_ready; bp.p1.out_free = h1_free; bp.p1.out_free_target = 1;
  bp.p2 = make_tparams(c2, pl.t2, n1, n2);
  bp.p2.in_ready = h1_ready; bp.p2.in_free = h1_free;
  bp.p2.out_ready = h2_ready; bp.p2.out_free = h2_free; bp.p2.out_free_target = pl.x3.tiles_per_frame;
  bp.p3 = pl.x3;
  bp.p3.in = c3.in;
  bp.p3.cta0 = n1 + n2; bp.p3.ncta = n3;
  bp.p3.in_ready = h2_ready; bp.p3.in_free = h2_free;
  static const bool want_stats = getenv("NSC_BLOCK_STATS") != nullptr;
  if (want_stats) {
    void* sp = nullptr;
    NSC_CUDA_OK(cudaGetSymbolAddress(&sp, g_block_stats));
    NSC_CUDA_OK(cudaMemsetAsync(sp, 0, sizeof(unsigned long long) * kStatCtas * kStatWords, st));
    bp.p1.stats = bp.p2.stats = bp.p3.stats = static_cast<unsigned long long*>(sp);
  }
  auto kern = plane_block_kernel<5, 3>;
  static size_t static_smem = 0;
  if (static_smem == 0) {
    cudaFuncAttributes fa;
    NSC_CUDA_OK(cudaFuncGetAttributes(&fa, kern));
    static_smem = fa.sharedSizeBytes;
  }
  NSC_CHECK_ARG(pl.smem + static_smem <= 227 * 1024, "plane engine: fused block needs %zu + %zu bytes of shared memory", pl.smem, static_smem);
  NSC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int Lout, padL;
  same_padding(c1.Lin, c1.K, 1, 1, &Lout, &padL);
  const double macs = (double)B * Lout * ((double)c1.K * c1.Cin * c1.Cout + (double)c2.K * c2.Cin * c2.Cout + (double)c3.K * c3.Cin * c3.Cout);
  char name[32];
  snprintf(name, sizeof(name), "pB%d_d%d_c%d_L%d", c1.planes, c2.dil, c1.Cin, c1.Lin);
  // algorithmic bytes: the block's input and output images, each moved once (the residual read and both intermediates stay in L2)
  ProfScope prof(st, name, 2.0 * macs, (double)B * (pt_real_bytes(c1.Lin, c1.Cin, c1.planes) + pt_real_bytes(Lout, c3.Cout, c3.planes)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n1 + n2 + n3));
  cfg.blockDim = dim3(kBlockThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NSC_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, bp));
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int plane_from_f32(const float* x, int x_cl, int64_t B, int L, int C, const PlaneTensor& t, cudaStream_t st) {
  if (B == 0) return NSC_OK;
  const int64_t total = B * (int64_t)pt_chunks_per_row(t) * L;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  ProfScope prof(st, "plane_from_f32", 0.0, (double)B * (4.0 * L * C + t.frame_bytes));
  plane_from_f32_kernel<<<blocks, 256, 0, st>>>(x, x_cl, B, L, C, t);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

int plane_to_f32(const PlaneTensor& t, float* y, int y_cl, int64_t B, int L, int C, cudaStream_t st) {
  if (B == 0) return NSC_OK;
  const int64_t total = B * (int64_t)((C + 7) / 8) * L;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  ProfScope prof(st, "plane_to_f32", 0.0, (double)B * (4.0 * L * C + t.frame_bytes));
  plane_to_f32_kernel<<<blocks, 256, 0, st>>>(t, y, y_cl, B, L, C);
  NSC_LAUNCH_OK();
  return NSC_OK;
}

}  // namespace nsc

// ---- conv1d on the plane engine with channels-last fp32 tensors at the API edge -------------------------------
namespace {

struct TcConvPlan {
  nsc::PlaneConv c;
  int Lout, Ly, Cy;
  int64_t in_bytes, out_bytes, res_bytes, w_bytes;
};

int make_tc_conv_plan(int64_t B, int Lin, int Cin, int Cout, int k, int dil, int stride, int act, int res_mode, int post_act,
                      int shuffle, int precision, TcConvPlan* pl) {
  using namespace nsc;
  NSC_CHECK_ARG(precision == 1 || precision == 2, "nsc_conv1d_tc: precision must be 1 (fp16 hi/lo) or 2 (fp16)");
  NSC_CHECK_ARG(Lin > 0 && Cin > 0 && Cout > 0 && k > 0 && dil > 0 && (stride == 1 || stride == 2) && shuffle >= 1, "nsc_conv1d_tc: bad shape");
  PlaneConv& c = pl->c;
  int padL;
  same_padding(Lin, k, dil, stride, &pl->Lout, &padL);
  c.kind = Cin == 1 ? PK_GEN : (((Cout == 20 && k == 9) || (Cout == 1 && k == 55)) && res_mode == RES_NONE && stride == 1 && shuffle == 1) ? PK_T : PK_X;
  if (c.kind == PK_T && Cin <= 32 && Cout == 20 && plane_narrow_kind() == PK_X) c.kind = PK_X;   // same choice as the codec program
  c.Lin = Lin; c.Cin = Cin; c.Cout = Cout; c.K = k; c.dil = dil; c.stride = stride;
  c.act = act; c.post_act = post_act; c.res_mode = res_mode; c.shuffle = shuffle;
  c.planes = precision == 1 ? 2 : 1;
  c.B = B;
  pl->Ly = pl->Lout * shuffle;
  pl->Cy = Cout / shuffle;
  if (Cin > 1) c.in = make_plane_tensor(nullptr, Lin, Cin, c.planes, stride == 2);
  if (Cout > 1) c.out = make_plane_tensor(nullptr, pl->Ly, pl->Cy, c.planes, 0);
  if (res_mode == RES_ADD) c.res = make_plane_tensor(nullptr, pl->Lout, Cout, c.planes, 0);
  const float dummy = 0.f;
  if (Cin == 1) c.xvec = &dummy;
  if (res_mode == RES_ADD_BCAST) c.resvec = &dummy;
  NSC_CHECK_ARG(plane_conv_supported(c), "nsc_conv1d_tc: shape not covered by the tensor engine (k%d d%d s%d %d->%d, L %d)", k, dil, stride, Cin, Cout, Lin);
  pl->in_bytes = Cin > 1 ? align_up(B * c.in.frame_bytes, 1024) : 0;
  pl->out_bytes = Cout > 1 ? align_up(B * c.out.frame_bytes, 1024) : 0;
  pl->res_bytes = res_mode == RES_ADD ? align_up(B * c.res.frame_bytes, 1024) : 0;
  pl->w_bytes = align_up(plane_wpack_bytes(c), 1024);
  return NSC_OK;
}

}  // namespace

extern "C" {

int64_t nsc_conv1d_tc_workspace_bytes(int64_t B, int32_t Lin, int32_t Cin, int32_t Cout, int32_t k, int32_t dilation,
                                      int32_t stride, int32_t res_mode, int32_t shuffle, int32_t precision) {
  TcConvPlan pl;
  if (make_tc_conv_plan(B < 1 ? 1 : B, Lin, Cin, Cout, k, dilation, stride, 0, res_mode, 0, shuffle, precision, &pl) != NSC_OK) return -1;
  return 1024 + pl.in_bytes + pl.out_bytes + pl.res_bytes + pl.w_bytes;
}

int nsc_conv1d_tc_plan_info(int64_t B, int32_t Lin, int32_t Cin, int32_t Cout, int32_t k, int32_t dilation, int32_t stride,
                            int32_t res_mode, int32_t shuffle, int32_t precision, int64_t* out12) {
  NSC_CHECK_ARG(out12 != nullptr, "nsc_conv1d_tc_plan_info: null output");
  TcConvPlan pl;
  NSC_TRY(make_tc_conv_plan(B < 1 ? 1 : B, Lin, Cin, Cout, k, dilation, stride, 0, res_mode, 0, shuffle, precision, &pl));
  NSC_CHECK_ARG(nsc::plane_plan_info(pl.c, out12), "nsc_conv1d_tc_plan_info: layer not planned");
  return NSC_OK;
}

// launch plan of the block's 20 -> 20 conv as the codec program runs it at L positions (folded where the frame is long enough,
// else the taps-in-N kernel); out12 as nsc_conv1d_tc_plan_info, out12[1] = 4 + form for the folded forms (5: k5, 6: block-diagonal k9)
int nsc_narrow_conv_plan_info(int64_t B, int32_t L, int32_t dilation, int64_t* out12) {
  using namespace nsc;
  NSC_CHECK_ARG(out12 != nullptr && L > 0 && L % 128 == 0 && (dilation == 1 || dilation == 2), "nsc_narrow_conv_plan_info: bad arguments");
  if (B < 1) B = 1;
  const bool by_parity = dilation == 2 && L / 4 >= 128;
  if (L / 2 < 128) return nsc_conv1d_tc_plan_info(B, L, 20, 20, 9, dilation, 1, 0, 1, 1, out12);
  const bool diag = dilation == 2 && !by_parity;
  PlaneConv c;
  c.kind = PK_X; c.Lin = by_parity ? L / 4 : L / 2; c.Cin = kFoldC; c.Cout = kFoldC; c.K = diag ? 9 : kFoldK; c.dil = 1; c.stride = 1;
  c.act = NSC_ACT_LRELU; c.planes = 2; c.B = by_parity ? 2 * B : B; c.fold2 = diag ? 2 : 1;
  c.in = make_plane_tensor(nullptr, c.Lin, kFoldC, 2, 0);
  c.out = make_plane_tensor(nullptr, L, 20, 2, 0);
  if (by_parity) { c.ileave = 1; c.bmul = 2; }
  NSC_CHECK_ARG(plane_plan_info(c, out12), "nsc_narrow_conv_plan_info: layer not planned");
  out12[1] = 4 + c.fold2;
  return NSC_OK;
}

int nsc_conv1d_tc(const float* x, const float* w, const float* b, const float* res, float* y, int64_t B, int32_t Lin,
                  int32_t Cin, int32_t Cout, int32_t k, int32_t dilation, int32_t stride, int32_t activation,
                  int32_t res_mode, int32_t post_activation, int32_t shuffle, int32_t precision, void* workspace,
                  int64_t workspace_bytes, void* stream) {
  using namespace nsc;
  if (B == 0) return NSC_OK;
  NSC_CHECK_ARG(x && w && y && workspace, "nsc_conv1d_tc: null pointer");
  NSC_CHECK_ARG(res_mode == RES_NONE || res != nullptr, "nsc_conv1d_tc: residual mode %d without a residual", res_mode);
  TcConvPlan pl;
  NSC_TRY(make_tc_conv_plan(B, Lin, Cin, Cout, k, dilation, stride, activation, res_mode, post_activation, shuffle, precision, &pl));
  const int64_t need = 1024 + pl.in_bytes + pl.out_bytes + pl.res_bytes + pl.w_bytes;
  if (workspace_bytes < need) {
    set_error("nsc_conv1d_tc: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return NSC_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  PlaneConv& c = pl.c;
  NSC_CUDA_OK(cudaMemsetAsync(p, 0, (size_t)(pl.in_bytes + pl.out_bytes + pl.res_bytes), st));   // the images' zero rows
  c.in.base = p; p += pl.in_bytes;
  c.out.base = p; p += pl.out_bytes;
  c.res.base = p; p += pl.res_bytes;
  c.wpack = p;
  c.w = w; c.bias = b;
  c.xvec = Cin == 1 ? x : nullptr;
  c.resvec = res_mode == RES_ADD_BCAST ? res : nullptr;
  c.yvec = Cout == 1 ? y : nullptr;
  if (Cin > 1) NSC_TRY(plane_from_f32(x, 1, B, Lin, Cin, c.in, st));
  if (res_mode == RES_ADD) NSC_TRY(plane_from_f32(res, 1, B, pl.Lout, Cout, c.res, st));
  NSC_TRY(plane_pack_weights(c, st));
  NSC_TRY(plane_launch(c, st));
  if (Cout > 1) NSC_TRY(plane_to_f32(c.out, y, 1, B, pl.Ly, pl.Cy, st));
  return NSC_OK;
}

// ---- the_bottleneck (nn_core_operator.py:57-79) on the plane engine, channels-last fp32 tensors at the edge -----------------
namespace {

struct TcBlockPlan {
  nsc::PlaneBlock b;
  int64_t x_bytes, y_bytes, n_bytes, w1, w2, w3, flag_bytes;
};

int make_tc_block_plan(int64_t B, int L, int wide, int narrow, int k_plain, int k_dilated, int dilation, int is_last_flat, int precision,
                       TcBlockPlan* pl, bool folded = false) {
  using namespace nsc;
  NSC_CHECK_ARG(precision == 1 || precision == 2, "nsc_bottleneck_block_tc: precision must be 1 (fp16 hi/lo) or 2 (fp16)");
  NSC_CHECK_ARG(B >= 1 && L > 0 && L % 128 == 0 && wide > 32 && wide <= 128 && narrow == 20 && k_plain == 9 && k_dilated == 9 &&
                    dilation >= 1 && dilation <= 2,
                "nsc_bottleneck_block_tc: shape not covered by the tensor engine (L %d, %d/%d, k %d/%d, dilation %d)", L, wide, narrow, k_plain,
                k_dilated, dilation);
  const int P = precision == 1 ? 2 : 1;
  PlaneBlock& b = pl->b;
  auto conv = [&](PlaneConv& c, int kind, int cin, int cout, int K, int dil, int act, int res_mode, int post) {
    c.kind = kind; c.Lin = L; c.Cin = cin; c.Cout = cout; c.K = K; c.dil = dil; c.stride = 1;
    c.act = act; c.post_act = post; c.res_mode = res_mode; c.shuffle = 1; c.planes = P; c.B = B;
  };
  conv(b.c1, PK_T, wide, narrow, k_plain, 1, NSC_ACT_LRELU, RES_NONE, NSC_ACT_NONE);
  conv(b.c2, PK_T, narrow, narrow, k_dilated, dilation, NSC_ACT_LRELU, RES_NONE, NSC_ACT_NONE);
  conv(b.c3, PK_X, narrow, wide, k_plain, 1, NSC_ACT_NONE, RES_ADD, is_last_flat ? NSC_ACT_NONE : NSC_ACT_LRELU);
  const PlaneTensor tx = make_plane_tensor(nullptr, L, wide, P, 0), tn = make_plane_tensor(nullptr, L, narrow, P, 0);
  b.c1.in = tx; b.c1.out = tn; b.c2.in = tn; b.c2.out = tn; b.c3.in = tn; b.c3.out = tx; b.c3.res = tx;
  // the 20 -> 20 conv on folded images (plane.cuh), as the codec program runs it where the frame is long enough
  const bool by_parity = dilation == 2 && L / 4 >= 128;      // dilation 2: per position parity, or block-diagonal k9 on short frames
  const PlaneTensor tf = make_plane_tensor(nullptr, L / 2, kFoldC, P, by_parity ? 1 : 0);
  const int64_t w2_plain = plane_wpack_bytes(b.c2);
  if (folded) {
    NSC_CHECK_ARG(P == 2 && L / 2 >= 128, "nsc_bottleneck_block_tc: the folded narrow conv needs hi/lo planes and 128 folded rows per frame");
    b.c1.fold_out = by_parity ? 2 : 1; b.c1.out = tf;
    const bool diag = dilation == 2 && !by_parity;
    conv(b.c2, PK_X, kFoldC, kFoldC, diag ? 9 : kFoldK, 1, NSC_ACT_LRELU, RES_NONE, NSC_ACT_NONE);
    b.c2.Lin = by_parity ? L / 4 : L / 2; b.c2.fold2 = diag ? 2 : 1; b.c2.in = tf; b.c2.out = tn;
    if (by_parity) { b.c2.ileave = 1; b.c2.bmul = 2; b.c2.B = 2 * B; b.c2.in.deint = 0; b.c2.in.rows = L / 4; b.c2.in.frame_bytes = tf.frame_bytes / 2; }
  }
  int ring = plane_block_ring_frames();
  while (ring > 2 && ring > B) ring /= 2;
  b.ring = ring;
  NSC_CHECK_ARG(plane_conv_supported(b.c1) && plane_conv_supported(b.c2) && plane_conv_supported(b.c3),
                "nsc_bottleneck_block_tc: a conv of the block is not covered by the tensor engine");
  pl->x_bytes = align_up(B * tx.frame_bytes, 1024);
  pl->y_bytes = pl->x_bytes;
  {
    const int64_t tf2 = make_plane_tensor(nullptr, L / 2, kFoldC, P, 1).frame_bytes;   // the largest of the narrow tensor's three forms
    pl->n_bytes = align_up(B * (tn.frame_bytes > tf2 ? tn.frame_bytes : tf2), 1024);
  }
  pl->w1 = align_up(plane_wpack_bytes(b.c1), 1024);
  {
    PlaneConv f;                                     // sized for either form of the narrow conv
    conv(f, PK_X, kFoldC, kFoldC, 9, 1, NSC_ACT_LRELU, RES_NONE, NSC_ACT_NONE);      // (the k9 form is the larger of the two folded ones)
    f.Lin = 128; f.fold2 = 2; f.in = make_plane_tensor(nullptr, 128, kFoldC, 2, 0); f.out = make_plane_tensor(nullptr, 256, narrow, 2, 0); f.planes = 2;
    const int64_t w2_fold = plane_wpack_bytes(f);
    const int64_t w2_now = plane_wpack_bytes(b.c2);
    int64_t m = w2_plain > w2_fold ? w2_plain : w2_fold;
    if (w2_now > m) m = w2_now;
    pl->w2 = align_up(m, 1024);
  }
  pl->w3 = align_up(plane_wpack_bytes(b.c3), 1024);
  pl->flag_bytes = align_up(plane_block_flag_words(B) * (int64_t)sizeof(uint32_t), 1024);
  return NSC_OK;
}

}  // namespace

int64_t nsc_bottleneck_block_tc_workspace_bytes(int64_t B, int32_t L, int32_t wide, int32_t narrow, int32_t precision) {
  TcBlockPlan pl;
  if (make_tc_block_plan(B < 1 ? 1 : B, L, wide, narrow, 9, 9, 1, 0, precision, &pl) != NSC_OK) return -1;
  return 2048 + pl.x_bytes + pl.y_bytes + 2 * pl.n_bytes + pl.w1 + pl.w2 + pl.w3 + pl.flag_bytes;
}

int nsc_bottleneck_block_tc(const float* x, const float* params, float* y, int64_t B, int32_t L, int32_t wide, int32_t narrow,
                            int32_t k_plain, int32_t k_dilated, int32_t dilation, int32_t is_last_flat, int32_t precision,
                            int32_t* fused_out, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace nsc;
  const bool want_folded = fused_out && *fused_out == -2;        // tests: three launches with the narrow conv on folded images
  const bool force_unfused = fused_out && (*fused_out == -1 || want_folded);      // tests: the same three kernels, one launch each
  if (fused_out) *fused_out = 0;
  if (B == 0) return NSC_OK;
  NSC_CHECK_ARG(x && params && y && workspace, "nsc_bottleneck_block_tc: null pointer");
  TcBlockPlan pl;
  NSC_TRY(make_tc_block_plan(B, L, wide, narrow, k_plain, k_dilated, dilation, is_last_flat, precision, &pl, want_folded));
  const int64_t need = 2048 + pl.x_bytes + pl.y_bytes + 2 * pl.n_bytes + pl.w1 + pl.w2 + pl.w3 + pl.flag_bytes;
  if (workspace_bytes < need) {
    set_error("nsc_bottleneck_block_tc: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return NSC_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // the images sit 1 KB into the workspace: the first slab's halo reads stay inside it
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023) + 1024;
  NSC_CUDA_OK(cudaMemsetAsync(p, 0, (size_t)(pl.x_bytes + pl.y_bytes + 2 * pl.n_bytes), st));   // the images' zero rows
  PlaneBlock& b = pl.b;
  b.c1.in.base = p; b.c3.res.base = p; p += pl.x_bytes;
  b.c3.out.base = p; p += pl.y_bytes;
  b.c1.out.base = p; b.c2.in.base = p; p += pl.n_bytes;
  b.c2.out.base = p; b.c3.in.base = p; p += pl.n_bytes;
  b.c1.wpack = p; p += pl.w1;
  b.c2.wpack = p; p += pl.w2;
  b.c3.wpack = p; p += pl.w3;
  b.flags = reinterpret_cast<uint32_t*>(p);
  // parameters in creation order: each kernel (k, cin, cout) then its bias (nn_core_operator.py:61-75)
  const float* w = params;
  b.c1.w = w; b.c1.bias = w + (int64_t)k_plain * wide * narrow; w = b.c1.bias + narrow;
  b.c2.w = w; b.c2.bias = w + (int64_t)k_dilated * narrow * narrow; w = b.c2.bias + narrow;
  b.c3.w = w; b.c3.bias = w + (int64_t)k_plain * narrow * wide;
  NSC_TRY(plane_from_f32(x, 1, B, L, wide, b.c1.in, st));
  NSC_TRY(plane_pack_weights(b.c1, st));
  NSC_TRY(plane_pack_weights(b.c2, st));
  NSC_TRY(plane_pack_weights(b.c3, st));
  if (!force_unfused && plane_block_supported(b)) {
    NSC_CUDA_OK(cudaMemsetAsync(b.flags, 0, (size_t)pl.flag_bytes, st));
    NSC_TRY(plane_block_launch(b, st));
    if (fused_out) *fused_out = 1;
  } else {   // an odd number of tiles (CTA pairs need an even one): the same three kernels, one launch each
    NSC_TRY(plane_launch(b.c1, st));
    NSC_TRY(plane_launch(b.c2, st));
    NSC_TRY(plane_launch(b.c3, st));
  }
  NSC_TRY(plane_to_f32(b.c3.out, y, 1, B, L, wide, st));
  return NSC_OK;
}

// ---- gated_bottleneck (nn_core_operator.py:82-112) on the plane engine, channels-last fp32 tensors at the edge ----------------
// k1 conv + leaky ReLU -> the two k15 gate convs as one layer with the gate product in its epilogue -> k9 conv + residual
// (+ leaky ReLU): the three launches of the codec program's gated block.  params in creation order: (w_1x1, b), (w_left, b),
// (w_right, b), (w_out, b).
namespace {

struct TcGatedPlan {
  nsc::PlaneConv c1, cg, c3;
  int64_t x_bytes, n_bytes, nd_bytes, w1, wg, w3;
};

int make_tc_gated_plan(int64_t B, int L, int wide, int narrow, int k_plain, int dilation, int is_last_flat, int precision, TcGatedPlan* pl) {
  using namespace nsc;
  NSC_CHECK_ARG(precision == 1 || precision == 2, "nsc_gated_block_tc: precision must be 1 (fp16 hi/lo) or 2 (fp16)");
  NSC_CHECK_ARG(B >= 1 && L > 0 && L % 128 == 0 && wide > 32 && wide <= 128 && narrow == 20 && k_plain == 9 && dilation >= 1 && dilation <= 2,
                "nsc_gated_block_tc: shape not covered by the tensor engine (L %d, %d/%d, k %d, dilation %d)", L, wide, narrow, k_plain, dilation);
  const int P = precision == 1 ? 2 : 1;
  const bool deint = dilation == 2 && (L / 2) % 128 == 0;      // else: plain image (16 halo rows for dilation 2)
  auto conv = [&](PlaneConv& c, int Lin, int cin, int cout, int K, int dil, int act, int res_mode, int post) {
    c.kind = PK_X; c.Lin = Lin; c.Cin = cin; c.Cout = cout; c.K = K; c.dil = dil; c.stride = 1;
    c.act = act; c.post_act = post; c.res_mode = res_mode; c.shuffle = 1; c.planes = P; c.B = B;
  };
  const PlaneTensor tx = make_plane_tensor(nullptr, L, wide, P, 0), tn = make_plane_tensor(nullptr, L, narrow, P, 0),
                    tnd = make_plane_tensor(nullptr, L, narrow, P, 1);
  conv(pl->c1, L, wide, narrow, 1, 1, NSC_ACT_LRELU, RES_NONE, NSC_ACT_NONE);
  pl->c1.in = tx; pl->c1.out = deint ? tnd : tn;
  if (deint) {
    conv(pl->cg, L / 2, narrow, 2 * narrow, 15, 1, NSC_ACT_NONE, RES_NONE, NSC_ACT_NONE);
    pl->cg.in = tnd; pl->cg.in.deint = 0; pl->cg.in.rows = L / 2; pl->cg.in.frame_bytes = tnd.frame_bytes / 2;
    pl->cg.ileave = 1; pl->cg.bmul = 2; pl->cg.B = 2 * B;
  } else {
    conv(pl->cg, L, narrow, 2 * narrow, 15, dilation, NSC_ACT_NONE, RES_NONE, NSC_ACT_NONE);
    pl->cg.in = tn;
  }
  pl->cg.glu = 1; pl->cg.out = tn;
  conv(pl->c3, L, narrow, wide, k_plain, 1, NSC_ACT_NONE, RES_ADD, is_last_flat ? NSC_ACT_NONE : NSC_ACT_LRELU);
  pl->c3.in = tn; pl->c3.out = tx; pl->c3.res = tx;
  NSC_CHECK_ARG(plane_conv_supported(pl->c1) && plane_conv_supported(pl->cg) && plane_conv_supported(pl->c3),
                "nsc_gated_block_tc: a conv of the block is not covered by the tensor engine");
  pl->x_bytes = align_up(B * tx.frame_bytes, 1024);
  pl->n_bytes = align_up(B * tn.frame_bytes, 1024);
  pl->nd_bytes = align_up(B * tnd.frame_bytes, 1024);
  pl->w1 = align_up(plane_wpack_bytes(pl->c1), 1024);
  pl->wg = align_up(plane_wpack_bytes(pl->cg), 1024);
  pl->w3 = align_up(plane_wpack_bytes(pl->c3), 1024);
  return NSC_OK;
}

}  // namespace

int64_t nsc_gated_block_tc_workspace_bytes(int64_t B, int32_t L, int32_t wide, int32_t narrow, int32_t dilation, int32_t precision) {
  TcGatedPlan pl;
  if (make_tc_gated_plan(B < 1 ? 1 : B, L, wide, narrow, 9, dilation, 0, precision, &pl) != NSC_OK) return -1;
  return 8192 + 2 * pl.x_bytes + 2 * pl.n_bytes + pl.nd_bytes + pl.w1 + pl.wg + pl.w3;
}

int nsc_gated_block_tc(const float* x, const float* params, float* y, int64_t B, int32_t L, int32_t wide, int32_t narrow, int32_t k_plain,
                       int32_t dilation, int32_t is_last_flat, int32_t precision, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace nsc;
  if (B == 0) return NSC_OK;
  NSC_CHECK_ARG(x && params && y && workspace, "nsc_gated_block_tc: null pointer");
  TcGatedPlan pl;
  NSC_TRY(make_tc_gated_plan(B, L, wide, narrow, k_plain, dilation, is_last_flat, precision, &pl));
  const int64_t need = 8192 + 2 * pl.x_bytes + 2 * pl.n_bytes + pl.nd_bytes + pl.w1 + pl.wg + pl.w3;
  if (workspace_bytes < need) {
    set_error("nsc_gated_block_tc: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
    return NSC_E_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // images sit 2 KB into the (zeroed) workspace and are followed by 2 KB of zeros: the 16-row halos of a dilation-2 gate layer on
  // the plain image read 1 KB beyond an image's own zero rows
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~(uintptr_t)1023);
  const int64_t img_bytes = 2 * pl.x_bytes + 2 * pl.n_bytes + pl.nd_bytes;
  NSC_CUDA_OK(cudaMemsetAsync(p, 0, (size_t)(img_bytes + 4096), st));
  p += 2048;
  uint8_t* xi = p; p += pl.x_bytes;
  uint8_t* yi = p; p += pl.x_bytes;
  uint8_t* n0 = p; p += pl.n_bytes;
  uint8_t* n1 = p; p += pl.n_bytes;
  uint8_t* n0d = p; p += pl.nd_bytes;
  p += 2048;
  pl.c1.in.base = xi; pl.c1.out.base = pl.c1.out.deint ? n0d : n0;
  pl.cg.in.base = pl.cg.ileave ? n0d : n0; pl.cg.out.base = n1;
  pl.c3.in.base = n1; pl.c3.out.base = yi; pl.c3.res.base = xi;
  pl.c1.wpack = p; p += pl.w1;
  pl.cg.wpack = p; p += pl.wg;
  pl.c3.wpack = p;
  const float* w = params;
  pl.c1.w = w; pl.c1.bias = w + (int64_t)wide * narrow; w = pl.c1.bias + narrow;
  pl.cg.w = w; pl.cg.bias = w + 15LL * narrow * narrow; w = pl.cg.bias + narrow;
  pl.cg.w2 = w; pl.cg.bias2 = w + 15LL * narrow * narrow; w = pl.cg.bias2 + narrow;
  pl.c3.w = w; pl.c3.bias = w + (int64_t)k_plain * narrow * wide;
  NSC_TRY(plane_from_f32(x, 1, B, L, wide, pl.c1.in, st));
  NSC_TRY(plane_pack_weights(pl.c1, st));
  NSC_TRY(plane_pack_weights(pl.cg, st));
  NSC_TRY(plane_pack_weights(pl.c3, st));
  NSC_TRY(plane_launch(pl.c1, st));
  NSC_TRY(plane_launch(pl.cg, st));
  NSC_TRY(plane_launch(pl.c3, st));
  NSC_TRY(plane_to_f32(pl.c3.out, y, 1, B, L, wide, st));
  return NSC_OK;
}

/* per-CTA counters of the most recent fused block launch (NSC_BLOCK_STATS=1): 8 words per CTA, see plane_conv.cu */
int nsc_debug_block_stats(unsigned long long* out_host, int32_t max_ctas) {
  NSC_CHECK_ARG(out_host != nullptr && max_ctas >= 1, "nsc_debug_block_stats: bad arguments");
  const int n = max_ctas < nsc::kStatCtas ? max_ctas : nsc::kStatCtas;
  NSC_CUDA_OK(cudaDeviceSynchronize());
  NSC_CUDA_OK(cudaMemcpyFromSymbol(out_host, nsc::g_block_stats, sizeof(unsigned long long) * n * nsc::kStatWords));
  return n;
}

}  // extern "C"
