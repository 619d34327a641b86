// Shared device-side PTX wrappers (mbarrier / TMA / tcgen05 / TMEM) and host-side tensor-map helpers for the
// tcgen05 kernels (tc_conv.cu, tc_slab.cu).  sm_100a only.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <mutex>

namespace mv2 {


// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]   (kind::f16: bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// ---- 2-CTA cluster helpers (weight-tile multicast) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 3-D box load delivered to the same shared-memory offset (and mbarrier offset) of every CTA in cta_mask
__device__ __forceinline__ void tma_load_3d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask) : "memory");
}
// tcgen05.commit arriving on the barrier at the same offset in every CTA of cta_mask
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask) : "memory");
}

// warp-uniform leader election (all 32 lanes must execute it); returns 1 in exactly one lane
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major operand with hardware swizzle.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (ignored for swizzled K-major)
//   bits [32,46) stride byte offset >> 4   (8 rows x row bytes)     bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr, uint32_t row_bytes) {
  const uint32_t sbo = 8 * row_bytes;
  const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}


// general: explicit SBO and row width (128 B rows -> SWIZZLE_128B, 64 B -> SWIZZLE_64B, 32 B -> SWIZZLE_32B)
__device__ __forceinline__ uint64_t make_kmajor_desc_rb(uint32_t saddr, uint32_t sbo, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

// same, with an explicit stride-byte-offset (distance between consecutive 8-row core groups), SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_desc_sbo(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}



// ------------------------------------------------------------------------------------------
// shared epilogue: fp32 accumulator chunk -> bias -> activation -> (GEGLU | shuffle) -> residual -> bf16 store
// ------------------------------------------------------------------------------------------
struct TcEpi {
  const float* bias;             // global, packed column order (only used to decide has-bias; values come from smem)
  const __nv_bfloat16* res;
  __nv_bfloat16* y;
  int act, shuffle, mode;        // mode 0 plain; 1 GEGLU: packed cols [16g, 16g+8) = x, [16g+8, 16g+16) = gate (M:466-469)
  int Co;                        // packed GEMM output columns
  int To, Ho, Wo;                // output volume before any depth-to-space/time shuffle
  int out_cf;                    // 1: y is channels-first (B, Co, To, Ho, Wo) -- EPI_RAGGED scalar stores only (conv_out)
  const float* oscale;           // [B][Co] or null: accumulator multiplier per (clip, output channel) before bias / activation
};

__device__ __forceinline__ void store8_bf16(__nv_bfloat16* dst, const float (&v)[8]) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
  o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = o;
}

// Epilogue flavours are compile-time so each kernel instance carries only the code it runs (the generic version was
// ~2300 SASS instructions per 32-column chunk and thrashed the instruction cache of the 8 epilogue warps).
// EPI_PLAIN (slab kernel only) additionally requires Co % 8 == 0 and stores through a shared-memory transpose;
// EPI_RAGGED is the direct per-row path with scalar tails (conv_out's 3 channels, and the tap kernel's plain mode).
// EPI_PLAIN_RES is EPI_PLAIN with a residual input: each lane reads its own row's residual (64 contiguous bytes per chunk)
// and act(conv + bias) + res is summed in fp32 and rounded to bf16 ONCE (the reference's bf16 `fn(x) + x` rounds twice; the
// single rounding is strictly closer to the fp32 result).
// EPI_FUSED_RU (slab kernel only): the whole conv half of a ResidualUnit in one launch -- the ELU'd 3x3x3 tile goes to
// shared memory as the A operand of a second tcgen05.mma against the 1x1x1 weights, and the second epilogue emits the
// SqueezeExcite online-softmax pool partials next to y (see tc_slab.cu).
// EPI_SHUFFLE_ST (slab kernel only): depth-to-space / depth-to-time stores through the same shared-memory transpose as
// EPI_PLAIN (64 contiguous bytes per output position and store instruction); needs Cy % 32 == 0 so that a 32-column chunk
// stays inside one sub-pixel phase.  EPI_SHUFFLE is the direct 16-byte-piece path for the other widths.
// EPI_DOWN_SPACE (slab kernel only): SpatialDownsample2x (3x3, stride 2) -- plain epilogue, but the slab is two row-parity
// sub-slabs of the input viewed as (W/2) x (2C) and the taps follow a small offset table (see tc_slab.cu).
enum { EPI_PLAIN = 0, EPI_GEGLU = 1, EPI_SHUFFLE = 2, EPI_RAGGED = 3, EPI_PLAIN_RES = 4, EPI_FUSED_RU = 5, EPI_SHUFFLE_ST = 6,
       EPI_DOWN_SPACE = 7 };

// Branch-free activations on the bare MUFU approximations (ex2/rcp with flush-to-zero): the results are rounded to
// bf16 right after, and __expf's denormal range handling costs ~5 extra instructions per element.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of the result): 2 MUFU + ~12 FP32
// instructions, about half of libdevice's erff; the GEGLU epilogue of the feed-forward runs 16 of them per 32 columns.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = ex2_approx(-1.4426950408889634f * ax * ax);
  return copysignf(fmaf(-p * t, e, 1.f), x);
}

// gelu(g) = g * Phi(g) with Phi(g) = 0.5 (1 + erf(g / sqrt 2)) and erf by Abramowitz & Stegun 7.1.26 in z = |g| / sqrt 2
// (|error| <= 1.5e-7): with q = 0.5 * t * poly(t) * exp(-g^2 / 2), t = 1 / (1 + 0.3275911 z),
//   Phi(g) = 1 - q (g >= 0), q (g < 0)   =>   gelu(g) = max(g, 0) - |g| * q.
// All constant factors (1 / sqrt 2, 0.5, log2 e) are folded into the coefficients: 2 MUFU + 12 FP32 instructions per value
// (the erf_fast form above costs 16); the fc1 + GEGLU epilogue is issue bound (ncu: ~39 instructions per output in total).
__device__ __forceinline__ float gelu_fast(float g) {
  const float ag = fabsf(g);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, ag, 1.f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float e = ex2_approx((g * g) * (-0.5f * 1.4426950408889634f));
  const float q = (p * t) * e;
  return fmaf(-ag, q, fmaxf(g, 0.f));
}

template <int ACT>
__device__ __forceinline__ float act_ct(float x) {
  if (ACT == MV2_ACT_ELU) {
    const float e = ex2_approx(x * 1.4426950408889634f) - 1.f;
    return x > 0.f ? x : e;
  }
  if (ACT == MV2_ACT_SILU) return x * rcp_approx(1.f + ex2_approx(-1.4426950408889634f * x));
  return x;
}

// r: 32 raw accumulator columns of ONE output row (position b,to,ho,wo); n = first packed column; sb = smem bias of
// these columns (always valid memory; zeros when there is no bias).
// row_base = linear position index * Co (plain mode), computed once per row by the caller.
template <int MODE, int ACT>
__device__ __forceinline__ void epi_chunk32_t(const TcEpi& e, const uint32_t (&r)[32], int ncols, int n, const float* sb,
                                              int b, int to, int ho, int wo, int64_t row_base) {
  if (MODE == EPI_GEGLU) {
    const int I = e.Co >> 1;
    const int64_t pos = (((int64_t)b * e.To + to) * e.Ho + ho) * e.Wo + wo;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (g * 16 >= ncols || n + g * 16 >= e.Co) break;
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float xv = __uint_as_float(r[g * 16 + q]) + sb[g * 16 + q];
        const float gt = __uint_as_float(r[g * 16 + 8 + q]) + sb[g * 16 + 8 + q];
        v[q] = gelu_fast(gt) * xv;
      }
      store8_bf16(e.y + pos * I + ((n + g * 16) >> 1), v);
    }
    return;
  }
  const int cy = MODE == EPI_SHUFFLE ? (e.shuffle == MV2_SHUFFLE_SPACE ? (e.Co >> 2) : (e.Co >> 1)) : e.Co;
  const bool vec_ok = (cy & 7) == 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const int ng = n + g * 8;
    if (g * 8 >= ncols || ng >= e.Co) break;
    float v[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(sb + g * 8), b1 = *reinterpret_cast<const float4*>(sb + g * 8 + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      if (MODE != EPI_SHUFFLE && e.oscale) {      // Conv3DMod demodulation (M:741-742)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float os = ng + q < e.Co ? e.oscale[(int64_t)b * e.Co + ng + q] : 0.f;
          v[q] = act_ct<ACT>(__uint_as_float(r[g * 8 + q]) * os + bb[q]);
        }
      } else
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = act_ct<ACT>(__uint_as_float(r[g * 8 + q]) + bb[q]);
    }
    int64_t off;
    if (MODE == EPI_SHUFFLE) {
      const int qd = ng / cy, c = ng - qd * cy;
      if (e.shuffle == MV2_SHUFFLE_SPACE) {
        const int p1 = qd >> 1, p2 = qd & 1;
        off = ((((int64_t)b * e.To + to) * (2 * e.Ho) + (2 * ho + p1)) * (2 * e.Wo) + (2 * wo + p2)) * cy + c;
      } else {
        off = ((((int64_t)b * (2 * e.To) + (2 * to + qd)) * e.Ho + ho) * e.Wo + wo) * cy + c;
      }
    } else {
      off = row_base + ng;
    }
    if (vec_ok && ng + 8 <= e.Co) {
      if (e.res) {
        const uint4 rv = *reinterpret_cast<const uint4*>(e.res + off);
        const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = __bfloat1622float2(rb[q]);
          v[2 * q] += f.x;
          v[2 * q + 1] += f.y;
        }
      }
      store8_bf16(e.y + off, v);
    } else {
      if (MODE == EPI_RAGGED && e.out_cf) {            // conv_out: the reconstruction goes out in torch's (B, C, T, H, W)
        const int64_t plane = (int64_t)e.Ho * e.Wo;
        const int64_t o0 = ((int64_t)b * e.Co * e.To + to) * plane + (int64_t)ho * e.Wo + wo;
        for (int q = 0; q < 8 && ng + q < e.Co; ++q) e.y[o0 + (int64_t)(ng + q) * e.To * plane] = __float2bfloat16_rn(v[q]);
      } else
      for (int q = 0; q < 8 && ng + q < e.Co; ++q) {   // scalar tail (Co % 8 != 0, e.g. conv_out's 3 channels)
        float x = v[q];
        if (e.res) x += __bfloat162float(e.res[off + q]);
        e.y[off + q] = __float2bfloat16_rn(x);
      }
    }
  }
}

// bias + activation + bf16 packing of one 32-column chunk (row-per-lane), for the staged epilogue
template <int ACT>
__device__ __forceinline__ void epi_pack32_t(const uint32_t (&r)[32], const float* sb, uint32_t (&pk)[16]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 b = *reinterpret_cast<const float4*>(sb + g * 4);
    pk[2 * g] = pack_bf16x2(act_ct<ACT>(__uint_as_float(r[4 * g]) + b.x), act_ct<ACT>(__uint_as_float(r[4 * g + 1]) + b.y));
    pk[2 * g + 1] = pack_bf16x2(act_ct<ACT>(__uint_as_float(r[4 * g + 2]) + b.z), act_ct<ACT>(__uint_as_float(r[4 * g + 3]) + b.w));
  }
}
// GEGLU of one packed 32-column chunk ([8 x | 8 gate] twice, M:466-469): 16 outputs -> 8 bf16x2 words
__device__ __forceinline__ void epi_geglu_pack32(const uint32_t (&r)[32], const float* sb, uint32_t* pk) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float bx[8], bg[8], v[8];
    *reinterpret_cast<float4*>(bx) = *reinterpret_cast<const float4*>(sb + g * 16);
    *reinterpret_cast<float4*>(bx + 4) = *reinterpret_cast<const float4*>(sb + g * 16 + 4);
    *reinterpret_cast<float4*>(bg) = *reinterpret_cast<const float4*>(sb + g * 16 + 8);
    *reinterpret_cast<float4*>(bg + 4) = *reinterpret_cast<const float4*>(sb + g * 16 + 12);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      v[q] = gelu_fast(__uint_as_float(r[g * 16 + 8 + q]) + bg[q]) * (__uint_as_float(r[g * 16 + q]) + bx[q]);
    pk[4 * g] = pack_bf16x2(v[0], v[1]); pk[4 * g + 1] = pack_bf16x2(v[2], v[3]);
    pk[4 * g + 2] = pack_bf16x2(v[4], v[5]); pk[4 * g + 3] = pack_bf16x2(v[6], v[7]);
  }
}
// bias + activation of one 32-column chunk, kept in fp32 (for the fp32-staged residual epilogue)
template <int ACT>
__device__ __forceinline__ void epi_act32_t(uint32_t (&r)[32], const float* sb) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 b = *reinterpret_cast<const float4*>(sb + g * 4);
    r[4 * g] = __float_as_uint(act_ct<ACT>(__uint_as_float(r[4 * g]) + b.x));
    r[4 * g + 1] = __float_as_uint(act_ct<ACT>(__uint_as_float(r[4 * g + 1]) + b.y));
    r[4 * g + 2] = __float_as_uint(act_ct<ACT>(__uint_as_float(r[4 * g + 2]) + b.z));
    r[4 * g + 3] = __float_as_uint(act_ct<ACT>(__uint_as_float(r[4 * g + 3]) + b.w));
  }
}
__device__ __forceinline__ void epi_act32(int act, uint32_t (&r)[32], const float* sb) {
  if (act == MV2_ACT_ELU) epi_act32_t<MV2_ACT_ELU>(r, sb);
  else if (act == MV2_ACT_SILU) epi_act32_t<MV2_ACT_SILU>(r, sb);
  else epi_act32_t<MV2_ACT_NONE>(r, sb);
}
__device__ __forceinline__ void epi_pack32(int act, const uint32_t (&r)[32], const float* sb, uint32_t (&pk)[16]) {
  if (act == MV2_ACT_ELU) epi_pack32_t<MV2_ACT_ELU>(r, sb, pk);
  else if (act == MV2_ACT_SILU) epi_pack32_t<MV2_ACT_SILU>(r, sb, pk);
  else epi_pack32_t<MV2_ACT_NONE>(r, sb, pk);
}

// activation is a kernel argument; dispatch once per chunk (warp uniform) into the compile-time variants
template <int MODE>
__device__ __forceinline__ void epi_chunk32(const TcEpi& e, const uint32_t (&r)[32], int ncols, int n, const float* sb,
                                            int b, int to, int ho, int wo, int64_t row_base) {
  if (MODE == EPI_GEGLU) { epi_chunk32_t<EPI_GEGLU, MV2_ACT_NONE>(e, r, ncols, n, sb, b, to, ho, wo, row_base); return; }
  if (e.act == MV2_ACT_ELU) epi_chunk32_t<MODE, MV2_ACT_ELU>(e, r, ncols, n, sb, b, to, ho, wo, row_base);
  else if (e.act == MV2_ACT_SILU) epi_chunk32_t<MODE, MV2_ACT_SILU>(e, r, ncols, n, sb, b, to, ho, wo, row_base);
  else epi_chunk32_t<MODE, MV2_ACT_NONE>(e, r, ncols, n, sb, b, to, ho, wo, row_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

static inline int pow2_ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }
static inline int floor_div(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }


}  // namespace mv2
