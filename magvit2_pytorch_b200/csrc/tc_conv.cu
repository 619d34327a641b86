// tcgen05 / TMA implicit-GEMM convolution for sm_100a (bf16 in, fp32 accumulate in TMEM).
//
// One kernel covers the whole dense-contraction family of the VideoTokenizer forward path:
// causal 3x3x3 convs, 1x1x1 convs / Linear layers, the strided compress_space / compress_time
// convs and the 1x1 up-samplers with their depth-to-space / depth-to-time stores.
//
// GEMM view:  D[m][n] = sum_{tap, c} X[pos(m) + off(tap)][c] * W[n][tap][c]
//   M = 128 output positions per CTA, laid out as a (bt, bh, bw) box of the output volume
//   N = up to 256 output channels per CTA
//   K = taps x Ci, walked in BK-channel slices of one tap at a time
// Operand staging: one TMA box load per (tap, slice) for A -- the box {BK, bw, bh, bt, 1} of the
// channels-last activation tensor shifted by the tap offset; out-of-bounds elements (the causal
// time halo, the spatial halo, ragged tile edges) are zero-filled by the TMA unit, so no padded
// copy of the activations is ever materialised (the reference does F.pad + conv, M:924-928).
// Strided convs read through "parity view" tensor maps (one per stride phase).  Both operands are
// K-major in shared memory with the hardware 128B/64B/32B swizzle (BK = 64/32/16 channels).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, warps 2-5 = epilogue (tcgen05.ld -> bias/activation/residual -> global).
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <mutex>
#include <algorithm>
#include <string.h>

namespace mv2 {

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
constexpr int TC_MAX_TAPS = 64;
constexpr int TC_MAX_MAPS = 8;
constexpr int TC_BM = 128;

struct alignas(64) TcParams {
  CUtensorMap amap[TC_MAX_MAPS];
  CUtensorMap wmap;
  int8_t tap_map[TC_MAX_TAPS], tap_dt[TC_MAX_TAPS], tap_dh[TC_MAX_TAPS], tap_dw[TC_MAX_TAPS];
  int ntaps, kchunks, ci_pad, bk;
  int B, To, Ho, Wo, Co;
  int bt, bh, bw, tt, th, tw;
  int bn, stages, tmem_cols;
  TcEpi epi;
};

template <int MODE>
__global__ void __launch_bounds__(192, 2) tc_conv_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bk = p.bk;
  const uint32_t row_bytes = bk * 2;
  const uint32_t a_bytes = TC_BM * row_bytes;
  const uint32_t b_bytes = p.bn * row_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;   // both multiples of 1024 when bn % 16 == 0 and bk >= 32; see host
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;
  // barriers: full[s] at +8s, empty[s] at +8(S+s), tmem_full at +16S, tmem slot at +16S+8
  const uint32_t full0 = bar_base, empty0 = bar_base + 8 * p.stages, tfull = bar_base + 16 * p.stages;
  const uint32_t tslot = tfull + 8;
  float* sbias = reinterpret_cast<float*>(smem_raw + (((tslot + 8 + 15) & ~15u) - smem_u32(smem_raw)));   // bn floats, 16-byte aligned

  // tile coordinates
  int tile = blockIdx.x;
  const int iw = tile % p.tw; tile /= p.tw;
  const int ih = tile % p.th; tile /= p.th;
  const int it = tile % p.tt; tile /= p.tt;
  const int b = tile;
  const int w0 = iw * p.bw, h0 = ih * p.bh, t0 = it * p.bt;
  const int n0 = blockIdx.y * p.bn;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.wmap);
    tma_prefetch_desc(&p.amap[0]);
  }
  if (warp == 1) tmem_alloc(tslot, p.tmem_cols);
  if (warp >= 2)
    for (int i = threadIdx.x - 64; i < p.bn; i += 128) sbias[i] = (p.epi.bias && n0 + i < p.Co) ? p.epi.bias[n0 + i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));
  // everything above overlapped the previous kernel's tail (PDL); activations may only be touched from here on
  pdl_wait();
  pdl_launch_dependents();

  const int n_iters = p.ntaps * p.kchunks;
  if (warp == 0) {
    if (lane == 0) {
      // incremental ring / tap bookkeeping: no division or modulo in the loop
      uint32_t s = 0, ph = 0;
      int tap = 0, kc = 0;
      for (int i = 0; i < n_iters; ++i) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, stage_bytes);
        const uint32_t sa = smem_base + s * stage_bytes;
        tma_load_5d(sa, &p.amap[p.tap_map[tap]], full0 + 8 * s, kc * bk, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap],
                    t0 + p.tap_dt[tap], b);
        tma_load_2d(sa + a_bytes, &p.wmap, full0 + 8 * s, tap * p.ci_pad + kc * bk, n0);
        if (++kc == p.kchunks) { kc = 0; ++tap; }
        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the loop (uniform operands); one elected lane issues the tcgen05 instructions
    // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major both, N>>3 at 17, M>>4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t leader = elect_one();
    const int ksteps = bk >> 4;
    const uint64_t d_hi = make_kmajor_desc(0, row_bytes);     // descriptor with a zero start address
    const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4, base16 = (smem_base & 0x3FFFF) >> 4;
    uint32_t s = 0, ph = 0, lo = base16;
    for (int i = 0; i < n_iters; ++i) {
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      const uint64_t ad = d_hi | (uint64_t)lo;
      const uint64_t bd = d_hi | (uint64_t)(lo + a16);
      if (leader) {
        // advancing 16 bf16 along K = +32 bytes = +2 in the (addr >> 4) field
        umma_bf16(tmem_base, ad, bd, idesc, i > 0 ? 1u : 0u);
        if (ksteps > 1) umma_bf16(tmem_base, ad + 2, bd + 2, idesc, 1u);
        if (ksteps > 2) {
          umma_bf16(tmem_base, ad + 4, bd + 4, idesc, 1u);
          umma_bf16(tmem_base, ad + 6, bd + 6, idesc, 1u);
        }
        umma_commit(empty0 + 8 * s);   // frees the smem slot once these MMAs retire
      }
      if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; lo = base16; } else { lo += stage16; }
    }
    if (leader) umma_commit(tfull);    // accumulator complete
  } else {
    // ---------------- epilogue warps 2..5 : TMEM lanes 32*(warp%4) .. +31 ----------------
    const int sub = warp & 3;
    const int row = sub * 32 + lane;
    const int lw = row % p.bw, lh = (row / p.bw) % p.bh, lt = row / (p.bw * p.bh);
    const int wo = w0 + lw, ho = h0 + lh, to = t0 + lt;
    const bool row_ok = wo < p.Wo && ho < p.Ho && to < p.To;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t tlane = tmem_base + ((uint32_t)(sub * 32) << 16);
    const int64_t row_base = ((((int64_t)b * p.To + to) * p.Ho + ho) * p.Wo + wo) * p.Co;
    for (int c0 = 0; c0 < p.bn; c0 += 32) {
      uint32_t r[32];
      if (p.bn - c0 >= 32) tmem_ld_32x32b_x32(tlane + c0, r);
      else tmem_ld_32x32b_x16(tlane + c0, r);
      tmem_ld_wait();
      if (row_ok) epi_chunk32<MODE>(p.epi, r, min(32, p.bn - c0), n0 + c0, sbias + c0, b, to, ho, wo, row_base);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
}  // namespace mv2

using namespace mv2;

extern "C" {

int mv2_tc_conv_supported(const mv2_tc_conv_args* a) {
  if (!a) return 0;
  if (a->Ci % 16 != 0) return 0;                    // TMA inner box = 32/64/128 B, global strides multiple of 16 B
  if (a->kt * a->kh * a->kw > TC_MAX_TAPS) return 0;
  if (a->st < 1 || a->st > 2 || a->sh < 1 || a->sh > 2 || a->sw < 1 || a->sw > 2) return 0;
  if (a->shuffle != MV2_SHUFFLE_NONE && ((a->shuffle == MV2_SHUFFLE_SPACE ? a->Co / 4 : a->Co / 2) % 8 != 0)) return 0;
  if (a->res && a->Co % 8 != 0) return 0;
  if (a->epi_mode == 1 && (a->Co % 32 != 0 || a->shuffle != MV2_SHUFFLE_NONE || a->res)) return 0;   // GEGLU pairs
  if (a->epi_mode != 0 && a->epi_mode != 1) return 0;
  if (a->out_layout != 0) return 0;
  if (a->oscale && (a->epi_mode != 0 || a->shuffle != MV2_SHUFFLE_NONE)) return 0;
  return 1;
}

int mv2_tc_conv_forward(const mv2_tc_conv_args* a, void* stream) {
  MV2_CHECK_ARG(a && a->x && a->w && a->y);
  if (!mv2_tc_conv_supported(a)) { set_error("mv2_tc_conv_forward: unsupported shape (Ci=%d Co=%d)", a->Ci, a->Co); return MV2_E_UNSUPPORTED; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return MV2_E_CUDA; }

  TcParams p;
  memset(&p, 0, sizeof(p));
  const int bk = (a->Ci % 64 == 0) ? 64 : ((a->Ci % 32 == 0) ? 32 : 16);
  const CUtensorMapSwizzle swz = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  p.bk = bk;
  p.ci_pad = a->Ci;
  p.kchunks = a->Ci / bk;
  p.B = a->B; p.To = a->To; p.Ho = a->Ho; p.Wo = a->Wo; p.Co = a->Co;
  // output tile box: bw * bh * bt = 128
  p.bw = std::min(128, pow2_ceil(a->Wo));
  p.bh = std::min(128 / p.bw, pow2_ceil(a->Ho));
  p.bt = 128 / (p.bw * p.bh);
  p.tw = ceil_div(a->Wo, p.bw); p.th = ceil_div(a->Ho, p.bh); p.tt = ceil_div(a->To, p.bt);
  // N tile: multiple of 16 (UMMA M=128 constraint), rows of B must keep 1024 B stage alignment
  int bn = std::min(256, (a->Co + 15) / 16 * 16);
  const int row_bytes = bk * 2;
  while ((bn * row_bytes) % 1024 != 0) bn += 16;      // bk=16 -> bn % 32 == 0, bk >= 32: already fine
  MV2_CHECK_ARG(bn <= 256);
  p.bn = bn;
  p.tmem_cols = std::max(32, pow2_ceil(bn));
  const int stage_bytes = TC_BM * row_bytes + bn * row_bytes;
  MV2_CHECK_ARG((TC_BM * row_bytes) % 1024 == 0);
  // ring depth: no deeper than the K loop (short-K layers then fit several CTAs per SM, overlapping one CTA's
  // epilogue with another's loads)
  int stages = (200 * 1024) / stage_bytes;
  stages = std::max(2, std::min(stages, 8));
  stages = std::max(1, std::min(stages, a->kt * a->kh * a->kw * (a->Ci / bk)));
  p.stages = stages;
  p.epi.bias = a->bias; p.epi.res = (const __nv_bfloat16*)a->res; p.epi.y = (__nv_bfloat16*)a->y;
  p.epi.act = a->act; p.epi.shuffle = a->shuffle; p.epi.mode = a->epi_mode; p.epi.Co = a->Co;
  p.epi.To = a->To; p.epi.Ho = a->Ho; p.epi.Wo = a->Wo; p.epi.out_cf = 0; p.epi.oscale = a->oscale;

  // ---- activation tensor maps: one per stride-parity phase ----
  const int st = a->st, sh = a->sh, sw = a->sw;
  const int nmaps = st * sh * sw;
  MV2_CHECK_ARG(nmaps <= TC_MAX_MAPS);
  const int64_t C = a->Ci, W = a->Wi, H = a->Hi, T = a->Ti;
  for (int pt = 0; pt < st; ++pt)
    for (int ph = 0; ph < sh; ++ph)
      for (int pw = 0; pw < sw; ++pw) {
        const int id = (pt * sh + ph) * sw + pw;
        const int64_t nW = (W - pw + sw - 1) / sw, nH = (H - ph + sh - 1) / sh, nT = (T - pt + st - 1) / st;
        if (nW <= 0 || nH <= 0 || nT <= 0) { set_error("empty stride phase (dimension smaller than stride)"); return MV2_E_UNSUPPORTED; }
        cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)nW, (cuuint64_t)nH, (cuuint64_t)nT, (cuuint64_t)a->B};
        cuuint64_t strides[4] = {(cuuint64_t)(sw * C * 2), (cuuint64_t)(sh * W * C * 2), (cuuint64_t)(st * H * W * C * 2),
                                 (cuuint64_t)(T * H * W * C * 2)};
        cuuint32_t box[5] = {(cuuint32_t)bk, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bt, 1};
        cuuint32_t es[5] = {1, 1, 1, 1, 1};
        char* base = (char*)a->x + ((int64_t)pt * H * W + (int64_t)ph * W + pw) * C * 2;
        CUresult r = enc(&p.amap[id], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations, phase %d) failed: %d", id, (int)r); return MV2_E_CUDA; }
      }
  // ---- taps ----
  p.ntaps = a->kt * a->kh * a->kw;
  for (int dt = 0; dt < a->kt; ++dt)
    for (int dh = 0; dh < a->kh; ++dh)
      for (int dw = 0; dw < a->kw; ++dw) {
        const int tap = (dt * a->kh + dh) * a->kw + dw;
        const int ot = dt - a->pt, oh = dh - a->ph, ow = dw - a->pw;
        const int pt = ((ot % st) + st) % st, ph = ((oh % sh) + sh) % sh, pw = ((ow % sw) + sw) % sw;
        p.tap_map[tap] = (int8_t)((pt * sh + ph) * sw + pw);
        p.tap_dt[tap] = (int8_t)floor_div(ot, st);
        p.tap_dh[tap] = (int8_t)floor_div(oh, sh);
        p.tap_dw[tap] = (int8_t)floor_div(ow, sw);
      }
  // ---- weights map: [Co][taps * Ci] K-major ----
  {
    const int64_t K = (int64_t)p.ntaps * a->Ci;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)a->Co};
    cuuint64_t strides[1] = {(cuuint64_t)(K * 2)};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)bn};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)a->w, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return MV2_E_CUDA; }
  }
  const size_t smem = (size_t)stages * stage_bytes + 16 * stages + 32 + (size_t)bn * 4 + 1024;
  static PerDeviceOnce attr_once;
  const cudaError_t attr_err = attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel<EPI_RAGGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_conv_kernel<EPI_GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_conv_kernel<EPI_SHUFFLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return e;
  });
  if (attr_err != cudaSuccess) { set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err)); return MV2_E_CUDA; }
  MV2_CHECK_ARG(smem <= 227 * 1024);
  dim3 grid((unsigned)((int64_t)a->B * p.tt * p.th * p.tw), (unsigned)ceil_div(a->Co, bn));
  if (a->epi_mode == 1) launch_k(tc_conv_kernel<EPI_GEGLU>, dim3(grid), dim3(192), smem, (cudaStream_t)stream, p);
  else if (a->shuffle != MV2_SHUFFLE_NONE) launch_k(tc_conv_kernel<EPI_SHUFFLE>, dim3(grid), dim3(192), smem, (cudaStream_t)stream, p);
  else launch_k(tc_conv_kernel<EPI_RAGGED>, dim3(grid), dim3(192), smem, (cudaStream_t)stream, p);
  MV2_CHECK_LAUNCH();
  return MV2_OK;
}

}  // extern "C"
