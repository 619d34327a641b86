/*
 * magvit2_b200.h -- C ABI of libmagvit2_b200.so, the sm_100a compute library behind
 * the VideoTokenizer forward path (tokenize / decode_from_code_indices / forward).
 *
 * The reference (lucidrains/magvit2-pytorch @ a00519fa) has NO native / FFI layer of
 * its own: its boundary is the Python class magvit2_pytorch.VideoTokenizer
 * (magvit2_pytorch/magvit2_pytorch.py:1045) whose forward dispatches ~730 ATen calls.
 * Each entry point below replaces the ATen call sequence of one reference module
 * (cited as M:line = magvit2_pytorch.py, A:line = attend.py).  The Python host class
 * magvit2_pytorch_b200.VideoTokenizer binds them with ctypes.
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs; no torch / C++ types cross the boundary.
 *   - every call returns 0 on success or a negative MV2_E_* code; mv2_last_error()
 *     returns a thread-local message.  Nothing throws across the boundary.
 *   - all device pointers are borrowed for the duration of the call; the library
 *     never allocates device memory: the caller supplies workspaces.
 *   - stream ordered, no implicit synchronisation; `stream` is a cudaStream_t passed
 *     as void*.
 *   - activations are channels-last ("NDHWC"): x[b][t][h][w][c], dtype MV2_F32 or
 *     MV2_BF16; accumulation is always fp32; biases / gammas / tiny SE + quantiser
 *     weights are fp32.
 *   - there is NO CPU fallback: every function launches sm_100a kernels.
 */
#ifndef MAGVIT2_B200_H
#define MAGVIT2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MV2_ABI_VERSION 3   /* 2: mv2_conv_args.oscale, mv2_tc_conv_args.{oscale,out_layout}; 3: num_codebooks / spherical in the quantiser entry points */

enum { MV2_F32 = 0, MV2_BF16 = 1,
       MV2_U8 = 2   /* source dtype of the two layout-in entry points only: decoded uint8 frames, normalised x / 255 */ };
enum { MV2_ACT_NONE = 0, MV2_ACT_ELU = 1, MV2_ACT_SILU = 2 };
enum { MV2_SHUFFLE_NONE = 0, MV2_SHUFFLE_SPACE = 1, MV2_SHUFFLE_TIME = 2 };
enum {
  MV2_OK = 0,
  MV2_E_ARG = -1,       /* bad argument / unsupported shape */
  MV2_E_CUDA = -2,      /* CUDA runtime / driver error (see mv2_last_error) */
  MV2_E_UNSUPPORTED = -3
};

int mv2_abi_version(void);
const char* mv2_last_error(void);
/* Compute capability of the current device as major*10+minor (100 on B200), <0 on error. */
int mv2_device_arch(void);
/* Programmatic dependent launch: when on, every kernel is launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization so its prologue (barrier init, TMEM allocation, bias staging,
 * block scheduling) overlaps the tail of the previous kernel of the stream; all kernels execute griddepcontrol.wait
 * before touching activations.  Returns the previous setting.  Off by default. */
int mv2_set_pdl(int on);

/* ---- layout: torch (B,C,T,H,W) <-> channels-last activations ------------------
 * mv2_to_channels_last : video ingest.  Replaces pad_at_dim (M:86-89, use M:1537) +
 *   the implicit layout of every later conv: dst[b][t+t_pad][h][w][c] = src[b][c][t][h][w],
 *   frames [0,t_pad) of dst are zero-filled.
 * mv2_to_channels_first: replaces the frame crop at M:1646-1647 and the layout return:
 *   dst[b][c][t][h][w] = src[b][t+t_crop][h][w][c],  dst has T - t_crop frames.            */
int mv2_to_channels_last(const void* src, int src_dtype, void* dst, int dst_dtype,
                         int B, int C, int T, int H, int W, int t_pad, void* stream);
int mv2_to_channels_first(const void* src, int src_dtype, void* dst, int dst_dtype,
                          int B, int C, int T, int H, int W, int t_crop, void* stream);
/* mv2_ingest_kwpack: ingest for the tensor-core conv_in (M:1109, 7x7x7 with C_in = 3): besides the layout change and
 *   the time_padding zero frames it packs the k_w taps into the channel axis,
 *   dst[b][t+t_pad][h][w][dw*C + c] = src[b][c][t][h][w + dw - pw]  (bf16, cpack channels, zero padded),
 *   so conv_in becomes a (k_t x k_h x 1)-tap implicit GEMM over cpack = 32 channels.                            */
int mv2_ingest_kwpack(const void* src, int src_dtype, void* dst, int B, int C, int T, int H, int W,
                      int t_pad, int kw, int pw, int cpack, void* stream);

/* mv2_copy_frames: frame-range copy between channels-last clips (device memcpy2D, no kernel):
 *   dst[b][dst_t0 + i] = src[b][src_t0 + i], i < n_frames; zero_front != 0 also zero-fills dst frames [0, dst_t0).
 * Used to split off / re-attach the first frame for separate_first_frame_encoding (reference M:1553-1561, M:1633-1639:
 * unpack / pack / pad_at_dim on the time axis).                                                                           */
int mv2_copy_frames(const void* src, void* dst, int B, int src_T, int dst_T, int src_t0, int dst_t0, int n_frames,
                    size_t frame_bytes, int zero_front, void* stream);

/* mv2_pad_cl: explicit causal / spatial padding for CausalConv3d with pad_mode != 'constant' (reference M:925-927:
 *   F.pad(x, (pw, pw, ph, ph, pt, 0), mode)): dst (B, T+pt, H+2ph, W+2pw, C) from src (B,T,H,W,C), channels-last;
 *   mode 1 = 'reflect', 2 = 'replicate', 3 = 'circular'.  The conv then runs with zero leading padding on dst.
 *   ('constant' never materialises its padding: TMA out-of-bounds fill / bounds checks.)                                   */
int mv2_pad_cl(const void* src, void* dst, int dtype, int B, int T, int H, int W, int C, int pt, int ph, int pw, int mode,
               void* stream);

/* ---- convolution family (CUDA-core fp32-accumulate path; any shape) -----------
 * One generic strided N-d convolution over channels-last activations with a fused
 * epilogue  y = shuffle(act(conv(x) + bias)) + res.   Replaces
 *   CausalConv3d.forward           F.pad + nn.Conv3d            M:924-928  (pt = kt-1, ph = kh/2, pw = kw/2)
 *   nn.Conv3d 1x1x1 / nn.Linear    M:939, M:352, M:367, M:493-495
 *   SpatialDownsample2x.forward    Conv2d k3 s2 p1 per frame    M:770-780
 *   TimeDownsample2x.forward       F.pad(2,0) + Conv1d k3 s2    M:796-807
 *   SpatialUpsample2x / TimeUpsample2x  1x1 conv + SiLU + depth-to-space/time  M:838-846, M:875-883
 *   Residual.forward               "+ x"                         M:173-174 (res)
 *   TokenShift.forward             half-channel one-frame delay  M:250-254 (x_token_shift)
 * Weights are packed by the host as w[tap][ci][co] (tap = (dt*kh + dh)*kw + dw) in the
 * activation dtype; bias is fp32[Co] or NULL.                                            */
typedef struct mv2_conv_args {
  const void* x;       /* (B, Ti, Hi, Wi, Ci) */
  const void* w;       /* [kt*kh*kw][Ci][Co]  */
  const float* bias;   /* [Co] or NULL */
  const void* res;     /* same shape as y, or NULL */
  void* y;             /* (B, To, Ho, Wo, Co) or its depth-to-space/time shuffle */
  int32_t dtype;
  int32_t B, Ti, Hi, Wi, Ci;
  int32_t To, Ho, Wo, Co;
  int32_t kt, kh, kw;
  int32_t st, sh, sw;
  int32_t pt, ph, pw;        /* leading zero padding; trailing padding is implied by To/Ho/Wo */
  int32_t act;               /* MV2_ACT_* */
  int32_t shuffle;           /* MV2_SHUFFLE_*: SPACE: y (B,To,2Ho,2Wo,Co/4), co=(c,p1,p2); TIME: y (B,2To,Ho,Wo,Co/2), co=(c,p) */
  int32_t x_token_shift;     /* 1: input channels >= ceil(Ci/2) are read from frame t-1 (zero at t = 0) */
  const float* oscale;       /* fp32 [B][Co] or NULL: per-(clip, output channel) multiplier applied to the accumulator BEFORE
                                bias / activation -- the demodulation factor of Conv3DMod (M:741-742), see mv2_mod_prepare */
} mv2_conv_args;
int mv2_conv_forward(const mv2_conv_args* a, void* stream);

/* ---- SqueezeExcite (M:221-240) --------------------------------------------------
 * se_pool  : per frame f (F = B*T frames of P = H*W positions, C channels):
 *            logit[n] = <y[f,n,:], wk> + bk; a = softmax_n(logit); pooled[f,c] = sum_n a[n] y[f,n,c]
 *            done as chunk partials (online softmax) + a combine fused into se_gate.
 * se_gate  : gate[f,:] = sigmoid(W2 leaky_relu_0.1(W1 pooled + b1) + b2)   (fp32 [F][C])
 * gate_residual : out = gate[f(m), c] * y[m, c] + x[m, c]                    (M:240 + M:174)
 * workspace for se_pool: mv2_se_workspace_bytes(F, P, C).                                   */
size_t mv2_se_workspace_bytes(int F, int P, int C);
int mv2_se_pool(const void* y, int dtype, int F, int P, int C, const float* wk, float bk,
                void* workspace, void* stream);
int mv2_se_gate(const void* workspace, int dtype /* of the y passed to mv2_se_pool */, int F, int P, int C, int Hd,
                const float* w1, const float* b1, const float* w2, const float* b2,
                float* gates, void* stream);
int mv2_gate_residual(const void* y, const void* x, const float* gates, void* out, int dtype,
                      int F, int P, int C, void* stream);

/* SqueezeExcite + residual for small frames in ONE launch (bf16 activations; P * C <= 131072 elements per frame, C and Hd
 * multiples of 8, <= 1024): pool, gate MLP and  out = gate * y + x  (M:221-240 + M:174) by one CTA per frame.  The gate MLP
 * weights are passed as bf16 (w1 [Hd][C], w2 [C][Hd]; exact for a bf16 model); wk / b1 / b2 fp32.                          */
int mv2_se_tail_supported(int F, int P, int C, int Hd);
int mv2_se_tail(const void* y, const void* x, void* out, int F, int P, int C, int Hd, const float* wk, float bk,
                const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2, void* stream);

/* ---- RMSNorm (M:275-276): out = x / max(||x||_2, 1e-12) * sqrt(C) * gamma over the channel
 * axis of channels-last tokens; token_shift as in mv2_conv_args (M:250-254).               */
int mv2_rmsnorm(const void* x, void* out, int dtype, const float* gamma,
                int B, int T, int P, int C, int token_shift, void* stream);

/* ---- axial softmax attention core (Attention.forward M:379-388 + Attend A:186-243) ------
 * qkv: [Ntok][3*heads*dim_head] laid out '(qkv h d)' (M:353); out: [Ntok][heads*dim_head] '(h d)'.
 * Sequence s = (o, n): token(i) = o*outer_stride + n*inner_stride + i*tok_stride, i in [0, L).
 *   space attention (M:444-454): n_outer = B*T, n_inner = 1, outer_stride = H*W, tok_stride = 1, L = H*W
 *   time  attention (M:456-464): n_outer = B, outer_stride = T*H*W, n_inner = H*W, inner_stride = 1,
 *                                tok_stride = H*W, L = T, causal = 1
 * n_mem learned key/values (mem_kv fp32 [2][heads][n_mem][dim_head], M:357, M:383-385) are
 * prepended; causal masking is right aligned (A:46-47, A:123-129): query i sees mem + keys <= i;
 * it is disabled when L == 1 (A:209-210).  dim_head must be a multiple of 32, <= 96.           */
typedef struct mv2_attn_args {
  const void* qkv; void* out; const float* mem_kv;
  int32_t dtype, heads, dim_head, n_mem, causal;
  int32_t n_outer, n_inner, L;
  int64_t outer_stride, inner_stride, tok_stride;
} mv2_attn_args;
int mv2_attention(const mv2_attn_args* a, void* stream);

/* ---- Taylor-series linear attention core (TaylorSeriesLinearAttn, un-vendored dependency;
 * SURVEY.md Appendix A.3; called at M:430).  q: [Ntok][heads*8], kv: [Ntok][2*heads*8] '(kv h d)',
 * out: [Ntok][heads*8]; sequences are n_seq contiguous runs of L tokens.  dim_head must be 8.
 * workspace: mv2_linattn_workspace_bytes(n_seq, heads, L).                                    */
size_t mv2_linattn_workspace_bytes(int n_seq, int heads, int L);
int mv2_linear_attention(const void* q, const void* kv, void* out, int dtype,
                         int n_seq, int L, int heads, int dim_head, void* workspace, void* stream);

/* ---- GEGLU (M:466-469): out[n][i] = gelu_erf(in[n][I + i]) * in[n][i] ------------------------- */
int mv2_geglu(const void* in, void* out, int dtype, int64_t N, int I, void* stream);

/* ---- quantisers (un-vendored vector-quantize-pytorch LFQ / FSQ; SURVEY.md Appendix A.1/A.2;
 * reference call sites M:1576, M:1593, M:1700, M:1705; constructor kwargs num_codebooks M:1057, lfq_spherical M:1070) ------
 * d = dims per codebook, num_codebooks = nc, D = d * nc <= 16 projected dims; one index per (token, codebook):
 *   indices [N][nc].
 * lfq_forward : x [N][C] -> p = tanh((Win x + bin)/clamp)*clamp (fp32), per codebook (L2-normalised first when
 *               spherical != 0): bit_j = p_j > 0, index = sum bit_j << (d-1-j) (int64), quantized [N][C] = Wout (+-1) + bout.
 *               presign (fp32 [N][D], after the optional normalisation) is optional (diagnostics / training losses).
 * lfq_decode  : indices -> quantized (LFQ.indices_to_codes).
 * fsq_*       : same with tanh-bound + round-half-even + mixed-radix int32 index (levels[d], shared by the codebooks).
 * win [D][C], bin [D], wout [C][D], bout [C] are fp32.                                         */
int mv2_lfq_forward(const void* x, int dtype, int64_t N, int C, int d, int num_codebooks,
                    const float* win, const float* bin, const float* wout, const float* bout,
                    float clamp, int spherical, int64_t* indices, void* quantized, float* presign, void* stream);
int mv2_lfq_decode(const void* indices, int index_is_i64, int64_t N, int C, int d, int num_codebooks,
                   const float* wout, const float* bout, void* quantized, int dtype, void* stream);
int mv2_fsq_forward(const void* x, int dtype, int64_t N, int C, int d, int num_codebooks, const int32_t* levels /* host */,
                    const float* win, const float* bin, const float* wout, const float* bout,
                    int32_t* indices, void* quantized, float* bounded, void* stream);
int mv2_fsq_decode(const void* indices, int index_is_i64, int64_t N, int C, int d, int num_codebooks, const int32_t* levels /* host */,
                   const float* wout, const float* bout, void* quantized, int dtype, void* stream);

/* ---- LFQ training-mode auxiliary terms (A.1 steps 7-8; the one collective on the path) -------
 * lfq_entropy_partials: from presign [N][nc][d] (d <= 12) accumulates, for this rank,
 *   stats[0] = sum_{tokens, codebooks} H(softmax_K(2*inv_temp*<p, code_k>)), stats[1] = sum (p - sign p)^2,
 *   avg_prob[nc][K] += sum_tokens prob (un-normalised; caller divides by the token count, then the cross-rank SUM
 *   all-reduce of avg_prob -- the 4 KiB NCCL all-reduce of cfg 3).
 * stats and avg_prob must be zeroed by the caller.                                             */
int mv2_lfq_entropy_partials(const float* presign, int64_t N, int d, int num_codebooks, float inv_temperature,
                             float* avg_prob, float* stats, void* stream);

/* mv2_lfq_aux_finalize: out4 = {per_sample_entropy, batch_entropy, commitment, aux_loss} from the partial sums above
 * (A.1 steps 7-10): per_sample = stats[0] / (n_tokens nc), commitment = stats[1] / (n_tokens nc d),
 * batch_entropy = mean over codebooks of sum_k -p_k log(max(p_k, 1e-5)), p = avg_prob_sum / n_tokens_global
 * (avg_prob_sum = the cross-rank SUM), aux = (per_sample - diversity_gamma * batch_entropy) * entropy_weight
 * + commitment * commitment_weight.                                                                                    */
int mv2_lfq_aux_finalize(const float* avg_prob_sum, const float* stats, int d, int num_codebooks, int64_t n_tokens,
                         int64_t n_tokens_global, float diversity_gamma, float entropy_weight, float commitment_weight,
                         float* out4, void* stream);

/* ---- gateloop_time (reference M:1216-1222: ToTimeSequence(Residual(SimpleGateLoopLayer(dim)))) -----------
 * qkva [B][T][P][3C] = Linear(dim, 3 dim) of the RMSNorm'ed activations (q | kv | a thirds), res / out [B][T][P][C]:
 *   s_t = sigmoid(a_t) * s_{t-1} + kv_t  (s_{-1} = 0, fp32 state),   out_t = q_t * s_t + res_t
 * per (b, pixel p, channel).  The norm and the projection run through mv2_rmsnorm and the conv entry points.   */
int mv2_gateloop_scan(const void* qkva, const void* res, void* out, int dtype, int B, int T, int P, int C, void* stream);

/* ---- reconstruction loss (reference M:1722 F.mse_loss(video, recon_video)) ------------------
 * out[0] = mean_i (a[i] - b[i])^2 over n elements of two same-layout tensors; a_dtype may be MV2_U8 (frames, x / 255).
 * Deterministic (fixed-order two-stage reduction); workspace: mv2_mse_workspace_bytes() bytes.                          */
int mv2_mse(const void* a, int a_dtype, const void* b, int b_dtype, int64_t n, void* workspace, float* out, void* stream);
size_t mv2_mse_workspace_bytes(void);

/* ---- tcgen05 / TMA implicit-GEMM convolution (bf16 in, fp32 accumulate in TMEM) ------------
 * Same operator family and epilogue as mv2_conv_forward, for bf16 activations, executed on the
 * 5th-generation tensor cores: TMA box loads with out-of-bounds zero fill implement the causal /
 * spatial halo (no padded copy, reference M:924-928), strided convs read through stride-phase
 * tensor maps, accumulators live in TMEM.  Weights are packed K-major: w[co][tap][ci] (bf16);
 * for depth-to-space / depth-to-time stores the host permutes the Co rows to co' = q*Cy + c
 * (q = p1*2+p2 or p) so that the shuffled stores are channel-contiguous; bias is permuted alike.
 * Requirements (mv2_tc_conv_supported): Ci % 16 == 0, strides in {1,2}, <= 64 taps.            */
typedef struct mv2_tc_conv_args {
  const void* x;       /* bf16 (B, Ti, Hi, Wi, Ci) */
  const void* w;       /* bf16 [Co][kt*kh*kw*Ci] */
  const float* bias;   /* fp32 [Co] or NULL */
  const void* res;     /* bf16, same shape as y, or NULL */
  void* y;             /* bf16 */
  int32_t B, Ti, Hi, Wi, Ci;
  int32_t To, Ho, Wo, Co;
  int32_t kt, kh, kw;
  int32_t st, sh, sw;
  int32_t pt, ph, pw;
  int32_t act;
  int32_t shuffle;
  int32_t epi_mode;    /* 0 plain; 1 fused GEGLU (M:466-469): packed columns come in groups of 16 = 8 x-columns then
                          their 8 gate-columns, output has Co/2 channels: y = gelu_erf(gate) * x */
  const float* oscale; /* as mv2_conv_args.oscale (plain / ragged epilogues only) */
  int32_t out_layout;  /* 0: y is channels-last (B,To,Ho,Wo,Co).  1: y is torch's channels-first (B,Co,To,Ho,Wo) -- the slab
                          kernel's conv_out (Co % 8 != 0) writes the reconstruction directly in the caller's layout; with
                          To < Ti and pt = kt - 1 - (Ti - To) the leading time_padding frames are never computed
                          (reference M:1642-1647: conv_out, then video[:, :, time_padding:]).                            */
} mv2_tc_conv_args;
int mv2_tc_conv_supported(const mv2_tc_conv_args* a);
int mv2_tc_conv_forward(const mv2_tc_conv_args* a, void* stream);
/* "Slab" variant for stride-1 convs with an in-plane kernel (the causal 3x3x3 residual convs): persistent
 * CTAs, one haloed activation slab per (frame, 64-channel slice) staged in shared memory once and reused by
 * all k_h*k_w in-plane taps, two 128-position M-tiles sharing each weight tile, double-buffered TMEM
 * accumulators.  Requirements (mv2_tc_slab_supported): stride 1, Ci % 64 == 0, Co % 32 == 0, no shuffle.   */
int mv2_tc_slab_supported(const mv2_tc_conv_args* a);
int mv2_tc_slab_forward(const mv2_tc_conv_args* a, void* stream);
/* SpatialDownsample2x (M:770-780: per-frame Conv2d k3 s2 p1) on the slab design: the input is read as (W/2) x (2C) with
 * row-parity sub-slabs, so no tap reloads its tile from L2.  `a` describes the conv as usual (kh = kw = 3, sh = sw = 2,
 * ph = pw = 1, kt = 1), but w is packed as bf16 [Co][6][2*Ci]: tap' = dh * 2 + q, q = 0: [zeros(Ci) | w[:, :, dh, 0]],
 * q = 1: [w[:, :, dh, 1] | w[:, :, dh, 2]].  Requirements: Hi, Wi even, Ci % 64 == 0, Co % 32 == 0.                       */
int mv2_tc_down_space_supported(const mv2_tc_conv_args* a);
int mv2_tc_down_space_forward(const mv2_tc_conv_args* a, void* stream);
/* Launch plan of mv2_tc_slab_forward for a layer shape on a device with n_sm SMs -- pure host arithmetic (no CUDA call,
 * the pointers in `a` are not dereferenced), exposed so the tiling rule and the static tile schedule can be checked
 * without a GPU.  mv2_tc_slab_plan: out6 = {M-tiles per weight tile (mw), N tile width (bn), N tiles, total tiles,
 * grid size, TMEM accumulator buffers}.  mv2_tc_slab_tile: the k-th tile that persistent CTA `cta` processes:
 * out6 = {tile id or -1 when the CTA has no k-th tile, clip b, frame t, h0, w0, first output column n0}. */
int mv2_tc_slab_plan(const mv2_tc_conv_args* a, int n_sm, int* out6);
int mv2_tc_slab_tile(const mv2_tc_conv_args* a, int n_sm, int cta, int k, int* out6);


/* ---- conditioning (cond_residual = ResidualUnitMod / Conv3DMod, reference M:680-753, M:946-988; stems M:1344-1352) -------
 * Conv3DMod applies per-clip weights  w_b = w * (cond_b + 1)  (over input channels), demodulated by
 * rsqrt(max(sum_{i,taps} w_b^2, eps)) per output channel, as a grouped conv.  The per-clip weights never need to exist:
 *     y[b, o] = inv_norm[b, o] * conv(x[b] * (cond[b] + 1), w)[o],   inv_norm[b,o] = rsqrt(max(sum_i (cond[b,i]+1)^2 S[o,i], eps)),
 * with S[o, i] = sum_taps w[o, i, tap]^2 packed once by the host.  So the shared-weight conv kernels run unchanged with
 *   mv2_dense_small   : y[b][n] = act(sum_k x[b][k] w[n][k] + bias[n])   (cond stems: Linear + SiLU; to_cond: Linear), fp32
 *   mv2_mod_prepare   : scale_in[b][i] = cond[b][i] + 1,  inv_norm[b][o] as above                                  , fp32
 *   mv2_scale_channels: out[b, pos, c] = x[b, pos, c] * scale[b][c]        (activation dtype)
 * and mv2_*conv_args.oscale = inv_norm.                                                                                   */
int mv2_dense_small(const float* x, const float* w, const float* bias, float* y, int B, int K, int N, int act, void* stream);
int mv2_mod_prepare(const float* cond, const float* S, float eps, float* scale_in, float* inv_norm, int B, int Ci, int Co,
                    void* stream);
int mv2_scale_channels(const void* x, const float* scale, void* out, int dtype, int B, int64_t positions_per_clip, int C,
                       void* stream);

/* ---- fused ResidualUnit front half (reference M:937-941 + the pooling half of SqueezeExcite M:229-233) ----------------
 * One launch computes  y = ELU(Conv3d_1x1x1(ELU(CausalConv3d_ktxkhxkw(x))))  for C -> C channels (C = 64 or 128, the
 * HBM-bound levels of the README config): the ELU'd 3x3x3 tile never leaves the SM -- it is written to shared memory as the
 * A operand of a second tcgen05.mma against the 1x1x1 weights -- and the second epilogue emits, next to y, one SqueezeExcite
 * pool record (max logit, sum e, sum e * y[C]; e = exp(logit - max), logit = <y, se_wk> + se_bk) per TMEM lane quarter of a tile,
 * which mv2_se_gate_records combines (replaces mv2_conv_forward x2 + mv2_se_pool for these layers).
 * w3: bf16 [C][kt*kh*kw*C] (K-major, as mv2_tc_conv_args.w); w1: bf16 [C][C]; b3 / b1 / se_wk: fp32 [C].
 * se_ws: workspace of mv2_tc_ru_workspace_bytes(a) bytes; records per frame = mv2_tc_ru_records(a).                     */
typedef struct mv2_tc_ru_args {
  const void* x;        /* bf16 (B, T, H, W, C) */
  const void* w3; const float* b3;
  const void* w1; const float* b1;
  const float* se_wk; float se_bk;
  void* y;              /* bf16 (B, T, H, W, C) */
  float* se_ws;
  int32_t B, T, H, W, C;
  int32_t kt, kh, kw;
} mv2_tc_ru_args;
int mv2_tc_ru_supported(const mv2_tc_ru_args* a);
int mv2_tc_ru_records(const mv2_tc_ru_args* a);
size_t mv2_tc_ru_workspace_bytes(const mv2_tc_ru_args* a);
int mv2_tc_ru_forward(const mv2_tc_ru_args* a, void* stream);
/* SE gate from pool records in the (max, sum, acc[C]) format, nrec records per frame laid out [F][nrec][C + 2], with the
 * hidden layer scratch (F * Hd floats) right behind them: gate[f,:] = sigmoid(W2 leaky_relu_0.1(W1 pooled + b1) + b2).   */
int mv2_se_gate_records(const void* workspace, int nrec, int F, int C, int Hd,
                        const float* w1, const float* b1, const float* w2, const float* b2,
                        float* gates, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAGVIT2_B200_H */
