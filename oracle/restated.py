"""TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may import this module, and only as the
checker / timed CPU baseline -- never as part of the product path.

CPU restatement (plain torch functional ops on CPU tensors) of the reference's
VideoTokenizer inference path: tokenize / decode_from_code_indices /
forward(return_recon).  It is driven purely by a ``state_dict`` with the
reference's key names plus the constructor kwargs, so it runs on the GPU box
where /root/reference does not exist.

Pinning: the reference has no tests and no golden vectors, so
this restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in
the build container by oracle/make_golden.py (reference source imported
unmodified through oracle/ref_loader.py) and committed under tests/golden/.
tests/test_oracle.py checks this file against those vectors bit-for-bit (codes)
and to fp32 round-off (activations).  The quantiser / Taylor-attention
arithmetic lives in un-vendored PyPI dependencies and is restated from their
published algorithm: that part is "parity unpinned".

Every function cites the reference lines it follows; M: = magvit2_pytorch/
magvit2_pytorch.py, A: = magvit2_pytorch/attend.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# layer schedule (M:1138-1318)
# ----------------------------------------------------------------------------

@dataclass
class Stage:
    kind: str                 # residual | cond_residual | compress_space | compress_time | attend_space | linear_attend_space | attend_time
    dim: int
    dim_out: int
    count: int = 1            # consecutive residual units
    nested: bool = False      # 'consecutive_residual' adds a '.{j}.' level to the keys


def schedule(layers, init_dim, max_dim) -> Tuple[List[Stage], int, int]:
    """Walks the ``layers`` spec exactly as the constructor loop does (M:1129-1318)
    and returns (stages, fmap_downsample_pow2, time_downsample_factor)."""
    dim = init_dim
    stages = []
    space_f, time_f = 1, 1
    for ld in layers:
        kind, *params = ld if isinstance(ld, tuple) else (ld,)
        dim_out = dim
        if kind == "residual":
            stages.append(Stage("residual", dim, dim, 1, False))
        elif kind == "consecutive_residual":
            stages.append(Stage("residual", dim, dim, int(params[0]), True))
        elif kind == "cond_residual":                                   # M:1150-1157
            stages.append(Stage("cond_residual", dim, dim))
        elif kind in ("compress_space", "compress_time"):
            dim_out = params[0] if len(params) > 0 else dim * 2      # M:1160-1162
            dim_out = int(min(dim_out, max_dim))
            stages.append(Stage(kind, dim, dim_out))
            if kind == "compress_space":
                space_f *= 2
            else:
                time_f *= 2
        elif kind in ("attend_space", "linear_attend_space", "attend_time", "gateloop_time"):
            stages.append(Stage(kind, dim, dim))
        else:
            raise ValueError(f"oracle: unsupported layer type {kind}")
        dim = dim_out
    return stages, space_f, time_f


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------

def causal_conv3d(x, w, b, pad_mode="constant"):
    """CausalConv3d.forward, stride 1 / dilation 1 (M:913-928): pad (k_t - 1) frames
    at the FRONT of time only, k//2 on both sides of H and W, then a plain conv3d."""
    kt, kh, kw = w.shape[2:]
    time_pad = kt - 1
    mode = pad_mode if time_pad < x.shape[2] else "constant"          # M:925
    x = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2, time_pad, 0), mode=mode)
    return F.conv3d(x, w, b)


def causal_conv_transpose3d(x, w, b, time_stride):
    """CausalConvTranspose3d.forward (M:990-1024): ConvTranspose3d stride (s, 1, 1), padding (0, k_h//2, k_w//2), then the
    output is cut to t * s frames.  w: (C_in, C_out, k_t, k_h, k_w)."""
    kh, kw = w.shape[3:]
    out = F.conv_transpose3d(x, w, b, stride=(time_stride, 1, 1), padding=(0, kh // 2, kw // 2))
    return out[:, :, :x.shape[2] * time_stride]


def squeeze_excite(x, sd, p):
    """SqueezeExcite.forward on video (M:221-240): per (b, f) frame, softmax over h*w of
    a 1x1 conv logit, pooled C-vector, 2-layer MLP with LeakyReLU(0.1), sigmoid gate."""
    b, c, f, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    logits = F.conv2d(xf, sd[p + "to_k.weight"], sd[p + "to_k.bias"])           # (bf,1,h,w)
    attn = logits.reshape(b * f, 1, h * w).softmax(dim=-1)
    pooled = torch.einsum("bin,bcn->bci", attn, xf.reshape(b * f, c, h * w))    # (bf,c,1)
    pooled = pooled[..., None]
    hid = F.leaky_relu(F.conv2d(pooled, sd[p + "net.0.weight"], sd[p + "net.0.bias"]), 0.1)
    gates = torch.sigmoid(F.conv2d(hid, sd[p + "net.2.weight"], sd[p + "net.2.bias"]))
    gates = gates.reshape(b, f, c, 1, 1).permute(0, 2, 1, 3, 4)
    return gates * x


def residual_unit(x, sd, p):
    """Residual(Sequential(CausalConv3d, ELU, Conv3d 1x1x1, ELU, SqueezeExcite)) (M:930-944, M:167-174).
    Residual units always use pad_mode 'constant' (M:1142-1148 never forwards pad_mode)."""
    y = F.elu(causal_conv3d(x, sd[p + "fn.0.conv.weight"], sd[p + "fn.0.conv.bias"]))
    y = F.elu(F.conv3d(y, sd[p + "fn.2.weight"], sd[p + "fn.2.bias"]))
    y = squeeze_excite(y, sd, p + "fn.4.")
    return y + x


def conv3d_mod(x, cond, w, eps=1e-8):
    """Conv3DMod.forward (M:718-753), StyleGAN2-style modulated causal conv, demod=True, zero padding.
    x (B,C,T,H,W), cond (B,C) already through the unit's ``to_cond``; w (O,I,kt,kh,kw) shared by the batch.
    Per-sample weights w_b = w * (cond_b + 1) over the input channels, demodulated by rsqrt(sum w_b^2) per output
    channel (clamped at eps), applied as one grouped conv over the batch (groups = B) -- restated literally."""
    b = x.shape[0]
    o, i, kt, kh, kw = w.shape
    wb = w[None] * (cond[:, None, :, None, None, None] + 1.)                                  # M:736-738
    inv_norm = (wb ** 2).sum(dim=(2, 3, 4, 5), keepdim=True).clamp(min=eps).rsqrt()          # M:741
    wb = wb * inv_norm                                                                        # M:742
    xg = x.reshape(1, b * i, *x.shape[2:])                                                    # M:744
    xg = F.pad(xg, (kw // 2, kw // 2, kh // 2, kh // 2, kt - 1, 0))                           # M:708-709, M:748 ('zeros' == constant 0)
    y = F.conv3d(xg, wb.reshape(b * o, i, kt, kh, kw), groups=b)                              # M:746-749
    return y.reshape(b, o, *y.shape[2:])                                                      # M:751


def residual_unit_mod(x, cond, sd, p):
    """ResidualUnitMod.forward (M:978-988): x + ELU(Conv1x1x1(ELU(Conv3DMod(x, to_cond(cond))))).  No SqueezeExcite."""
    c = F.linear(cond, sd[p + "to_cond.weight"], sd[p + "to_cond.bias"])                      # M:983
    y = F.elu(conv3d_mod(x, c, sd[p + "conv.weights"]))                                       # M:985-986
    y = F.elu(F.conv3d(y, sd[p + "conv_out.weight"], sd[p + "conv_out.bias"]))                # M:987-988
    return y + x


def spatial_down(x, sd, p):
    """SpatialDownsample2x (M:757-780): per-frame Conv2d k3 s2 p1, no blur (antialias False)."""
    b, c, t, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    o = F.conv2d(xf, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2, padding=1)
    return o.reshape(b, t, *o.shape[1:]).permute(0, 2, 1, 3, 4)


def time_down(x, sd, p):
    """TimeDownsample2x (M:782-807): per pixel, F.pad(t,(2,0)) then Conv1d k3 s2."""
    b, c, t, h, w = x.shape
    xs = x.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, t)
    xs = F.pad(xs, (2, 0))
    o = F.conv1d(xs, sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2)
    return o.reshape(b, h, w, o.shape[1], o.shape[2]).permute(0, 3, 4, 1, 2)


def spatial_up(x, sd, p):
    """SpatialUpsample2x (M:811-846): Conv2d 1x1 C->4*Cout, SiLU, 'b (c p1 p2) h w -> b c (h p1) (w p2)'."""
    b, c, t, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    o = F.silu(F.conv2d(xf, sd[p + "net.0.weight"], sd[p + "net.0.bias"]))
    co = o.shape[1] // 4
    o = o.reshape(b * t, co, 2, 2, h, w).permute(0, 1, 4, 2, 5, 3).reshape(b * t, co, h * 2, w * 2)
    return o.reshape(b, t, co, h * 2, w * 2).permute(0, 2, 1, 3, 4)


def time_up(x, sd, p):
    """TimeUpsample2x (M:848-883): Conv1d 1x1 C->2*Cout, SiLU, 'b (c p) t -> b c (t p)'."""
    b, c, t, h, w = x.shape
    xs = x.permute(0, 3, 4, 1, 2).reshape(b * h * w, c, t)
    o = F.silu(F.conv1d(xs, sd[p + "net.0.weight"], sd[p + "net.0.bias"]))
    co = o.shape[1] // 2
    o = o.reshape(b * h * w, co, 2, t).permute(0, 1, 3, 2).reshape(b * h * w, co, t * 2)
    return o.reshape(b, h, w, co, t * 2).permute(0, 3, 4, 1, 2)


def rmsnorm_last(x, gamma):
    """RMSNorm over the last dim (M:275-276): F.normalize (eps 1e-12 on the L2 norm) * sqrt(C) * gamma."""
    c = x.shape[-1]
    return F.normalize(x, dim=-1) * (c ** 0.5) * gamma.reshape(-1)


def token_shift(x):
    """TokenShift (M:250-254): second half of the channels delayed by one frame, zero at t=0."""
    a, s = x.chunk(2, dim=1)
    s = F.pad(s, (0, 0, 0, 0, 1, -1))
    return torch.cat((a, s), dim=1)


def softmax_attention(q, k, v, causal):
    """Attend.forward non-flash branch (A:218-241) which the flash branch reproduces with a
    right-aligned bool mask (A:123-129): query i attends keys j <= i + (k_len - q_len)."""
    scale = q.shape[-1] ** -0.5
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    i, j = dots.shape[-2:]
    if causal and i > 1:                                                # A:209-210
        mask = torch.ones((i, j), dtype=torch.bool).triu(j - i + 1)     # A:46-47
        dots = dots.masked_fill(mask, -torch.finfo(dots.dtype).max)
    return torch.einsum("bhij,bhjd->bhid", dots.softmax(dim=-1), v)


def attention_tokens(x, sd, p, heads, causal):
    """Attention.forward on (batch, n, C) tokens (M:371-388): RMSNorm, qkv Linear (no bias),
    4 learned memory key/values prepended, attend, output Linear (no bias)."""
    b, n, c = x.shape
    xn = rmsnorm_last(x, sd[p + "norm.gamma"])
    qkv = F.linear(xn, sd[p + "to_qkv.0.weight"])
    qkv = qkv.reshape(b, n, 3, heads, -1).permute(2, 0, 3, 1, 4)       # 'b n (qkv h d) -> qkv b h n d'
    q, k, v = qkv[0], qkv[1], qkv[2]
    mem = sd[p + "mem_kv"].to(x.dtype)
    mk = mem[0][None].expand(b, -1, -1, -1)
    mv = mem[1][None].expand(b, -1, -1, -1)
    k = torch.cat((mk, k), dim=-2)
    v = torch.cat((mv, v), dim=-2)
    o = softmax_attention(q, k, v, causal)
    o = o.permute(0, 2, 1, 3).reshape(b, n, -1)
    return F.linear(o, sd[p + "to_out.1.weight"])


def space_attention(x, sd, p, heads):
    """SpaceAttention (M:444-454): fold time into batch, tokens = h*w."""
    b, c, t, h, w = x.shape
    tok = x.permute(0, 2, 3, 4, 1).reshape(b * t, h * w, c)
    o = attention_tokens(tok, sd, p, heads, causal=False)
    return o.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3)


def time_attention(x, sd, p, heads):
    """TimeAttention (M:456-464): fold space into batch, tokens = t, causal."""
    b, c, t, h, w = x.shape
    tok = x.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)
    o = attention_tokens(tok, sd, p, heads, causal=True)
    return o.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)


def taylor_linear_attention(x, sd, p, heads, dim_head):
    """TaylorSeriesLinearAttn, non-causal (un-vendored dependency):
    phi(x) = [1, x, x(x)x/sqrt2]; out = phi(q) . sum_n phi(k)_n (x) v_n / max(phi(q).sum_n phi(k)_n, 1e-5)."""
    b, n, c = x.shape
    q = F.linear(x, sd[p + "attn.to_q.0.weight"]).reshape(b, n, heads, dim_head).permute(0, 2, 1, 3)
    kv = F.linear(x, sd[p + "attn.to_kv.0.weight"]).reshape(b, n, 2, heads, dim_head).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    q = q * dim_head ** -0.5

    def phi(z):
        one = z.new_ones((*z.shape[:-1], 1))
        z2 = (z[..., :, None] * z[..., None, :]) * (0.5 ** 0.5)
        return torch.cat((one, z, z2.reshape(*z.shape[:-1], -1)), dim=-1)

    q, k = phi(q), phi(k)
    kvs = torch.einsum("bhnd,bhne->bhde", k, v)
    ksum = k.sum(dim=-2)
    num = torch.einsum("bhnd,bhde->bhne", q, kvs)
    den = torch.einsum("bhnd,bhd->bhn", q, ksum)[..., None]
    o = num / den.clamp(min=1e-5)
    o = o.permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
    return F.linear(o, sd[p + "attn.to_out.0.weight"])


def linear_space_attention(x, sd, p, heads, dim_head):
    """LinearSpaceAttention (M:421-442): RMSNorm then Taylor attention over h*w tokens per frame."""
    b, c, t, h, w = x.shape
    tok = x.permute(0, 2, 3, 4, 1).reshape(b * t, h * w, c)
    tok = rmsnorm_last(tok, sd[p + "norm.gamma"])
    o = taylor_linear_attention(tok, sd, p, heads, dim_head)
    return o.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3)


def gateloop_time(x, sd, p):
    """ToTimeSequence(SimpleGateLoopLayer) (M:178-191, M:1220; un-vendored dependency gateloop-transformer, restated in
    oracle/shims/gateloop.py): per pixel, over time: RMSNorm, q / kv / a = Linear(dim, 3 dim) chunks, the gated recurrence
    s_t = sigmoid(a_t) s_{t-1} + kv_t, out_t = q_t s_t."""
    b, c, t, h, w = x.shape
    tok = x.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)
    tok = rmsnorm_last(tok, sd[p + "norm.gamma"])
    q, kv, a = F.linear(tok, sd[p + "to_qkva.0.weight"]).chunk(3, dim=-1)
    a = a.sigmoid()
    s = torch.zeros_like(kv[:, 0])
    outs = []
    for i in range(t):
        s = a[:, i] * s + kv[:, i]
        outs.append(q[:, i] * s)
    o = torch.stack(outs, dim=1)
    return o.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)


def feed_forward(x, sd, p):
    """FeedForward (M:471-508): channel-first RMSNorm, Conv3d 1x1x1 C->2I, GEGLU
    (x, gate = chunk; gelu(gate) * x, M:466-469), Conv3d 1x1x1 I->C."""
    xl = x.permute(0, 2, 3, 4, 1)
    xl = rmsnorm_last(xl, sd[p + "norm.gamma"])
    xn = xl.permute(0, 4, 1, 2, 3)
    hdn = F.conv3d(xn, sd[p + "net.0.weight"], sd[p + "net.0.bias"])
    a, gate = hdn.chunk(2, dim=1)
    hdn = F.gelu(gate) * a
    return F.conv3d(hdn, sd[p + "net.2.weight"], sd[p + "net.2.bias"])


# ----------------------------------------------------------------------------
# quantisers (un-vendored dependency)
# ----------------------------------------------------------------------------

def lfq_presign(x, sd, clamp=10., nc=1, spherical=False):
    """project_in + tanh soft clamp (+ per-codebook L2 normalisation when spherical, A.1 step 4);
    returns fp32 (B, N, nc * d) pre-sign values."""
    b, c = x.shape[:2]
    tok = x.permute(0, 2, 3, 4, 1).reshape(b, -1, c)
    p = F.linear(tok, sd["quantizers.project_in.weight"], sd["quantizers.project_in.bias"])
    if clamp is not None:
        p = (p / clamp).tanh() * clamp
    if spherical:
        p = F.normalize(p.reshape(b, p.shape[1], nc, -1), dim=-1).reshape(p.shape)
    return p.float()


def lfq_quantize(x, sd, clamp=10., nc=1, spherical=False):
    """LFQ eval forward (A.1 steps 1-6, 9): returns (quantized (B,C,T,H,W), indices int64 (B,T,H,W[,nc]), presign)."""
    b, c, t, h, w = x.shape
    p = lfq_presign(x, sd, clamp, nc, spherical)
    q = torch.where(p > 0, torch.ones_like(p), -torch.ones_like(p))
    mask = sd["quantizers.mask"].to(torch.int32)
    idx = ((q.reshape(b, -1, nc, mask.numel()) > 0).int() * mask).sum(dim=-1)      # int64 (torch.sum promotes)
    out = F.linear(q.to(x.dtype), sd["quantizers.project_out.weight"], sd["quantizers.project_out.bias"])
    out = out.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3)
    idx = idx.reshape(b, t, h, w, nc)
    return out, (idx[..., 0] if nc == 1 else idx), p                               # keep_num_codebooks_dim = nc > 1


def lfq_indices_to_codes(idx, sd, dtype, nc=1):
    """LFQ.indices_to_codes (A.1 last line): bits -> +-1 -> project_out -> channel first."""
    mask = sd["quantizers.mask"]
    if nc == 1:
        idx = idx[..., None]
    bits = ((idx[..., None].to(torch.int64) & mask) != 0).to(dtype)                # (..., nc, d)
    codes = (bits * 2 - 1).reshape(*idx.shape[:-1], -1)
    out = F.linear(codes, sd["quantizers.project_out.weight"].to(dtype), sd["quantizers.project_out.bias"].to(dtype))
    return out.movedim(-1, 1)


def lfq_train_losses(p, d_bits, world_reduce=None, inv_temperature=100., diversity_gamma=2.5,
                     entropy_w=0.1, commit_w=1.0, nc=1):
    """LFQ training-mode auxiliary terms (A.1 steps 7, 8, 10) from fp32 pre-sign values p (B, N, nc * d).
    ``world_reduce`` optionally maps the local avg_prob (nc, K) to the cross-rank mean (the 4 KiB all-reduce)."""
    K = 2 ** d_bits
    codes = torch.arange(K)
    mask = 2 ** torch.arange(d_bits - 1, -1, -1)
    codebook = ((codes[:, None] & mask) != 0).float() * 2 - 1
    x = p.reshape(-1, nc, d_bits).float()
    logits = 2 * inv_temperature * torch.einsum("tcd,kd->tck", x, codebook)
    prob = logits.softmax(dim=-1)

    def ent(pr):
        return (-pr * torch.log(pr.clamp(min=1e-5))).sum(dim=-1)

    per_sample = ent(prob).mean()
    avg = prob.mean(dim=0)                                            # (nc, K)
    if nc == 1:
        avg = avg[0]
    if world_reduce is not None:
        avg = world_reduce(avg)
    batch_ent = ent(avg).mean()
    q = torch.where(x > 0, torch.ones_like(x), -torch.ones_like(x))
    commit = ((x - q) ** 2).mean()
    aux = (per_sample - diversity_gamma * batch_ent) * entropy_w + commit * commit_w
    return per_sample, batch_ent, commit, aux, avg


def _fsq_consts(levels, dtype=torch.float32):
    lv = torch.tensor(levels, dtype=torch.int32)
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0).to(torch.int32)
    return lv, basis


def fsq_quantize(x, sd, levels, nc=1):
    """FSQ forward (A.2): project_in, tanh bound, round-half-even, mixed-radix int32 index per codebook, project_out."""
    b, c, t, h, w = x.shape
    lv, basis = _fsq_consts(levels)
    tok = x.permute(0, 2, 3, 4, 1).reshape(b, -1, c)
    z = F.linear(tok, sd["quantizers.project_in.weight"], sd["quantizers.project_in.bias"])
    zf = z if z.dtype in (torch.float32, torch.float64) else z.float()
    zf = zf.reshape(b, -1, nc, len(levels))
    half_l = (lv - 1) * (1 + 1e-3) / 2
    offset = torch.where(lv % 2 == 0, 0.5, 0.0)
    shift = (offset / half_l).atanh()
    bounded = (zf + shift).tanh() * half_l - offset
    half_w = lv // 2
    codes = (bounded + (bounded.round() - bounded).detach()) / half_w              # round with the straight-through gradient (A.2)
    idx = ((codes * half_w + half_w) * basis).sum(dim=-1).to(torch.int32)          # (b, n, nc)
    out = F.linear(codes.reshape(b, -1, nc * len(levels)).to(x.dtype), sd["quantizers.project_out.weight"],
                   sd["quantizers.project_out.bias"])
    out = out.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3)
    idx = idx.reshape(b, t, h, w, nc)
    return out, (idx[..., 0] if nc == 1 else idx), bounded.reshape(b, -1, nc * len(levels))


def fsq_indices_to_codes(idx, sd, levels, dtype, nc=1):
    lv, basis = _fsq_consts(levels)
    if nc == 1:
        idx = idx[..., None]
    nonneg = (idx[..., None] // basis) % lv
    half_w = lv // 2
    codes = ((nonneg - half_w) / half_w).to(dtype).reshape(*idx.shape[:-1], -1)
    out = F.linear(codes, sd["quantizers.project_out.weight"].to(dtype), sd["quantizers.project_out.bias"].to(dtype))
    return out.movedim(-1, 1)


# ----------------------------------------------------------------------------
# the tokenizer
# ----------------------------------------------------------------------------

class OracleTokenizer:
    """Functional restatement of VideoTokenizer's inference path (M:1045-1720)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], *, image_size, layers=("residual",) * 3,
                 codebook_size=None, channels=3, init_dim=64, max_dim=float("inf"),
                 use_fsq=False, fsq_levels=None, attn_dim_head=32, attn_heads=8,
                 linear_attn_dim_head=8, linear_attn_heads=16, pad_mode="constant",
                 lfq_soft_clamp_input_value=10., dim_cond=None, dim_cond_expansion_factor=4.,
                 separate_first_frame_encoding=False, num_codebooks=1, lfq_spherical=False, dtype=torch.float32, **unused):
        self.dtype = dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in state_dict.items()
                   if not k.startswith("discr.")}
        self.image_size = image_size
        self.layers = tuple(layers)
        self.stages, self.space_f, self.time_f = schedule(self.layers, init_dim, max_dim)
        self.time_padding = self.time_f - 1                                  # M:1328-1329
        self.fmap_size = image_size // self.space_f                         # M:1168, M:1331
        self.use_fsq = use_fsq
        self.fsq_levels = fsq_levels
        self.codebook_size = codebook_size
        self.heads, self.dim_head = attn_heads, attn_dim_head
        self.lin_heads, self.lin_dim_head = linear_attn_heads, linear_attn_dim_head
        self.pad_mode = pad_mode
        self.clamp = lfq_soft_clamp_input_value
        self.nc, self.spherical = int(num_codebooks), bool(lfq_spherical)            # M:1057, M:1070
        self.channels = channels
        # conditioning (M:1134-1153, M:1318, M:1336-1352): ``has_cond`` is set by the first cond layer and never reset, so
        # every later layer is called with ``cond=`` too -- the reference's plain layers then raise TypeError.  Only specs
        # whose conditioned layers are the trailing ones run in the reference; anything else is rejected here as well.
        kinds = [st.kind for st in self.stages]
        self.has_cond = "cond_residual" in kinds
        if self.has_cond:
            first = kinds.index("cond_residual")
            if any(k != "cond_residual" for k in kinds[first:]):
                raise TypeError("a non-conditioned layer after a cond_* layer receives cond= in the reference (M:1153, M:1318) and fails")
            assert dim_cond is not None, "dim_cond must be passed into VideoTokenizer, if tokenizer is to be conditioned"   # M:1151
        self.dim_cond = dim_cond
        self.sep_first = bool(separate_first_frame_encoding)                     # M:1113-1120

    def _cond_in(self, cond, which):
        """encoder_cond_in / decoder_cond_in (M:1344-1352): Linear + SiLU stem."""
        assert cond is not None, "`cond` must be passed into tokenizer forward method since conditionable layers were specified"  # M:1542
        assert tuple(cond.shape[1:]) == (self.dim_cond,)                                     # M:1545
        return F.silu(F.linear(cond.to(self.dtype), self.sd[f"{which}_cond_in.0.weight"], self.sd[f"{which}_cond_in.0.bias"]))

    # one encoder/decoder stage --------------------------------------------------
    def _apply(self, x, st: Stage, p: str, decoder: bool, taps=None, cond=None):
        sd = self.sd
        if st.kind == "cond_residual":
            x = residual_unit_mod(x, cond, sd, p)
        elif st.kind == "residual":
            if st.nested:
                for j in range(st.count):
                    x = residual_unit(x, sd, f"{p}{j}.")
            else:
                x = residual_unit(x, sd, p)
        elif st.kind == "compress_space":
            x = spatial_up(x, sd, p) if decoder else spatial_down(x, sd, p)
        elif st.kind == "compress_time":
            x = time_up(x, sd, p) if decoder else time_down(x, sd, p)
        elif st.kind == "attend_space":                                      # M:1189-1197
            x = space_attention(x, sd, p + "0.fn.", self.heads) + x
            x = feed_forward(x, sd, p + "1.fn.") + x
        elif st.kind == "linear_attend_space":                               # M:1206-1214
            x = linear_space_attention(x, sd, p + "0.fn.", self.lin_heads, self.lin_dim_head) + x
            x = feed_forward(x, sd, p + "1.fn.") + x
        elif st.kind == "gateloop_time":                                     # M:1216-1222
            x = gateloop_time(x, sd, p + "fn.fn.") + x
        elif st.kind == "attend_time":                                       # M:1234-1242
            x = time_attention(token_shift(x), sd, p + "0.fn.fn.", self.heads) + x
            x = feed_forward(token_shift(x), sd, p + "1.fn.fn.") + x
        else:
            raise ValueError(st.kind)
        return x

    def _encode(self, video, taps=None, cond=None, video_contains_first_frame=True):
        """VideoTokenizer.encode (M:1523-1576).  NB the final LayerNorm (M:1322-1326) is never
        executed: zip() with has_cond_across_layers truncates it (M:1565).  Without a first frame
        (video_contains_first_frame=False) the clip is neither front-padded nor split (M:1530-1537)."""
        x = video.to(self.dtype)
        if not video_contains_first_frame:
            x = causal_conv3d(x, self.sd["conv_in.conv.weight"], self.sd["conv_in.conv.bias"], self.pad_mode)
        elif self.sep_first:
            # M:1553-1561: the first frame goes through its own 2-D conv (SameConv2d, M:887-890), the remaining frames
            # through the causal conv_in on their own (their causal padding starts at frame 1), then the feature map is
            # re-padded with time_padding zero frames
            w2, b2 = self.sd["conv_in_first_frame.weight"], self.sd["conv_in_first_frame.bias"]
            first = F.conv2d(x[:, :, 0], w2, b2, padding=(w2.shape[2] // 2, w2.shape[3] // 2))
            rest = causal_conv3d(x[:, :, 1:], self.sd["conv_in.conv.weight"], self.sd["conv_in.conv.bias"], self.pad_mode)
            x = torch.cat((first[:, :, None], rest), dim=2)
            x = F.pad(x, (0, 0, 0, 0, self.time_padding, 0))                  # M:1561
        else:
            x = F.pad(x, (0, 0, 0, 0, self.time_padding, 0))                  # M:1537
            x = causal_conv3d(x, self.sd["conv_in.conv.weight"], self.sd["conv_in.conv.bias"], self.pad_mode)
        if taps is not None:
            taps["conv_in"] = x
        c = self._cond_in(cond, "encoder") if self.has_cond else None               # M:1544-1548
        for i, st in enumerate(self.stages):
            x = self._apply(x, st, f"encoder_layers.{i}.", decoder=False, cond=c)
            if taps is not None:
                taps[f"enc{i}"] = x
        return x

    def _decode(self, quantized, taps=None, cond=None, video_contains_first_frame=True):
        """VideoTokenizer.decode (M:1598-1649): decoder layers are the encoder's in reverse
        (insert(0), M:1315); conv_out; drop the first time_padding frames (M:1646-1647) when the clip had a first frame."""
        x = quantized.to(self.dtype)
        n = len(self.stages)
        c = self._cond_in(cond, "decoder") if self.has_cond else None               # M:1612-1616
        for j, st in enumerate(reversed(self.stages)):
            x = self._apply(x, st, f"decoder_layers.{j}.", decoder=True, cond=c)
            if taps is not None:
                taps[f"dec{j}"] = x
        if self.sep_first and video_contains_first_frame:                      # M:1633-1639
            tp = self.time_padding
            w2, b2 = self.sd["conv_out_first_frame.weight"], self.sd["conv_out_first_frame.bias"]
            first = F.conv2d(x[:, :, tp], w2, b2, padding=(w2.shape[2] // 2, w2.shape[3] // 2))
            rest = causal_conv3d(x[:, :, tp + 1:], self.sd["conv_out.conv.weight"], self.sd["conv_out.conv.bias"], self.pad_mode)
            return torch.cat((first[:, :, None], rest), dim=2)
        x = causal_conv3d(x, self.sd["conv_out.conv.weight"], self.sd["conv_out.conv.bias"], self.pad_mode)
        return x[:, :, self.time_padding:] if video_contains_first_frame else x

    @torch.no_grad()
    def encode(self, *a, **k):
        return self._encode(*a, **k)

    @torch.no_grad()
    def decode(self, *a, **k):
        return self._decode(*a, **k)

    def loss_forward(self, video, train: bool, world_reduce=None, lfq_entropy_loss_weight=0.1, lfq_commitment_loss_weight=1.,
                     lfq_diversity_gamma=2.5, quantizer_aux_loss_weight=1., cond=None):
        """forward(video, return_loss=True) of a tokenizer built with use_gan=False, perceptual_loss_weight=0 (M:1695-1727,
        M:1868-1896): total_loss = recon_loss + aux_loss * quantizer_aux_loss_weight.  Differentiable: with state_dict tensors
        that require grad, ``out["total_loss"].backward()`` yields the parameter gradients the trainer consumes (T:356-363).
        train=True is the LFQ training branch (M:1705; A.1 steps 6-8, 10: straight-through output, entropy + commitment terms);
        train=False (and FSQ) has zero auxiliary loss (M:1700-1703)."""
        video, ff = self._check_video(video, True)
        x = self._encode(video, cond=cond, video_contains_first_frame=ff)
        out = {}
        if self.use_fsq or not train:
            q, idx, _ = self.quantize.__wrapped__(self, x)
            aux = torch.zeros((), dtype=video.dtype)
        else:
            b, c, t, h, w = x.shape
            p = lfq_presign(x, self.sd, self.clamp, self.nc, self.spherical)  # (B, N, nc d) fp32, differentiable
            qd = torch.where(p > 0, torch.ones_like(p), -torch.ones_like(p))
            st = p + (qd - p).detach()                                       # straight-through (A.1 step 6)
            d_bits = p.shape[-1] // self.nc
            ps, be, cm, aux, _ = lfq_train_losses(p, d_bits, world_reduce, 100., lfq_diversity_gamma, lfq_entropy_loss_weight,
                                                  lfq_commitment_loss_weight, self.nc)
            out.update(per_sample_entropy=ps, batch_entropy=be, commitment=cm)
            q = F.linear(st.to(x.dtype), self.sd["quantizers.project_out.weight"], self.sd["quantizers.project_out.bias"])
            q = q.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3)
            mask = self.sd["quantizers.mask"].to(torch.int32)
            idx = ((qd.reshape(b, -1, self.nc, d_bits) > 0).int() * mask).sum(dim=-1).reshape(b, t, h, w, self.nc)
            if self.nc == 1:
                idx = idx[..., 0]
        recon = self._decode(q, cond=cond, video_contains_first_frame=ff)
        recon_loss = F.mse_loss(video.to(recon.dtype), recon)                # M:1722
        out.update(codes=idx, recon=recon, recon_loss=recon_loss, aux=aux, total_loss=recon_loss + aux * quantizer_aux_loss_weight)
        return out

    @torch.no_grad()
    def quantize(self, x):
        if self.use_fsq:
            return fsq_quantize(x, self.sd, self.fsq_levels, self.nc)
        return lfq_quantize(x, self.sd, self.clamp, self.nc, self.spherical)

    def _check_video(self, video, video_contains_first_frame=True):
        assert video.ndim in (4, 5)                                            # M:1675
        assert tuple(video.shape[-2:]) == (self.image_size, self.image_size)   # M:1677
        if video.ndim == 4:
            video = video[:, :, None]                                          # M:1684
            video_contains_first_frame = True                                  # M:1685
        assert (video.shape[2] - int(video_contains_first_frame)) % self.time_f == 0     # M:1691
        return video, bool(video_contains_first_frame)

    @torch.no_grad()
    def tokenize(self, video, taps=None, return_presign=False, cond=None, video_contains_first_frame=True):
        """VideoTokenizer.tokenize (M:1651-1654) = forward(return_codes=True) (M:1695-1708).  NB the reference's
        tokenize() does not forward ``cond`` / ``video_contains_first_frame``; such calls go through
        forward(video, cond, return_codes=True, video_contains_first_frame=...)."""
        video, ff = self._check_video(video, video_contains_first_frame)
        x = self.encode(video, taps, cond=cond, video_contains_first_frame=ff)
        _, idx, pre = self.quantize(x)
        return (idx, pre) if return_presign else idx

    @torch.no_grad()
    def decode_from_code_indices(self, codes, taps=None, cond=None, video_contains_first_frame=True):
        """M:1579-1595: flat (b, f*h*w) ids are un-flattened with the fmap size."""
        assert codes.dtype in (torch.long, torch.int32)
        if codes.ndim == 2:
            assert codes.shape[-1] % (self.fmap_size ** 2) == 0
            codes = codes.reshape(codes.shape[0], -1, self.fmap_size, self.fmap_size)
        if self.use_fsq:
            q = fsq_indices_to_codes(codes, self.sd, self.fsq_levels, self.dtype, self.nc)
        else:
            q = lfq_indices_to_codes(codes, self.sd, self.dtype, self.nc)
        return self.decode(q, taps, cond=cond, video_contains_first_frame=video_contains_first_frame)

    @torch.no_grad()
    def forward(self, video, return_codes=False, return_recon=False, cond=None, video_contains_first_frame=True):
        """forward up to M:1720 (inference returns only); ``cond`` reaches encode and decode (M:1695, M:1710)."""
        video, ff = self._check_video(video, video_contains_first_frame)
        x = self.encode(video, cond=cond, video_contains_first_frame=ff)
        q, idx, _ = self.quantize(x)
        if return_codes and not return_recon:
            return idx
        rec = self.decode(q, cond=cond, video_contains_first_frame=ff)
        if return_codes:
            return idx, rec
        return rec
