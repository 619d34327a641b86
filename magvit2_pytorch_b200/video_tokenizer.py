"""B200-native drop-in for the reference ``VideoTokenizer`` inference path.

Mirrors the reference class's public surface (magvit2_pytorch/magvit2_pytorch.py:1045-1720 =
M:): the keyword-only constructor and ``layers=(...)`` spec (M:1047-1092, M:1138-1318),
``tokenize`` (M:1651), ``decode_from_code_indices`` (M:1579), ``forward`` inference returns
(M:1657-1720), ``encode`` / ``decode`` (M:1523, M:1598), ``parameters()`` as a list (M:1460),
``state_dict`` key layout, ``save`` / ``load`` / ``init_and_load_from``
(M:1447-1458, M:1495-1520), ``copy_for_eval`` (M:1476), ``device`` (M:1443).

All arithmetic runs in hand-written sm_100a kernels behind the C ABI of libmagvit2_b200.so;
the compute dtype follows the parameters' dtype (``.float()`` -> fp32 CUDA-core path,
``.bfloat16()`` -> bf16 tcgen05 path).  There is no CPU / eager fallback.

``cond_residual`` layers (ResidualUnitMod / Conv3DMod, M:680-753, M:946-988) run on the device through the
factorisation in include/magvit2_b200.h; the other ``cond_*`` types raise in the reference itself.
``separate_first_frame_encoding`` (M:1113-1120, M:1553-1561, M:1633-1639) runs through the same conv kernels.
``forward(return_loss=True)`` (reconstruction + quantiser auxiliary loss; models built with ``use_gan=False,
perceptual_loss_weight=0``) runs on the device; in ``model.train()`` with gradients enabled it returns a loss with a
``grad_fn`` (train.py: forward by the same kernels, backward by library code).
Out of scope (raise at construction / call): the
GAN / perceptual training losses (``return_discr_loss``, ``return_loss`` with a discriminator or VGG).
"""
from __future__ import annotations

import contextlib
import copy
import functools
import pickle
from collections import namedtuple
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import modules as M
from .engine import Engine

__version__ = "0.1.0"


# reference M:1028-1037
LossBreakdown = namedtuple("LossBreakdown", [
    "recon_loss", "lfq_aux_loss", "quantizer_loss_breakdown", "perceptual_loss", "adversarial_gen_loss",
    "adaptive_adversarial_weight", "multiscale_gen_losses", "multiscale_gen_adaptive_weights"])


@dataclass
class Stage:
    kind: str          # residual | compress_space | compress_time | attend_space | linear_attend_space | attend_time
    dim: int
    dim_out: int
    count: int = 1
    nested: bool = False


def _on_model_device(fn):
    """Runs the method with the model's CUDA device current: the C ABI launches on the current device (kernels, TMA
    descriptors, function attributes), so a tokenizer on cuda:1 must not be driven while cuda:0 is current."""
    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        dev = self.device
        ctx = torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()
        with ctx:
            return fn(self, *a, **k)
    return wrapper


_UNSUPPORTED_LAYERS = {
    "cond_attend_space": "raises in the reference itself (SURVEY.md 2 row 9)",
    "cond_linear_attend_space": "raises in the reference itself (SURVEY.md 2 row 9)",
    "cond_attend_time": "raises in the reference itself (SURVEY.md 2 row 9)",
}


class VideoTokenizer(nn.Module):
    def __init__(
        self,
        *,
        image_size,
        layers: Tuple = ("residual", "residual", "residual"),
        residual_conv_kernel_size=3,
        num_codebooks=1,
        codebook_size: Optional[int] = None,
        channels=3,
        init_dim=64,
        max_dim=float("inf"),
        dim_cond=None,
        dim_cond_expansion_factor=4.,
        input_conv_kernel_size: Tuple[int, int, int] = (7, 7, 7),
        output_conv_kernel_size: Tuple[int, int, int] = (3, 3, 3),
        pad_mode: str = "constant",
        lfq_entropy_loss_weight=0.1,
        lfq_commitment_loss_weight=1.,
        lfq_diversity_gamma=2.5,
        lfq_spherical=False,
        quantizer_aux_loss_weight=1.,
        lfq_soft_clamp_input_value=10.,
        lfq_activation=None,
        use_fsq=False,
        fsq_levels: Optional[List[int]] = None,
        attn_dim_head=32,
        attn_heads=8,
        attn_dropout=0.,
        linear_attn_dim_head=8,
        linear_attn_heads=16,
        vgg=None,
        vgg_weights=None,
        perceptual_loss_weight=1e-1,
        discr_kwargs: Optional[dict] = None,
        multiscale_discrs: Tuple = tuple(),
        use_gan=True,
        adversarial_loss_weight=1.,
        grad_penalty_loss_weight=10.,
        multiscale_adversarial_loss_weight=1.,
        flash_attn=True,
        separate_first_frame_encoding=False,
    ):
        super().__init__()
        cfg = dict(locals())
        cfg.pop("self")
        cfg.pop("__class__", None)
        for k in ("vgg", "lfq_activation"):      # modules are not part of the pickled config here
            cfg[k] = None
        cfg["multiscale_discrs"] = tuple()
        self._configs = pickle.dumps(cfg)       # M:1097-1100

        if not isinstance(layers, tuple):
            raise TypeError("layers must be a tuple")
        if pad_mode not in ("constant", "reflect", "replicate", "circular"):
            raise ValueError(f"unknown pad_mode {pad_mode!r}")
        if attn_dropout != 0.:
            raise NotImplementedError("attention dropout is a training feature")
        ks = residual_conv_kernel_size

        self.channels = channels
        self.image_size = image_size
        self.conv_in = M.CausalConv3d(channels, init_dim, tuple(input_conv_kernel_size), pad_mode)
        self.conv_in_first_frame = nn.Identity()
        self.conv_out_first_frame = nn.Identity()
        self.separate_first_frame_encoding = bool(separate_first_frame_encoding)
        if separate_first_frame_encoding:                                          # M:1113-1120: SameConv2d (M:887-890)
            ik, ok = tuple(input_conv_kernel_size)[-2:], tuple(output_conv_kernel_size)[-2:]
            self.conv_in_first_frame = nn.Conv2d(channels, init_dim, ik, padding=(ik[0] // 2, ik[1] // 2))
            self.conv_out_first_frame = nn.Conv2d(init_dim, channels, ok, padding=(ok[0] // 2, ok[1] // 2))
        self.encoder_layers = nn.ModuleList([])
        self.decoder_layers = nn.ModuleList([])
        self.conv_out = M.CausalConv3d(init_dim, channels, tuple(output_conv_kernel_size), pad_mode)

        # ---- layer schedule (M:1129-1318) ----
        dim = init_dim
        fmap = image_size
        tdf = 1
        has_cond = False
        stages: List[Stage] = []
        for layer_def in layers:
            kind, *params = layer_def if isinstance(layer_def, tuple) else (layer_def,)
            dim_out = dim
            if kind in _UNSUPPORTED_LAYERS:
                raise NotImplementedError(f"layer type {kind!r}: {_UNSUPPORTED_LAYERS[kind]}")
            if has_cond and kind != "cond_residual":
                # has_cond is never reset in the reference (M:1153, M:1318): every later layer is called with cond=, and its
                # plain layers then raise TypeError -- only specs whose conditioned layers are the trailing ones run there
                raise TypeError(f"layer {kind!r} after a cond_* layer: the reference passes cond= to it and fails (M:1153, M:1318)")
            if kind == "residual":
                enc, dec = M.residual_unit(dim, ks), M.residual_unit(dim, ks)
                stages.append(Stage("residual", dim, dim, 1, False))
            elif kind == "cond_residual":                                        # M:1150-1157
                assert dim_cond is not None, "dim_cond must be passed into VideoTokenizer, if tokenizer is to be conditioned"
                has_cond = True
                dc = int(dim_cond * dim_cond_expansion_factor)
                enc, dec = M.ResidualUnitMod(dim, ks, dc), M.ResidualUnitMod(dim, ks, dc)
                stages.append(Stage("cond_residual", dim, dim))
            elif kind == "consecutive_residual":
                n, = params
                enc = nn.Sequential(*[M.residual_unit(dim, ks) for _ in range(n)])
                dec = nn.Sequential(*[M.residual_unit(dim, ks) for _ in range(n)])
                stages.append(Stage("residual", dim, dim, int(n), True))
            elif kind == "compress_space":
                dim_out = params[0] if len(params) > 0 else dim * 2
                dim_out = int(min(dim_out, max_dim))
                enc, dec = M.SpatialDownsample2x(dim, dim_out), M.SpatialUpsample2x(dim_out, dim)
                assert fmap > 1
                fmap //= 2
                stages.append(Stage(kind, dim, dim_out))
            elif kind == "compress_time":
                dim_out = params[0] if len(params) > 0 else dim * 2
                dim_out = int(min(dim_out, max_dim))
                enc, dec = M.TimeDownsample2x(dim, dim_out), M.TimeUpsample2x(dim_out, dim)
                tdf *= 2
                stages.append(Stage(kind, dim, dim_out))
            elif kind == "attend_space":
                def mk():
                    return nn.Sequential(M.Residual(M.Attention(dim, attn_dim_head, attn_heads, causal=False)),
                                         M.Residual(M.FeedForward(dim)))
                enc, dec = mk(), mk()
                stages.append(Stage(kind, dim, dim))
            elif kind == "linear_attend_space":
                def mk():
                    return nn.Sequential(M.Residual(M.LinearSpaceAttention(dim, linear_attn_dim_head, linear_attn_heads)),
                                         M.Residual(M.FeedForward(dim)))
                enc, dec = mk(), mk()
                stages.append(Stage(kind, dim, dim))
            elif kind == "gateloop_time":                                        # M:1216-1222
                enc = M.ToTimeSequence(M.Residual(M.SimpleGateLoopLayer(dim)))
                dec = M.ToTimeSequence(M.Residual(M.SimpleGateLoopLayer(dim)))
                stages.append(Stage(kind, dim, dim))
            elif kind == "attend_time":
                def mk():
                    return nn.Sequential(
                        M.Residual(M.TokenShift(M.Attention(dim, attn_dim_head, attn_heads, causal=True))),
                        M.Residual(M.TokenShift(M.FeedForward(dim))))
                enc, dec = mk(), mk()
                stages.append(Stage(kind, dim, dim))
            else:
                raise ValueError(f"unknown layer type {kind}")          # M:1311-1312
            self.encoder_layers.append(enc)
            self.decoder_layers.insert(0, dec)
            dim = dim_out

        # final LayerNorm: constructed and present in state_dict but never executed by the reference
        # (zip truncation at M:1565) -- kept for checkpoint compatibility only.
        self.encoder_layers.append(nn.Sequential(M.Marker("to channels-last"), nn.LayerNorm(dim), M.Marker("to channels-first")))

        self.stages = stages
        self.time_downsample_factor = tdf
        self.time_padding = tdf - 1
        self.fmap_size = fmap
        self.has_cond = has_cond
        self.has_cond_across_layers = [st.kind == "cond_residual" for st in stages]
        self.dim_cond = dim_cond
        self.encoder_cond_in = nn.Identity()
        self.decoder_cond_in = nn.Identity()
        if has_cond:                                                              # M:1340-1352: Linear + SiLU stems
            dc = int(dim_cond * dim_cond_expansion_factor)
            self.encoder_cond_in = nn.Sequential(nn.Linear(dim_cond, dc), M.Marker("SiLU"))
            self.decoder_cond_in = nn.Sequential(nn.Linear(dim_cond, dc), M.Marker("SiLU"))

        # ---- quantiser (M:1356-1384) ----
        self.use_fsq = use_fsq
        if not use_fsq:
            assert codebook_size is not None and fsq_levels is None, \
                "if use_fsq is set to False, `codebook_size` must be set (and not `fsq_levels`)"
            self.quantizers = M.LFQ(dim, codebook_size, lfq_entropy_loss_weight, lfq_commitment_loss_weight,
                                    lfq_diversity_gamma, lfq_soft_clamp_input_value, num_codebooks, lfq_spherical)   # M:1364-1373
        else:
            assert codebook_size is None and fsq_levels is not None, \
                "if use_fsq is set to True, `fsq_levels` must be set (and not `codebook_size`)"
            self.quantizers = M.FSQ(fsq_levels, dim, num_codebooks)                        # M:1378-1382
        self.quantizer_aux_loss_weight = quantizer_aux_loss_weight
        self.register_buffer("zero", torch.tensor(0.), persistent=False)

        # training-only branches of the reference are not built
        self.vgg = None
        self.use_vgg = False
        self.perceptual_loss_weight = perceptual_loss_weight
        self.use_gan = use_gan
        self.has_gan = False
        self.has_multiscale_gan = False
        self.has_multiscale_discrs = False
        self.multiscale_discrs = nn.ModuleList([])
        self.adversarial_loss_weight = adversarial_loss_weight
        self.grad_penalty_loss_weight = grad_penalty_loss_weight
        self.multiscale_adversarial_loss_weight = multiscale_adversarial_loss_weight

        self.quantizer_loss_breakdown = None     # set by a train-mode forward (LFQ): (per_sample_entropy, batch_entropy, commitment)
        self.quantizer_aux_loss = None
        self._engine: Optional[Engine] = None
        # opt-in: replay each (entry point, input shape) as one CUDA graph after a warm-up call -- the forward path is
        # a static launch plan (~170 kernels), so this removes the per-launch host overhead.  Outputs are cloned out
        # of the graph's static buffers.
        self.cuda_graphs = False
        self._graphs = {}
        # graph-cache lane: calls issued under different lanes (host_io.StreamLanes: one CUDA stream per lane) replay
        # separate graph instances with their own static buffers and memory pools, so they may overlap on the device
        self._lane = 0
        # opt-in: programmatic dependent launch for every kernel of the path (mv2_set_pdl); process-wide library state
        self.pdl = False

    # ------------------------------------------------------------------ module plumbing
    @property
    def device(self):
        return self.zero.device

    def parameters(self, recurse: bool = True):
        # list, as the reference returns (M:1460-1471)
        return [*self.conv_in.parameters(), *self.conv_in_first_frame.parameters(), *self.conv_out_first_frame.parameters(),
                *self.conv_out.parameters(), *self.encoder_layers.parameters(),
                *self.decoder_layers.parameters(), *self.encoder_cond_in.parameters(), *self.decoder_cond_in.parameters(),
                *self.quantizers.parameters()]

    def discr_parameters(self):
        return []

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        # reference checkpoints carry discriminator weights (always constructed, M:1422); they are
        # not part of the inference path and are dropped here.
        sd = {k: v for k, v in state_dict.items() if not (k.startswith("discr.") or k.startswith("multiscale_discrs."))}
        return super().load_state_dict(sd, strict=strict, **kw)

    def __deepcopy__(self, memo):
        # the engine (packed weights, ctypes handles) and captured CUDA graphs (static buffers, private pools) belong
        # to THIS instance: a copy starts without them and re-packs / re-captures on first use
        eng, self._engine = self._engine, None
        graphs, self._graphs = self._graphs, {}
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                new.__dict__[k] = copy.deepcopy(v, memo)
        finally:
            self._engine = eng
            self._graphs = graphs
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_engine"] = None
        st["_graphs"] = {}
        return st

    def copy_for_eval(self):
        dev = self.device
        c = copy.deepcopy(self.cpu())
        self.to(dev)
        c.eval()
        return c.to(dev)

    @classmethod
    def init_and_load_from(cls, path, strict=True):
        path = Path(path)
        assert path.exists()
        pkg = torch.load(str(path), map_location="cpu", weights_only=False)
        assert "config" in pkg, "model configs were not found in this saved checkpoint"
        config = pickle.loads(pkg["config"])
        # reference checkpoints pickle module-valued kwargs we do not build
        for k in ("vgg", "lfq_activation"):
            config[k] = None
        config["multiscale_discrs"] = tuple()
        tok = cls(**config)
        tok.load(path, strict=strict)
        return tok

    def save(self, path, overwrite=True):
        path = Path(path)
        assert overwrite or not path.exists(), f"{str(path)} already exists"
        torch.save(dict(model_state_dict=self.state_dict(), version=__version__, config=self._configs), str(path))

    def load(self, path, strict=True):
        path = Path(path)
        assert path.exists()
        pkg = torch.load(str(path), map_location="cpu", weights_only=False)
        sd = pkg.get("model_state_dict")
        assert sd is not None
        self.load_state_dict(sd, strict=strict)

    # ------------------------------------------------------------------ engine access
    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self)
        self._engine.prepare()
        self._engine.lib.mv2_set_pdl(1 if self.pdl else 0)
        return self._engine

    def _graph_call(self, name, fn, *tensors):
        """fn(*tensors) -> tensor | tuple of tensors, replayed through a cached CUDA graph when enabled."""
        if not self.cuda_graphs:
            return fn(*tensors)
        eng = self.engine
        key = (name, eng._sig_id, tuple((tuple(t.shape), t.dtype) for t in tensors), self._lane)
        if any(k[1] != eng._sig_id for k in self._graphs):      # parameters were re-packed: old graphs read stale weights
            self._graphs = {k: v for k, v in self._graphs.items() if k[1] == eng._sig_id}
        ent = self._graphs.get(key)
        if ent is None:                       # first call: plain run (warms up lazy init: attributes, entry points)
            self._graphs[key] = "warm"
            return fn(*tensors)
        if ent == "warm":                     # second call: capture
            static_in = [torch.empty_like(t) for t in tensors]
            for s_, t in zip(static_in, tensors):
                s_.copy_(t)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            l0 = eng.launches
            with torch.cuda.graph(g):
                out = fn(*static_in)
            ent = (g, static_in, out, eng.launches - l0)
            self._graphs[key] = ent
            g.replay()
        else:
            g, static_in, out, n_launch = ent
            for s_, t in zip(static_in, tensors):
                s_.copy_(t, non_blocking=True)
            g.replay()
            eng.launches += n_launch
        out = ent[2]
        if isinstance(out, tuple):
            return tuple(o.clone() for o in out)
        return out.clone()

    def _check_video(self, v, video_contains_first_frame=True):
        assert v.ndim in {4, 5}                                                   # M:1675
        assert tuple(v.shape[-2:]) == (self.image_size, self.image_size)          # M:1677
        if v.ndim == 4:                                                           # M:1681-1685
            v = v[:, :, None]
            video_contains_first_frame = True
        frames = v.shape[2]
        ff = int(bool(video_contains_first_frame))
        assert (frames - ff) % self.time_downsample_factor == 0, \
            f"number of frames {frames} minus the first frame ({frames - ff}) must be divisible by the total " \
            f"downsample factor across time {self.time_downsample_factor}"      # M:1691
        assert v.shape[1] == self.channels
        if v.device != self.device:
            raise RuntimeError(f"input is on {v.device} but the tokenizer is on {self.device}")
        return v, bool(ff)

    # ------------------------------------------------------------------ reference API
    @torch.no_grad()
    @_on_model_device
    def encode(self, video, quantize=False, cond=None, video_contains_first_frame=True):
        """M:1523-1576.  Returns (B, C, T', H', W') like the reference."""
        video, ff = self._check_video(video, video_contains_first_frame)
        cond = self._check_cond(cond, video.shape[0])
        eng = self.engine
        x = eng.encode_cl(video, ff, cond)
        if quantize:
            q, idx, _ = eng.quantize_cl(x)
            out = eng.to_channels_first(q)
            if self.use_fsq:
                return out, idx
            return out, idx, self.zero
        return eng.to_channels_first(x)

    def _check_cond(self, cond, batch):
        """M:1542-1545 / M:1610-1613."""
        assert (not self.has_cond) or cond is not None, \
            "`cond` must be passed into tokenizer forward method since conditionable layers were specified"
        if cond is None:
            return None
        if not self.has_cond:
            return None                      # the reference runs the (Identity) stem and never uses the result
        assert tuple(cond.shape) == (batch, self.dim_cond)
        self._check_on_device(cond, "cond")
        return cond

    def _check_on_device(self, t, what):
        if t.device != self.device:
            raise RuntimeError(f"{what} is on {t.device} but the tokenizer is on {self.device}")

    @torch.no_grad()
    @_on_model_device
    def decode(self, quantized, cond=None, video_contains_first_frame=True):
        """M:1598-1649.  quantized: (B, C, T', H', W')."""
        assert quantized.ndim == 5 and quantized.shape[1] == self.quantizers.dim, \
            f"quantized must be (B, {self.quantizers.dim}, T, H, W), got {tuple(quantized.shape)}"
        self._check_on_device(quantized, "quantized")
        cond = self._check_cond(cond, quantized.shape[0])
        eng = self.engine
        return eng.decode_cl(eng.to_channels_last(quantized), bool(video_contains_first_frame), cond)

    @torch.no_grad()
    @_on_model_device
    def decode_from_code_indices(self, codes, cond=None, video_contains_first_frame=True):
        """M:1579-1595."""
        assert codes.dtype in (torch.long, torch.int32)                           # M:1585
        if codes.ndim == 2:                                                       # M:1587-1591
            n = codes.shape[-1]
            assert n % (self.fmap_size ** 2) == 0, \
                f"flattened video ids must have a length ({n}) that is divisible by the fmap size " \
                f"({self.fmap_size}) squared ({self.fmap_size ** 2})"
            codes = codes.reshape(codes.shape[0], -1, self.fmap_size, self.fmap_size)
        nc = self.quantizers.num_codebooks
        assert codes.ndim == (4 if nc == 1 else 5) and (nc == 1 or codes.shape[-1] == nc), \
            f"codes must be (B, T, H, W{'' if nc == 1 else ', num_codebooks'}) or flat (B, N), got {tuple(codes.shape)}"
        self._check_on_device(codes, "codes")
        eng = self.engine
        ff = bool(video_contains_first_frame)
        cond = self._check_cond(cond, codes.shape[0])
        if cond is not None:
            return self._graph_call("decode_codes_cond" + ("" if ff else "_noff"),
                                    lambda c, cd: eng.decode_cl(eng.codes_to_quantized_cl(c), ff, cd), codes.contiguous(), cond.contiguous())
        return self._graph_call("decode_codes" if ff else "decode_codes_noff",
                                lambda c: eng.decode_cl(eng.codes_to_quantized_cl(c), ff), codes.contiguous())

    @torch.no_grad()
    @_on_model_device
    def lfq_loss_breakdown(self, video, group=None):
        """Training-mode LFQ auxiliary terms the reference computes at M:1705 (``quantizer_loss_breakdown``):
        returns ``(codes, (per_sample_entropy, batch_entropy, commitment), aux_loss)``.  ``batch_entropy`` uses the
        cross-rank mean code probability -- the single 4 KiB all-reduce of the path (dist.LfqBatchEntropy), issued
        on a side stream.  (The losses' backward and the GAN/perceptual terms are out of scope)"""
        from .dist import LfqBatchEntropy
        assert not self.use_fsq, "FSQ has no auxiliary loss (reference M:1702)"
        video, _ = self._check_video(video)
        eng = self.engine
        x = eng.encode_cl(video)
        _, codes, pre = eng.quantize_cl(x, want_quantized=False, want_aux=True)
        q = self.quantizers
        be = LfqBatchEntropy(eng, num_codebooks=q.num_codebooks)
        be.start(pre, group)
        ps, bent, commit, aux = be.finish(q.diversity_gamma, q.entropy_loss_weight, q.commitment_loss_weight, group)
        return codes, (ps, bent, commit), aux

    def _forward_train_mode(self, eng, video, need_recon, ff=True, group=None):
        """``model.train()`` forward of the LFQ tokenizer (reference M:1705: ``self.quantizers(x, return_loss_breakdown=True)``
        in training mode): besides codes / reconstruction the quantiser's auxiliary terms are computed -- per-sample entropy,
        batch (codebook) entropy of the CROSS-RANK mean code probability, commitment -- and kept in
        ``self.quantizer_loss_breakdown`` = (per_sample_entropy, batch_entropy, commitment) / ``self.quantizer_aux_loss``.
        The batch-entropy term needs the one collective of the path: a 4 KiB SUM all-reduce of avg_prob (A.1 step 7), issued
        on a side stream so that it overlaps the decoder.  The decoder is fed the quantised value q itself; the reference's
        straight-through ``x + (q - x).detach()`` equals q up to one rounding (SURVEY 8d cfg 3).  No autograd (N2)."""
        from .dist import LfqBatchEntropy
        qz = self.quantizers

        def enc(v):
            x = eng.encode_cl(v, ff)
            q, codes_, pre = eng.quantize_cl(x, want_quantized=need_recon, want_aux=True)
            return (codes_, pre, q) if need_recon else (codes_, pre)

        sfx = "" if ff else "_noff"
        res = self._graph_call(("train_enc_q" if need_recon else "train_enc") + sfx, enc, video)
        codes, pre = res[0], res[1]
        be = LfqBatchEntropy(eng, num_codebooks=qz.num_codebooks)
        be.start(pre, group)                                   # partial sums + all-reduce on the side stream ...
        recon = self._graph_call("train_dec" + sfx, lambda t: eng.decode_cl(t, ff), res[2]) if need_recon else None   # ... under the decoder
        ps, bent, commit, aux = be.finish(qz.diversity_gamma, qz.entropy_loss_weight, qz.commitment_loss_weight, group)
        self.quantizer_loss_breakdown = (ps, bent, commit)
        self.quantizer_aux_loss = aux
        return (codes, recon) if need_recon else codes

    @torch.no_grad()
    def tokenize(self, video):
        """M:1651-1654."""
        self.eval()
        return self.forward(video, return_codes=True)

    @_on_model_device
    def forward(self, video_or_images, cond=None, return_loss=False, return_codes=False, return_recon=False,
                return_discr_loss=False, return_recon_loss_only=False, apply_gradient_penalty=True,
                video_contains_first_frame=True, adversarial_loss_weight=None,
                multiscale_adversarial_loss_weight=None):
        """The reference forward (M:1657-1896): inference returns (codes / reconstruction), ``return_recon_loss_only`` and --
        for models without the GAN / perceptual branches (``use_gan=False, perceptual_loss_weight=0``) -- ``return_loss``:
        ``(total_loss, LossBreakdown)`` with ``total_loss = recon_loss + aux_loss * quantizer_aux_loss_weight`` (M:1868-1896)."""
        assert (return_loss + return_codes + return_discr_loss) <= 1               # M:1674
        if return_discr_loss:
            raise NotImplementedError("the GAN discriminator losses (reference M:1728-1786) are outside the accelerated path "
                                      "(SURVEY.md 8f N2)")
        if return_loss and self._needs_gan_or_vgg():
            raise NotImplementedError(
                "return_loss with the GAN / perceptual / adaptive-weighting terms (reference M:1788-1866) is outside the "
                "accelerated path (SURVEY.md 8f N2): construct with use_gan=False, perceptual_loss_weight=0.")
        video, ff = self._check_video(video_or_images, video_contains_first_frame)
        cond = self._check_cond(cond, video.shape[0])
        if return_loss and self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # the trainer's generator step (T:356-363): loss with a grad_fn.  Forward = the same kernels; backward = train.py
            from .train import train_forward
            recon, aux, _, qlb = train_forward(self, video.contiguous(), ff, cond)
            target = video.float() / 255. if video.dtype == torch.uint8 else video
            recon_loss = torch.nn.functional.mse_loss(target.to(recon.dtype), recon)          # M:1722
            self.quantizer_loss_breakdown, self.quantizer_aux_loss = qlb, aux.detach()
            aux_losses = aux.to(recon_loss.dtype)
            total_loss = recon_loss + aux_losses * self.quantizer_aux_loss_weight                # M:1868-1871
            return total_loss, LossBreakdown(recon_loss, aux_losses, qlb, self.zero, self.zero, 0., [], [])
        with torch.no_grad():
            eng = self.engine
            need_recon = return_recon or return_recon_loss_only or return_loss or not return_codes
            aux = None

            def run(v, cd=None):
                x = eng.encode_cl(v, ff, cd)
                q, codes_, _ = eng.quantize_cl(x, want_quantized=need_recon)
                if not need_recon:
                    return codes_
                return codes_, eng.decode_cl(q, ff, cd)

            if cond is not None:
                out = self._graph_call(("fwd_recon_cond" if need_recon else "fwd_codes_cond") + ("" if ff else "_noff"), run,
                                       video.contiguous(), cond.contiguous())
            elif self.training and not self.use_fsq:
                out = self._forward_train_mode(eng, video.contiguous(), need_recon, ff)
                aux = self.quantizer_aux_loss
            else:
                out = self._graph_call(("fwd_recon" if need_recon else "fwd_codes") + ("" if ff else "_noff"), run, video.contiguous())
            if return_codes and not return_recon:
                return out                                                         # M:1707-1708
            codes, recon = out
            if return_codes:
                return codes, recon                                                # M:1714-1715
            if not (return_loss or return_recon_loss_only):
                return recon                                                       # M:1719-1720
            recon_loss = eng.mse(video, recon).to(self.dtype)                      # M:1722
            if return_recon_loss_only:                                             # M:1726-1727
                return recon_loss, recon
            # M:1868-1896 with perceptual_loss = gen_loss = zero, adaptive_weight = 0., no multiscale discriminators
            zero = self.zero
            aux_losses = zero if aux is None else aux.to(recon_loss.dtype)         # eval mode / FSQ: M:1700-1703
            total_loss = recon_loss + aux_losses * self.quantizer_aux_loss_weight
            qlb = None if (self.use_fsq or aux is None) else self.quantizer_loss_breakdown
            return total_loss, LossBreakdown(recon_loss, aux_losses, qlb, zero, zero, 0., [], [])

    def _needs_gan_or_vgg(self) -> bool:
        """True when the reference constructor would have built a VGG (M:1392) or discriminators (M:1427, M:1435)."""
        return bool((self.channels in {1, 3, 4} and self.perceptual_loss_weight > 0.)
                    or (self.use_gan and self.adversarial_loss_weight > 0.) or self.has_multiscale_discrs)

    @property
    def dtype(self):
        return self.conv_in.conv.weight.dtype
