"""Deterministic synthetic weights (keyed by ``state_dict`` name + shape) and videos for tests and bench.py.
Neutral data-synthesis helper: it imports neither the product package nor ``oracle/``.

The reference ships no checkpoints and the README-size generator has 117.8 M
parameters (470 MB fp32), far too large to commit.  Instead every test / bench
fills a model's generator tensors from a per-key seeded generator, so the
reference (in this container), the restated oracle and the CUDA product (on
the GPU box) all see bit-identical weights without shipping them.

The fill also de-degenerates the reference's default init:
``SqueezeExcite`` zero-inits its last conv with bias -10 (reference
magvit2_pytorch.py:218-219) which makes every ResidualUnit ~identity and would
hide conv errors; the up-samplers use a repeated-kernel init (M:829-836,
M:866-873).  Here every weight is i.i.d. normal with fan-in scaling.
"""
from __future__ import annotations

import hashlib

import torch


def _seed_for(key: str, seed: int) -> int:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF


def synth_tensor(key: str, shape, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(_seed_for(key, seed))
    shape = tuple(shape)
    t = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "gamma":
        t = 1.0 + 0.1 * t
    elif leaf == "mem_kv":
        pass  # reference init is randn (M:357)
    elif leaf == "bias":
        t = 0.05 * t
    elif leaf == "weight":
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = t * (fan_in ** -0.5)
        else:  # 1-d weight (the dead LayerNorm)
            t = 1.0 + 0.1 * t
    return t.to(dtype)


def is_generator_key(key: str) -> bool:
    return not (key.startswith("discr.") or key.startswith("multiscale_discrs.") or key.startswith("vgg."))


@torch.no_grad()
def fill_state_dict_(module, seed: int = 0):
    """Overwrite every floating-point generator tensor of ``module`` in place."""
    sd = module.state_dict()
    for k, v in sd.items():
        if not is_generator_key(k) or not v.is_floating_point():
            continue
        v.copy_(synth_tensor(k, v.shape, seed).to(v.dtype))
    return module


def synth_video(batch, channels, frames, size, seed=1234):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(batch, channels, frames, size, size, generator=g)
