"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of ``gateloop_transformer.SimpleGateLoopLayer`` (un-vendored dependency
``gateloop-transformer>=0.2.2``, reference setup.py:28; imported at magvit2_pytorch.py:36, constructed at M:1220-1221 as
``ToTimeSequence(Residual(SimpleGateLoopLayer(dim = dim)))`` -- note the reference builds ``gateloop_kwargs`` (M:1216-1218)
but never passes it, so the layer runs with its defaults).  Restated from the published algorithm (simplified gate loop,
defaults prenorm=True, post_ln=False, reverse=False):

    x  = RMSNorm(x)                      F.normalize(x, -1) * sqrt(dim) * gamma
    q, kv, a = chunk(Linear(dim, 3 dim, bias=False)(x), 3)       'b n (qkva d) -> qkva (b d) n 1'
    a  = sigmoid(a)
    s_t = a_t * s_{t-1} + kv_t           (associative scan over the sequence, per channel, s_{-1} = 0)
    out_t = q_t * s_t

"parity unpinned": the upstream source is not available in this container.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class RMSNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return F.normalize(x, dim=-1) * self.scale * self.gamma


def gate_loop_scan(q, kv, a):
    """q, kv, a: (..., n, d) with a already in (0, 1).  Sequential form of the associative scan."""
    s = torch.zeros_like(kv[..., 0, :])
    outs = []
    for t in range(kv.shape[-2]):
        s = a[..., t, :] * s + kv[..., t, :]
        outs.append(q[..., t, :] * s)
    return torch.stack(outs, dim=-2)


class SimpleGateLoopLayer(nn.Module):
    def __init__(self, dim, prenorm=True, use_heinsen=False, use_jax_associative_scan=False, post_ln=False, reverse=False):
        super().__init__()
        assert not post_ln and not reverse, "only the defaults the reference uses are restated"
        self.dim = dim
        self.norm = RMSNorm(dim) if prenorm else None
        self.to_qkva = nn.Sequential(nn.Linear(dim, dim * 3, bias=False), nn.Identity())

    def forward(self, x, cache=None, return_cache=False):
        assert cache is None and not return_cache
        if self.norm is not None:
            x = self.norm(x)
        q, kv, a = self.to_qkva(x).chunk(3, dim=-1)
        return gate_loop_scan(q, kv, a.sigmoid())
