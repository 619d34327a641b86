"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of ``taylor_series_linear_attention.TaylorSeriesLinearAttn``
(un-vendored dependency ``taylor-series-linear-attention>=0.1.5``, reference
setup.py:35; imported at magvit2_pytorch.py:34, constructed at M:415-419,
called at M:430).  Restated from the published algorithm (SURVEY.md Appendix
A.3): non-causal, second-order Taylor feature map, no prenorm / gating.
"parity unpinned": the upstream source is not available in this container.
"""
from __future__ import annotations

import torch
from torch import nn


def second_taylor_expansion(x):
    # phi(x) = [1, x, (x outer x) / sqrt(2)]
    lead = x.shape[:-1]
    d = x.shape[-1]
    x0 = x.new_ones((*lead, 1))
    x2 = (x[..., :, None] * x[..., None, :]) * (0.5 ** 0.5)
    return torch.cat((x0, x, x2.reshape(*lead, d * d)), dim=-1)


class _Split(nn.Module):
    def __init__(self, parts, heads):
        super().__init__()
        self.parts, self.heads = parts, heads

    def forward(self, t):
        b, n, _ = t.shape
        t = t.reshape(b, n, self.parts, self.heads, -1).permute(2, 0, 3, 1, 4)
        return t[0] if self.parts == 1 else t


class TaylorSeriesLinearAttn(nn.Module):
    def __init__(self, dim, *, dim_head=16, heads=8, dropout=0.):
        super().__init__()
        self.scale = dim_head ** -0.5
        inner = dim_head * heads
        self.heads = heads
        self.to_q = nn.Sequential(nn.Linear(dim, inner, bias=False), _Split(1, heads))
        self.to_kv = nn.Sequential(nn.Linear(dim, inner * 2, bias=False), _Split(2, heads))
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.Dropout(dropout))

    def forward(self, x, eps=1e-5):
        q = self.to_q(x)
        k, v = self.to_kv(x)
        q = q * self.scale
        q, k = second_taylor_expansion(q), second_taylor_expansion(k)
        kv = torch.einsum("bhnd,bhne->bhde", k, v)
        k_sum = k.sum(dim=-2)
        num = torch.einsum("bhnd,bhde->bhne", q, kv)
        den = torch.einsum("bhnd,bhd->bhn", q, k_sum)[..., None]
        out = num / den.clamp(min=eps)
        b, h, n, d = out.shape
        out = out.permute(0, 2, 1, 3).reshape(b, n, h * d)
        return self.to_out(out)
