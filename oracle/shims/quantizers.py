"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the two quantizers the reference pulls from the un-vendored
PyPI dependency ``vector-quantize-pytorch>=1.14.39`` (reference setup.py:34,
import at magvit2_pytorch.py:21, constructed at magvit2_pytorch.py:1364-1382).
The dependency's source is NOT under /root/reference and is not installable
here (no network), so this file restates its *published algorithm* from
SURVEY.md Appendix A.1/A.2:  "parity unpinned" for this arithmetic -- there is
no upstream golden vector to check it against.  The reference's own call sites
(M:1576, M:1593, M:1700, M:1705) and its README round-trip (README.md:85-90)
anchor the behaviour.

Module/parameter names (project_in, project_out, mask) follow the upstream
package so that ``state_dict`` keys of the loaded reference model are the ones a
real checkpoint has.
"""
from __future__ import annotations

from collections import namedtuple
from math import log2

import torch
import torch.nn.functional as F
from torch import nn

LFQReturn = namedtuple("Return", ["quantized", "indices", "entropy_aux_loss"])
LFQLossBreakdown = namedtuple("LossBreakdown", ["per_sample_entropy", "batch_entropy", "commitment"])


def _entropy(prob, eps=1e-5):
    return (-prob * torch.log(prob.clamp(min=eps))).sum(dim=-1)


def _maybe_distributed_mean(t):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    t = t.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t / dist.get_world_size()


class LFQ(nn.Module):
    """Lookup-free quantiser (Appendix A.1)."""

    def __init__(self, *, dim=None, codebook_size=None, entropy_loss_weight=0.1,
                 commitment_loss_weight=0.25, diversity_gamma=1., num_codebooks=1,
                 soft_clamp_input_value=None, spherical=False, codebook_scale=1.,
                 inv_temperature=100.):
        super().__init__()
        assert codebook_size is not None and log2(codebook_size).is_integer()
        self.codebook_size = codebook_size
        self.codebook_dim = int(log2(codebook_size))
        self.num_codebooks = num_codebooks
        cdims = self.codebook_dim * num_codebooks
        dim = dim if dim is not None else cdims
        self.dim = dim
        has_proj = dim != cdims
        self.project_in = nn.Linear(dim, cdims) if has_proj else nn.Identity()
        self.project_out = nn.Linear(cdims, dim) if has_proj else nn.Identity()
        self.has_projections = has_proj
        self.entropy_loss_weight = entropy_loss_weight
        self.commitment_loss_weight = commitment_loss_weight
        self.diversity_gamma = diversity_gamma
        self.soft_clamp_input_value = soft_clamp_input_value
        self.spherical = spherical
        self.codebook_scale = codebook_scale
        self.inv_temperature = inv_temperature
        self.register_buffer("mask", 2 ** torch.arange(self.codebook_dim - 1, -1, -1))
        self.register_buffer("zero", torch.tensor(0.), persistent=False)
        all_codes = torch.arange(codebook_size)
        bits = ((all_codes[..., None].int() & self.mask) != 0).float()
        self.register_buffer("codebook", bits * codebook_scale * 2 - codebook_scale, persistent=False)

    @property
    def dtype(self):
        return self.codebook.dtype

    def indices_to_codes(self, indices, project_out=True):
        is_img_or_video = indices.ndim >= (3 + int(self.num_codebooks > 1))
        if self.num_codebooks == 1:       # keep_num_codebooks_dim = num_codebooks > 1: otherwise add the codebook axis back
            indices = indices[..., None]
        bits = ((indices[..., None].int() & self.mask) != 0).to(self.dtype)
        codes = bits * self.codebook_scale * 2 - self.codebook_scale
        codes = codes.reshape(*codes.shape[:-2], -1)
        if project_out:
            codes = self.project_out(codes)
        if is_img_or_video:
            codes = codes.movedim(-1, 1)
        return codes

    def forward(self, x, return_loss_breakdown=False, mask=None):
        is_img_or_video = x.ndim >= 4
        if is_img_or_video:
            x = x.movedim(1, -1)
            lead = x.shape[1:-1]
            x = x.reshape(x.shape[0], -1, x.shape[-1])
        x = self.project_in(x)
        if self.soft_clamp_input_value is not None:
            cv = self.soft_clamp_input_value
            x = (x / cv).tanh() * cv
        b, n = x.shape[:2]
        x = x.reshape(b, n, self.num_codebooks, self.codebook_dim)
        if self.spherical:
            x = F.normalize(x, dim=-1)
        orig_dtype = x.dtype
        x = x.float()
        original_input = x
        cbv = torch.ones_like(x) * self.codebook_scale
        quantized = torch.where(x > 0, cbv, -cbv)
        indices = ((quantized > 0).int() * self.mask.int()).sum(dim=-1)
        if self.training:
            x = x + (quantized - x).detach()
        else:
            x = quantized
        if self.training:
            distance = -2 * torch.einsum("...id,jd->...ij", original_input, self.codebook.float())
            prob = (-distance * self.inv_temperature).softmax(dim=-1)
            per_sample_probs = prob.reshape(b * n, self.num_codebooks, -1)
            per_sample_entropy = _entropy(per_sample_probs).mean()
            avg_prob = per_sample_probs.mean(dim=0)
            avg_prob = _maybe_distributed_mean(avg_prob)
            codebook_entropy = _entropy(avg_prob).mean()
            entropy_aux_loss = per_sample_entropy - self.diversity_gamma * codebook_entropy
        else:
            entropy_aux_loss = per_sample_entropy = codebook_entropy = self.zero
        if self.training and self.commitment_loss_weight > 0.:
            commit_loss = F.mse_loss(original_input, quantized.detach(), reduction="none").mean()
        else:
            commit_loss = self.zero
        x = x.to(orig_dtype)
        x = x.reshape(b, n, -1)
        x = self.project_out(x)
        if is_img_or_video:
            x = x.reshape(b, *lead, x.shape[-1]).movedim(-1, 1)
            indices = indices.reshape(b, *lead, self.num_codebooks)
        if self.num_codebooks == 1:
            indices = indices[..., 0]
        aux_loss = entropy_aux_loss * self.entropy_loss_weight + commit_loss * self.commitment_loss_weight
        ret = LFQReturn(x, indices, aux_loss)
        if not return_loss_breakdown:
            return ret
        return ret, LFQLossBreakdown(per_sample_entropy, codebook_entropy, commit_loss)


class FSQ(nn.Module):
    """Finite scalar quantiser (Appendix A.2)."""

    def __init__(self, levels, dim=None, num_codebooks=1):
        super().__init__()
        _levels = torch.tensor(levels, dtype=torch.int32)
        self.register_buffer("_levels", _levels, persistent=False)
        _basis = torch.cumprod(torch.tensor([1] + list(levels[:-1])), dim=0, dtype=torch.int32)
        self.register_buffer("_basis", _basis, persistent=False)
        self.codebook_dim = len(levels)
        self.num_codebooks = num_codebooks
        eff = self.codebook_dim * num_codebooks
        self.dim = dim if dim is not None else eff
        has_proj = self.dim != eff
        self.project_in = nn.Linear(self.dim, eff) if has_proj else nn.Identity()
        self.project_out = nn.Linear(eff, self.dim) if has_proj else nn.Identity()
        self.codebook_size = int(_levels.prod().item())

    def bound(self, z, eps=1e-3):
        half_l = (self._levels - 1) * (1 + eps) / 2
        offset = torch.where(self._levels % 2 == 0, 0.5, 0.0)
        shift = (offset / half_l).atanh()
        return (z + shift).tanh() * half_l - offset

    def quantize(self, z):
        b = self.bound(z)
        q = b + (b.round() - b).detach()        # round_ste (A.2: "straight-through"): the value is round(b) exactly
        half_width = self._levels // 2
        return q / half_width

    def codes_to_indices(self, zhat):
        half_width = self._levels // 2
        zhat = zhat * half_width + half_width
        return (zhat * self._basis).sum(dim=-1).to(torch.int32)

    def indices_to_codes(self, indices):
        is_img_or_video = indices.ndim >= (3 + int(self.num_codebooks > 1))
        if self.num_codebooks == 1:
            indices = indices[..., None]
        ind = indices[..., None]
        nonneg = (ind // self._basis) % self._levels
        half_width = self._levels // 2
        codes = (nonneg - half_width) / half_width
        codes = codes.reshape(*codes.shape[:-2], -1)
        codes = self.project_out(codes.to(self.project_out.weight.dtype) if isinstance(self.project_out, nn.Linear) else codes)
        if is_img_or_video:
            codes = codes.movedim(-1, 1)
        return codes

    def forward(self, z):
        is_img_or_video = z.ndim >= 4
        if is_img_or_video:
            z = z.movedim(1, -1)
            lead = z.shape[1:-1]
            z = z.reshape(z.shape[0], -1, z.shape[-1])
        z = self.project_in(z)
        b, n = z.shape[:2]
        z = z.reshape(b, n, self.num_codebooks, self.codebook_dim)
        orig_dtype = z.dtype
        if z.dtype not in (torch.float32, torch.float64):
            z = z.float()
        codes = self.quantize(z)
        indices = self.codes_to_indices(codes)
        codes = codes.reshape(b, n, -1).to(orig_dtype)
        out = self.project_out(codes)
        if is_img_or_video:
            out = out.reshape(b, *lead, out.shape[-1]).movedim(-1, 1)
            indices = indices.reshape(b, *lead, self.num_codebooks)
        if self.num_codebooks == 1:
            indices = indices[..., 0]
        return out, indices
