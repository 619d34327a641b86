"""The torch restatements train.py differentiates in the backward (channels-last, driven by the parameter containers) against the
oracle's restatements of the same reference blocks (channels-first, driven by a state_dict) -- CPU only.  The gradients through
them are checked end to end on the GPU (tests/test_train_gpu.py); this file keeps the forward definitions pinned in the CPU suite."""
import pytest
import torch

from magvit2_pytorch_b200 import modules as M
from magvit2_pytorch_b200 import train as T
from oracle import restated as R


def _cl(x):   # (B,C,T,H,W) -> (B,T,H,W,C)
    return x.permute(0, 2, 3, 4, 1).contiguous()


def _cf(x):
    return x.permute(0, 4, 1, 2, 3).contiguous()


def _sd(mod, prefix=""):
    return {prefix + k: v.detach() for k, v in mod.state_dict().items()}


def _randomise(mod, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in mod.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.ndim > 1 else 0.1))
    return mod


@pytest.fixture
def x():
    return torch.randn(2, 16, 5, 6, 6, generator=torch.Generator().manual_seed(0))


def test_squeeze_excite(x):
    se = _randomise(M.SqueezeExcite(16), 1)
    want = R.squeeze_excite(x, _sd(se, "p."), "p.")
    assert torch.allclose(_cf(T._squeeze_excite(_cl(x), se)), want, atol=1e-5)


@pytest.mark.parametrize("axis", ["space", "time"])
def test_attention_block(x, axis):
    at = _randomise(M.Attention(16, dim_head=8, heads=2, causal=(axis == "time")), 2)
    sd = _sd(at, "p.")
    if axis == "space":
        want = R.space_attention(x, sd, "p.", 2) + x
    else:
        want = R.time_attention(R.token_shift(x), sd, "p.", 2) + x
    assert torch.allclose(_cf(T._attention_block(_cl(x), at, axis)), want, atol=1e-5)


def test_linear_attention_block(x):
    la = _randomise(M.LinearSpaceAttention(16, dim_head=4, heads=3), 3)
    want = R.linear_space_attention(x, _sd(la, "p."), "p.", 3, 4) + x
    assert torch.allclose(_cf(T._linear_attention_block(_cl(x), la)), want, atol=1e-5)


@pytest.mark.parametrize("shift", [False, True])
def test_feed_forward_block(x, shift):
    ff = _randomise(M.FeedForward(16), 4)
    xin = R.token_shift(x) if shift else x
    want = R.feed_forward(xin, _sd(ff, "p."), "p.") + x
    assert torch.allclose(_cf(T._feed_forward_block(_cl(x), ff, shift)), want, atol=1e-5)


def test_gateloop_block(x):
    gl = _randomise(M.SimpleGateLoopLayer(16), 5)
    want = R.gateloop_time(x, _sd(gl, "p."), "p.") + x
    assert torch.allclose(_cf(T._gateloop_block(_cl(x), gl)), want, atol=1e-5)


def test_upsamplers(x):
    us = _randomise(M.SpatialUpsample2x(16, 8), 6)
    assert torch.allclose(_cf(T._upsample_space(_cl(x), us.net[0])), R.spatial_up(x, _sd(us, "p."), "p."), atol=1e-5)
    ut = _randomise(M.TimeUpsample2x(16, 8), 7)
    assert torch.allclose(_cf(T._upsample_time(_cl(x), ut.net[0])), R.time_up(x, _sd(ut, "p."), "p."), atol=1e-5)


def test_residual_unit_mod(x):
    mod = _randomise(M.ResidualUnitMod(16, (3, 3, 3), 12), 8)
    cond = torch.randn(2, 12, generator=torch.Generator().manual_seed(9))
    want = R.residual_unit_mod(x, cond, _sd(mod, "p."), "p.")
    assert torch.allclose(_cf(T._residual_unit_mod(_cl(x), cond, mod)), want, atol=1e-5)


@pytest.mark.parametrize("nc,spherical", [(1, False), (2, True)])
def test_lfq_train_terms(x, nc, spherical):
    """Straight-through output and auxiliary loss of the LFQ training branch (A.1 steps 2-10) against the oracle's
    lfq_presign / lfq_train_losses, single process (avg_global = the local mean)."""
    qz = _randomise(M.LFQ(16, 16, 0.1, 1.0, 2.5, 10., num_codebooks=nc, spherical=spherical), 10)
    sd = _sd(qz, "quantizers.")
    p = R.lfq_presign(x, sd, 10., nc, spherical)
    ps, be, cm, aux, avg = R.lfq_train_losses(p, 4, None, 100., 2.5, 0.1, 1.0, nc)
    out, aux_t = T._lfq_train(_cl(x), qz, avg.reshape(-1))
    q_ref, _, _ = R.lfq_quantize(x, sd, 10., nc, spherical)
    assert torch.allclose(_cf(out), q_ref, atol=1e-5)               # the straight-through value equals the quantised output
    assert abs(aux_t.item() - aux.item()) < 1e-5


def test_fsq_train_value(x):
    qz = _randomise(M.FSQ([8, 5, 5], 16, num_codebooks=2), 11)
    q_ref, _, _ = R.fsq_quantize(x, _sd(qz, "quantizers."), [8, 5, 5], 2)
    assert torch.allclose(_cf(T._fsq_train(_cl(x), qz)), q_ref, atol=1e-5)
