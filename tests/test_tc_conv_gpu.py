"""tcgen05 implicit-GEMM convolution vs the CUDA-core kernel on identical bf16 inputs (both accumulate in
fp32), shape by shape, through the engine's conv() entry (C ABI underneath)."""
import pytest
import torch

from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import ACT_ELU, ACT_NONE, ACT_SILU, SHUFFLE_NONE, SHUFFLE_SPACE, SHUFFLE_TIME
from magvit2_pytorch_b200.engine import Engine, pack_conv

pytestmark = pytest.mark.gpu

# (name, weight shape (Co,Ci,*k), k3, x shape (B,T,H,W), conv kwargs, shuffle_q)
CASES = [
    ("res3x3x3_c64", (64, 64, 3, 3, 3), None, (1, 3, 32, 32), dict(act=ACT_ELU), 1),
    ("res3x3x3_c128", (128, 128, 3, 3, 3), None, (2, 4, 16, 16), dict(act=ACT_ELU), 1),
    ("res3x3x3_c256", (256, 256, 3, 3, 3), None, (1, 3, 8, 16), dict(), 1),
    ("res3x3x3_c512_2ntiles", (512, 512, 3, 3, 3), None, (1, 2, 8, 8), dict(act=ACT_ELU), 1),
    ("res3x3x3_wide_w128", (64, 64, 3, 3, 3), None, (1, 2, 4, 128), dict(), 1),
    ("res3x3x3_ragged", (64, 64, 3, 3, 3), None, (1, 3, 12, 24), dict(act=ACT_ELU), 1),
    ("pointwise_c64", (64, 64, 1, 1, 1), None, (2, 3, 16, 16), dict(act=ACT_ELU), 1),
    ("linear_512_768", (768, 512, 1, 1, 1), None, (1, 2, 16, 16), dict(), 1),
    ("ff1_ragged_n_256_1364", (1364, 256, 1, 1, 1), None, (1, 2, 8, 8), dict(), 1),
    ("residual_add", (256, 128, 1, 1, 1), None, (1, 2, 16, 16), dict(res=True), 1),
    ("conv_out_co3", (3, 64, 3, 3, 3), None, (1, 3, 16, 32), dict(), 1),
    ("down_space", (128, 64, 3, 3), None, (2, 3, 32, 32), dict(stride=(1, 2, 2), pad=(0, 1, 1), down="space"), 1),
    ("down_time", (512, 512, 3), (3, 1, 1), (1, 8, 8, 8), dict(stride=(2, 1, 1), pad=(2, 0, 0), down="time"), 1),
    ("down_time_oddT", (64, 64, 3), (3, 1, 1), (1, 5, 8, 16), dict(stride=(2, 1, 1), pad=(2, 0, 0), down="time"), 1),
    ("up_space", (256, 128, 1, 1), None, (1, 3, 16, 16), dict(act=ACT_SILU, shuffle=SHUFFLE_SPACE), 4),
    ("up_time", (1024, 512, 1), (1, 1, 1), (1, 3, 8, 8), dict(act=ACT_SILU, shuffle=SHUFFLE_TIME), 2),
    ("bk32_c32", (32, 32, 3, 3, 3), None, (1, 3, 8, 8), dict(act=ACT_ELU), 1),
    ("bk16_c16", (16, 16, 3, 3, 3), None, (1, 3, 16, 16), dict(act=ACT_ELU), 1),
    ("bk16_c48_tinyspatial", (64, 48, 3, 3, 3), None, (2, 3, 4, 4), dict(), 1),
    ("up_space_small_cy16", (64, 32, 1, 1), None, (1, 2, 8, 8), dict(act=ACT_SILU, shuffle=SHUFFLE_SPACE), 4),
]


def _engine():
    m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
    return m.engine


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tc_matches_cuda_core(case):
    assert torch.cuda.is_available()
    name, wshape, k3, xshape, kw, q = case
    kw = dict(kw)
    g = torch.Generator(device="cpu").manual_seed(sum(map(ord, name)))
    fan_in = 1
    for v in wshape[1:]:
        fan_in *= v
    w = (torch.randn(wshape, generator=g) * fan_in ** -0.5).cuda()
    bias = (torch.randn(wshape[0], generator=g) * 0.1).cuda()
    B, T, H, W = xshape
    x = torch.randn((B, T, H, W, wshape[1]), generator=g).cuda().to(torch.bfloat16)
    eng = _engine()
    pk = pack_conv(w, bias, torch.bfloat16, k=k3, shuffle_q=q)
    down = kw.pop("down", None)
    if down == "space":
        kw["out_spatial"] = (T, (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1)
    elif down == "time":
        kw["out_spatial"] = ((T + 2 - 3) // 2 + 1, H, W)
    want_res = kw.pop("res", False)

    def run(use_tc):
        eng.use_tc = use_tc
        res = None
        if want_res:
            To, Ho, Wo = kw.get("out_spatial", (T, H, W))
            gg = torch.Generator(device="cpu").manual_seed(7)
            res = torch.randn((B, To, Ho, Wo, wshape[0]), generator=gg).cuda().to(torch.bfloat16)
        return eng.conv(x, pk, res=res, **kw)

    eng.tc_calls = 0
    y_tc = run(True)
    assert eng.tc_calls == 1, "tcgen05 path was not taken"
    y_ref = run(False)
    torch.cuda.synchronize()
    assert y_tc.shape == y_ref.shape
    a, b = y_tc.float(), y_ref.float()
    assert torch.isfinite(a).all()
    tol = 0.008 * b.abs().max().item() + 1e-3          # one bf16 ulp at the output scale
    err = (a - b).abs().max().item()
    assert err <= tol, f"{name}: max-abs diff {err} > {tol}"
    assert (a - b).abs().mean().item() < 0.002 * b.abs().mean().item() + 1e-4
