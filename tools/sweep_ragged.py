"""Sweep ragged N-tile widths (Co not a multiple of bn) for the deep conv3 layers, where 128x256 tiles quantise badly
over 148 SMs.  Each configuration is checked bit-for-bit against the default tiling.  Writes gpurun_out/sweep_ragged.json"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import ACT_ELU
from magvit2_pytorch_b200.engine import pack_conv

B = 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
eng.tc_variant = "slab"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
LAYERS = [
    ("res3 c512 T20 16", (512, 512, 3, 3, 3), (20, 16, 16)),
    ("res3 c512 T10 16", (512, 512, 3, 3, 3), (10, 16, 16)),
    ("res3 c512 T5 16", (512, 512, 3, 3, 3), (5, 16, 16)),
    ("res3 c256 T20 32", (256, 256, 3, 3, 3), (20, 32, 32)),
    ("res3 c1024 T5 16", (1024, 1024, 3, 3, 3), (5, 16, 16)),
    ("pw c512 T20 16", (512, 512, 1, 1, 1), (20, 16, 16)),
    ("pw c512 T5 16", (512, 512, 1, 1, 1), (5, 16, 16)),
]
CFGS = ["auto"] + [f"{mw},{bn}" for mw in (1, 2) for bn in (256, 208, 192, 176, 144, 128, 112, 96, 80)]


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


out = []
for name, wshape, (T, H, W) in LAYERS:
    w = torch.randn(wshape, device="cuda") * 0.02
    pk = pack_conv(w, torch.randn(wshape[0], device="cuda") * 0.1, torch.bfloat16)
    x = torch.randn((B, T, H, W, wshape[1]), device="cuda").to(torch.bfloat16)
    os.environ.pop("MV2_SLAB_CFG", None)
    ref = eng.conv(x, pk, act=ACT_ELU).clone()
    res, bad = {}, []
    for cfg in CFGS:
        if cfg == "auto":
            os.environ.pop("MV2_SLAB_CFG", None)
        else:
            mw, bn = map(int, cfg.split(","))
            if bn > wshape[0] or mw * bn > 512:
                continue
            os.environ["MV2_SLAB_CFG"] = cfg
        y = eng.conv(x, pk, act=ACT_ELU)
        if not torch.equal(y, ref):
            bad.append((cfg, (y.float() - ref.float()).abs().max().item()))
        res[cfg] = round(timeit(lambda: eng.conv(x, pk, act=ACT_ELU)) * 1e3, 1)
    best = min(res, key=res.get)
    print(f"{name:22s} auto {res['auto']:7.1f} us   best {best} {res[best]:7.1f} us   mismatches {bad}")
    print("      ", "  ".join(f"{k}:{v:.0f}" for k, v in res.items()))
    out.append(dict(layer=name, us=res, best=best, mismatches=bad))
os.environ.pop("MV2_SLAB_CFG", None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_ragged.json", "w"), indent=1)
