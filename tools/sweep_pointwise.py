"""Sweep (mw, bn) tilings of the slab kernel for the README-config 1x1x1 layers (pointwise convs, attention projections,
FeedForward fc2 with residual) -- the HBM / latency bound part of the step.  Each configuration is
checked bit-for-bit against the default tiling.  Writes gpurun_out/sweep_pointwise.json
Usage: python tools/sweep_pointwise.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import ACT_ELU, ACT_NONE
from magvit2_pytorch_b200.engine import pack_conv

B = 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
eng.tc_variant = "slab"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
# (name, Co, Ci, (T, H, W), act, residual)
LAYERS = [
    ("pw c64 @128", 64, 64, (20, 128, 128), ACT_ELU, False),
    ("pw c128 @64", 128, 128, (20, 64, 64), ACT_ELU, False),
    ("pw c256 @32", 256, 256, (20, 32, 32), ACT_ELU, False),
    ("pw c512 @16 T20", 512, 512, (20, 16, 16), ACT_ELU, False),
    ("pw c512 @16 T5", 512, 512, (5, 16, 16), ACT_ELU, False),
    ("qkv 512->768 @16", 768, 512, (20, 16, 16), ACT_NONE, False),
    ("attn out 256->512 +res", 512, 256, (20, 16, 16), ACT_NONE, True),
    ("linattn out 128->256 +res", 256, 128, (20, 32, 32), ACT_NONE, True),
    ("fc2 704->256 +res @32", 256, 704, (20, 32, 32), ACT_NONE, True),
    ("fc2 1408->512 +res @16", 512, 1408, (20, 16, 16), ACT_NONE, True),
]
CFGS = ["auto"] + [f"{mw},{bn}" for mw in (1, 2, 4) for bn in (256, 192, 128, 64)]


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
for name, co, ci, (T, H, W), act, with_res in LAYERS:
    w = torch.randn((co, ci, 1, 1, 1), device="cuda") * ci ** -0.5
    pk = pack_conv(w, torch.randn(co, device="cuda") * 0.1, torch.bfloat16)
    x = torch.randn((B, T, H, W, ci), device="cuda").to(torch.bfloat16)
    res = torch.randn((B, T, H, W, co), device="cuda").to(torch.bfloat16) if with_res else None
    run = lambda: eng.conv(x, pk, act=act, res=res)
    os.environ.pop("MV2_SLAB_CFG", None)
    ref = run().clone()
    times, bad = {}, []
    for cfg in CFGS:
        if cfg == "auto":
            os.environ.pop("MV2_SLAB_CFG", None)
        else:
            mw, bn = map(int, cfg.split(","))
            if bn > co or mw * bn > 512 or (mw >= 2 and W <= 8) or (mw == 4 and W <= 16):
                continue
            os.environ["MV2_SLAB_CFG"] = cfg
        y = run()
        if not torch.equal(y, ref):
            bad.append((cfg, (y.float() - ref.float()).abs().max().item()))
        times[cfg] = round(timeit(run) * 1e3, 1)
    best = min(times, key=times.get)
    mb = (x.numel() + ref.numel() * (2 if with_res else 1)) * 2 / 1e6
    print(f"{name:28s} {mb:6.0f} MB  auto {times['auto']:6.1f} us ({mb / times['auto']:.2f} TB/s)  best {best} {times[best]:6.1f} us   mismatches {bad}")
    print("      ", "  ".join(f"{k}:{v:.0f}" for k, v in times.items()))
    out.append(dict(layer=name, MB=mb, us=times, best=best, mismatches=bad))
os.environ.pop("MV2_SLAB_CFG", None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_pointwise.json", "w"), indent=1)
