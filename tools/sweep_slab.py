"""Sweep tap-wise vs slab (mw, bn) configurations for the README-config conv shapes.  Writes gpurun_out/sweep.json"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import ACT_ELU, ACT_NONE
from magvit2_pytorch_b200.engine import pack_conv, pack_conv_in_kwpack

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
LAYERS = [
    ("res3 c64 T20 128", (64, 64, 3, 3, 3), (20, 128, 128)),
    ("res3 c128 T20 64", (128, 128, 3, 3, 3), (20, 64, 64)),
    ("res3 c256 T20 32", (256, 256, 3, 3, 3), (20, 32, 32)),
    ("res3 c512 T20 16", (512, 512, 3, 3, 3), (20, 16, 16)),
    ("res3 c512 T10 16", (512, 512, 3, 3, 3), (10, 16, 16)),
    ("res3 c512 T5 16", (512, 512, 3, 3, 3), (5, 16, 16)),
    ("pw c64 T20 128", (64, 64, 1, 1, 1), (20, 128, 128)),
    ("pw c128 T20 64", (128, 128, 1, 1, 1), (20, 64, 64)),
    ("pw c256 T20 32", (256, 256, 1, 1, 1), (20, 32, 32)),
    ("pw c512 T20 16", (512, 512, 1, 1, 1), (20, 16, 16)),
    ("pw c512 T5 16", (512, 512, 1, 1, 1), (5, 16, 16)),
    ("qkv 512->768 T20 16", (768, 512, 1, 1, 1), (20, 16, 16)),
    ("conv_out 64->3", (3, 64, 3, 3, 3), (20, 128, 128)),
    ("up 128->64x4 T20 64", (256, 128, 1, 1), (20, 64, 64)),
    ("up 512->256x4 T20 16", (1024, 512, 1, 1), (20, 16, 16)),
    ("conv_in kwpack", None, (20, 128, 128)),
]
CFGS = ["tap"] + [f"{mw},{bn}" for mw in (1, 2, 4) for bn in (256, 128, 64, 32)]


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


out = []
for name, wshape, (T, H, W) in LAYERS:
    if wshape is None:
        w = torch.randn((64, 3, 7, 7, 7), device="cuda") * 0.03
        pk = pack_conv_in_kwpack(w, torch.zeros(64, device="cuda"))
        x = torch.randn((B, T, H, W, 32), device="cuda").to(torch.bfloat16)
        kw = dict(pad=(6, 3, 0))
        flops = 2.0 * B * T * H * W * 64 * 3 * 343
        co = 64
    else:
        w = torch.randn(wshape, device="cuda") * 0.02
        is_up = name.startswith("up ")
        pk = pack_conv(w, torch.zeros(wshape[0], device="cuda"), torch.bfloat16, shuffle_q=4 if is_up else 1)
        x = torch.randn((B, T, H, W, wshape[1]), device="cuda").to(torch.bfloat16)
        kw = dict(act=ACT_ELU, shuffle=1) if is_up else dict(act=ACT_ELU)
        taps = 1
        for v_ in wshape[2:]:
            taps *= v_
        flops = 2.0 * B * T * H * W * wshape[0] * wshape[1] * taps
        co = wshape[0]
    res = {}
    for cfg in CFGS:
        if cfg == "tap":
            eng.tc_variant = "tap"
            os.environ.pop("MV2_SLAB_CFG", None)
        else:
            mw, bn = map(int, cfg.split(","))
            co_pad = (co + 31) // 32 * 32
            if bn > co_pad or co_pad % bn or (mw >= 2 and W <= 8) or mw * bn > 512:
                continue
            eng.tc_variant = "slab"
            os.environ["MV2_SLAB_CFG"] = cfg
        try:
            res[cfg] = timeit(lambda: eng.conv(x, pk, **kw))
        except Exception as e:  # unsupported combination
            res[cfg] = None
    os.environ.pop("MV2_SLAB_CFG", None)
    eng.tc_variant = "slab"
    res["auto"] = timeit(lambda: eng.conv(x, pk, **kw))
    best = min((v, k) for k, v in res.items() if v)
    out.append(dict(layer=name, gflop=flops / 1e9, ms=res, best=best[1], best_tflops=flops / best[0] / 1e9))
    print(f"{name:22s} " + " ".join(f"{k}={v*1e3:6.0f}us" if v else f"{k}=n/a" for k, v in res.items()) +
          f"  -> best {best[1]} {flops / best[0] / 1e9:6.0f} TF/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)
