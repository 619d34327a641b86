"""Per-layer timing of the README-config dense contractions on one GPU (CUDA events, L2 flushed between
iterations).  Usage: python tools/bench_layers.py [B]   -> prints a table + writes gpurun_out/layers.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer  # noqa: E402
from magvit2_pytorch_b200._lib import ACT_ELU, ACT_NONE, ACT_SILU, SHUFFLE_SPACE, SHUFFLE_TIME  # noqa: E402
from magvit2_pytorch_b200.engine import pack_conv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

LAYERS = [
    # name, weight shape, k3, (T,H,W), kwargs
    ("res3 c64 20x128x128", (64, 64, 3, 3, 3), None, (20, 128, 128), dict(act=ACT_ELU)),
    ("res3 c128 20x64x64", (128, 128, 3, 3, 3), None, (20, 64, 64), dict(act=ACT_ELU)),
    ("res3 c256 20x32x32", (256, 256, 3, 3, 3), None, (20, 32, 32), dict(act=ACT_ELU)),
    ("res3 c512 20x16x16", (512, 512, 3, 3, 3), None, (20, 16, 16), dict(act=ACT_ELU)),
    ("res3 c512 10x16x16", (512, 512, 3, 3, 3), None, (10, 16, 16), dict(act=ACT_ELU)),
    ("res3 c512 5x16x16", (512, 512, 3, 3, 3), None, (5, 16, 16), dict(act=ACT_ELU)),
    ("pw c64 20x128x128", (64, 64, 1, 1, 1), None, (20, 128, 128), dict(act=ACT_ELU)),
    ("pw c512 20x16x16", (512, 512, 1, 1, 1), None, (20, 16, 16), dict(act=ACT_ELU)),
    ("down_space 64->128", (128, 64, 3, 3), None, (20, 128, 128), dict(stride=(1, 2, 2), pad=(0, 1, 1), down="space")),
    ("up_space 128->64x4", (256, 128, 1, 1), None, (20, 64, 64), dict(act=ACT_SILU, shuffle=SHUFFLE_SPACE, q=4)),
    ("conv_out 64->3", (3, 64, 3, 3, 3), None, (20, 128, 128), dict()),
    ("ff1 512->2730 20x16x16", (2730, 512, 1, 1, 1), None, (20, 16, 16), dict()),
]
rows = []
for name, wshape, k3, (T, H, W), kw in LAYERS:
    kw = dict(kw)
    q = kw.pop("q", 1)
    down = kw.pop("down", None)
    if down == "space":
        kw["out_spatial"] = (T, H // 2, W // 2)
    w = torch.randn(wshape, device="cuda") * 0.02
    bias = torch.zeros(wshape[0], device="cuda")
    pk = pack_conv(w, bias, torch.bfloat16, k=k3, shuffle_q=q)
    x = torch.randn((B, T, H, W, wshape[1]), device="cuda").to(torch.bfloat16)
    To, Ho, Wo = kw.get("out_spatial", (T, H, W))
    taps = 1
    for v in wshape[2:]:
        taps *= v
    flops = 2.0 * B * To * Ho * Wo * wshape[0] * wshape[1] * taps
    res = {}
    for mode in ("tc", "slab"):
        eng.use_tc = True
        eng.tc_variant = "tap" if mode == "tc" else "slab"
        if mode == "slab" and not (len(wshape) == 5 and wshape[2] * wshape[3] * wshape[4] > 1 and wshape[0] > 3):
            continue
        for _ in range(3):
            eng.conv(x, pk, **kw)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.conv(x, pk, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[mode] = min(ts)
    row = dict(layer=name, B=B, gflop=flops / 1e9, tc_ms=res.get("tc"), slab_ms=res.get("slab"),
               tc_tflops=flops / res["tc"] / 1e9 if "tc" in res else None,
               slab_tflops=flops / res["slab"] / 1e9 if "slab" in res else None)
    rows.append(row)
    print(f"{name:28s} {flops/1e9:9.1f} GF  tap {res.get('tc', float('nan')):8.3f} ms = {row['tc_tflops']:7.1f} TF/s"
          f"   slab {res.get('slab', float('nan')):8.3f} ms = {(row['slab_tflops'] or float('nan')):7.1f} TF/s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/layers.json", "w"), indent=1)
