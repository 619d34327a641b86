"""Sweep (mw, bn) tilings / the 2-CTA weight multicast of the slab kernel for the README-config FeedForward fc1 + GEGLU layers
(M:492, M:466-469).  Each configuration is checked bit-for-bit against the default.  Writes gpurun_out/sweep_ff.json"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200.engine import pack_ff

B = 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
LAYERS = [("fc1 512->2x1365 @16 T20", 512, (20, 16, 16)), ("fc1 256->2x682 @32 T20", 256, (20, 32, 32)), ("fc1 512->2x1365 @16 T5", 512, (5, 16, 16))]
CFGS = [("auto", None), ("1,256", None), ("2,128", None)]
STAGES = [None, "4"]


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
for name, C_, (T, H, W) in LAYERS:
    inner = int(C_ * 4 * 2 / 3)
    fc1w = torch.randn((2 * inner, C_, 1, 1, 1), device="cuda") * C_ ** -0.5
    fc2w = torch.randn((C_, inner, 1, 1, 1), device="cuda") * inner ** -0.5
    fc1, fc2 = pack_ff(fc1w, torch.randn(2 * inner, device="cuda") * 0.1, fc2w, torch.zeros(C_, device="cuda"), torch.bfloat16)
    x = torch.randn((B, T, H, W, C_), device="cuda").to(torch.bfloat16)
    run = lambda: eng.conv(x, fc1)
    for k in ("MV2_SLAB_CFG", "MV2_SLAB_CLUSTER"):
        os.environ.pop(k, None)
    ref = run().clone()
    gf = 2.0 * B * T * H * W * C_ * fc1.Co_tc / 1e9
    times, bad = {}, []
    for cfg, cl in [(c, st) for c, _ in CFGS for st in STAGES]:
        for k in ("MV2_SLAB_CFG", "MV2_SLAB_CLUSTER", "MV2_SLAB_STAGES"):
            os.environ.pop(k, None)
        if cfg != "auto":
            mw, bn = map(int, cfg.split(","))
            if mw * bn > 512 or (mw >= 2 and W <= 8) or (mw == 4 and W <= 16):
                continue
            os.environ["MV2_SLAB_CFG"] = cfg
        if cl:
            os.environ["MV2_SLAB_STAGES"] = cl
        key = cfg + (f" s{cl}" if cl else "")
        y = run()
        if not torch.equal(y, ref):
            bad.append((key, (y.float() - ref.float()).abs().max().item()))
        times[key] = round(timeit(run) * 1e3, 1)
    best = min(times, key=times.get)
    print(f"{name:28s} {gf:6.1f} GFLOP (padded)  auto {times['auto']:6.1f} us ({gf / times['auto'] / 1e-3 / 1e3:.0f} TFLOP/s)  best {best} {times[best]:6.1f} us   mismatches {bad}")
    print("      ", "  ".join(f"{k}:{v:.0f}" for k, v in times.items()))
    out.append(dict(layer=name, gflop=gf, us=times, best=best, mismatches=bad))
for k in ("MV2_SLAB_CFG", "MV2_SLAB_CLUSTER", "MV2_SLAB_STAGES"):
    os.environ.pop(k, None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_ff.json", "w"), indent=1)
