"""Run a few single-layer convs once each (after 1 warm-up) so `ncu --set full` can capture them.
Usage: python tools/prof_layers.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import ACT_ELU, ACT_NONE
from magvit2_pytorch_b200.engine import pack_conv

B = 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
CASES = [
    ("tap", (512, 512, 1, 1, 1), (20, 16, 16), ACT_ELU),     # pointwise c512
    ("tap", (64, 64, 1, 1, 1), (20, 128, 128), ACT_ELU),     # pointwise c64
    ("tap", (512, 512, 3, 3, 3), (20, 16, 16), ACT_ELU),     # res3 c512 tap-wise
    ("slab", (64, 64, 3, 3, 3), (20, 128, 128), ACT_ELU),
    ("slab", (128, 128, 3, 3, 3), (20, 64, 64), ACT_ELU),
    ("slab", (512, 512, 3, 3, 3), (20, 16, 16), ACT_ELU),
]
for variant, wshape, (T, H, W), act in CASES:
    w = torch.randn(wshape, device="cuda") * 0.02
    pk = pack_conv(w, torch.zeros(wshape[0], device="cuda"), torch.bfloat16)
    x = torch.randn((B, T, H, W, wshape[1]), device="cuda").to(torch.bfloat16)
    eng.use_tc, eng.tc_variant = True, variant
    for _ in range(2):
        eng.conv(x, pk, act=act)
    torch.cuda.synchronize()
print("done")
