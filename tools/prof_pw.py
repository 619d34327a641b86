"""Single pointwise (1x1x1, C=64, 4x20x128x128) conv through the slab kernel, for an ncu --set full capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import ACT_ELU
from magvit2_pytorch_b200.engine import pack_conv
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
for C_, (T, H, W) in ((64, (20, 128, 128)), (128, (20, 64, 64))):
    w = torch.randn((C_, C_, 1, 1, 1), device="cuda") * 0.1
    pk = pack_conv(w, torch.zeros(C_, device="cuda"), torch.bfloat16)
    x = torch.randn((4, T, H, W, C_), device="cuda").to(torch.bfloat16)
    eng.use_tc, eng.tc_variant = True, "slab"
    for _ in range(2):
        eng.conv(x, pk, act=ACT_ELU)
torch.cuda.synchronize()
print("done")
