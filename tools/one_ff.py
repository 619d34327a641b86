"""One FeedForward fc1 + GEGLU launch (README 256-channel level by default) for ncu captures.  Usage: one_ff.py [C] [T] [HW]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200.engine import pack_ff

C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 32
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
inner = int(C_ * 4 * 2 / 3)
fc1, fc2 = pack_ff(torch.randn((2 * inner, C_, 1, 1, 1), device="cuda") * C_ ** -0.5, torch.randn(2 * inner, device="cuda") * 0.1,
                   torch.randn((C_, inner, 1, 1, 1), device="cuda") * inner ** -0.5, torch.zeros(C_, device="cuda"), torch.bfloat16)
x = torch.randn((4, T, HW, HW, C_), device="cuda").to(torch.bfloat16)
for _ in range(3):
    y = eng.conv(x, fc1)
torch.cuda.synchronize()
print(tuple(y.shape))
