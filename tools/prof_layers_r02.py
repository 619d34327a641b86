"""Runs every distinct tcgen05 layer shape of the README config (B = 4) exactly once so that `ncu --set full` captures one launch
per shape (tools/capture_r02.sh).  Prints the launch order.  Usage: python tools/prof_layers_r02.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import ACT_ELU, ACT_NONE, ACT_SILU, SHUFFLE_SPACE
from magvit2_pytorch_b200.engine import pack_conv, pack_conv_down_space, pack_conv_in_kwpack, pack_ff

B = 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
eng.use_tc, eng.tc_variant = True, "auto"
g = torch.Generator(device="cpu").manual_seed(0)


def rn(*s, scale=0.05):
    return (torch.randn(s, generator=g) * scale).cuda()


def xin(T, H, W, C):
    return rn(B, T, H, W, C, scale=1.0).to(torch.bfloat16)


order = []
# fused ResidualUnit front half, C = 64 and 128
for C_, (T, H, W) in ((64, (20, 128, 128)), (128, (20, 64, 64))):
    hd = max(16, C_ // 2)
    p = dict(conv3=pack_conv(rn(C_, C_, 3, 3, 3, scale=(27 * C_) ** -0.5), rn(C_), torch.bfloat16),
             conv1=pack_conv(rn(C_, C_, 1, 1, 1, scale=C_ ** -0.5), rn(C_), torch.bfloat16), wk=rn(C_), bk=0.05,
             w1=rn(hd, C_), b1=rn(hd), w2=rn(C_, hd), b2=rn(C_), hidden=hd)
    eng.residual_unit(xin(T, H, W, C_), p)
    order.append(f"fused_ru c{C_} {T}x{H}x{W}")
# plain 3x3x3
for C_, (T, H, W) in ((256, (20, 32, 32)), (512, (20, 16, 16)), (512, (10, 16, 16)), (512, (5, 16, 16))):
    eng.conv(xin(T, H, W, C_), pack_conv(rn(C_, C_, 3, 3, 3, scale=(27 * C_) ** -0.5), rn(C_), torch.bfloat16), act=ACT_ELU)
    order.append(f"conv3 c{C_} {T}x{H}x{W}")
# conv_in (kw-packed) and conv_out (channels-first, cropped)
pin = pack_conv_in_kwpack(rn(64, 3, 7, 7, 7, scale=0.03), rn(64))
eng.conv(rn(B, 20, 128, 128, 32, scale=1.0).to(torch.bfloat16), pin, pad=(6, 3, 0))
order.append("conv_in kwpack")
eng.conv(xin(20, 128, 128, 64), pack_conv(rn(3, 64, 3, 3, 3), rn(3), torch.bfloat16), pad=(2 - 3, 1, 1), out_spatial=(17, 128, 128), out_cf=True)
order.append("conv_out cf")
# down-samplers
for Ci, Co, (T, H, W) in ((64, 128, (20, 128, 128)), (128, 256, (20, 64, 64)), (256, 512, (20, 32, 32))):
    w = rn(Co, Ci, 3, 3, scale=(9 * Ci) ** -0.5)
    pk = pack_conv(w, rn(Co), torch.bfloat16)
    pack_conv_down_space(pk, w)
    eng.conv(xin(T, H, W, Ci), pk, stride=(1, 2, 2), pad=(0, 1, 1), out_spatial=(T, H // 2, W // 2))
    order.append(f"down_space {Ci}->{Co}")
eng.conv(xin(20, 16, 16, 512), pack_conv(rn(512, 512, 3, scale=0.03), rn(512), torch.bfloat16, k=(3, 1, 1)), stride=(2, 1, 1), pad=(2, 0, 0),
         out_spatial=(10, 16, 16))
order.append("down_time 512")
# up-sampler (depth-to-space, staged), FeedForward, residual projection, pointwise
eng.conv(xin(20, 64, 64, 128), pack_conv(rn(256, 128, 1, 1), rn(256), torch.bfloat16, shuffle_q=4), act=ACT_SILU, shuffle=SHUFFLE_SPACE)
order.append("up_space 128->64x4")
fc1, fc2 = pack_ff(rn(2730, 512, 1, 1, 1), rn(2730), rn(512, 1365, 1, 1, 1), rn(512), torch.bfloat16)
x = xin(20, 16, 16, 512)
h = eng.conv(x, fc1)
order.append("ff fc1+geglu 512")
eng.conv(h, fc2, res=x)
order.append("ff fc2+res 512")
eng.conv(xin(20, 16, 16, 512), pack_conv(rn(512, 512, 1, 1, 1), rn(512), torch.bfloat16), act=ACT_ELU)
order.append("pointwise c512")
torch.cuda.synchronize()
print("slab launches in order:", "; ".join(order))
