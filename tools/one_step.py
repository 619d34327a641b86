"""Tokenize+decode steps of the README config (bf16) for ncu launch lists / captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import README_KW
from magvit2_pytorch_b200 import VideoTokenizer
import synth_data as Wt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
m = VideoTokenizer(**README_KW)
Wt.fill_state_dict_(m, 0)
m = m.cuda().bfloat16().eval()
v = Wt.synth_video(B, 3, 17, 128, seed=5).cuda()
for _ in range(steps):
    codes = m.tokenize(v)
    rec = m.decode_from_code_indices(codes)
torch.cuda.synchronize()
e = m.engine
print("launches per step:", e.launches // steps, "tc convs:", e.tc_calls // steps, "simt convs:", e.simt_conv_calls // steps)
