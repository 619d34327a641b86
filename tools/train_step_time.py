"""Time of the training-mode generator step (forward + backward + SGD step) on the README config, bf16, B clips (default 4):
forward through the engine kernels, backward = own dgrad kernels + aten.convolution_backward + torch restatements (train.py).
Usage: python tools/train_step_time.py [B] [own_dgrad 0|1]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import README_KW
from magvit2_pytorch_b200 import VideoTokenizer, train as T
import synth_data as Wt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
own = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
orig = T.TrainRunner.__init__


def patched(self, m):
    orig(self, m)
    self.own_dgrad = own


T.TrainRunner.__init__ = patched
torch.manual_seed(0)
m = VideoTokenizer(**dict(README_KW, use_gan=False, perceptual_loss_weight=0.))
Wt.fill_state_dict_(m, 0)
m = m.cuda().bfloat16().train()
opt = torch.optim.SGD(m.parameters(), lr=1e-6)
v = Wt.synth_video(B, 3, 17, 128, seed=5).cuda().bfloat16()


def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = m(v, return_loss=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
with torch.no_grad():
    m.eval()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m(v, return_codes=True, return_recon=True)
    torch.cuda.synchronize()
    f0.record()
    for _ in range(n):
        m(v, return_codes=True, return_recon=True)
    f1.record()
    torch.cuda.synchronize()
out = {"clips": B, "own_dgrad": own, "train_step_ms": ms, "eval_forward_ms": f0.elapsed_time(f1) / n, "loss": float(loss),
       "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
print(json.dumps(out))
