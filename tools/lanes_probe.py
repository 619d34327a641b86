"""Throughput of the README step with 1..L concurrent stream lanes (device-resident inputs)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import README_KW
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200.host_io import StreamLanes
import synth_data as Wt

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
torch.manual_seed(0)
m = VideoTokenizer(**README_KW)
Wt.fill_state_dict_(m, 0)
m = m.cuda().bfloat16().eval()
m.cuda_graphs = True
NB = 12
vs = [Wt.synth_video(B, 3, 17, 128, seed=5 + i).cuda() for i in range(NB)]

def step(v):
    codes = m.tokenize(v)
    return codes, m.decode_from_code_indices(codes)

res = {}
ref = None
for L in (1, 2, 3, 4):
    lanes = StreamLanes(m, L)
    for i in range(3 * L):
        out, _ = lanes.run(step, vs[i % NB])
    lanes.join()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        out, _ = lanes.run(step, vs[i % NB])
    lanes.join()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # same result whatever the lane count
    o, _ = lanes.run(step, vs[0]); lanes.join(); torch.cuda.synchronize()
    if ref is None:
        ref = (o[0].clone(), o[1].clone())
    same = bool(torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]))
    res[L] = {"ms_per_step": ms, "frames_per_s": B * 17 / ms * 1e3, "identical_to_1_lane": same}
    print(L, res[L], flush=True)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "lanes_probe.json"), "w"), indent=1)
