"""Per-kernel-class time of one README tokenize+decode step (bf16, 4 clips), from a CUPTI trace (torch.profiler) of
eager launches: warm-cache kernel durations, no launch gaps.  tcgen05 conv launches are matched in order with the
engine's shape log so they can be grouped by layer class.
Usage: python tools/step_breakdown.py [out.json] [readme|cfg4|fsq]"""
import json, os, sys, collections
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import README_KW, WORKLOADS
from magvit2_pytorch_b200 import VideoTokenizer
import synth_data as Wt

STEPS = 3
WL = WORKLOADS[sys.argv[2]] if len(sys.argv) > 2 else WORKLOADS["readme"]
m = VideoTokenizer(**WL["kw"])
Wt.fill_state_dict_(m, 0)
m = m.cuda().bfloat16().eval()
m.cuda_graphs = False
v = Wt.synth_video(WL["clips"], 3, 17, WL["size"], seed=5).cuda()


def step():
    return m.decode_from_code_indices(m.tokenize(v))


for _ in range(2):
    step()
torch.cuda.synchronize()
eng = m.engine
eng.conv_log = []
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        step()
    torch.cuda.synchronize()
log, eng.conv_log = eng.conv_log, None
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "mem" not in e.name.lower()[:6]),
             key=lambda e: e.time_range.start)
tc = [e for e in evs if "tc_slab_kernel" in e.name or "tc_conv_kernel" in e.name]
assert len(tc) == len(log), (len(tc), len(log))


def conv_class(r):
    k = r["k"]
    taps = k[0] * k[1] * k[2]
    if r["geglu"]:
        return "ff fc1+geglu"
    if r["shuffle"]:
        return "upsample conv"
    if r["kind"] == "tap":
        return "strided conv (tap kernel)"
    if taps == 27:
        return f"conv3x3x3 c{r['Ci']}"
    if taps > 27:
        return "conv_in 7x7x7"
    if taps > 1:
        return f"conv {k} {r['Ci']}->{r['Co']}"
    if r["Co"] <= 8:
        return "conv_out"
    return f"pointwise {'(+res) ' if r['res'] else ''}c{r['Ci']}->{r['Co']} @{r['out'][2]}"


agg = collections.defaultdict(lambda: [0.0, 0])
ids = {id(e): conv_class(r) for e, r in zip(tc, log)}
for e in evs:
    name = ids.get(id(e))
    if name is None:
        name = e.name.split("(")[0].replace("void ", "").replace("mv2::", "")
        name = name.split("<")[0]
    agg[name][0] += e.device_time_total / STEPS
    agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
for k, (us, n) in rows:
    print(f"{k:48s} {us:9.1f} us  {n // STEPS:4d} launches  {100 * us / tot:5.1f} %")
print(f"{'total kernel time / step':48s} {tot:9.1f} us")
# per-launch durations (last step, launch order) of the small CUDA-core kernels
last = evs[len(evs) * (STEPS - 1) // STEPS:]
seqs = collections.defaultdict(list)
for e in last:
    name = ids.get(id(e)) or e.name.split("(")[0].replace("void ", "").replace("mv2::", "").split("<")[0]
    seqs[name].append(round(e.device_time_total, 1))
for k in ("se_pool_online_kernel", "se_hidden_kernel", "se_out_kernel", "gate_residual_bf16x8_kernel", "rmsnorm_bf16x8_kernel",
          "pointwise c512->512 @16", "conv3x3x3 c512", "ff fc1+geglu", "strided conv (tap kernel)", "upsample conv"):
    print(f"{k}: {seqs.get(k)}")
if len(sys.argv) > 1:
    json.dump({"total_us": tot, "classes": {k: {"us": us, "launches": n // STEPS} for k, (us, n) in rows}}, open(sys.argv[1], "w"), indent=1)
