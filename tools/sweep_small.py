"""Tuning sweep of the small memory-bound kernels on the README shapes (B = 4): se_pool chunk rows (MV2_SE_ROWS) incl. the SE
gate MLP that combines the records, rmsnorm tokens per warp (MV2_RN_TPW).  Warm L2 (these tensors are L2 resident in the real
step), median of 20.  Writes gpurun_out/sweep_small.json.  Usage: python tools/sweep_small.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import check

m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
lib = eng.lib


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return round(ts[len(ts) // 2] * 1e3, 2)


out = {"se": [], "rmsnorm": []}
g = torch.Generator(device="cpu").manual_seed(0)
rn = lambda *s: (torch.randn(s, generator=g) * 0.1).cuda()
for (C_, F_, P) in [(512, 80, 256), (512, 40, 256), (512, 20, 256), (256, 80, 1024)]:
    Hd = max(16, C_ // 2)
    y = rn(F_, P, C_).to(torch.bfloat16)
    wk, b1, b2, w1, w2 = rn(C_), rn(Hd), rn(C_), rn(Hd, C_), rn(C_, Hd)
    ws = torch.empty(lib.mv2_se_workspace_bytes(F_, P, C_) // 4, device="cuda", dtype=torch.float32)
    gates = torch.empty((F_, C_), device="cuda", dtype=torch.float32)
    st = eng._stream()
    rec = {"C": C_, "F": F_, "P": P, "us": {}}
    for rows in ("default", 32, 64, 128, 256, 512, 1024):
        if rows == "default":
            os.environ.pop("MV2_SE_ROWS", None)
        else:
            if rows > P:
                continue
            os.environ["MV2_SE_ROWS"] = str(rows)
        pool = lambda: check(lib.mv2_se_pool(y.data_ptr(), 1, F_, P, C_, wk.data_ptr(), 0.1, ws.data_ptr(), st))
        gate = lambda: check(lib.mv2_se_gate(ws.data_ptr(), 1, F_, P, C_, Hd, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gates.data_ptr(), st))
        pool(); 
        rec["us"][str(rows)] = {"se_pool": timeit(pool), "se_gate(hidden+out)": timeit(gate)}
    os.environ.pop("MV2_SE_ROWS", None)
    out["se"].append(rec)
    print(json.dumps(rec), flush=True)
for (C_, T, P, shift) in [(256, 20, 1024, 0), (512, 20, 256, 0), (512, 5, 256, 1)]:
    x = rn(4, T, P, C_).to(torch.bfloat16)
    o = torch.empty_like(x)
    gamma = rn(C_) + 1
    st = eng._stream()
    rec = {"C": C_, "T": T, "P": P, "us": {}}
    for tpw in ("default", 1, 2, 4):
        if tpw == "default":
            os.environ.pop("MV2_RN_TPW", None)
        else:
            os.environ["MV2_RN_TPW"] = str(tpw)
        rec["us"][str(tpw)] = timeit(lambda: check(lib.mv2_rmsnorm(x.data_ptr(), o.data_ptr(), 1, gamma.data_ptr(), 4, T, P, C_, shift, st)))
    os.environ.pop("MV2_RN_TPW", None)
    out["rmsnorm"].append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_small.json", "w"), indent=1)
