"""SE tail kernel vs the 4-launch general path on the deep-level shapes (C = 512, 16 x 16 frames).  Also the target of
`ncu --set full -k regex:se_tail_kernel`.  Usage: python tools/prof_se_tail.py"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200._lib import check

m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
lib = eng.lib
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, flush_l2=False, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


out = []
for (C_, F_, P) in [(512, 80, 256), (512, 40, 256), (512, 20, 256), (256, 80, 1024)]:
    Hd = max(16, C_ // 2)
    g = torch.Generator(device="cpu").manual_seed(C_)
    rn = lambda *s: (torch.randn(s, generator=g) * 0.1).cuda()
    y, x = rn(F_, P, C_).to(torch.bfloat16), rn(F_, P, C_).to(torch.bfloat16)
    o = torch.empty_like(y)
    wk, b1, b2 = rn(C_), rn(Hd), rn(C_)
    w1, w2 = rn(Hd, C_).to(torch.bfloat16), rn(C_, Hd).to(torch.bfloat16)
    w1f, w2f = w1.float().contiguous(), w2.float().contiguous()
    st = eng._stream()
    rec = {"C": C_, "F": F_, "P": P}
    if lib.mv2_se_tail_supported(F_, P, C_, Hd):
        fn = lambda: check(lib.mv2_se_tail(y.data_ptr(), x.data_ptr(), o.data_ptr(), F_, P, C_, Hd, wk.data_ptr(), 0.1, w1.data_ptr(), b1.data_ptr(),
                                           w2.data_ptr(), b2.data_ptr(), st))
        rec["se_tail_us_warm"] = timeit(fn)
        rec["se_tail_us_l2_flushed"] = timeit(fn, True, 8)
    ws = torch.empty(lib.mv2_se_workspace_bytes(F_, P, C_) // 4, device="cuda", dtype=torch.float32)
    gates = torch.empty((F_, C_), device="cuda", dtype=torch.float32)

    def general():
        check(lib.mv2_se_pool(y.data_ptr(), 1, F_, P, C_, wk.data_ptr(), 0.1, ws.data_ptr(), st))
        check(lib.mv2_se_gate(ws.data_ptr(), 1, F_, P, C_, Hd, w1f.data_ptr(), b1.data_ptr(), w2f.data_ptr(), b2.data_ptr(), gates.data_ptr(), st))
        check(lib.mv2_gate_residual(y.data_ptr(), x.data_ptr(), gates.data_ptr(), o.data_ptr(), 1, F_, P, C_, st))
    rec["general_4_launches_us_warm"] = timeit(general)
    rec["general_4_launches_us_l2_flushed"] = timeit(general, True, 8)
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/se_tail.json", "w"), indent=1)
