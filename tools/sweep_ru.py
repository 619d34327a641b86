"""Fused ResidualUnit kernel (mv2_tc_ru_forward) vs the unfused launches it replaces (conv3x3x3, conv1x1x1, se_pool), per
README layer shape, over MV2_RU_CFG = "mw,nh,tpw,slab_stages,w_stages".  Writes gpurun_out/sweep_ru.json.
Usage: python tools/sweep_ru.py [B]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magvit2_pytorch_b200 import VideoTokenizer
from magvit2_pytorch_b200.engine import pack_conv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m = VideoTokenizer(image_size=32, init_dim=16, codebook_size=1024, layers=("residual",)).cuda().bfloat16()
eng = m.engine
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
LAYERS = [("ru c64 T20 128", 64, (20, 128, 128)), ("ru c128 T20 64", 128, (20, 64, 64))]
CFGS = {64: ["4,2,1,2,0", "4,1,1,2,0", "2,2,1,3,0", "2,2,3,3,0", "2,2,1,2,0", "2,1,1,3,0", "4,1,1,2,5", "4,1,1,2,4", "4,1,1,2,3"],
        128: ["2,1,1,2,0", "2,1,1,2,3", "2,1,1,2,4", "2,2,1,2,0", "2,1,1,3,0"]}


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def pack(C_):
    g = torch.Generator(device="cpu").manual_seed(C_)
    w3 = (torch.randn((C_, C_, 3, 3, 3), generator=g) * (27 * C_) ** -0.5).cuda()
    w1 = (torch.randn((C_, C_, 1, 1, 1), generator=g) * C_ ** -0.5).cuda()
    hd = max(16, C_ // 2)
    z = lambda *s: (torch.randn(s, generator=g) * 0.1).cuda()
    return dict(conv3=pack_conv(w3, z(C_), torch.bfloat16), conv1=pack_conv(w1, z(C_), torch.bfloat16), wk=z(C_), bk=0.05,
                w1=z(hd, C_), b1=z(hd), w2=z(C_, hd), b2=z(C_), hidden=hd)


import ctypes as C
from magvit2_pytorch_b200._lib import TcRuArgs, check, ACT_ELU


def components(x, p, C_):
    """Per-launch times (ms) of the pieces of one ResidualUnit, fused and unfused."""
    Bx, T, H, W, _ = x.shape
    F_, Pn = Bx * T, H * W
    st = eng._stream()
    lib = eng.lib
    res = {}
    c3, c1 = p["conv3"], p["conv1"]
    ra = TcRuArgs(x=x.data_ptr(), w3=c3.w_tc.data_ptr(), b3=c3.bias_tc.data_ptr(), w1=c1.w_tc.data_ptr(), b1=c1.bias_tc.data_ptr(),
                  se_wk=p["wk"].data_ptr(), se_bk=p["bk"], y=None, se_ws=None, B=Bx, T=T, H=H, W=W, C=C_, kt=3, kh=3, kw=3)
    y = torch.empty_like(x)
    ws = torch.empty(lib.mv2_tc_ru_workspace_bytes(C.byref(ra)) // 4, device="cuda", dtype=torch.float32)
    ra.y, ra.se_ws = y.data_ptr(), ws.data_ptr()
    nrec = lib.mv2_tc_ru_records(C.byref(ra))
    gates = torch.empty((F_, C_), device="cuda", dtype=torch.float32)
    out = torch.empty_like(x)
    res["fused: mv2_tc_ru_forward"] = timeit(lambda: check(lib.mv2_tc_ru_forward(C.byref(ra), st)))
    res[f"fused: se_gate_records (nrec={nrec})"] = timeit(lambda: check(lib.mv2_se_gate_records(
        ws.data_ptr(), nrec, F_, C_, p["hidden"], p["w1"].data_ptr(), p["b1"].data_ptr(), p["w2"].data_ptr(), p["b2"].data_ptr(), gates.data_ptr(), st)))
    res["gate_residual"] = timeit(lambda: check(lib.mv2_gate_residual(y.data_ptr(), x.data_ptr(), gates.data_ptr(), out.data_ptr(), 1, F_, Pn, C_, st)))
    res["unfused: conv3x3x3"] = timeit(lambda: eng.conv(x, c3, act=ACT_ELU))
    res["unfused: conv1x1x1"] = timeit(lambda: eng.conv(y, c1, act=ACT_ELU))
    ws2 = torch.empty(lib.mv2_se_workspace_bytes(F_, Pn, C_) // 4, device="cuda", dtype=torch.float32)
    res["unfused: se_pool"] = timeit(lambda: check(lib.mv2_se_pool(y.data_ptr(), 1, F_, Pn, C_, p["wk"].data_ptr(), p["bk"], ws2.data_ptr(), st)))
    res["unfused: se_gate"] = timeit(lambda: check(lib.mv2_se_gate(ws2.data_ptr(), 1, F_, Pn, C_, p["hidden"], p["w1"].data_ptr(), p["b1"].data_ptr(),
                                                                   p["w2"].data_ptr(), p["b2"].data_ptr(), gates.data_ptr(), st)))
    return res


out = []
for name, C_, (T, H, W) in LAYERS:
    p = pack(C_)
    x = torch.randn((B, T, H, W, C_), device="cuda").to(torch.bfloat16)
    rec = {"layer": name, "B": B, "ms": {}}
    os.environ.pop("MV2_RU_CFG", None)
    rec["components_ms"] = components(x, p, C_)
    eng.fuse_ru = False
    os.environ.pop("MV2_RU_CFG", None)
    rec["ms"]["unfused whole unit"] = timeit(lambda: eng.residual_unit(x, p))
    eng.fuse_ru = True
    ref = None
    for cfg in ["default"] + CFGS[C_]:
        if cfg == "default":
            os.environ.pop("MV2_RU_CFG", None)
        else:
            os.environ["MV2_RU_CFG"] = cfg
        try:
            rec["ms"]["fused unit " + cfg] = timeit(lambda: eng.residual_unit(x, p))
            y = eng.residual_unit(x, p)
            torch.cuda.synchronize()
            if ref is None:
                ref = y
            else:
                rec.setdefault("same_as_default", {})[cfg] = bool(torch.equal(y, ref))
        except Exception as e:  # noqa: BLE001
            rec["ms"]["fused unit " + cfg] = f"error: {e}"
    os.environ.pop("MV2_RU_CFG", None)
    out.append(rec)
    print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_ru.json", "w"), indent=1)
