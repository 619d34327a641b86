"""The C-ABI library loads and exports every symbol include/magvit2_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from magvit2_pytorch_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "magvit2_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mv2_[a-z0-9_]+)\s*\(", src)))


def test_library_built():
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    for name in _lib.SIGNATURES:
        assert name in declared, f"{name} bound in _lib.py but not declared in the header"


def test_load_and_version():
    lib = _lib.load()
    assert lib.mv2_abi_version() == 3
    assert lib.mv2_se_workspace_bytes(2, 256, 64) == (2 * 8 * 66 + 2 * 80) * 4
    assert lib.mv2_linattn_workspace_bytes(3, 16, 1024) == 3 * 16 * 4 * 657 * 4 + 3 * 16 * 2 * 16 * 88 * 2


def test_struct_sizes_match_header_layout():
    # 5 pointers + 22 int32 (conv), 3 pointers + 8 int32 + 3 int64 (attention)
    assert ctypes.sizeof(_lib.ConvArgs) == 5 * 8 + 22 * 4 + 8            # + oscale pointer
    assert ctypes.sizeof(_lib.TcConvArgs) == 5 * 8 + 22 * 4 + 8 + 8      # + oscale pointer, out_layout (+ tail padding)
    assert ctypes.sizeof(_lib.AttnArgs) == 3 * 8 + 8 * 4 + 3 * 8


def test_sass_is_sm100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_tcgen05_issue_loops_stay_in_uniform_registers():
    """Regression guard for the single most expensive lesson of the slab kernel: the warp that issues tcgen05.mma must
    keep its loop nest (descriptors, ring indices, predicates) in uniform registers.  When the compiler cannot prove the
    tile id / trip counts warp-uniform it re-materialises them with R2UR right in front of every UTCHMMA, which cost
    5-8 % on every layer (profiles/r01_bench_v29.json vs the table-driven schedule).  Static check on the SASS."""
    import re
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    kernels = re.split(r"\n\s*Function : ", sass)[1:]
    checked = 0
    for k in kernels:
        name = k.split("\n", 1)[0]
        if "tc_slab_kernel" not in name and "tc_conv_kernel" not in name:
            continue
        ins = [l for l in k.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        mma = [i for i, l in enumerate(ins) if "UTCHMMA" in l]
        assert mma, f"{name}: no tcgen05.mma (UTCHMMA) in the SASS"
        window = ins[max(0, mma[0] - 45):mma[0]]
        assert not any("R2UR" in l for l in window), f"{name}: MMA operands are converted from vector registers per issue"
        assert any("UTMALDG" in l for l in ins), f"{name}: no TMA tensor loads (UTMALDG)"
        checked += 1
    assert checked >= 6          # 4 slab + >= 2 tap-kernel epilogue flavours
