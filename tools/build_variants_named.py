#!/usr/bin/env python3
"""Tuning variants of any built-in process library:  python tools/build_variants_named.py <process name> TAG:KEY=VAL,... ...
(the companion of build_variants.py, which only knows g g > t t~ + k g).  Output: tools/bin/libmfp_<proc>_<TAG>.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from madflow_b200 import build, codegen  # noqa: E402

BIN = os.path.join(ROOT, "tools", "bin")
os.makedirs(BIN, exist_ok=True)
ir = [b for b in build.builtin_irs() if b["name"] == sys.argv[1]][0]
for spec in sys.argv[2:]:
    tag, _, envs = spec.partition(":")
    extra = []
    for kv in filter(None, envs.split(",")):
        key, val = kv.split("=")
        if key == "FLAGS":
            extra += val.split("+")
        else:
            os.environ["MADFLOW_B200_HP_" + key] = val
    src = os.path.join(BIN, f"var_{ir['name']}_{tag}.cu")
    open(src, "w").write(codegen.emit_process_source(ir))
    for kv in filter(None, envs.split(",")):
        os.environ.pop("MADFLOW_B200_HP_" + kv.split("=")[0], None)
    out = os.path.join(BIN, f"libmfp_{ir['name']}_{tag}.so")
    res = subprocess.run(["nvcc", "-Xptxas=-v"] + codegen.NVCC_FLAGS + extra + ["-I", codegen.CSRC, "-o", out, src], capture_output=True, text=True)
    print(tag, out if res.returncode == 0 else "FAILED\n" + res.stderr[-2000:])
    for line in res.stderr.splitlines():
        if "smatrix_kernel_hp" in line and "Function properties" in line:
            i = res.stderr.splitlines().index(line)
            print("   ", " ".join(res.stderr.splitlines()[i + 1:i + 3]).strip())
