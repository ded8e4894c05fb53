#!/usr/bin/env python3
"""Build tuning variants of a process library: python tools/build_variants.py <k final gluons> TAG:ENV=VAL,ENV=VAL ...
(ENV without the MADFLOW_B200_HP_ prefix, e.g.  a:E=1,NCG=2,MINBLOCKS=4).  Output: tools/bin/libmfp_<proc>_<TAG>.so"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from madflow_b200 import codegen, procgen  # noqa: E402

BIN = os.path.join(ROOT, "tools", "bin")
os.makedirs(BIN, exist_ok=True)
k = int(sys.argv[1])
ir = procgen.generate_ir(k)
procs = []
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
    cmd = ["nvcc", "-Xptxas=-v"] + codegen.NVCC_FLAGS + extra + ["-I", codegen.CSRC, "-o", out, src]
    procs.append((tag, out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for tag, out, pr in procs:
    log = pr.communicate()[0]
    if pr.returncode:
        print(tag, "FAILED\n", log[-3000:])
        continue
    m = re.search(r"smatrix_kernel_hp.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", log, re.S)
    print(f"{tag}: {out}  regs {m.group(3)} stack {m.group(1)} spill {m.group(2)}" if m else f"{tag}: built")
