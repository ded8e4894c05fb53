#!/usr/bin/env python3
"""Phase breakdown of the helicity-parallel matrix-element kernel (clock64 timers, -DMF_HP_PROFILE):
    python tools/profile_phases.py build            # here: compile tools/bin/libmfp_<proc>_prof.so
    python tools/profile_phases.py run <nevents>    # on the GPU box
Prints SM cycles per phase summed over blocks and their shares."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIN = os.path.join(ROOT, "tools", "bin")
PROCS = sys.argv[3:] if len(sys.argv) > 3 else ["1_gg_ttxg", "1_gg_ttxgg"]
PHASES = ["externals", "currents", "pair objects", "amplitude tiles (DMMA)", "JAMP", "colour + reduction"]

if sys.argv[1] == "build":
    from madflow_b200 import build, codegen
    os.makedirs(BIN, exist_ok=True)
    for ir in build.builtin_irs():
        if ir["name"] not in PROCS:
            continue
        src = os.path.join(codegen.GENDIR, f"proc_{ir['name']}.cu")
        out = os.path.join(BIN, f"libmfp_{ir['name']}_prof.so")
        codegen.compile_source(src, out, extra_flags=["-DMF_HP_PROFILE"])
        print(out)
else:
    import numpy as np
    import torch
    from madflow_b200 import _runtime as rt
    from madflow_b200 import phasespace as ps

    nev = int(sys.argv[2])
    for name in PROCS:
        lib = rt.ProcessLib(name if name.endswith(".so") else os.path.join(BIN, f"libmfp_{name}_prof.so"))
        lib.set_variant("hp")
        n = lib.info.nexternal
        gen = ps.PhaseSpaceGenerator(n, 13e3, [173.0, 173.0] + [0.0] * (n - 4), com_output=False)
        x = torch.rand((nev, lib.info.ndim), dtype=torch.float64, device="cuda")
        p, w, x1, x2, _ = gen(x)
        g = 1.2177157847767195
        coup = torch.tensor([[complex(re, im) * g**k] for re, im, k in lib.coupling_defs], dtype=torch.complex128, device="cuda")
        out = torch.empty(nev, dtype=torch.float64, device="cuda")
        args = (p, 0, nev, [173.0, 1.4915000200271606], coup, 0, 0.7071067690849304, out)
        lib.smatrix(*args)
        lib.lib.mfp_profile_read(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.smatrix(*args)
        e1.record()
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 8)()
        lib.lib.mfp_profile_read(buf, 1)
        tot = sum(buf[:6])
        print(f"{name}: {nev} events, {e0.elapsed_time(e1):.3f} ms, {nev / e0.elapsed_time(e1) * 1e3:.4g} ev/s; block cycles per event {tot / nev:.0f}")
        for ph, c in zip(PHASES, buf[:6]):
            print(f"   {ph:20s} {c / nev:12.1f} cycles/event  {100.0 * c / tot:5.1f} %")
