#!/usr/bin/env python3
"""A/B timing of process libraries: python tools/time_smatrix.py <nevents> <lib.so> [<lib.so> ...]
Times mfp_smatrix (device-resident lab-frame RAMBO momenta, per-event couplings) with CUDA events."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madflow_b200 import _runtime as rt
from madflow_b200 import phasespace as ps

nev = int(sys.argv[1])
for path in sys.argv[2:]:
    lib = rt.ProcessLib(path)
    n = lib.info.nexternal
    masses = [173.0, 173.0] + [0.0] * (n - 4)
    gen = ps.PhaseSpaceGenerator(n, 13e3, masses, com_output=False)
    x = torch.rand((nev, lib.info.ndim), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    p, w, x1, x2, _ = gen(x)
    g = 1.2177157847767195
    defs = lib.coupling_defs
    coup = torch.tensor([[complex(re, im) * g**k] for re, im, k in defs], dtype=torch.complex128, device="cuda")
    out = torch.empty(nev, dtype=torch.float64, device="cuda")
    par = [173.0, 1.4915000200271606]
    sqh = 0.7071067690849304
    for _ in range(2):
        lib.smatrix(p, 0, nev, par, coup, 0, sqh, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        lib.smatrix(p, 0, nev, par, coup, 0, sqh, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{os.path.basename(path):50s} {lib.variant:6s} block={lib.info.block_threads:4d} {nev / ms * 1e3:12.4g} ev/s  "
          f"{ms:9.3f} ms  checksum {float(out.sum()):.12e}", flush=True)
