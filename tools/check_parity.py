#!/usr/bin/env python3
"""Parity of process libraries against the oracle on a few RAMBO points (GPU box):
    python tools/check_parity.py <k final-state gluons> <nevents> <lib.so> [<lib.so> ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import MT, WT, sm_params  # noqa: E402
from madflow_b200 import _runtime as rt  # noqa: E402
from madflow_b200 import procgen  # noqa: E402
from oracle import SQH_REF  # noqa: E402
from oracle import matrix as omatrix  # noqa: E402
from oracle import phasespace as ops  # noqa: E402

k, nev = int(sys.argv[1]), int(sys.argv[2])
ir = procgen.generate_ir(k)
n = ir["nexternal"]
x = np.random.default_rng(5).random((nev, 4 * (n - 2) + 2))
p, w, x1, x2 = ops.ramboflow(x, n, 13e3, [MT, MT] + [0.0] * (n - 4), xfactor="converged")
p = ops.boost_to_lab(p, x1, x2)
a_s = 0.09 + 0.05 * np.random.default_rng(3).random(nev)
params = sm_params(alpha_s=a_s)
coup = np.stack([params[c] for c in ir["couplings"]])
ref = omatrix.smatrix(ir, p, params)
row = min(77, ir["ncomb"] - 1)
ref_row = omatrix.matrix(ir, p, ir["helicities"][row], params)
d_p = torch.as_tensor(p).cuda().contiguous()
d_c = torch.as_tensor(coup).cuda().contiguous()
for path in sys.argv[3:]:
    lib = rt.ProcessLib(path)
    out = torch.empty(nev, dtype=torch.float64, device="cuda")
    lib.smatrix(d_p, 0, nev, [MT, WT], d_c, 1, SQH_REF, out)
    one = torch.empty(nev, dtype=torch.float64, device="cuda")
    lib.smatrix(d_p, 0, nev, [MT, WT], d_c, 1, SQH_REF, one, only_comb=row)
    torch.cuda.synchronize()
    print(f"{os.path.basename(path):40s} smatrix max rel err {np.max(np.abs(out.cpu().numpy() / ref - 1)):.2e}   "
          f"helicity row {row}: {np.max(np.abs(one.cpu().numpy() / ref_row - 1)):.2e}", flush=True)
