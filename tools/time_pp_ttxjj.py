#!/usr/bin/env python3
"""`p p > t t~ j j` on the GPU: the twelve subprocess libraries on the same events, luminosity-weighted sum
(madflow_exec.py:141-155, 444-455), with the synthetic lhagrid1 set of oracle/pdf.py (no real grid offline):
    python tools/time_pp_ttxjj.py [events per iteration]
Prints generated events/s of one VEGAS iteration (CUDA events) and, per subprocess, its matrix-element rate alone."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madflow_b200 import integrand, matrix, pdf as mpdf, procgen, vegas
from oracle import pdf as opdf

nev = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = tempfile.mkdtemp()
opdf.write_toy_set(d)
pd = mpdf.mkPDF("ToyPDF/0", dirname=d)
MT = 173.0
masses = [MT, MT, 0.0, 0.0]


def run(label, fi):
    v = vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v.compile(fi)
    v.run_iteration()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        r = v.run_iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:44s} {v.last_me_events / ms * 1e3:12.4g} ME events/s  {nev / ms * 1e3:12.4g} generated/s  {ms:8.2f} ms  "
          f"sigma {r[0]:.6g} +/- {r[1]:.3g}", flush=True)


names = procgen.MULTI_PROCESSES["p p > t t~ j j"]
parts = []
for name in names:
    m, model = matrix.get_process(name)
    fi = integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, running=True, pdf=pd)
    run(f"{name} alone (toy PDF)", fi)
    parts.append(fi)
multi = integrand.MultiProcessIntegrand(parts)
run("p p > t t~ j j (12 subprocesses, toy PDF)", multi)
multi.release()
