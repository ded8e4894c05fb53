#!/usr/bin/env python3
"""Cost of the parton luminosity and of the multi-subprocess sum in the fused integrand:
python tools/time_pdf.py [nevents].  Uses the synthetic lhagrid1 set of oracle/pdf.py (no real grid offline).
Prints matrix-element events/s of one VEGAS iteration (CUDA events) with and without the PDF table."""
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madflow_b200 import integrand, matrix, pdf as mpdf, vegas
from oracle import pdf as opdf

nev = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
d = tempfile.mkdtemp()
opdf.write_toy_set(d)
pd = mpdf.mkPDF("ToyPDF/0", dirname=d)
MT = 173.0


def run(label, fi, ndim):
    v = vegas.VegasFlow(ndim, nev, seed=4)
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


# the interpolation kernels alone (PDF.xfxQ2 / .alphasQ2): points/s and the HBM traffic they imply
npt = 1 << 24
g = torch.Generator("cuda").manual_seed(1)
xs = 10 ** (-4.0 * torch.rand(npt, dtype=torch.float64, device="cuda", generator=g))
q2s = 10 ** (2.0 + 5.0 * torch.rand(npt, dtype=torch.float64, device="cuda", generator=g))
for label, fn, nout in (("xfxQ2, 1 flavour", lambda: pd.xfxQ2([21], xs, q2s), 1), ("xfxQ2, 5 flavours", lambda: pd.xfxQ2([21, 1, 2, -1, -2], xs, q2s), 5),
                        ("xfxQ2, 11 flavours", lambda: pd.xfxQ2_allpid(xs, q2s), 11), ("alphasQ2", lambda: pd.alphasQ2(q2s), 1)):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nin = 1 if label == "alphasQ2" else 2
    print(f"{label:44s} {npt / ms * 1e3:12.4g} points/s  {ms:8.3f} ms  {(nin + nout) * 8 * npt / ms / 1e6:8.1f} GB/s of x, q2 in + values out", flush=True)

for name, k, variant in (("1_gg_ttx", 0, "thread"), ("1_gg_ttx", 0, "hp"), ("1_gg_ttxg", 1, "hp")):
    m, model = matrix.get_process(name)
    m.set_variant(variant)
    masses = [MT, MT] + [0.0] * k
    for label, p in (("no_pdf, one-loop alpha_s", None), ("toy PDF luminosity + alpha_s table", pd)):
        fi = integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, running=True, pdf=p)
        run(f"{name} [{variant}] {label}", fi, fi.n_dim)
    m.set_variant("default")
parts = []
for name in ("1_gg_ttx", "1_uux_ttx"):
    m, model = matrix.get_process(name)
    parts.append(integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT], pt_cut=30.0, running=True, pdf=pd))
multi = integrand.MultiProcessIntegrand(parts)
run("p p > t t~ (g g + q q~, toy PDF)", multi, 10)
multi.release()
parts = []
for name in ("1_gg_ttxg", "1_gu_ttxu", "1_gux_ttxux", "1_uux_ttxg"):
    m, model = matrix.get_process(name)
    parts.append(integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT, 0.0], pt_cut=30.0, running=True, pdf=pd))
multi = integrand.MultiProcessIntegrand(parts)
run("p p > t t~ j (4 subprocesses, toy PDF)", multi, 14)
multi.release()
