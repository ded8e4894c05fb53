"""CPU timing of the oracle integrand on all host cores.  TEST/BENCH INFRASTRUCTURE.

Used only by bench.py's `cpu_baseline` leg and `--impl reference` arm: TensorFlow, vegasflow and
pdfflow are not installable offline, so "the reference's CPU implementation" is this package's
restatement of it, executed the way TF-CPU executes the reference -- op by op, vectorised over the
event axis in complex128 -- with one worker process per host core.
"""
import json
import multiprocessing as mp
import os
import time

import numpy as np

_state = {}


def _init(ir_json, sqrts, masses, pt_cut, lab, running, alpha_s, b0, mz2, pdf_spec=None, pdf_dir=None):
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    from oracle import model, vegas

    ir = json.loads(ir_json)
    ir["jamp"] = [[tuple(t) for t in terms] for terms in ir["jamp"]]

    def params(a):
        a = alpha_s if a is None else a
        c = model.sm_qcd_couplings(a)
        return dict(c, mdl_MT=173.0, mdl_WT=1.4915000200271606)

    a_fn = (lambda q2: alpha_s / (1 + alpha_s * b0 * np.log(q2 / mz2))) if running else None
    grid = None
    if pdf_spec:   # luminosity and (when running) alpha_s from the LHAPDF set, as madflow does without --no_pdf
        from oracle import pdf as opdf

        grid = opdf.GridPDF.from_set(pdf_spec, pdf_dir)
        if running and len(grid.as_q2):
            a_fn = grid.alphasQ2
    _state["xs"] = vegas.make_cross_section(ir, params, sqrts, masses, pt_cut=pt_cut, lab_frame=lab, alpha_s_fn=a_fn,
                                            pdf=grid)
    _state["ndim"] = 4 * (ir["nexternal"] - 2) + 2


def _work(job):
    from oracle import philox, vegas

    seed, iteration, first, n = job
    ndim = _state["ndim"]
    u = vegas.confine(philox.uniforms(seed, iteration, first, n, ndim))
    x, k, w = vegas.map_to_grid(u, vegas.uniform_grid(ndim))
    f = _state["xs"](x)
    a, b, c = vegas.accumulate(f, w / n, k)
    return int(np.count_nonzero(f)), float(a)


class CpuIntegrand:
    """Pool of workers evaluating the oracle's cross_section on event chunks."""

    def __init__(self, ir, sqrts, masses, pt_cut, lab, running, alpha_s=0.118, b0=0.0, mz2=1.0, cores=None,
                 pdf_spec=None, pdf_dir=None):
        self.cores = cores or os.cpu_count() or 1
        args = (json.dumps(ir), sqrts, masses, pt_cut, lab, running, alpha_s, b0, mz2, pdf_spec, pdf_dir)
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_init, initargs=args)

    def step(self, n_events, iteration=0, seed=4, chunk=None):
        """Evaluate n_events generated events; returns (ME-evaluated events, seconds)."""
        chunk = chunk or max(1, min(20000, (n_events + self.cores - 1) // self.cores))
        jobs = [(seed, iteration, first, min(chunk, n_events - first)) for first in range(0, n_events, chunk)]
        t0 = time.perf_counter()
        out = self.pool.map(_work, jobs)
        dt = time.perf_counter() - t0
        return sum(o[0] for o in out), dt

    def close(self):
        self.pool.close()
        self.pool.join()
