#!/usr/bin/env python3
"""Benchmark of the madflow hot path (matrix element + RAMBO + VEGAS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--process NAME] [--impl ours|reference]

One "step" = one VEGAS iteration of the fused integrand kernel over `--events` generated events
per GPU (Philox -> VEGAS map -> RAMBO -> cuts -> boost -> alpha_s -> smatrix -> accumulate), then
the deterministic reduction, the single all-reduce (N > 1) and the grid refinement.  Prints ONE
JSON line (rank 0); see DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MT = 173.0
METRIC = "fp64_matrix_element_events_per_sec"
UNIT = "events/s"

# per-process workload: BASELINE.json configs (sqrts 13 TeV, m_t 173, pt > 30 GeV on every outgoing leg)
WORKLOADS = {
    "1_gg_ttx": dict(label="g g > t t~ LO, --no_pdf, RAMBO + VEGAS, pt>30, alpha_s frozen 0.118", masses=[MT, MT],
                     pt_cut=30.0, running=False, events=1 << 26, e2e_events=1 << 22),
    "1_gg_ttxg": dict(label="g g > t t~ g LO, --no_pdf, pt>30 cuts, alpha_s frozen", masses=[MT, MT, 0.0],
                      pt_cut=30.0, running=False, events=1 << 24, e2e_events=1 << 21),
    "1_gg_ttxgg": dict(label="g g > t t~ g g LO, --no_pdf, pt>30 cuts, running g_s (one-loop alpha_s at (sum mT/2)^2)",
                       masses=[MT, MT, 0.0, 0.0], pt_cut=30.0, running=True, events=1 << 22, e2e_events=1 << 20),
    "1_gg_ttxggg": dict(label="g g > t t~ g g g LO, --no_pdf, pt>30 cuts, running g_s", masses=[MT, MT, 0.0, 0.0, 0.0],
                        pt_cut=30.0, running=True, events=1 << 18, e2e_events=1 << 16),
}
# the light-quark subprocesses of p p > t t~ (j): same phase space and cuts as their gluon-fusion counterparts
WORKLOADS["1_uux_ttx"] = dict(WORKLOADS["1_gg_ttx"], label="u u~ > t t~ LO, --no_pdf, RAMBO + VEGAS, pt>30, alpha_s frozen 0.118")
for _name, _proc in (("1_uux_ttxg", "u u~ > t t~ g"), ("1_gu_ttxu", "g u > t t~ u"), ("1_gux_ttxux", "g u~ > t t~ u~")):
    WORKLOADS[_name] = dict(WORKLOADS["1_gg_ttxg"], label=f"{_proc} LO, --no_pdf, pt>30 cuts, alpha_s frozen")
PREFERENCE = ["1_gg_ttxgg", "1_gg_ttxg", "1_gg_ttx"]


def default_process():
    libdir = os.path.join(ROOT, "madflow_b200", "lib")
    for name in PREFERENCE:
        if os.path.exists(os.path.join(libdir, f"libmfp_{name}.so")):
            return name
    return "1_gg_ttx"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


def cpu_integrand(proc, wl):
    from madflow_b200 import integrand as mfi
    from madflow_b200 import matrix as mfm
    from oracle.cpu_bench import CpuIntegrand

    ir = mfm.load_ir(proc)
    if ir is None:
        from madflow_b200 import process_ir

        ir = process_ir.gg_ttx_pinned()
    return CpuIntegrand(ir, 13e3, wl["masses"], wl["pt_cut"], True, wl["running"], alpha_s=0.118,
                        b0=mfi.one_loop_b0(), mz2=mfi.MZ**2, pdf_spec=wl.get("pdf"), pdf_dir=wl.get("pdf_dir"))


def time_cpu(cpu, target_s=12.0):
    """Size a sample for ~target_s of wall time, run it, return (ME events/s, n, n_me, seconds)."""
    cpu.step(8 * cpu.cores, iteration=999)  # start the workers (imports) outside any timing
    probe, dt = 64 * cpu.cores, 0.0
    for _ in range(8):
        n_me, dt = cpu.step(probe, iteration=1000)
        if dt >= 1.0:
            break
        probe *= 4
    rate = probe / max(dt, 1e-6)
    n = int(max(probe, min(rate * target_s, 2e8)))
    n_me, dt = cpu.step(n, iteration=1001)
    return n_me / dt, n, n_me, dt


def run_reference(args, wl, proc):
    """--impl reference: the oracle restatement of the reference's CPU path on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = cpu_integrand(proc, wl)
    _, n, _, _ = time_cpu(cpu, target_s=6.0)
    for i in range(args.warmup):
        cpu.step(max(n // 4, 64), iteration=i)
    tot_me, tot_t = 0, 0.0
    for i in range(args.steps):
        n_me, dt = cpu.step(n, iteration=100 + i)
        tot_me += n_me
        tot_t += dt
    cpu.close()
    value = tot_me / tot_t
    sample = f"{n} generated events per step ({tot_me // max(args.steps, 1)} reach the matrix element), Philox points, uniform grid"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "process": proc,
                   "note": "oracle/ numpy restatement of the reference's TF graph (TensorFlow/vegasflow/pdfflow not "
                           "installable offline), op-by-op vectorised over events, one process per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--process", default=None)
    ap.add_argument("--events", type=int, default=None, help="generated events per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="default", choices=["default", "thread", "hp"])
    ap.add_argument("--pdf", default=None, help="LHAPDF set (member 0) for the parton luminosity and alpha_s, as madflow "
                    "without --no_pdf; needs the set on disk (--pdf_dir / LHAPDF_DATA_PATH).  Default: --no_pdf")
    ap.add_argument("--pdf_dir", default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    proc = args.process or default_process()
    wl = dict(WORKLOADS[proc])
    if args.events:
        wl["events"] = args.events
    if args.pdf:   # BASELINE config 2 names the PDF; no grid exists offline, so this is opt-in
        wl["pdf"], wl["pdf_dir"] = args.pdf + "/0", args.pdf_dir
        wl["running"] = True
        wl["label"] = (wl["label"].split(", --no_pdf")[0] + f", PDF {args.pdf} (luminosity + alpha_s of the set at "
                       "q2 = (sum mT/2)^2), pt>30 cuts")

    if args.impl == "reference":
        run_reference(args, wl, proc)
        return

    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU implementation of the product path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import ctypes

    from madflow_b200 import _runtime as rt
    from madflow_b200 import integrand as mfi
    from madflow_b200 import matrix as mfm
    from madflow_b200 import phasespace as mfps
    from madflow_b200 import vegas as mfv

    m, model = mfm.get_process(proc)
    m.set_variant(args.variant)
    pdf = None
    if wl.get("pdf"):
        from madflow_b200.pdf import mkPDF

        pdf = mkPDF(wl["pdf"], dirname=wl["pdf_dir"])
    fi = mfi.FusedIntegrand(m, model, sqrts=13e3, masses=wl["masses"], pt_cut=wl["pt_cut"], lab_frame=True,
                            running=wl["running"], pdf=pdf)
    n_per_gpu = wl["events"]
    vegas = mfv.VegasFlow(fi.n_dim, n_per_gpu * world, seed=4)
    vegas.compile(fi)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 FMA peak of this device, measured in this run (MEASURED_PEAKS.json has no FP64 entry)
    lib = rt.core()
    tf_, ms_ = ctypes.c_double(), ctypes.c_double()
    rt.check(lib, lib.mf_fp64_peak(20000, ctypes.byref(tf_), ctypes.byref(ms_)))
    fp64_burst = tf_.value

    for _ in range(args.warmup):
        vegas.run_iteration()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    n_me = 0
    results = []
    for _ in range(args.steps):
        results.append(vegas.run_iteration())
        n_me += vegas.last_me_events
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n_me / (ms * 1e-3)
    generated = n_per_gpu * world * args.steps

    # dominant kernel alone: K launches of the fused integrand kernel on the launching stream
    nblocks = fi.nblocks()
    partial = vegas._partial_buf(nblocks)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    first, count = mfv.shard_events(n_per_gpu * world, rank, world)
    fi.launch(vegas.divisions, 4, 9999, first, count, 1.0 / (n_per_gpu * world), partial, nblocks, True)
    torch.cuda.synchronize()
    k0.record()
    for i in range(args.steps):
        fi.launch(vegas.divisions, 4, 10000 + i, first, count, 1.0 / (n_per_gpu * world), partial, nblocks, True)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / args.steps
    me_per_launch = n_me / args.steps / world
    flops = m.flops_per_event
    achieved = flops * me_per_launch / (kernel_ms * 1e-3) / 1e12
    # sustained FP64 probe right after the long kernels (clocks settle under load)
    rt.check(lib, lib.mf_fp64_peak(200000, ctypes.byref(tf_), ctypes.byref(ms_)))
    fp64_sustained = tf_.value

    # end to end through the public API with HOST buffers: pinned momenta -> H2D -> smatrix -> D2H
    n_e2e = wl["e2e_events"]
    psg = mfps.PhaseSpaceGenerator(int(m.nexternal), 13e3, wl["masses"], com_output=False)
    for i in range(2, int(m.nexternal)):
        psg.register_cut("pt", particle=i, min_val=wl["pt_cut"])
    xr = torch.rand((int(n_e2e * 2.5), fi.n_dim), dtype=torch.float64, device="cuda")
    ps, wts, x1, x2, idx = psg(xr)
    ps = ps[:n_e2e].contiguous()
    n_e2e = ps.shape[0]
    if wl["running"]:
        mt_sum = torch.sum(psg.mt(ps[:, 2:, :]), dim=-1)
        alpha = mfi.alpha_s_one_loop((mt_sum / 2.0) ** 2, 0.118, fi.mz2, fi.b0)
        params = model.evaluate(alpha)
    else:
        if not model.frozen:
            model.freeze_alpha_s(0.118)
        params = model.evaluate(None)
    npar = len(m.param_names)
    h_ps = ps.cpu().pin_memory()
    h_coup = [c.cpu().pin_memory() if c.numel() > 1 else c for c in params[npar:]]
    h_out = torch.empty(n_e2e, dtype=torch.float64).pin_memory()
    del xr, ps, wts, x1, x2, idx

    def e2e_step():
        d_ps = h_ps.to("cuda", non_blocking=True)
        d_c = [c.to("cuda", non_blocking=True) for c in h_coup]
        out = m.smatrix(d_ps, *params[:npar], *d_c)
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(h_out[0])

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_e2e * world * args.steps / float(t.item())
    h2d = h_ps.numel() * 8 + sum(c.numel() * 16 for c in h_coup if c.numel() > 1)
    d2h = n_e2e * 8

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_integrand(proc, wl)
        v, n, nme, dt = time_cpu(cpu, target_s=12.0)
        cpu.close()
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cpu.cores, "kind": "port",
                        "sample": f"{n} generated events of the same workload ({nme} reach the matrix element), "
                                  f"{dt:.1f} s wall; oracle/ numpy restatement, one process per core"}

    peaks, executed = {}, {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    try:  # executed FP64 instruction counts of the dominant kernel, from the committed ncu captures
        executed = json.load(open(os.path.join(ROOT, "profiles", "executed_flops.json"))).get(proc, {})
    except OSError:
        pass
    exec_flops = executed.get("flops_per_me_event")
    exec_tflops = exec_flops * me_per_launch / (kernel_ms * 1e-3) / 1e12 if exec_flops else None
    final, err, chi2 = mfv.combine_iterations(results)
    # kernels of this package inside the timed region, per step: the integrand (1 fused kernel, or generate +
    # matrix element + accumulate for the helicity-parallel pipeline) and the block-partial reduction per launch
    # chunk, plus the grid refinement
    chunks = -(-n_per_gpu // min(n_per_gpu, fi.max_events_per_launch))
    gpu_launches = args.steps * (chunks * ((3 if m.variant == "hp" else 1) + 1) + 1)
    bytes_per_event = 0.0  # the fused kernel reads no per-event input from HBM
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": wl["label"], "process": proc, "events_generated_per_gpu_per_step": n_per_gpu,
            "events_reaching_matrix_element_per_step": n_me // args.steps,
            "generated_events_per_sec": generated / (ms * 1e-3),
            "l2": "inputs are generated in-kernel from Philox counters (no per-event HBM input); e2e inputs "
                  f"{h2d / 2**20:.0f} MiB per step exceed the 126 MB L2",
            "sigma_pb": final, "sigma_err_pb": err, "constants": "reference", "kernel_variant": m.variant,
        },
        "roofline": {
            "bound": "fp64", "achieved": achieved, "peak": fp64_sustained, "unit": "TFLOP/s",
            "frac": achieved / fp64_sustained, "traffic": executed.get("dram_bytes_per_launch"),
            "traffic_source": executed.get("dram_source"),
            "kernel": "integrand_kernel_hp<Proc>" if m.variant == "hp" else "integrand_kernel<Proc>", "kernel_ms": kernel_ms,
            "flops_per_event_algorithmic": flops,
            "peak_source": "mf_fp64_peak DFMA probe in this run (sustained, after the timed kernels); burst "
                           f"{fp64_burst:.1f} TFLOP/s; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2; "
                           "MEASURED_PEAKS.json has no FP64 entry",
            "hbm_gbs_measured": peaks.get("hbm_gbs"), "hbm_bytes_per_event": bytes_per_event,
            "note": "achieved = F_alg (the reference's operation count over all helicities, SURVEY 8d) x events / time; "
                    "the kernel executes fewer operations (wavefunctions and vertices are evaluated once per helicity "
                    "variant of their own legs), so this fraction can exceed 1; the executed figures below are the "
                    "hardware-side view",
            "executed_flops_per_event": exec_flops, "executed_tflops": exec_tflops,
            "executed_frac": exec_tflops / fp64_sustained if exec_tflops else None,
            "executed_source": executed.get("source"),
        },
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "Matrix.smatrix on pinned host momenta (+ per-event couplings): H2D, fused kernel, D2H of |M|^2",
                "events_per_step": n_e2e},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
