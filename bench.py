#!/usr/bin/env python3
"""Benchmark of the madflow hot path (matrix element + RAMBO + VEGAS) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--process NAME] [--impl ours|reference]

One "step" = one VEGAS iteration of the device-resident integrand over `--events` generated events IN TOTAL
(Philox -> VEGAS map -> RAMBO -> cuts -> boost -> alpha_s -> smatrix -> accumulate), sharded over the N GPUs
(strong scaling: BASELINE config 3 is "1e8 events/iteration at 1/2/4/8 B200") and launched in chunks of at most
`max_events_per_launch` events, then the deterministic reduction, the single all-reduce (N > 1) and the grid
refinement.  Prints ONE JSON line (rank 0); see DESIGN.md "Measurement" for every field.  The line of the default
process also carries `other_configs`: BASELINE configs 1, 2 and 4 measured the same way in the same run.
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
# `events` = generated events per iteration IN TOTAL (all GPUs together)
WORKLOADS = {
    "1_gg_ttx": dict(label="g g > t t~ LO, --no_pdf, RAMBO + VEGAS, pt>30, alpha_s frozen 0.118", masses=[MT, MT],
                     pt_cut=30.0, running=False, events=1 << 27, e2e_events=1 << 22, config=1),
    "1_gg_ttxg": dict(label="g g > t t~ g LO, --no_pdf, pt>30 cuts, alpha_s frozen", masses=[MT, MT, 0.0],
                      pt_cut=30.0, running=False, events=1 << 25, e2e_events=1 << 21, config=2),
    "1_gg_ttxgg": dict(label="g g > t t~ g g LO, --no_pdf, pt>30 cuts, running g_s (one-loop alpha_s at (sum mT/2)^2), "
                             "1e8 events/iteration", masses=[MT, MT, 0.0, 0.0], pt_cut=30.0, running=True,
                       events=100_000_000, e2e_events=1 << 22, config=3),
    "1_gg_ttxggg": dict(label="g g > t t~ g g g LO, --no_pdf, pt>30 cuts, running g_s", masses=[MT, MT, 0.0, 0.0, 0.0],
                        pt_cut=30.0, running=True, events=1 << 21, e2e_events=1 << 16, config=4),
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
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "process": proc,
                   "note": "oracle/ numpy restatement of the reference's TF graph (TensorFlow/vegasflow/pdfflow not "
                           "installable offline), op-by-op vectorised over events, one process per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def measure_iterations(vegas, steps, warmup, barrier, dist, torch, sampler=None):
    """W untimed + K timed VEGAS iterations: (ME events, ms = max over ranks, clocks)."""
    for _ in range(warmup):
        vegas.run_iteration()
    barrier()
    if sampler is not None:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    n_me = 0
    for _ in range(steps):
        vegas.run_iteration()
        n_me += vegas.last_me_events
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler is not None else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return n_me, float(t.item()), clocks


def time_kernel_alone(fi, vegas, n_total, rank, world, reps, torch):
    """The integrand kernels of ONE launch chunk, alone on the launching stream: (ms per launch, events of the chunk)."""
    from madflow_b200 import vegas as mfv

    nblocks = fi.nblocks()
    partial = vegas._partial_buf(nblocks)
    first, count = mfv.shard_events(n_total, rank, world)
    count = min(count, fi.max_events_per_launch)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fi.launch(vegas.divisions, 4, 9999, first, count, 1.0 / n_total, partial, nblocks, True)
    torch.cuda.synchronize()
    k0.record()
    for i in range(reps):
        fi.launch(vegas.divisions, 4, 10000 + i, first, count, 1.0 / n_total, partial, nblocks, True)
    k1.record()
    torch.cuda.synchronize()
    return k0.elapsed_time(k1) / reps, count


def load_executed(proc):
    try:  # executed FP64 instruction counts of the dominant kernel, from the committed ncu captures
        return json.load(open(os.path.join(ROOT, "profiles", "executed_flops.json"))).get(proc, {})
    except OSError:
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--process", default=None)
    ap.add_argument("--events", type=int, default=None, help="generated events per step in total (all GPUs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs 1, 2, 4 in the default run")
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

    def build(proc_, wl_):
        m_, model_ = mfm.get_process(proc_)
        m_.set_variant(args.variant)
        pdf = None
        if wl_.get("pdf"):
            from madflow_b200.pdf import mkPDF

            pdf = mkPDF(wl_["pdf"], dirname=wl_["pdf_dir"])
        fi_ = mfi.FusedIntegrand(m_, model_, sqrts=13e3, masses=wl_["masses"], pt_cut=wl_["pt_cut"], lab_frame=True,
                                 running=wl_["running"], pdf=pdf)
        v_ = mfv.VegasFlow(fi_.n_dim, wl_["events"], seed=4)
        v_.compile(fi_)
        return m_, model_, fi_, v_

    m, model, fi, vegas = build(proc, wl)
    n_total = wl["events"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # FP64 peak of this device, measured in this run (MEASURED_PEAKS.json has no FP64 entry): the DFMA rate and the rate
    # of the FP64 tensor instruction (they share one pipe); the roofline denominator is the larger of the two
    lib = rt.core()
    tf_, ms_ = ctypes.c_double(), ctypes.c_double()
    rt.check(lib, lib.mf_fp64_peak(20000, ctypes.byref(tf_), ctypes.byref(ms_)))
    fp64_burst = tf_.value

    n_me, ms, clocks = measure_iterations(vegas, args.steps, args.warmup, barrier, dist, torch,
                                          ClockSampler(local) if rank == 0 else None)
    value = n_me / (ms * 1e-3)
    generated = n_total * args.steps

    # dominant kernel alone: launches of the integrand kernels over one launch chunk on the launching stream
    kernel_ms, chunk_events = time_kernel_alone(fi, vegas, n_total, rank, world, min(args.steps, 3), torch)
    me_per_launch = chunk_events * (n_me / generated)
    flops = m.flops_per_event
    achieved = flops * me_per_launch / (kernel_ms * 1e-3) / 1e12
    # sustained FP64 probes right after the long kernels (clocks settle under load)
    rt.check(lib, lib.mf_fp64_peak(200000, ctypes.byref(tf_), ctypes.byref(ms_)))
    dfma_sustained = tf_.value
    rt.check(lib, lib.mf_dmma_peak(100000, ctypes.byref(tf_), ctypes.byref(ms_)))
    dmma_sustained = tf_.value
    fp64_peak = max(dfma_sustained, dmma_sustained)

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
        # the public host-buffer call: chunks over two streams, copies overlapped with the kernel, result on the host
        m.smatrix_pinned(h_ps, *params[:npar], *h_coup, out=h_out)
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
    del h_ps, h_out

    # BASELINE configs 1, 2 and 4 the same way (all ranks take part: the iterations all-reduce)
    others = []
    if args.process is None and not args.no_other_configs:
        for oproc in ("1_gg_ttx", "1_gg_ttxg", "1_gg_ttxggg"):
            if not os.path.exists(os.path.join(ROOT, "madflow_b200", "lib", f"libmfp_{oproc}.so")):
                continue
            owl = dict(WORKLOADS[oproc])
            om_, _, ofi, ov = build(oproc, owl)
            o_me, o_ms, _ = measure_iterations(ov, 3, 2, barrier, dist, torch)
            o_kms, o_chunk = time_kernel_alone(ofi, ov, owl["events"], rank, world, 2, torch)
            o_exec = load_executed(oproc)
            o_mel = o_chunk * (o_me / (3.0 * owl["events"]))
            o_tf = o_exec.get("flops_per_me_event", 0.0) * o_mel / (o_kms * 1e-3) / 1e12
            others.append({
                "baseline_config": owl["config"], "process": oproc, "workload": owl["label"], "value": o_me / (o_ms * 1e-3),
                "unit": UNIT, "steps": 3, "warmup": 2, "ms_per_step": o_ms / 3, "events_generated_per_step": owl["events"],
                "kernel_variant": om_.variant, "kernel_ms": o_kms,
                "algorithmic_frac": om_.flops_per_event * o_mel / (o_kms * 1e-3) / 1e12 / fp64_peak,
                "executed_flops_per_event": o_exec.get("flops_per_me_event"),
                "executed_frac": o_tf / fp64_peak if o_exec.get("flops_per_me_event") else None})
            del ofi, ov

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

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    executed = load_executed(proc)
    exec_flops = executed.get("flops_per_me_event")
    exec_tflops = exec_flops * me_per_launch / (kernel_ms * 1e-3) / 1e12 if exec_flops else None
    # kernels of this package inside the timed region, per step: the integrand (1 fused kernel, or generate +
    # matrix element + accumulate for the helicity-parallel pipeline) and the block-partial reduction per launch
    # chunk, plus the grid refinement
    per_rank = mfv.shard_events(n_total, 0, world)[1]
    chunks = -(-per_rank // min(per_rank, fi.max_events_per_launch))
    gpu_launches = args.steps * (chunks * ((3 if m.variant == "hp" else 1) + 1) + 1)
    bytes_per_event = 0.0  # the fused kernel reads no per-event input from HBM
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": wl["label"], "process": proc, "events_generated_per_step": n_total,
            "events_generated_per_gpu_per_step": per_rank, "launch_chunks_per_gpu_per_step": chunks,
            "events_reaching_matrix_element_per_step": n_me // args.steps,
            "generated_events_per_sec": generated / (ms * 1e-3),
            "l2": "inputs are generated in-kernel from Philox counters (no per-event HBM input); e2e inputs "
                  f"{h2d / 2**20:.0f} MiB per step exceed the 126 MB L2",
            "constants": "reference", "kernel_variant": m.variant,
            "note": "no cross section is quoted on this line: with the reference's cuts (pt only) the g || g collinear "
                    "region of this process is not regulated and single iterations fluctuate by tens of percent "
                    "(DESIGN.md section 8); integrated-sigma checks against the oracle live in tests/ and in "
                    "profiles/ (Delta R regulated)",
        },
        "roofline": {
            "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak, "traffic": executed.get("dram_bytes_per_launch"),
            "traffic_source": executed.get("dram_source"),
            "kernel": "smatrix_kernel_hp<Proc> (+ ps_generate_kernel, accumulate_kernel)" if m.variant == "hp" else "integrand_kernel<Proc>",
            "kernel_ms": kernel_ms, "events_per_launch": chunk_events, "flops_per_event_algorithmic": flops,
            "peak_source": f"max of the in-run probes after the timed kernels: DFMA {dfma_sustained:.1f}, FP64 tensor "
                           f"instruction (mma.sync.m8n8k4, same pipe) {dmma_sustained:.1f} TFLOP/s; DFMA burst before: "
                           f"{fp64_burst:.1f}; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2; MEASURED_PEAKS.json has "
                           "no FP64 entry",
            "hbm_gbs_measured": peaks.get("hbm_gbs"), "hbm_bytes_per_event": bytes_per_event,
            "note": "achieved = F_alg (the reference's operation count over all helicities and diagrams, SURVEY 8d) x "
                    "events / time; the kernel executes far fewer operations (colour-reduced recursion, every object once "
                    "per helicity variant of its own legs), so this fraction exceeds 1; the executed figures are the "
                    "hardware-side view",
            "executed_flops_per_event": exec_flops, "executed_tflops": exec_tflops,
            "executed_frac": exec_tflops / fp64_peak if exec_tflops else None,
            "executed_source": executed.get("source"),
        },
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "path": "Matrix.smatrix_pinned on pinned host momenta (+ per-event couplings): H2D, fused kernel, D2H of |M|^2, "
                        "chunks of 2^18 events on two streams",
                "events_per_step": n_e2e},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
        "other_configs": others,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
