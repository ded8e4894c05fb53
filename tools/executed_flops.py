#!/usr/bin/env python3
"""Executed FP64 operations of the dominant kernel from an ncu --set full report:
    python tools/executed_flops.py <report.ncu-rep> <matrix-element events of the profiled launch>
CUDA-core part = 2 x DFMA + DADD + DMUL thread instructions (smsp__sass_thread_inst_executed_op_*_pred_on.sum.per_cycle_elapsed
x sm__cycles_elapsed.avg), tensor part = sm__ops_path_tensor_src_fp64.sum (512 flop per DMMA.8x8x4 warp instruction)."""
import csv
import json
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
val = {h: v for h, v in zip(rows[0], rows[2])}


def f(name):
    return float(val[name].replace(",", ""))


nev = float(sys.argv[2])
cyc = f("sm__cycles_elapsed.avg")
dfma = f("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") * cyc
dadd = f("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") * cyc
dmul = f("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") * cyc
tensor = f("sm__ops_path_tensor_src_fp64.sum") if "sm__ops_path_tensor_src_fp64.sum" in val else 0.0   # flop of the DMMAs
dmma = tensor / 512.0
cuda = 2 * dfma + dadd + dmul
res = {
    "flops_per_me_event": round((cuda + tensor) / nev), "cuda_core_flops_per_me_event": round(cuda / nev),
    "tensor_core_flops_per_me_event": round(tensor / nev),
    "fp64_pipe_active_pct": round(f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
    "dmma_pipe_active_pct": round(f("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active"), 1),
    "dram_bytes_per_launch": round(f("dram__bytes_read.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[dict(zip(rows[0], rows[1]))["dram__bytes_read.sum"]]
                                   + f("dram__bytes_write.sum") * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[dict(zip(rows[0], rows[1]))["dram__bytes_write.sum"]]),
    "kernel_ms_under_ncu": round(f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}[dict(zip(rows[0], rows[1]))["gpu__time_duration.sum"]], 3),
    "thread_inst_per_cycle": {"dfma": round(dfma / cyc, 1), "dadd": round(dadd / cyc, 1), "dmul": round(dmul / cyc, 1)},
    "cycles": round(cyc), "dmma_warp_inst": round(dmma), "me_events": int(nev),
}
print(json.dumps(res, indent=1))
