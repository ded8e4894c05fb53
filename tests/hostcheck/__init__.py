"""TEST INFRASTRUCTURE: build and load CPU executions of the device code (see the .cu headers)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "madflow_b200", "csrc")
BUILD = os.path.join(HERE, "_build")

c_dp = ctypes.POINTER(ctypes.c_double)


def _dp(a):
    return a.ctypes.data_as(c_dp) if a is not None else None


def _nvcc(src, out, defs=()):
    os.makedirs(BUILD, exist_ok=True)
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
           "-I", CSRC, "-o", out, src] + list(defs)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError(r.stdout + r.stderr)
    return out


def core():
    out = os.path.join(BUILD, "libhc_core.so")
    src = os.path.join(HERE, "hostcheck_core.cu")
    newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    if not os.path.exists(out) or os.path.getmtime(out) < max(newest, os.path.getmtime(src)):
        _nvcc(src, out)
    return ctypes.CDLL(out)


_names = None


def _builtin_names():
    global _names
    if _names is None:
        from madflow_b200 import build

        _names = {b["name"] for b in build.builtin_irs()}
    return _names


def process(ir):
    from madflow_b200 import codegen

    os.makedirs(codegen.GENDIR, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    builtin = ir["name"] in _builtin_names()
    # processes that are not compiled into the package keep their emitted source with the other test artefacts
    src = os.path.join(codegen.GENDIR if builtin else BUILD, f"proc_{ir['name']}.cu")
    text = codegen.emit_process_source(ir)
    if not os.path.exists(src) or open(src).read() != text:
        open(src, "w").write(text)
    out = os.path.join(BUILD, f"libhc_{ir['name']}.so")
    newest = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    if not os.path.exists(out) or os.path.getmtime(out) < max(newest, os.path.getmtime(src)):
        _nvcc(os.path.join(HERE, "hostcheck_proc.cu"), out, [f'-DMF_PROC_SOURCE="{src}"'])
    return ctypes.CDLL(out)


def smatrix(lib, ir, p, par, coup, sqh, only_comb=-1, hp=False):
    """p (nevt,n,4); coup (ncoup,) or (ncoup,nevt) complex.  hp=True runs the helicity-parallel flavour."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    nevt = p.shape[0]
    coup = np.ascontiguousarray(np.asarray(coup, dtype=np.complex128))
    stride = 1 if coup.ndim == 2 and coup.shape[1] == nevt and nevt > 1 else 0
    cflat = coup.view(np.float64)
    out = np.empty(nevt)
    par = np.ascontiguousarray(par, dtype=np.float64)
    fn = lib.hostcheck_smatrix_hp if hp else lib.hostcheck_smatrix
    rc = fn(_dp(p), ctypes.c_longlong(nevt), _dp(par), _dp(cflat), ctypes.c_longlong(stride),
                               ctypes.c_double(sqh), ctypes.c_int(only_comb), _dp(out))
    assert rc == 0
    return out
