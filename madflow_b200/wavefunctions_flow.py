"""HELAS external wavefunctions -- same names and arguments as
python_package/madflow/wavefunctions_flow.py (sxxxxx :33, ixxxxx :55, oxxxxx :88, vxxxxx :119),
evaluated by the FP64 device functions of csrc/helas.cuh through the C ABI (mf_wavefunction).

Inputs: p (nevt,4) as (E,px,py,pz) -- torch CUDA tensor, torch CPU tensor or numpy array;
mass/helicity/state scalars.  Output: torch complex128 CUDA tensor of shape (6,nevt)
((3,nevt) for sxxxxx), the reference's layout.
"""
import ctypes

import torch

from . import _runtime as rt
from . import config

SQH = config.get_constants().SQH


def _scalar(x):
    if isinstance(x, torch.Tensor):
        return x.item()
    return float(x)


def _call(kind, p, mass, nhel, ns, rows):
    p = rt.to_device(p)
    if p.ndim != 2 or p.shape[1] != 4:
        raise ValueError("momenta must have shape (nevents, 4)")
    nevt = p.shape[0]
    out = torch.empty((rows, nevt), dtype=torch.complex128, device=p.device)
    lib = rt.core()
    rc = lib.mf_wavefunction(kind, rt.ptr(p), ctypes.c_int64(nevt), ctypes.c_double(_scalar(mass)),
                             int(_scalar(nhel)), int(_scalar(ns)), ctypes.c_double(config.get_constants().SQH),
                             rt.ptr(out), rt.stream_ptr())
    rt.check(lib, rc)
    return out


def sxxxxx(p, nss):
    """Scalar wavefunction (reference: wavefunctions_flow.py:33-51; its body cannot run)."""
    return _call(3, p, 0.0, 0, nss, 3)


def ixxxxx(p, fmass, nhel, nsf):
    """Incoming-flow fermion wavefunction |fi> (wavefunctions_flow.py:55-85)."""
    return _call(0, p, fmass, nhel, nsf, 6)


def oxxxxx(p, fmass, nhel, nsf):
    """Outgoing-flow fermion wavefunction <fo| (wavefunctions_flow.py:88-116)."""
    return _call(1, p, fmass, nhel, nsf, 6)


def vxxxxx(p, vmass, nhel, nsv):
    """Vector wavefunction; nhel=4 gives the BRST-check polarisation (wavefunctions_flow.py:119-154)."""
    return _call(2, p, vmass, nhel, nsv, 6)
