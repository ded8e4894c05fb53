"""Phase space -- same functions and class as python_package/madflow/phasespace.py
(rambo :145, ramboflow :257, _boost_to_lab :322, PhaseSpaceGenerator :359), evaluated by the FP64
device functions of csrc/phasespace.cuh through the C ABI (mf_rambo, mf_phasespace,
mf_boost_to_lab).  Momenta are (nevents, nparticles, 4) as (E,px,py,pz); tensors are torch CUDA.

Differences from the reference (DESIGN.md "Phase space"): the massive rescaling factor is iterated
per event until converged instead of until the first event of the batch converges
(phasespace.py:90-92); cuts support a named particle only (`particle=None` has no well-defined
meaning in the reference either, its mask would be two-dimensional).
"""
import ctypes
import logging

import numpy as np
import math

import torch

from . import _runtime as rt
from . import config

logger = logging.getLogger(__name__)
PI = config.get_constants().PI
ACC = config.get_constants().ACC


def _psconst():
    k = config.get_constants()
    return rt.mf_ps_const(k.PI, k.ACC, k.GEV2PB)


def _masses_arr(masses, n):
    vals = [0.0] * rt.MFP_MAX_OUT
    if masses is not None:
        for i, m in enumerate(masses):
            vals[i] = float(m.item() if isinstance(m, torch.Tensor) else m)
    return (ctypes.c_double * rt.MFP_MAX_OUT)(*vals)


def _fourdot(f1, f2):
    """phasespace.py:26-30"""
    return f1[..., 0] * f2[..., 0] - torch.sum(f1[..., 1:] * f2[..., 1:], dim=-1)


def _invariant_mass(fm):
    """phasespace.py:33-35 (squared invariant mass, as in the reference)"""
    return _fourdot(fm, fm)


def rambo(xrand, n_particles, sqrts, masses=None, check_physical=False):
    """RAMBO (phasespace.py:145-212): xrand (nevents, 4*n) -> momenta (nevents, n, 4), weights."""
    x = rt.to_device(xrand)
    nevt = x.shape[0]
    if x.ndim != 2 or x.shape[1] != 4 * n_particles:
        raise ValueError(f"xrand must have shape (nevents, {4 * n_particles})")
    if masses is not None and float(np.sum([float(m) for m in masses])) == 0.0:
        masses = None
    d_sqrts = None
    s_val = 0.0
    if isinstance(sqrts, (float, int)):
        s_val = float(sqrts)
        if check_physical and masses is not None and sum(float(m) for m in masses) > s_val:
            raise ValueError(f"Not enough energy ({sqrts}) to generate particles of mass: {masses}")
    else:
        d_sqrts = rt.to_device(sqrts).reshape(-1)
        if d_sqrts.numel() == 1:
            s_val, d_sqrts = float(d_sqrts.item()), None
        elif d_sqrts.numel() != nevt:
            raise ValueError("per-event sqrts must have one value per event")
    p = torch.empty((nevt, n_particles, 4), dtype=torch.float64, device=x.device)
    w = torch.empty(nevt, dtype=torch.float64, device=x.device)
    k = _psconst()
    lib = rt.core()
    rc = lib.mf_rambo(int(n_particles), rt.ptr(x), ctypes.c_int64(nevt), ctypes.c_double(s_val), rt.ptr(d_sqrts),
                      _masses_arr(masses, n_particles) if masses is not None else None, ctypes.byref(k),
                      rt.ptr(p), rt.ptr(w), rt.stream_ptr())
    rt.check(lib, rc)
    return p, w


def _phasespace(xrand, nparticles, com_sqrts, masses, cuts, lab):
    x = rt.to_device(xrand)
    nevt = x.shape[0]
    ndim = 2 if nparticles == 3 else 4 * (nparticles - 2) + 2
    if x.ndim != 2 or x.shape[1] < ndim:
        raise ValueError(f"xrand must have shape (nevents, {ndim})")
    if x.shape[1] != ndim:
        x = x[:, :ndim].contiguous()
    dev = x.device
    p = torch.empty((nevt, nparticles, 4), dtype=torch.float64, device=dev)
    w = torch.empty(nevt, dtype=torch.float64, device=dev)
    x1 = torch.empty(nevt, dtype=torch.float64, device=dev)
    x2 = torch.empty(nevt, dtype=torch.float64, device=dev)
    ok = torch.empty(nevt, dtype=torch.uint8, device=dev) if cuts else None
    carr = (rt.mf_cut * max(len(cuts), 1))()
    for i, (var, particle, lo, hi) in enumerate(cuts):
        carr[i] = rt.mf_cut(rt.CUT_VARS[var], rt.cut_particle(var, particle), lo is not None, hi is not None,
                            float(lo) if lo is not None else 0.0, float(hi) if hi is not None else 0.0)
    k = _psconst()
    lib = rt.core()
    rc = lib.mf_phasespace(int(nparticles), rt.ptr(x), ctypes.c_int64(nevt), ctypes.c_double(float(com_sqrts)),
                           _masses_arr(masses, nparticles - 2), ctypes.byref(k), carr, len(cuts), int(lab),
                           rt.ptr(p), rt.ptr(w), rt.ptr(x1), rt.ptr(x2), rt.ptr(ok), rt.stream_ptr())
    rt.check(lib, rc)
    return p, w, x1, x2, ok


def ramboflow(xrand, nparticles, com_sqrts, masses=None):
    """phasespace.py:257-319: xrand (nevents, 4*(nparticles-2)+2) -> p (nevents, nparticles, 4) in the
    partonic centre-of-mass frame, wgt, x1, x2."""
    p, w, x1, x2, _ = _phasespace(xrand, nparticles, com_sqrts, masses, [], False)
    return p, w, x1, x2


def _boost_to_lab(p_com, x1, x2):
    """phasespace.py:322-356 (returns a new tensor)."""
    p = rt.to_device(p_com).clone()
    a, b = rt.to_device(x1), rt.to_device(x2)
    lib = rt.core()
    rc = lib.mf_boost_to_lab(int(p.shape[1]), rt.ptr(p), rt.ptr(a), rt.ptr(b), ctypes.c_int64(p.shape[0]),
                             rt.stream_ptr())
    rt.check(lib, rc)
    return p


class PhaseSpaceGenerator:
    """Phase space generator with registrable cuts (phasespace.py:359-520)."""

    def __init__(self, nparticles, com_sqrts, masses=None, com_output=True, algorithm="ramboflow"):
        if masses is None:
            masses = [0.0] * (nparticles - 2)
        if len(masses) != (nparticles - 2):
            raise ValueError(
                "Missmatch in PhaseSpaceGenerator between particles and masses"
                f" {len(masses)} given for {nparticles-2} outgoing particles"
            )
        self._sqrts = float(com_sqrts)
        self._masses = [float(m.item() if isinstance(m, torch.Tensor) else m) for m in masses]
        self._nparticles = nparticles
        self._cuts = []
        self._cuts_info = []
        self._com_output = com_output
        if algorithm != "ramboflow":
            raise ValueError(f"PS algorithm {algorithm} not understood")

    def clear_cuts(self):
        """Clear all cuts"""
        self._cuts = []
        self._cuts_info = []

    @staticmethod
    def mt2(ps_point):
        """Transverse mass squared of the given ps point (nevents, 4) (phasespace.py:405-410)"""
        pt2 = PhaseSpaceGenerator.pt(ps_point) ** 2
        return _invariant_mass(ps_point) + pt2

    @staticmethod
    def mt(ps_point):
        """Transverse mass of the given ps point (phasespace.py:412-415)"""
        return torch.sqrt(PhaseSpaceGenerator.mt2(ps_point))

    @staticmethod
    def pt(ps_point):
        """pt of the ps point (nevents, [:], 4) (phasespace.py:417-422)"""
        return torch.sqrt(ps_point[..., 1] ** 2 + ps_point[..., 2] ** 2)

    @staticmethod
    def mij(ps_a, ps_b):
        """Invariant mass of a pair of ps points (nevents, 4) -- extension, see register_cut."""
        s = ps_a + ps_b
        return torch.sqrt(torch.clamp(_invariant_mass(s), min=0.0))

    @staticmethod
    def dr(ps_a, ps_b):
        """Delta R = sqrt(d eta^2 + d phi^2) of a pair of ps points (nevents, 4) -- extension, see register_cut."""
        def eta(p):
            pabs = torch.sqrt(p[..., 1] ** 2 + p[..., 2] ** 2 + p[..., 3] ** 2)
            return 0.5 * torch.log((pabs + p[..., 3]) / (pabs - p[..., 3]))
        dphi = torch.abs(torch.atan2(ps_a[..., 2], ps_a[..., 1]) - torch.atan2(ps_b[..., 2], ps_b[..., 1]))
        dphi = torch.where(dphi > math.pi, 2.0 * math.pi - dphi, dphi)
        return torch.sqrt((eta(ps_a) - eta(ps_b)) ** 2 + dphi**2)

    def register_cut(self, variable, particle=None, min_val=None, max_val=None):
        """Register min_val < variable(particle) < max_val (phasespace.py:424-478).

        Extension: the pair variables "mij" (invariant mass) and "dr" (Delta R) take particle=(i, j).  The
        reference only cuts on single particles, which leaves the collinear singularity between final-state
        gluons of g g > t t~ g g (g) unregulated (its cross section is not finite with pt cuts alone)."""
        if not hasattr(self, variable) or variable not in rt.CUT_VARS:
            raise ValueError(f"{variable} is not implemented")
        if variable in rt.PAIR_CUTS:
            if (not isinstance(particle, (tuple, list)) or len(particle) != 2 or max(particle) >= self._nparticles
                    or min(particle) < 0):
                raise ValueError(f"{variable} cuts need particle=(i, j) with valid indices")
            particle = (int(particle[0]), int(particle[1]))
        elif particle is not None and particle >= self._nparticles:
            raise ValueError(f"Cannot apply cuts to particle {particle}, python idx starts at 0!")
        if particle is None:
            raise ValueError("madflow_b200 cuts need the `particle` argument")
        if min_val is None and max_val is None:
            logger.warning(f"Cut for {variable} has no min or max val, ignoring")
            return
        if len(self._cuts) >= rt.MFP_MAX_CUTS:
            raise ValueError(f"at most {rt.MFP_MAX_CUTS} cuts are supported")
        self._cuts.append((variable, particle if variable in rt.PAIR_CUTS else int(particle), min_val, max_val))
        self._cuts_info.append(f"{min_val} < {variable}({particle}) < {max_val}")

    @property
    def cuts(self):
        return list(self._cuts)

    def __call__(self, xrand):
        """phasespace.py:480-520: -> ps, wgt, x1, x2, idx.  With cuts only the passing events are
        returned and idx (npass,1) holds their positions; without cuts idx is the scalar 1."""
        p, w, x1, x2, ok = _phasespace(xrand, self._nparticles, self._sqrts, self._masses, self._cuts,
                                       not self._com_output)
        if self._cuts:
            mask = ok.bool()
            idx = torch.nonzero(mask).to(config.DTYPEINT)
            p, w, x1, x2 = p[mask], w[mask], x1[mask], x2[mask]
        else:
            idx = torch.tensor(1, dtype=config.DTYPEINT, device=p.device)
        return p, w, x1, x2, idx
