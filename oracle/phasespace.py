"""RAMBO phase space, x1/x2 sampling, cuts and boost: numpy restatement.  TEST INFRASTRUCTURE.

Follows python_package/madflow/phasespace.py.  Momenta are (nevt, nparticles, 4) as (E,px,py,pz).

Two switches select between the reference's literal behaviour and the batch-independent one the
CUDA path implements:
  const    oracle.REFERENCE (float32-rounded PI, ACC, GeV->pb constant: what the reference's
           float_me(<python float>) produces) or oracle.EXACT;
  xfactor  "reference": phasespace.py:38-105 verbatim -- the Newton loop on the massive rescaling
           factor stops for the WHOLE batch as soon as one event has converged, and the energies
           returned lag one Newton step behind the factor;
           "converged": every event is iterated on its own until f <= ACC (at most 10 steps) and
           the energies are recomputed from the final factor.
"""
import math

import numpy as np

from . import REFERENCE


def fourdot(a, b):
    """phasespace.py:26-30."""
    return a[..., 0] * b[..., 0] - np.sum(a[..., 1:] * b[..., 1:], axis=-1)


def invariant_mass2(p):
    """phasespace.py:33-35 (the reference calls the squared mass `_invariant_mass`)."""
    return fourdot(p, p)


def pt(p):
    """phasespace.py:417-422."""
    return np.sqrt(p[..., 1] ** 2 + p[..., 2] ** 2)


def mt2(p):
    """phasespace.py:405-410."""
    return invariant_mass2(p) + pt(p) ** 2


def mt(p):
    """phasespace.py:412-415."""
    return np.sqrt(mt2(p))


def massive_xfactor(sqrts, masses, e0, const=REFERENCE, mode="reference"):
    """phasespace.py:38-105.  sqrts (nevt,1), masses (n,), e0 (nevt,n) massless energies.
    Returns xfactor (nevt,1), new_E (nevt,n)."""
    total_mass = np.sum(masses)
    e2 = np.square(e0)
    m2 = np.square(masses)
    x = np.sqrt(1 - (total_mass / sqrts) ** 2)
    new_E = e0
    if mode == "reference":
        go = True
        it = 0
        while go and it < 10:
            new_E = np.sqrt(m2 + e2 * x**2)
            f0 = np.sum(new_E, axis=1, keepdims=True) - sqrts
            g0 = np.sum(e2 / new_E, axis=1, keepdims=True)
            nxt = x - f0 / (x * g0)
            not_done = f0 > const.ACC
            x = np.where(not_done, nxt, x)
            go = bool(np.all(not_done))
            it += 1
        return x, new_E
    assert mode == "converged"
    active = np.ones_like(x, dtype=bool)
    for _ in range(10):
        new_E = np.sqrt(m2 + e2 * x**2)
        f0 = np.sum(new_E, axis=1, keepdims=True) - sqrts
        g0 = np.sum(e2 / new_E, axis=1, keepdims=True)
        active = active & (f0 > const.ACC)
        if not active.any():
            break
        x = np.where(active, x - f0 / (x * g0), x)
    new_E = np.sqrt(m2 + e2 * x**2)
    return x, new_E


def gen_unconstrained_momenta(xr, const=REFERENCE):
    """phasespace.py:121-142.  xr (nevt,4)."""
    costh = 2.0 * xr[:, 0] - 1.0
    sinth = np.sqrt(1.0 - costh**2)
    phi = 2 * const.PI * xr[:, 1]
    energy = -1.0 * np.log(xr[:, 2] * xr[:, 3])
    return np.stack([energy, energy * sinth * np.sin(phi), energy * sinth * np.cos(phi), energy * costh], axis=1)


def conformal_transformation(q, bquad):
    """phasespace.py:108-118."""
    bvec = bquad[:, 1:]
    gamma = -bquad[:, 0:1]
    a = 1.0 / (1.0 + gamma)
    bq = np.sum(q[:, 1:] * bvec, axis=1, keepdims=True)
    tmp = bq * a + q[:, 0:1]
    pvec = q[:, 1:] + bvec * tmp
    pnrg = q[:, 0:1] * gamma + bq
    return np.concatenate([pnrg, pvec], axis=1)


def rambo(xrand, n, sqrts, masses=None, const=REFERENCE, xfactor="reference"):
    """phasespace.py:145-212.  xrand (nevt,4n); sqrts float or (nevt,).  -> (nevt,n,4), (nevt,)."""
    xrand = np.asarray(xrand, dtype=np.float64)
    nev = xrand.shape[0]
    sqrts = np.broadcast_to(np.asarray(sqrts, dtype=np.float64), (nev,)).reshape(-1, 1)
    if masses is not None and np.sum(masses) == 0:
        masses = None
    all_q = [gen_unconstrained_momenta(xrand[:, 4 * i : 4 * i + 4], const) for i in range(n)]
    sum_q = np.sum(np.stack(all_q), axis=0)
    sq2 = sum_q**2
    qmass = np.sqrt(sq2[:, 0:1] - np.sum(sq2[:, 1:], axis=1, keepdims=True))
    x = sqrts / qmass
    bquad = -sum_q / qmass
    tmp_p = np.stack([conformal_transformation(q, bquad) for q in all_q], axis=1)
    all_p = tmp_p * x[:, :, None]

    wt = math.log(const.PI / 2.0) * (n - 1)
    wt = wt - 2.0 * math.lgamma(n - 1)
    wt = wt - math.log(n - 1)
    wt = wt + (2 * n - 4) * np.log(sqrts[:, 0])
    norm = np.power(2 * const.PI, 3 * n - 4)
    if masses is None:
        return all_p, np.exp(wt) / norm

    masses = np.asarray(masses, dtype=np.float64)
    xf, new_E = massive_xfactor(sqrts, masses, all_p[:, :, 0], const, xfactor)
    pvec = all_p[:, :, 1:] * xf[:, :, None]
    massive_p = np.concatenate([new_E[:, :, None], pvec], axis=-1)
    v = all_p[:, :, 0] * xf
    wt2 = np.prod(v / new_E, axis=1)
    wt3 = np.sum(v**2 / new_E, axis=1)
    wt = wt + (2 * n - 3) * np.log(xf[:, 0]) + np.log(wt2 / wt3 * sqrts[:, 0])
    return massive_p, np.exp(wt) / norm


def get_x1x2(xarr, shat_min, s_in):
    """phasespace.py:215-233."""
    taumin = shat_min / s_in
    delta = 1.0 - taumin
    tau = xarr[:, 0] * delta + taumin
    x1 = np.power(tau, xarr[:, 1])
    x2 = tau / x1
    wgt = delta * (-1.0 * np.log(tau))
    shat = x1 * x2 * s_in
    return shat, wgt, x1, x2


def get_x1x2_onshell(xr, mass, s_in):
    """phasespace.py:236-254 (2 -> 1)."""
    ratio = mass / np.sqrt(s_in)
    tau_max = np.log(ratio)
    wgt = -2.0 * tau_max / s_in
    tau = tau_max - 2.0 * xr * tau_max
    x1 = ratio * np.exp(tau)
    x2 = ratio * np.exp(-tau)
    oo = np.ones_like(xr)
    return oo * mass**2, oo * wgt, x1, x2


def ramboflow(xrand, nparticles, com_sqrts, masses=None, const=REFERENCE, xfactor="reference"):
    """phasespace.py:257-319.  -> p (nevt,nparticles,4) in the partonic COM frame, wgt, x1, x2."""
    xrand = np.asarray(xrand, dtype=np.float64)
    shat_min = 0.0 if masses is None else float(np.sum(masses) ** 2)
    if nparticles == 3:
        shat, wgt, x1, x2 = get_x1x2_onshell(xrand[:, 0], masses[0], com_sqrts**2)
        roots = np.sqrt(shat)
        zeros = np.zeros_like(x1)
        p_out = np.stack([roots, zeros, zeros, zeros], axis=-1)[:, None, :]
    else:
        shat, wgt, x1, x2 = get_x1x2(xrand[:, :2], shat_min, com_sqrts**2)
        roots = np.sqrt(shat)
        zeros = np.zeros_like(x1)
        p_out, wtps = rambo(xrand[:, 2:], nparticles - 2, roots, masses, const, xfactor)
        wgt = wgt * wtps
    ein = roots / 2.0
    pa = np.stack([ein, zeros, zeros, ein], axis=1)[:, None, :]
    pb = np.stack([ein, zeros, zeros, -ein], axis=1)[:, None, :]
    final_p = np.concatenate([pa, pb, p_out], axis=1)
    wgt = wgt * const.GEV2PB
    wgt = wgt / (2 * shat)
    return final_p, wgt, x1, x2


def boost_to_lab(p_com, x1, x2):
    """phasespace.py:322-356: E' = E cosh(eta) - pz sinh(eta), pz' = -E sinh(eta) + pz cosh(eta)."""
    eta = -0.5 * np.log(x1 / x2)
    cth, sth = np.cosh(eta)[:, None], np.sinh(eta)[:, None]
    E, pz = p_com[..., 0], p_com[..., 3]
    # batch_dot sums the four products in index order; the two zero terms do not change the value
    out = p_com.copy()
    out[..., 0] = E * cth + pz * (-1.0 * sth)
    out[..., 3] = E * (-1.0 * sth) + pz * cth
    return out


class PhaseSpaceGenerator:
    """phasespace.py:359-520."""

    def __init__(self, nparticles, com_sqrts, masses=None, com_output=True, algorithm="ramboflow",
                 const=REFERENCE, xfactor="reference"):
        if masses is None:
            masses = [0.0] * (nparticles - 2)
        if len(masses) != nparticles - 2:
            raise ValueError("Missmatch in PhaseSpaceGenerator between particles and masses")
        if algorithm != "ramboflow":
            raise ValueError(f"PS algorithm {algorithm} not understood")
        self._sqrts, self._masses, self._n = float(com_sqrts), masses, nparticles
        self._cuts = []
        self._com_output = com_output
        self._const, self._xfactor = const, xfactor

    def clear_cuts(self):
        self._cuts = []

    pt = staticmethod(pt)
    mt = staticmethod(mt)
    mt2 = staticmethod(mt2)

    # pair variables (an extension of this package's PhaseSpaceGenerator, not in the reference): particle=(i, j)
    @staticmethod
    def mij(a, b):
        s = a + b
        return np.sqrt(np.maximum(s[..., 0] ** 2 - s[..., 1] ** 2 - s[..., 2] ** 2 - s[..., 3] ** 2, 0.0))

    @staticmethod
    def dr(a, b):
        def eta(p):
            pabs = np.sqrt(p[..., 1] ** 2 + p[..., 2] ** 2 + p[..., 3] ** 2)
            return 0.5 * np.log((pabs + p[..., 3]) / (pabs - p[..., 3]))
        dphi = np.abs(np.arctan2(a[..., 2], a[..., 1]) - np.arctan2(b[..., 2], b[..., 1]))
        dphi = np.where(dphi > np.pi, 2.0 * np.pi - dphi, dphi)
        return np.sqrt((eta(a) - eta(b)) ** 2 + dphi**2)

    def register_cut(self, variable, particle=None, min_val=None, max_val=None):
        """phasespace.py:424-478: min < var(p_particle) < max, strict inequalities."""
        try:
            fun = getattr(self, variable)
        except AttributeError:
            raise ValueError(f"{variable} is not implemented")
        if particle is not None and not isinstance(particle, tuple) and particle >= self._n:
            raise ValueError(f"Cannot apply cuts to particle {particle}, python idx starts at 0!")
        self._cuts.append((fun, particle, min_val, max_val))

    def __call__(self, xrand):
        """phasespace.py:480-520: cuts on COM-frame momenta, compaction, then optional boost."""
        ps, wgt, x1, x2 = ramboflow(xrand, self._n, self._sqrts, self._masses, self._const, self._xfactor)
        if self._cuts:
            ok = np.ones(ps.shape[0], dtype=bool)
            for fun, particle, lo, hi in self._cuts:
                if isinstance(particle, tuple):
                    val = fun(ps[:, particle[0], :], ps[:, particle[1], :])
                else:
                    val = fun(ps[:, particle, :] if particle is not None else ps)
                if lo is not None:
                    ok &= val > lo
                if hi is not None:
                    ok &= val < hi
            ps, wgt, x1, x2 = ps[ok], wgt[ok], x1[ok], x2[ok]
            idx = np.argwhere(ok).astype(np.int32)
        else:
            idx = np.int32(1)
        if not self._com_output:
            ps = boost_to_lab(ps, x1, x2)
        return ps, wgt, x1, x2, idx


def massless_volume(n, e):
    """Analytic n-body massless phase-space volume (reference tests/test_ps.py:9-16)."""
    return pow(e, 2 * n - 4) / 2.0 / math.factorial(n - 1) / math.factorial(n - 2) / pow(4 * np.pi, 2 * n - 3)
