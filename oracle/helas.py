"""HELAS external wavefunctions, numpy restatement.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows python_package/madflow/wavefunctions_flow.py.  Inputs: p (nevt,4) float64 as (E,px,py,pz);
mass, nhel, nsf/nsv Python scalars (uniform over the batch, as in the reference where they are
shape-() tensors).  Output: (6,nevt) complex128 (3,nevt for sxxxxx): rows 0-1 carry the momentum,
rows 2-5 the spinor / polarisation components.
"""
import numpy as np

from . import REFERENCE

_err = dict(divide="ignore", invalid="ignore")


def _cplx(re, im):
    re, im = np.broadcast_arrays(np.asarray(re, dtype=np.float64), np.asarray(im, dtype=np.float64))
    out = np.empty(re.shape, dtype=np.complex128)
    out.real, out.imag = re, im
    return out


def sign(x, y):
    """wavefunctions_flow.py:19-29: x*sign(y), which is 0 at y == 0 (not Fortran's +|x|)."""
    return x * np.sign(y)


def _momentum_rows(p, s):
    # wavefunctions_flow.py:72-73 (i, s=-nsf), :105-106 (o, s=+nsf), :136-137 (v, s=+nsv)
    return _cplx(p[:, 0] * s, p[:, 3] * s), _cplx(p[:, 1] * s, p[:, 2] * s)


def sxxxxx(p, nss):
    """wavefunctions_flow.py:33-51.  (The reference's own body fails at :49 -- expand_dims of a
    scalar on axis 1 -- so this states the evident intent: (p*nss as two complex, 1+0j).)"""
    w0, w1 = _momentum_rows(p, nss)
    return np.stack([w0, w1, np.ones_like(w0)])


def _pp_zero_spinor(fmass, nsf, ip, im):
    """wavefunctions_flow.py:372-387 (_ox_massive_pp_zero): rest-frame spinor, 4 real numbers."""
    sqm0 = np.sqrt(abs(fmass))
    sqm = [sqm0, sign(sqm0, fmass)]
    return [
        im * sqm[abs(im)],
        ip * nsf * sqm[abs(im)],
        im * nsf * sqm[abs(ip)],
        ip * sqm[abs(ip)],
    ]


def _massive_building_blocks(p, fmass, nsf, nh, ysign):
    """Shared by wavefunctions_flow.py:204-229 (ix, ysign=+1) and :405-432 (ox, ysign=-1)."""
    E, px, py, pz = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    pp = np.minimum(E, np.sqrt(px**2 + py**2 + pz**2))
    sf = [(1 + nsf + (1 - nsf) * nh) * 0.5, (1 + nsf - (1 - nsf) * nh) * 0.5]
    with np.errstate(**_err):
        omega = [np.sqrt(E + pp), fmass / np.sqrt(E + pp)]
        ip, im = int((1 + nh) // 2), int((1 - nh) // 2)
        sfomeg = [sf[0] * omega[ip], sf[1] * omega[im]]
        pp3 = np.maximum(pp + pz, 0.0)
        den = np.sqrt(2.0 * pp * pp3)
        chi1 = np.where(pp3 == 0, _cplx(-nh, 0.0), _cplx(nh * px / den, ysign * py / den))
        chi2 = _cplx(np.sqrt(pp3 * 0.5 / pp), 0.0)
    chi = [chi2, chi1]
    return pp, sfomeg, chi, ip, im


def ixxxxx(p, fmass, nhel, nsf, const=REFERENCE):
    """wavefunctions_flow.py:55-85 with helpers :159-330."""
    p = np.asarray(p, dtype=np.float64)
    w0, w1 = _momentum_rows(p, -nsf)
    nh = nhel * nsf
    if fmass != 0:
        pp, sfomeg, chi, ip, im = _massive_building_blocks(p, fmass, nsf, nh, +1.0)
        moving = [sfomeg[0] * chi[im], sfomeg[0] * chi[ip], sfomeg[1] * chi[im], sfomeg[1] * chi[ip]]
        # :181-183: rest frame uses _ox_massive_pp_zero with (ip, im) := (im, ip)
        rest = _pp_zero_spinor(fmass, nsf, im, ip)
        v = [np.where(pp == 0, _cplx(r, 0.0), m) for r, m in zip(rest, moving)]
    else:
        E, px, py, pz = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
        with np.errstate(**_err):
            sqp0p3 = np.sqrt(np.maximum(E + pz, 0.0)) * nsf
            chi1 = np.where(sqp0p3 == 0, _cplx(-nhel * np.sqrt(2.0 * E), 0.0),
                            _cplx(nh * px / sqp0p3, py / sqp0p3))
        chi0 = _cplx(sqp0p3, 0.0)
        zero = np.zeros_like(chi0)
        v = [zero, zero, chi0, chi1] if nh == 1 else [chi1, chi0, zero, zero]  # :296-330
    return np.stack([w0, w1] + v)


def oxxxxx(p, fmass, nhel, nsf, const=REFERENCE):
    """wavefunctions_flow.py:88-116 with helpers :334-462."""
    p = np.asarray(p, dtype=np.float64)
    w0, w1 = _momentum_rows(p, nsf)
    nh = nhel * nsf
    if fmass != 0:
        pp, sfomeg, chi, _, _ = _massive_building_blocks(p, fmass, nsf, nh, -1.0)
        ipl, iml = int((1 + nh) // 2), int((1 - nh) // 2)
        moving = [sfomeg[1] * chi[iml], sfomeg[1] * chi[ipl], sfomeg[0] * chi[iml], sfomeg[0] * chi[ipl]]
        ip = int(-((1 - nh) // 2) * nhel)  # :353-354
        im = int((1 + nh) // 2 * nhel)
        rest = _pp_zero_spinor(fmass, nsf, ip, im)
        v = [np.where(pp == 0, _cplx(r, 0.0), m) for r, m in zip(rest, moving)]
    else:
        E, px, py, pz = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
        with np.errstate(**_err):
            sqp0p3 = np.sqrt(np.maximum(E + pz, 0.0)) * nsf
            chi0 = np.where(sqp0p3 == 0, _cplx(-nhel * np.sqrt(2.0 * E), 0.0),
                            _cplx(nh * px / sqp0p3, -py / sqp0p3))  # :452-455: p*[1,1,-1,1]
        chi1 = _cplx(sqp0p3, 0.0)
        zero = np.zeros_like(chi1)
        v = [chi1, chi0, zero, zero] if nh == 1 else [zero, zero, chi0, chi1]  # :459-462, roles inverted
    return np.stack([w0, w1] + v)


def vxxxxx(p, vmass, nhel, nsv, const=REFERENCE):
    """wavefunctions_flow.py:119-154 with helpers :467-747."""
    p = np.asarray(p, dtype=np.float64)
    SQH = const.SQH
    w0, w1 = _momentum_rows(p, nsv)
    E, px, py, pz = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    with np.errstate(**_err):
        if nhel == 4:  # BRST check polarisation, :467-515
            d = E if vmass == 0 else vmass
            v = [_cplx(c / d, 0.0) for c in (E, px, py, pz)]
            return np.stack([w0, w1] + v)
        pt2 = px**2 + py**2
        hel0 = 1.0 - abs(nhel)
        nsvahl = nsv * abs(nhel)
        if vmass != 0:  # :548-678
            pp = np.minimum(E, np.sqrt(pt2 + pz**2))
            pt = np.minimum(pp, np.sqrt(pt2))
            emp = E / (vmass * pp)
            v2 = _cplx(hel0 * pp / vmass, 0.0)
            v5 = _cplx(hel0 * pz * emp + nhel * pt / pp * SQH, 0.0)
            pzpt = pz / (pp * pt) * SQH * nhel
            v3a = _cplx(hel0 * px * emp - px * pzpt, -nsvahl * py / pt * SQH)
            v4a = _cplx(hel0 * py * emp - py * pzpt, nsvahl * px / pt * SQH)
            v3b = _cplx(-nhel * SQH, 0.0) * np.ones_like(E)
            v4b = _cplx(0.0, nsvahl * sign(SQH, pz))
            v3 = np.where(pt != 0, v3a, v3b)
            v4 = np.where(pt != 0, v4a, v4b)
            rest = [1.0 + 0j, -nhel * SQH + 0j, 1j * nsvahl * SQH, hel0 + 0j]  # :576-593, v[0] stays 1
            v = [np.where(pp == 0, r, m) for r, m in zip(rest, [v2, v3, v4, v5])]
        else:  # :683-747
            pp = E
            pt = np.sqrt(pt2)
            v2 = np.zeros_like(w0)
            v5 = _cplx(nhel * pt / pp * SQH, 0.0)
            pzpt = pz / (pp * pt) * SQH * nhel
            v3a = _cplx(-px * pzpt, -nsv * py / pt * SQH)
            v4a = _cplx(-py * pzpt, nsv * px / pt * SQH)
            v3b = _cplx(-nhel * SQH, 0.0) * np.ones_like(E)
            v4b = _cplx(0.0, nsv * sign(SQH, pz))
            v = [v2, np.where(pt != 0, v3a, v3b), np.where(pt != 0, v4a, v4b), v5]
    return np.stack([w0, w1] + v)
