"""PDF and alpha_s interpolation of an LHAPDF `lhagrid1` set.  TEST INFRASTRUCTURE.

The reference gets both from pdfflow (`mkPDF(args.pdf + "/0")`, `pdf.xfxQ2(pids, x, q2)`,
`pdf.alphasQ2(q2)`: python_package/madflow/scripts/madflow_exec.py:342, 382, 412-413, 431).  pdfflow is a
third-party dependency (unpinned, python_package/setup.py:11) that is absent from /root/reference and not
installable offline, and so are LHAPDF and every grid file: PARITY UNPINNED.  What is restated here is the
published algorithm pdfflow implements -- LHAPDF 6's `LogBicubicInterpolator` (cubic Hermite splines in
log x and log Q2 with finite-difference derivatives, one-sided at the edges of a subgrid) and
`AlphaS_Ipol` (cubic Hermite in log Q2 on the AlphaS_Qs / AlphaS_Vals table, split at flavour
thresholds) -- with plain numpy, vectorised over events.  Outside the grid x and Q2 are frozen at the
edge (LHAPDF's "nearest" extrapolation); the hot path never gets there (x >= (sum m)^2/s ~ 7e-4,
Q in [m_t, sqrt(s)]).

Checked by properties (tests/test_oracle.py): knots are reproduced exactly, quadratics in (log x, log Q2)
on uniform log grids are reproduced to rounding, continuity across cells, subgrid selection at thresholds.
"""
import os

import numpy as np


# ---------------------------------------------------------------------------------- file format
def parse_info(path):
    """`<set>.info`: a YAML dictionary (Flavors, AlphaS_Qs, AlphaS_Vals, NumMembers, XMin, ...)."""
    import yaml

    with open(path) as fh:
        return yaml.safe_load(fh)


def parse_member(path):
    """`<set>_NNNN.dat` in lhagrid1 format: header block, then per subgrid a line of x knots, a line of Q
    knots (NOT Q2), a line of flavour ids and nx*nq rows of xf values (x-major, one column per flavour),
    blocks separated by `---`.  Returns [dict(x, q2, pids, xf[nx, nq, nfl])]."""
    with open(path) as fh:
        blocks = fh.read().split("---")
    subgrids = []
    for blk in blocks[1:]:
        lines = [ln for ln in blk.strip().splitlines() if ln.strip()]
        if len(lines) < 4:
            continue
        x = np.array(lines[0].split(), dtype=np.float64)
        q = np.array(lines[1].split(), dtype=np.float64)
        pids = [int(t) for t in lines[2].split()]
        vals = np.array([ln.split() for ln in lines[3:]], dtype=np.float64)
        assert vals.shape == (len(x) * len(q), len(pids)), (vals.shape, len(x), len(q), len(pids))
        subgrids.append(dict(x=x, q2=q * q, pids=pids, xf=vals.reshape(len(x), len(q), len(pids))))
    assert subgrids, f"{path}: no subgrid found"
    return subgrids


def find_set(name, dirname=None):
    """Directory of PDF set `name`: `dirname`, else PDFFLOW_DATA_PATH / LHAPDF_DATA_PATH (what pdfflow's
    mkPDF consults before asking `lhapdf-config --datadir`)."""
    roots = [dirname] if dirname else []
    for var in ("PDFFLOW_DATA_PATH", "LHAPDF_DATA_PATH", "LHA_PATH"):
        roots += [p for p in os.environ.get(var, "").split(":") if p]
    for r in roots:
        d = os.path.join(r, name)
        if os.path.isdir(d):
            return d
    raise FileNotFoundError(f"PDF set {name} not found in {roots or '(no search path: pass dirname or set LHAPDF_DATA_PATH)'}")


def load_set(spec, dirname=None):
    """'Name/member' (pdfflow's mkPDF convention, madflow_exec.py:342) -> (info, subgrids)."""
    name, _, member = spec.partition("/")
    member = int(member or 0)
    d = find_set(name, dirname)
    info = parse_info(os.path.join(d, f"{name}.info"))
    return info, parse_member(os.path.join(d, f"{name}_{member:04d}.dat"))


# ---------------------------------------------------------------------------------- interpolation
def _cubic(t, vl, vdl, vh, vdh):
    """LHAPDF `_interpolateCubic`: cubic Hermite on [0, 1]."""
    t2 = t * t
    t3 = t2 * t
    p0 = (2 * t3 - 3 * t2 + 1) * vl
    m0 = (t3 - 2 * t2 + t) * vdl
    p1 = (-2 * t3 + 3 * t2) * vh
    m1 = (t3 - t2) * vdh
    return p0 + m0 + p1 + m1


def _ddlogx(xf, logx, ix, iq):
    """LHAPDF `_dxf_dlogx`: central = mean of the two one-sided differences; one-sided at the edges."""
    nx = len(logx)
    lo = np.maximum(ix - 1, 0)
    hi = np.minimum(ix + 1, nx - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        ldd = (xf[ix, iq] - xf[lo, iq]) / (logx[ix] - logx[lo])
        rdd = (xf[hi, iq] - xf[ix, iq]) / (logx[hi] - logx[ix])
    return np.where(ix == 0, rdd, np.where(ix == nx - 1, ldd, (ldd + rdd) / 2.0))


def _interp_subgrid(sg, ifl, x, q2):
    """LHAPDF `LogBicubicInterpolator::_interpolateXQ2` on one subgrid, vectorised."""
    logxs, logq2s = np.log(sg["x"]), np.log(sg["q2"])
    xf = sg["xf"][:, :, ifl]
    nx, nq = len(logxs), len(logq2s)
    assert nx >= 4 and nq >= 4, "bicubic interpolation needs 4 knots per direction"
    x = np.clip(x, sg["x"][0], sg["x"][-1])
    q2 = np.clip(q2, sg["q2"][0], sg["q2"][-1])
    logx, logq2 = np.log(x), np.log(q2)
    ix = np.clip(np.searchsorted(sg["x"], x, side="right") - 1, 0, nx - 2)
    iq = np.clip(np.searchsorted(sg["q2"], q2, side="right") - 1, 0, nq - 2)
    dlogx = logxs[ix + 1] - logxs[ix]
    tx = (logx - logxs[ix]) / dlogx
    dq1 = logq2s[iq + 1] - logq2s[iq]
    tq = (logq2 - logq2s[iq]) / dq1

    def row(j):
        return _cubic(tx, xf[ix, j], _ddlogx(xf, logxs, ix, j) * dlogx, xf[ix + 1, j], _ddlogx(xf, logxs, ix + 1, j) * dlogx)

    vl, vh = row(iq), row(iq + 1)
    lowedge, highedge = iq == 0, iq + 1 == nq - 1
    jl, jh = np.maximum(iq - 1, 0), np.minimum(iq + 2, nq - 1)
    vll, vhh = row(jl), row(jh)
    with np.errstate(divide="ignore", invalid="ignore"):
        dq0 = logq2s[iq] - logq2s[jl]
        dq2 = logq2s[jh] - logq2s[iq + 1]
        fwd = (vh - vl) / dq1
        vdl_c = (fwd + (vl - vll) / dq0) / 2.0
        vdh_c = (fwd + (vhh - vh) / dq2) / 2.0
    vdl = np.where(lowedge, fwd, vdl_c)
    vdh = np.where(highedge, fwd, vdh_c)
    return _cubic(tq, vl, vdl * dq1, vh, vdh * dq1)


class GridPDF:
    """One member of an lhagrid1 set: `xfxQ2(pids, x, q2)` -> (nevt, len(pids)), `alphasQ2(q2)` -> (nevt,)
    (the two pdfflow calls of the hot path, madflow_exec.py:412-413, 431)."""

    def __init__(self, info, subgrids):
        self.info, self.subgrids = info, subgrids
        self.pids = list(subgrids[0]["pids"])
        q = np.asarray(info.get("AlphaS_Qs", []), dtype=np.float64)
        self.as_q2 = q * q
        self.as_vals = np.asarray(info.get("AlphaS_Vals", []), dtype=np.float64)

    @classmethod
    def from_set(cls, spec, dirname=None):
        return cls(*load_set(spec, dirname))

    def flavour_index(self, pid):
        pid = 21 if pid == 0 else int(pid)   # LHAPDF: 0 is an alias of the gluon
        return self.pids.index(pid)

    def xfxQ2(self, pids, x, q2):
        x, q2 = np.broadcast_arrays(np.asarray(x, dtype=np.float64), np.asarray(q2, dtype=np.float64))
        x, q2 = x.reshape(-1), q2.reshape(-1)
        out = np.zeros((x.shape[0], len(pids)))
        lo = np.array([sg["q2"][0] for sg in self.subgrids])
        # the subgrid whose [q2min, q2max) holds q2; a threshold value belongs to the upper subgrid
        which = np.clip(np.searchsorted(lo, q2, side="right") - 1, 0, len(self.subgrids) - 1)
        for s, sg in enumerate(self.subgrids):
            sel = which == s
            if not np.any(sel):
                continue
            for c, pid in enumerate(pids):
                out[sel, c] = _interp_subgrid(sg, self.flavour_index(pid), x[sel], q2[sel])
        return out

    def alpha_subgrids(self):
        """AlphaS_Ipol::_setup_grids: the table split where a Q value is repeated (flavour threshold)."""
        cuts = [0] + [i + 1 for i in range(len(self.as_q2) - 1) if self.as_q2[i] == self.as_q2[i + 1]] + [len(self.as_q2)]
        return [(self.as_q2[a:b], self.as_vals[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b - a >= 2]

    def alphasQ2(self, q2):
        """LHAPDF `AlphaS_Ipol::calcAlphasQ2`."""
        q2 = np.asarray(q2, dtype=np.float64).reshape(-1)
        subs = self.alpha_subgrids()
        assert subs, "the set has no AlphaS_Qs / AlphaS_Vals table"
        out = np.empty_like(q2)
        lo = np.array([s[0][0] for s in subs])
        which = np.clip(np.searchsorted(lo, q2, side="right") - 1, 0, len(subs) - 1)
        for s, (kq2, kas) in enumerate(subs):
            sel = which == s
            if not np.any(sel):
                continue
            lq = np.log(kq2)
            n = len(lq)
            z = np.log(q2[sel])
            i = np.clip(np.searchsorted(kq2, q2[sel], side="right") - 1, 0, n - 2)
            fwd = lambda j: (kas[np.minimum(j + 1, n - 1)] - kas[j]) / np.where(j + 1 < n, lq[np.minimum(j + 1, n - 1)] - lq[j], 1.0)
            bwd = lambda j: (kas[j] - kas[np.maximum(j - 1, 0)]) / np.where(j > 0, lq[j] - lq[np.maximum(j - 1, 0)], 1.0)
            cen = lambda j: 0.5 * (fwd(j) + bwd(j))
            d0 = np.where(i == 0, fwd(i), cen(i))
            d1 = np.where(i == n - 2, bwd(i + 1), cen(i + 1))
            dl = lq[i + 1] - lq[i]
            out[sel] = _cubic((z - lq[i]) / dl, kas[i], d0 * dl, kas[i + 1], d1 * dl)
        # below the first knot: constant gradient in log-log; above the last: frozen
        below = q2 < self.as_q2[0]
        if np.any(below):
            nxt = 1
            while self.as_q2[nxt] == self.as_q2[0]:
                nxt += 1
            grad = np.log10(self.as_vals[nxt] / self.as_vals[0]) / np.log10(self.as_q2[nxt] / self.as_q2[0])
            out[below] = self.as_vals[0] * (q2[below] / self.as_q2[0]) ** grad
        out[q2 > self.as_q2[-1]] = self.as_vals[-1]
        return out


# ---------------------------------------------------------------------------------- synthetic set
def write_toy_set(root, name="ToyPDF", nx=60, q_knots=((1.65, 2.5, 3.5, 4.75), (4.75, 8.0, 20.0, 60.0, 173.0, 500.0, 2000.0, 1.0e4)),
                  pids=(-5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 21), alpha_mz=0.118):
    """Write a small synthetic lhagrid1 set (no real grid is available offline): smooth CTEQ-like shapes
    x f(x, Q) = A x^-a (1-x)^b (1 + c log(Q/Q0)), two Q subgrids joined at the b threshold, one-loop alpha_s
    table.  The numbers are NOT physical; the set exercises the file format, the subgrid logic and the
    interpolation."""
    import math

    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    xs = np.concatenate([np.logspace(-7, -1, nx // 2, endpoint=False), np.linspace(0.1, 1.0, nx - nx // 2)])

    def shape(pid, x, q):
        if pid == 21:
            a, b, A, c = 0.25, 5.0, 3.0, 0.30
        elif pid in (1, 2):
            a, b, A, c = 0.15, 3.0 + 0.5 * pid, 1.2 / pid, -0.05
        else:
            a, b, A, c = 0.20, 7.0 + 0.3 * abs(pid), 0.25 / (1 + 0.2 * abs(pid)), 0.15
        return A * x ** (-a) * (1.0 - x) ** b * (1.0 + c * math.log(q / 1.65))

    with open(os.path.join(d, f"{name}_0000.dat"), "w") as fh:
        fh.write("PdfType: central\nFormat: lhagrid1\n---\n")
        for qs in q_knots:
            fh.write(" ".join(f"{x:.10e}" for x in xs) + "\n")
            fh.write(" ".join(f"{q:.10e}" for q in qs) + "\n")
            fh.write(" ".join(str(p) for p in pids) + "\n")
            for x in xs:
                for q in qs:
                    fh.write(" ".join(f"{shape(p, x, q):.10e}" for p in pids) + "\n")
            fh.write("---\n")
    qa = [1.65, 2.0, 3.0, 4.75, 4.75, 7.0, 15.0, 40.0, 91.1876, 200.0, 600.0, 2000.0, 1.0e4]
    b5, b4 = (33 - 10) / (12 * math.pi), (33 - 8) / (12 * math.pi)

    def a_s(q, below):
        a5 = lambda qq: alpha_mz / (1 + alpha_mz * b5 * math.log(qq * qq / 91.1876 ** 2))
        if not below:
            return a5(q)
        ab = a5(4.75)
        return ab / (1 + ab * b4 * math.log(q * q / 4.75 ** 2))

    vals = [a_s(q, i < 4) for i, q in enumerate(qa)]
    with open(os.path.join(d, f"{name}.info"), "w") as fh:
        fh.write(f'SetDesc: "synthetic test set written by oracle/pdf.py (not physical)"\nNumMembers: 1\n'
                 f"Flavors: [{', '.join(str(p) for p in pids)}]\nFormat: lhagrid1\n"
                 f"XMin: {xs[0]:.6e}\nXMax: 1.0\nQMin: 1.65\nQMax: 1.0e4\n"
                 f"AlphaS_MZ: {alpha_mz}\nAlphaS_Type: ipol\n"
                 f"AlphaS_Qs: [{', '.join(f'{q:.10e}' for q in qa)}]\n"
                 f"AlphaS_Vals: [{', '.join(f'{v:.10e}' for v in vals)}]\n")
    return d
