"""PDFs and alpha_s from LHAPDF `lhagrid1` sets, evaluated on the GPU -- the part of pdfflow's interface that
madflow uses (python_package/madflow/scripts/madflow_exec.py:342 `mkPDF(args.pdf + "/0")`, :412-413
`pdf.xfxQ2(int_me(flavours), x, q2)`, :382/:431 `pdf.alphasQ2(q2)`).

pdfflow is a third-party dependency of the reference that is not in its tree (unpinned in setup.py:11) and
cannot be installed offline; the algorithm restated in csrc/pdf.cuh is the one it implements, LHAPDF 6's
log-bicubic interpolation and `AlphaS_Ipol` (parity unpinned).  The grid of one member is read
from the standard LHAPDF directory layout (`<dir>/<set>/<set>.info`, `<set>_NNNN.dat`), packed into one array of
doubles (knots, their logarithms, values; layout in csrc/pdf.cuh) and kept in device memory; the fused integrand
reads it through `mfp_integrand_args.d_pdf`, the methods below through `mf_pdf_xfxq2` / `mf_pdf_alphasq2`
(include/madflow_b200.h).  No grid ships with this repository and none can be downloaded here: the sets have
to be on disk (`dirname=`, PDFFLOW_DATA_PATH or LHAPDF_DATA_PATH).
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _runtime as rt

PDF_HEADER = 8   # csrc/pdf.cuh


class PDFError(rt.MadflowB200Error):
    pass


def _search_path(dirname):
    roots = [dirname] if dirname else []
    for var in ("PDFFLOW_DATA_PATH", "LHAPDF_DATA_PATH", "LHA_PATH"):
        roots += [p for p in os.environ.get(var, "").split(":") if p]
    return roots


def read_info(path):
    import yaml

    with open(path) as fh:
        return yaml.safe_load(fh)


def read_member(path):
    """One `lhagrid1` member file -> list of subgrids dict(x, q2, pids, xf (nx, nq, nfl))."""
    with open(path) as fh:
        text = fh.read()
    blocks = text.split("---")
    if "lhagrid1" not in blocks[0]:
        raise PDFError(f"{path}: not an lhagrid1 file")
    subgrids = []
    for blk in blocks[1:]:
        lines = [ln for ln in blk.strip().splitlines() if ln.strip()]
        if len(lines) < 4:
            continue
        x = np.array(lines[0].split(), dtype=np.float64)
        q = np.array(lines[1].split(), dtype=np.float64)
        pids = [int(t) for t in lines[2].split()]
        vals = np.loadtxt(lines[3:], dtype=np.float64, ndmin=2)
        if vals.shape != (len(x) * len(q), len(pids)):
            raise PDFError(f"{path}: subgrid of {len(x)} x {len(q)} knots and {len(pids)} flavours has {vals.shape} values")
        if len(x) < 4 or len(q) < 4:
            raise PDFError(f"{path}: the bicubic interpolation needs at least 4 knots per direction")
        subgrids.append(dict(x=x, q2=q * q, pids=pids, xf=vals.reshape(len(x), len(q), len(pids))))
    if not subgrids:
        raise PDFError(f"{path}: no subgrid found")
    if any(sg["pids"] != subgrids[0]["pids"] for sg in subgrids):
        raise PDFError(f"{path}: the subgrids list different flavours")
    return subgrids


def pack_table(info, subgrids):
    """The member as one float64 array in the layout csrc/pdf.cuh documents."""
    nsub, nfl = len(subgrids), len(subgrids[0]["pids"])
    q = np.asarray(info.get("AlphaS_Qs") or [], dtype=np.float64)
    av = np.asarray(info.get("AlphaS_Vals") or [], dtype=np.float64)
    aq2 = q * q
    cuts = [0] + [i + 1 for i in range(len(aq2) - 1) if aq2[i] == aq2[i + 1]] + [len(aq2)]
    asubs = [(aq2[a:b], av[a:b]) for a, b in zip(cuts[:-1], cuts[1:]) if b - a >= 2]
    head = [float(nsub), float(nfl), float(len(asubs)), 0.0, 0.0, 0.0, 0.0, 0.0]
    if len(aq2):
        nxt = 1
        while nxt < len(aq2) - 1 and aq2[nxt] == aq2[0]:
            nxt += 1
        grad = math.log10(av[nxt] / av[0]) / math.log10(aq2[nxt] / aq2[0])
        head[3:8] = [aq2[0], av[0], grad, aq2[-1], av[-1]]
    off = PDF_HEADER + 8 * nsub + 4 * len(asubs)
    desc, data = [], []
    for sg in subgrids:
        nx, nq = len(sg["x"]), len(sg["q2"])
        o_x, o_lx, o_q, o_lq, o_xf = off, off + nx, off + 2 * nx, off + 2 * nx + nq, off + 2 * nx + 2 * nq
        desc += [nx, nq, o_x, o_lx, o_q, o_lq, o_xf, sg["q2"][0]]
        data += [sg["x"], np.log(sg["x"]), sg["q2"], np.log(sg["q2"]), sg["xf"].reshape(-1)]
        off = o_xf + nx * nq * nfl
    for kq2, kas in asubs:
        n = len(kq2)
        desc += [n, off, off + n, off + 2 * n]
        data += [kq2, np.log(kq2), kas]
        off += 3 * n
    table = np.concatenate([np.asarray(head), np.asarray(desc, dtype=np.float64)] + [np.asarray(d, dtype=np.float64) for d in data])
    assert table.shape[0] == off
    return table


class PDF:
    """One member of a PDF set on the GPU.  Method names and argument order are pdfflow's."""

    def __init__(self, dirname, fname, member=0):
        self.dirname, self.fname, self.member = dirname, fname, int(member)
        self.info = read_info(os.path.join(dirname, fname, f"{fname}.info"))
        path = os.path.join(dirname, fname, f"{fname}_{self.member:04d}.dat")
        if not os.path.exists(path):
            raise PDFError(f"{path} not found (NumMembers = {self.info.get('NumMembers')})")
        self.subgrids = read_member(path)
        self.flavor_scheme = list(self.subgrids[0]["pids"])
        self.has_alphas = bool(self.info.get("AlphaS_Qs")) and bool(self.info.get("AlphaS_Vals"))
        self._host_table = pack_table(self.info, self.subgrids)
        self._table = None

    # ---- device side
    @property
    def table(self):
        """The packed member in device memory (uploaded on first use)."""
        if self._table is None:
            rt._require_cuda()
            self._table = rt.to_device(self._host_table)
        return self._table

    def column(self, pid):
        """Column of flavour `pid` in the table; 0 is LHAPDF's alias of the gluon (21)."""
        pid = int(pid)
        pid = 21 if pid == 0 else pid
        try:
            return self.flavor_scheme.index(pid)
        except ValueError:
            raise PDFError(f"flavour {pid} is not in {self.fname} ({self.flavor_scheme})") from None

    # ---- pdfflow's interface
    @property
    def nmembers(self):
        return 1

    @property
    def active_members(self):
        return [self.member]

    @property
    def q2min(self):
        return float(self.subgrids[0]["q2"][0])

    @property
    def q2max(self):
        return float(self.subgrids[-1]["q2"][-1])

    @property
    def xmin(self):
        return float(self.subgrids[0]["x"][0])

    def trace(self):
        """pdfflow traces its TensorFlow graphs here; the CUDA kernels are compiled ahead of time."""

    alphas_trace = trace

    def xfxQ2(self, pid, x, q2):
        """x f(x, Q2) for the flavours `pid` (PDG ids): (nevt, len(pid)) float64 on the GPU, squeezed like
        pdfflow's result (the reference reshapes it to (-1, nflavours) anyway, madflow_exec.py:415-416)."""
        pids = [int(p) for p in (pid.tolist() if hasattr(pid, "tolist") else pid)] if not isinstance(pid, int) else [pid]
        asked = pids
        pids = sorted(set(pids))     # every flavour once (a `p p > ..` luminosity list repeats them), gathered below
        cols = (ctypes.c_int32 * len(pids))(*[self.column(p) for p in pids])
        x, q2 = rt.to_device(x).reshape(-1), rt.to_device(q2).reshape(-1)
        if q2.numel() == 1 and x.numel() > 1:
            q2 = q2.expand(x.numel()).contiguous()
        if x.numel() != q2.numel():
            raise PDFError("xfxQ2: x and q2 must have the same number of points")
        n = x.numel()
        out = torch.empty((n, len(pids)), dtype=torch.float64, device=x.device)
        lib = rt.core()
        rt.check(lib, lib.mf_pdf_xfxq2(rt.ptr(self.table), cols, len(pids), rt.ptr(x), rt.ptr(q2), ctypes.c_int64(n),
                                       rt.ptr(out), rt.stream_ptr()))
        if asked != pids:
            out = out[:, [pids.index(p) for p in asked]]
        return out.squeeze()

    def xfxQ2_allpid(self, x, q2):
        return self.xfxQ2(self.flavor_scheme, x, q2)

    def xfxQ(self, pid, x, q):
        return self.xfxQ2(pid, x, rt.to_device(q) ** 2)

    def alphasQ2(self, q2):
        if not self.has_alphas:
            raise PDFError(f"{self.fname}.info has no AlphaS_Qs / AlphaS_Vals table")
        q2 = rt.to_device(q2).reshape(-1)
        out = torch.empty_like(q2)
        lib = rt.core()
        rt.check(lib, lib.mf_pdf_alphasq2(rt.ptr(self.table), rt.ptr(q2), ctypes.c_int64(q2.numel()), rt.ptr(out), rt.stream_ptr()))
        return out.squeeze()

    def alphasQ(self, q):
        return self.alphasQ2(rt.to_device(q) ** 2)

    # pdfflow's "python" variants take lists / numpy arrays; the methods above already do
    py_xfxQ2, py_xfxQ2_allpid, py_xfxQ, py_alphasQ2, py_alphasQ = xfxQ2, xfxQ2_allpid, xfxQ, alphasQ2, alphasQ

    def __repr__(self):
        return f"PDF({self.fname}/{self.member}, {len(self.subgrids)} subgrids, flavours {self.flavor_scheme})"


def mkPDF(fname, dirname=None):
    """pdfflow.mkPDF: `fname` = "<set>/<member>" (madflow_exec.py:342), `dirname` = the LHAPDF data directory
    (default: PDFFLOW_DATA_PATH, LHAPDF_DATA_PATH)."""
    name, _, member = str(fname).partition("/")
    roots = _search_path(dirname)
    for r in roots:
        if os.path.isdir(os.path.join(r, name)):
            return PDF(r, name, int(member or 0))
    raise PDFError(f"PDF set '{name}' not found in {roots or 'an empty search path'}: no grid ships with madflow_b200 and none "
                   "can be downloaded here; install the LHAPDF set and pass its directory (dirname= / --pdf_dir / "
                   "LHAPDF_DATA_PATH), or run with --no_pdf")


def initial_state_channels(matrix, pdf, initial_states=None, mirror=None):
    """The flavour pairs whose luminosities are summed for one subprocess: `initial_states` plus, when
    `mirror_initial_states`, the same pairs with the hadrons exchanged (madflow_exec.py:141-155, 446-454),
    as columns of the PDF table.  `initial_states` / `mirror` override what the process library recorded."""
    initials = [tuple(int(f) for f in pair) for pair in (initial_states if initial_states is not None else matrix.initial_states)]
    pairs = list(initials)
    if (getattr(matrix, "mirror_initial_states", False) if mirror is None else mirror):
        pairs += [(b, a) for a, b in initials]
    return [pdf.column(a) for a, _ in pairs], [pdf.column(b) for _, b in pairs]
