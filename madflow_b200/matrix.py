"""Matrix_<proc> classes backed by the compiled process libraries, and `get_processes`.

A `Matrix` has the attributes and the `smatrix(all_ps, *params)` signature of the class the
reference generates from madgraph_plugin/template_files/matrix_method_python.inc:60-104:
nexternal, ndiags, ncomb, initial_states, mirror_initial_states, helicities, denominator,
__str__, smatrix.  `smatrix` hands the momenta and couplings to ONE fused CUDA kernel
(include/madflow_b200_process.h: mfp_smatrix); there is no other implementation behind it.
"""
import collections
import json
import os

import numpy as np
import torch

from . import _runtime as rt
from . import config
from .parameters import Model


class Matrix:
    """Drop-in for the generated `Matrix_<process_string>` (matrix_method_python.inc:60-104)."""

    def __init__(self, name, ir=None, lib_path=None):
        self._lib = rt.process_lib(name, lib_path)
        info = self._lib.info
        self._name = self._lib.name
        self.nexternal = float(info.nexternal)
        self.ndiags = float(info.ndiags)
        self.ncomb = float(info.ncomb)
        self.ncolor = int(info.ncolor)
        self.helicities = [list(map(float, h)) for h in self._lib.helicities]
        self.denominator = float(info.denominator)
        self.param_names = list(self._lib.param_names)
        self.coupling_names = list(self._lib.coupling_names)
        self.flops_per_event = float(info.flops_per_event)
        self.ir = ir
        self.initial_states = [tuple(s) for s in ir["initial_states"]] if ir else []
        self.mirror_initial_states = bool(ir["mirror_initial_states"]) if ir else False

    def __str__(self):
        return self._name

    def set_variant(self, variant):
        """Select the kernel flavour ('default', 'thread', 'hp'); see DESIGN.md "Kernel mapping"."""
        self._lib.set_variant(variant)

    @property
    def variant(self):
        return self._lib.variant

    def clean(self):
        pass

    # -- helpers
    def _split_params(self, params):
        npar, ncoup = len(self.param_names), len(self.coupling_names)
        if len(params) != npar + ncoup:
            raise TypeError(f"smatrix expects {npar} real parameters {self.param_names} followed by {ncoup} "
                            f"couplings {self.coupling_names}; got {len(params)} values")
        par = [float(p.item()) if isinstance(p, torch.Tensor) else float(np.asarray(p).reshape(-1)[0])
               for p in params[:npar]]
        return par, params[npar:]

    def _pack_couplings(self, coups, nevt):
        dev = config.device()
        cols = []
        stride = 0
        for c in coups:
            t = c if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c))
            t = t.to(device=dev, dtype=torch.complex128).reshape(-1)
            if t.numel() not in (1, nevt):
                raise ValueError("couplings must have shape (1,) or (nevents,)")
            if t.numel() == nevt and nevt > 1:
                stride = 1
            cols.append(t)
        if not cols:
            return None, 0
        if stride:
            cols = [c.expand(nevt) if c.numel() == 1 else c for c in cols]
        return torch.stack(cols).contiguous(), stride

    def smatrix(self, all_ps, *params, layout="aos"):
        """|M|^2 summed over helicities and colours, averaged: shape (nevents,) float64 CUDA tensor.

        all_ps: (nevents, nexternal, 4) momenta (E,px,py,pz) [layout="aos", the reference's] or
        (nexternal, 4, nevents) [layout="soa"]; params: masses/widths then couplings, in the order
        of `param_names + coupling_names` (= Model.evaluate's order)."""
        n = int(self.nexternal)
        p = rt.to_device(all_ps)
        if layout == "aos":
            if p.ndim != 3 or p.shape[1:] != (n, 4):
                raise ValueError(f"all_ps must have shape (nevents, {n}, 4)")
            nevt, lay = p.shape[0], rt.LAYOUT_AOS
        elif layout == "soa":
            if p.ndim != 3 or p.shape[:2] != (n, 4):
                raise ValueError(f"all_ps must have shape ({n}, 4, nevents)")
            nevt, lay = p.shape[2], rt.LAYOUT_SOA
        else:
            raise ValueError("layout must be 'aos' or 'soa'")
        par, coups = self._split_params(params)
        d_coup, stride = self._pack_couplings(coups, nevt)
        out = torch.empty(nevt, dtype=torch.float64, device=p.device)
        if nevt:
            self._lib.smatrix(p, lay, nevt, par, d_coup, stride, config.get_constants().SQH, out)
        return out

    def smatrix_pinned(self, h_ps, *params, out=None, chunk=1 << 18):
        """smatrix for momenta (and per-event couplings) that live in PINNED host memory: the events go through the
        device in chunks on two streams, so that the host->device copy of chunk i + 1 and the device->host copy of the
        results of chunk i - 1 overlap the kernel of chunk i.  Returns a pinned host tensor (nevents,) (or fills `out`);
        synchronises before returning.  This is the end-to-end path bench.py times."""
        n = int(self.nexternal)
        if not (isinstance(h_ps, torch.Tensor) and h_ps.device.type == "cpu" and h_ps.is_pinned()):
            raise ValueError("smatrix_pinned: all_ps must be a pinned CPU tensor (torch.Tensor.pin_memory())")
        if h_ps.ndim != 3 or tuple(h_ps.shape[1:]) != (n, 4) or h_ps.dtype != torch.float64:
            raise ValueError(f"all_ps must be float64 with shape (nevents, {n}, 4)")
        nevt = int(h_ps.shape[0])
        par, coups = self._split_params(params)
        coups = [c if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c)) for c in coups]
        coups = [c.reshape(-1).to(torch.complex128) for c in coups]
        if any(c.numel() not in (1, nevt) for c in coups):
            raise ValueError("couplings must have shape (1,) or (nevents,)")
        if out is None:
            out = torch.empty(nevt, dtype=torch.float64).pin_memory()
        dev = config.device()
        cur = torch.cuda.current_stream(dev)
        if not hasattr(self, "_pipe_streams"):
            self._pipe_streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        sqh = config.get_constants().SQH
        for s_ in self._pipe_streams:
            s_.wait_stream(cur)
        for ci, lo in enumerate(range(0, nevt, chunk)):
            hi = min(lo + chunk, nevt)
            with torch.cuda.stream(self._pipe_streams[ci & 1]):
                d_ps = h_ps[lo:hi].to(dev, non_blocking=True)
                part = [c if c.numel() == 1 else c[lo:hi] for c in coups]
                d_coup, stride = self._pack_couplings([c.to(dev, non_blocking=True) for c in part], hi - lo)
                d_out = torch.empty(hi - lo, dtype=torch.float64, device=dev)
                self._lib.smatrix(d_ps, rt.LAYOUT_AOS, hi - lo, par, d_coup, stride, sqh, d_out)
                out[lo:hi].copy_(d_out, non_blocking=True)
        for s_ in self._pipe_streams:
            cur.wait_stream(s_)
            s_.synchronize()
        return out

    def matrix(self, all_ps, hel, *params):
        """One helicity configuration (matrix_method_python.inc:106-138); `hel` is a row of
        self.helicities (or its index)."""
        n = int(self.nexternal)
        p = rt.to_device(all_ps)
        if isinstance(hel, (int, np.integer)):
            ic = int(hel)
        else:
            row = [int(round(float(h))) for h in hel]
            ic = [list(map(int, h)) for h in self.helicities].index(row)
        nevt = p.shape[0]
        par, coups = self._split_params(params)
        d_coup, stride = self._pack_couplings(coups, nevt)
        out = torch.empty(nevt, dtype=torch.float64, device=p.device)
        if nevt:
            self._lib.smatrix(p, rt.LAYOUT_AOS, nevt, par, d_coup, stride, config.get_constants().SQH, out, ic)
        return out

    def smatrix_host(self, all_ps, *params):
        """numpy in, numpy out through the host-buffer C entry point (mfp_smatrix_host)."""
        p = np.ascontiguousarray(all_ps, dtype=np.float64)
        nevt = p.shape[0]
        par, coups = self._split_params(params)
        cs = [np.asarray(c.cpu() if isinstance(c, torch.Tensor) else c, dtype=np.complex128).reshape(-1) for c in coups]
        stride = 1 if any(c.size == nevt and nevt > 1 for c in cs) else 0
        if stride:
            cs = [np.broadcast_to(c, (nevt,)) for c in cs]
        h_coup = np.ascontiguousarray(np.stack(cs)) if cs else np.zeros((0, 1), dtype=np.complex128)
        out = np.empty(nevt)
        self._lib.smatrix_host(p, rt.LAYOUT_AOS, par, h_coup, stride, config.get_constants().SQH, out)
        return out


# ----------------------------------------------------------------------------------------------
# SM parameters of the reference's frozen test matrix element (tests/mockup_debug_me.py:22-26)
SM_PARAMS = {"mdl_MT": 173.0, "mdl_WT": 1.4915000200271606}
_COUPLING_FUNCS = {
    "GC_10": lambda G: -G,                       # mockup_debug_me.py:24
    "GC_11": lambda G: complex(0, 1) * G,        # mockup_debug_me.py:25
    "GC_12": lambda G: complex(0, 1) * G**2,     # [EXT] models/sm
}


def get_model_param(matrix, param_values=None):
    """Model for a compiled process: the counterpart of the generated `get_model_param(model,
    param_card_path)` (matrix_method_python.inc:36-41) with the SM values instead of a param_card
    read through MG5 (absent offline).  `param_values` overrides masses/widths by name."""
    vals = dict(SM_PARAMS)
    vals.update(param_values or {})
    C = collections.namedtuple("constants", matrix.param_names)
    F = collections.namedtuple("functions", matrix.coupling_names)
    return Model(C(*[vals[n] for n in matrix.param_names]), F(*[_COUPLING_FUNCS[n] for n in matrix.coupling_names]))


def available_processes():
    names = []
    if os.path.isdir(rt.LIBDIR):
        for f in sorted(os.listdir(rt.LIBDIR)):
            if f.startswith("libmfp_") and f.endswith(".so"):
                names.append(f[len("libmfp_"):-3])
    return names


def load_ir(name):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "generated", f"proc_{name}.json")
    if os.path.exists(path):
        from . import process_ir

        return process_ir.loads(open(path).read())
    return None


# where the masses and widths of the built-in (SM) processes sit in an SLHA param_card
PARAM_CARD_ENTRIES = {"mdl_MT": ("MASS", 6), "mdl_WT": ("DECAY", 6), "mdl_MB": ("MASS", 5), "mdl_MZ": ("MASS", 23),
                      "mdl_WZ": ("DECAY", 23), "mdl_MW": ("MASS", 24), "mdl_WW": ("DECAY", 24), "mdl_MH": ("MASS", 25),
                      "mdl_WH": ("DECAY", 25)}


def param_values_from_card(path, names):
    """Masses / widths `names` from an SLHA param_card (reference: get_model_param(model, param_card_path),
    matrix_method_python.inc:36-41, madflow_exec.py:130-136); entries the card does not have keep their defaults."""
    from .param_card import ParamCard

    card = ParamCard(str(path))
    out = {}
    for n in names:
        if n in PARAM_CARD_ENTRIES:
            block, code = PARAM_CARD_ENTRIES[n]
            entry = card.get(block, {}).get(code) if block in card else None
            if entry is not None:
                out[n] = float(entry.value)
    return out


def get_process(name, param_card=None, param_values=None):
    """(Matrix, Model) for one compiled process, e.g. "1_gg_ttx".  param_card: path of an SLHA card whose masses and
    widths replace the built-in SM values; param_values: {name: value} overrides on top."""
    m = Matrix(name, load_ir(name))
    vals = param_values_from_card(param_card, m.param_names) if param_card is not None else {}
    vals.update(param_values or {})
    return m, get_model_param(m, vals)


def get_processes(names=None):
    """All compiled processes as (matrices, models) -- the shape `_import_matrices` returns in the
    reference (scripts/madflow_exec.py:115-138)."""
    names = names or available_processes()
    pairs = [get_process(n) for n in names]
    return [p[0] for p in pairs], [p[1] for p in pairs]
