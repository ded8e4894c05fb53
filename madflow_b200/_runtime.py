"""ctypes bindings of the C-ABI libraries (include/madflow_b200.h, include/madflow_b200_process.h).

PyTorch is used for device memory and streams only: tensors are handed to the kernels as raw
device pointers (`data_ptr()`) together with the current CUDA stream.  There is no fallback: a
missing library or a missing GPU raises.
"""
import ctypes
import os

import numpy as np
import torch

from . import config

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")

c_dp = ctypes.c_void_p
MFP_MAX_PARAMS, MFP_MAX_COUPLINGS, MFP_MAX_OUT, MFP_MAX_CUTS, MFP_MAX_CHANNELS = 8, 8, 8, 16, 32
LAYOUT_AOS, LAYOUT_SOA = 0, 1
CUT_VARS = {"pt": 0, "mt": 1, "mt2": 2, "mij": 3, "dr": 4}
PAIR_CUTS = ("mij", "dr")   # extension: cuts on a pair of particles, particle = (i, j)


def cut_particle(variable, particle):
    """The `particle` field of mf_cut: an index, or i + 256 * j for the pair variables."""
    if variable in PAIR_CUTS:
        i, j = particle
        return int(i) + 256 * int(j)
    return int(particle)


class MadflowB200Error(RuntimeError):
    pass


class mfp_info(ctypes.Structure):
    _fields_ = [
        ("name", ctypes.c_char * 64),
        ("nexternal", ctypes.c_int32), ("ninitial", ctypes.c_int32), ("ncomb", ctypes.c_int32),
        ("ncolor", ctypes.c_int32), ("ndiags", ctypes.c_int32), ("namps", ctypes.c_int32),
        ("nwavefuncs", ctypes.c_int32), ("nparams", ctypes.c_int32), ("ncouplings", ctypes.c_int32),
        ("ndim", ctypes.c_int32), ("block_threads", ctypes.c_int32),
        ("denominator", ctypes.c_double), ("flops_per_event", ctypes.c_double),
    ]


class mf_cut(ctypes.Structure):
    _fields_ = [("var", ctypes.c_int32), ("particle", ctypes.c_int32), ("has_min", ctypes.c_int32),
                ("has_max", ctypes.c_int32), ("vmin", ctypes.c_double), ("vmax", ctypes.c_double)]


class mf_ps_const(ctypes.Structure):
    _fields_ = [("pi", ctypes.c_double), ("acc", ctypes.c_double), ("gev2pb", ctypes.c_double)]


class mfp_integrand_args(ctypes.Structure):
    _fields_ = [
        ("d_grid", ctypes.c_void_p), ("seed", ctypes.c_uint64), ("iteration", ctypes.c_uint32),
        ("first_event", ctypes.c_uint64), ("nevents", ctypes.c_int64), ("inv_total_events", ctypes.c_double),
        ("com_sqrts", ctypes.c_double), ("masses", ctypes.c_double * MFP_MAX_OUT),
        ("lab_frame", ctypes.c_int32), ("ncuts", ctypes.c_int32), ("cuts", mf_cut * MFP_MAX_CUTS),
        ("pi", ctypes.c_double), ("acc", ctypes.c_double), ("gev2pb", ctypes.c_double),
        ("par", ctypes.c_double * MFP_MAX_PARAMS), ("alpha_mode", ctypes.c_int32), ("alpha_s", ctypes.c_double),
        ("mz2", ctypes.c_double), ("b0", ctypes.c_double), ("sqh", ctypes.c_double),
        ("d_partial", ctypes.c_void_p), ("nblocks", ctypes.c_int32), ("accumulate_hist", ctypes.c_int32),
        ("d_workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
        ("d_pdf", ctypes.c_void_p), ("nchannels", ctypes.c_int32),
        ("chan_fl1", ctypes.c_int8 * MFP_MAX_CHANNELS), ("chan_fl2", ctypes.c_int8 * MFP_MAX_CHANNELS),
        ("fixed_q2", ctypes.c_double), ("skip_accumulate", ctypes.c_int32),
    ]


class mfp_event_view(ctypes.Structure):
    _fields_ = [("d_mom", ctypes.c_void_p), ("d_weight", ctypes.c_void_p), ("d_me", ctypes.c_void_p),
                ("d_alpha_s", ctypes.c_void_p), ("capacity", ctypes.c_int64), ("d_bins", ctypes.c_void_p)]


def _require_cuda():
    if not torch.cuda.is_available():
        raise MadflowB200Error("no CUDA device visible: madflow_b200 has no CPU implementation")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def to_device(x, dtype=torch.float64):
    """numpy / torch-CPU / torch-CUDA -> contiguous CUDA tensor of `dtype` on the current device."""
    dev = config.device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(dev).contiguous()


_core = None


def core():
    """libmadflow_b200.so (process-independent kernels)."""
    global _core
    if _core is None:
        path = os.path.join(LIBDIR, "libmadflow_b200.so")
        if not os.path.exists(path):
            raise MadflowB200Error(f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = ctypes.CDLL(path)
        lib.mf_last_error.restype = ctypes.c_char_p
        _core = lib
    return _core


def check(lib, rc, errfn="mf_last_error"):
    if rc != 0:
        fn = getattr(lib, errfn)
        fn.restype = ctypes.c_char_p
        raise MadflowB200Error((fn() or b"unknown error").decode())


class ProcessLib:
    """One libmfp_<process>.so."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise MadflowB200Error(f"process library {path} not found (build it with madflow_b200.build)")
        self.path = path
        self.lib = ctypes.CDLL(path)
        self.lib.mfp_last_error.restype = ctypes.c_char_p
        self.lib.mfp_param_name.restype = ctypes.c_char_p
        self.lib.mfp_coupling_name.restype = ctypes.c_char_p
        info = mfp_info()
        self._check(self.lib.mfp_get_info(ctypes.byref(info)))
        self.info = info
        self.name = info.name.decode()
        self.param_names = [self.lib.mfp_param_name(i).decode() for i in range(info.nparams)]
        self.coupling_names = [self.lib.mfp_coupling_name(i).decode() for i in range(info.ncouplings)]
        self.coupling_defs = []
        for i in range(info.ncouplings):
            re, im, pw = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
            self._check(self.lib.mfp_coupling_def(i, ctypes.byref(re), ctypes.byref(im), ctypes.byref(pw)))
            self.coupling_defs.append((re.value, im.value, pw.value))
        self.helicities = [[self.lib.mfp_helicity(ic, leg) for leg in range(info.nexternal)]
                           for ic in range(info.ncomb)]
        if os.environ.get("MADFLOW_B200_VARIANT"):
            self.set_variant(os.environ["MADFLOW_B200_VARIANT"])

    def _check(self, rc):
        if rc != 0:
            raise MadflowB200Error((self.lib.mfp_last_error() or b"unknown error").decode())

    def set_variant(self, variant):
        """Kernel flavour: 0/'default', 1/'thread' (one event per thread), 2/'hp' (helicity-parallel)."""
        v = {"default": 0, "thread": 1, "hp": 2}.get(variant, variant)
        self._check(self.lib.mfp_set_variant(int(v)))

    @property
    def variant(self):
        return {1: "thread", 2: "hp"}[int(self.lib.mfp_get_variant())]

    def smatrix(self, d_p, layout, nevt, par, d_coup, coup_stride, sqh, d_out, only_comb=None):
        _require_cuda()
        par_arr = (ctypes.c_double * MFP_MAX_PARAMS)(*([float(v) for v in par] + [0.0] * (MFP_MAX_PARAMS - len(par))))
        if only_comb is None:
            rc = self.lib.mfp_smatrix(ptr(d_p), layout, ctypes.c_int64(nevt), par_arr, ptr(d_coup),
                                      ctypes.c_int64(coup_stride), ctypes.c_double(sqh), ptr(d_out), stream_ptr())
        else:
            rc = self.lib.mfp_matrix_hel(ptr(d_p), layout, ctypes.c_int64(nevt), int(only_comb), par_arr,
                                         ptr(d_coup), ctypes.c_int64(coup_stride), ctypes.c_double(sqh),
                                         ptr(d_out), stream_ptr())
        self._check(rc)

    def smatrix_host(self, h_p, layout, par, h_coup, coup_stride, sqh, h_out):
        """numpy in / numpy out through the host-buffer entry point (copies inside the call)."""
        _require_cuda()
        nevt = h_out.shape[0]
        par_arr = (ctypes.c_double * MFP_MAX_PARAMS)(*([float(v) for v in par] + [0.0] * (MFP_MAX_PARAMS - len(par))))
        rc = self.lib.mfp_smatrix_host(ctypes.c_void_p(h_p.ctypes.data), layout, ctypes.c_int64(nevt), par_arr,
                                       ctypes.c_void_p(h_coup.ctypes.data), ctypes.c_int64(coup_stride),
                                       ctypes.c_double(sqh), ctypes.c_void_p(h_out.ctypes.data))
        self._check(rc)

    def integrand_blocks(self):
        _require_cuda()
        return int(self.lib.mfp_integrand_blocks())

    def set_integrand_blocks(self, nblocks):
        """Fix the grid size / event-buffer segmentation of the helicity-parallel integrand (0 = automatic)."""
        self._check(self.lib.mfp_set_integrand_blocks(int(nblocks)))

    def integrand_workspace(self, nevents):
        self.lib.mfp_integrand_workspace.restype = ctypes.c_int64
        return int(self.lib.mfp_integrand_workspace(ctypes.c_int64(int(nevents))))

    def integrand(self, args):
        _require_cuda()
        self._check(self.lib.mfp_integrand(ctypes.byref(args), stream_ptr()))

    def integrand_events(self, d_workspace, nevents):
        """Device pointers of the events of the last integrand call on this workspace (mfp_event_view)."""
        view = mfp_event_view()
        self._check(self.lib.mfp_integrand_events(ctypes.c_void_p(d_workspace), ctypes.c_int64(int(nevents)),
                                                  ctypes.byref(view)))
        return view


_process_cache = {}


def process_lib(name, path=None):
    path = path or os.path.join(LIBDIR, f"libmfp_{name}.so")
    if path not in _process_cache:
        _process_cache[path] = ProcessLib(path)
    return _process_cache[path]
