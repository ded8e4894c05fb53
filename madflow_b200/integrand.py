"""The cross-section integrand of scripts/madflow_exec.py:422-470 evaluated entirely on the device: ONE fused kernel
in the one-event-per-thread flavour (g g > t t~, .. g), a three-kernel pipeline (generate -> matrix element ->
accumulate over an event buffer in HBM, csrc/pipeline_kernels.cuh) in the helicity-parallel flavour.

    sigma-integrand(xrand) = luminosity(x1, x2, q2) * smatrix(ps(xrand); couplings(alpha_s(q2))) * ps_weight

with the parton luminosity and alpha_s from an LHAPDF grid (madflow_b200.pdf, the reference's pdfflow calls) or
luminosity 1 and a frozen / one-loop alpha_s (--no_pdf).

`FusedIntegrand` describes it (process, collider energy, masses, cuts, frame, alpha_s mode) and
`VegasFlow` launches it (include/madflow_b200_process.h: mfp_integrand).  `python_integrand` gives
the same function assembled from the separate API calls, exactly as the reference's
`cross_section` closure is -- it is what the tests compare the fused kernel with.
"""
import ctypes
import math

import torch

from . import _runtime as rt
from . import config
from .phasespace import PhaseSpaceGenerator

MZ = 91.188


def one_loop_b0(nf=5):
    return (33.0 - 2.0 * nf) / (12.0 * math.pi)


def alpha_s_one_loop(q2, alpha_mz=0.118, mz2=MZ * MZ, b0=None):
    """alpha_s(q2) = a/(1 + a*b0*log(q2/mz2)).  Stands in for pdfflow.alphasQ2 (reference:
    scripts/madflow_exec.py:431), which needs an LHAPDF grid that is not available offline."""
    b0 = one_loop_b0() if b0 is None else b0
    return alpha_mz / (1.0 + alpha_mz * b0 * torch.log(q2 / mz2))


class FusedIntegrand:
    def __init__(self, matrix, model, sqrts=13e3, masses=None, pt_cut=None, cuts=None, lab_frame=True,
                 alpha_s=None, running=False, alpha_mz=0.118, mz=MZ, nf=5, pdf=None, fixed_scale=None,
                 initial_states=None, mirror_initial_states=None):
        """alpha_s: frozen value (default: 0.118 as `madflow --no_pdf -q`, madflow_exec.py:379-380);
        running=True: q2 = (sum mT/2)^2 per event (madflow_exec.py:428-431) with one-loop alpha_s.
        pdf: a madflow_b200.pdf.PDF -- the event weight gets the parton luminosity of the process's initial
        states at muF^2 = q2 (madflow_exec.py:410-417, 450-454) and, with running=True, alpha_s comes from the
        set's table (`pdf.alphasQ2`, :431); fixed_scale (GeV): muF = muR fixed and alpha_s frozen at
        `pdf.alphasQ2(fixed_scale^2)` (`madflow -q`, :376-386).  initial_states / mirror_initial_states: the
        flavour pairs of the luminosity when they differ from the ones the process library recorded."""
        self.matrix, self.model = matrix, model
        self._lib = matrix._lib
        n = int(matrix.nexternal)
        self.nexternal = n
        self.n_dim = 4 * (n - 2) + 2
        self.sqrts = float(sqrts)
        self.masses = [float(m) for m in (masses if masses is not None else [0.0] * (n - 2))]
        if len(self.masses) != n - 2:
            raise ValueError("one mass per outgoing particle")
        self.cuts = list(cuts or [])
        if pt_cut is not None:  # madflow_exec.py:392-395
            self.cuts += [("pt", i, float(pt_cut), None) for i in range(2, n)]
        self.lab_frame = bool(lab_frame)
        self.event_sink = None                 # madflow_b200.events.EventSink
        self.max_events_per_launch = 1 << 23   # bounds the HBM scratch of the pipeline flavour (~2 GB)
        self.running = bool(running)
        self.pdf = pdf
        self.fixed_q2 = float(fixed_scale) ** 2 if fixed_scale is not None else 0.0
        if fixed_scale is not None and running:
            raise ValueError("a fixed scale freezes alpha_s: running=True contradicts fixed_scale")
        self.channels = None
        if pdf is not None:
            from .pdf import initial_state_channels

            self.channels = initial_state_channels(matrix, pdf, initial_states, mirror_initial_states)
            self.initial_pairs = [tuple(int(f) for f in pr) for pr in
                                  (initial_states if initial_states is not None else matrix.initial_states)]
            if matrix.mirror_initial_states if mirror_initial_states is None else mirror_initial_states:
                self.initial_pairs += [(b, a) for a, b in list(self.initial_pairs)]
            if not 0 < len(self.channels[0]) <= rt.MFP_MAX_CHANNELS:
                raise ValueError(f"{len(self.channels[0])} initial-state channels (1..{rt.MFP_MAX_CHANNELS} supported): "
                                 "does the process know its initial_states?")
            if fixed_scale is not None and alpha_s is None and pdf.has_alphas:
                alpha_s = float(pdf.alphasQ2([self.fixed_q2]))   # madflow_exec.py:382
        if alpha_s is None:
            alpha_s = 0.118
        if not running and config.get_constants().mode == "reference":
            import numpy as np

            alpha_s = float(np.float32(alpha_s))  # Model.freeze_alpha_s goes through float_me([a]) (parameters.py:53)
        self.alpha_s = float(alpha_mz if running else alpha_s)
        self.mz2, self.b0 = float(mz) ** 2, one_loop_b0(nf)
        consts = list(model._constants)
        self.par = [float(c.item() if isinstance(c, torch.Tensor) else c) for c in consts[: len(matrix.param_names)]]

    def nblocks(self):
        return self._lib.integrand_blocks()

    def _workspace(self, nevents):
        need = self._lib.integrand_workspace(nevents)
        if need == 0:
            return None
        ws = getattr(self, "_ws", None)
        if ws is None or ws.numel() < need:
            self._ws = ws = torch.empty(need, dtype=torch.uint8, device=config.device())
        return ws

    def _args(self):
        a = rt.mfp_integrand_args()
        a.com_sqrts = self.sqrts
        for i, m in enumerate(self.masses):
            a.masses[i] = m
        a.lab_frame = int(self.lab_frame)
        a.ncuts = len(self.cuts)
        for i, (var, particle, lo, hi) in enumerate(self.cuts):
            a.cuts[i] = rt.mf_cut(rt.CUT_VARS[var], rt.cut_particle(var, particle), lo is not None, hi is not None,
                                  float(lo) if lo is not None else 0.0, float(hi) if hi is not None else 0.0)
        k = config.get_constants()
        a.pi, a.acc, a.gev2pb, a.sqh = k.PI, k.ACC, k.GEV2PB, k.SQH
        for i, v in enumerate(self.par):
            a.par[i] = v
        a.alpha_mode = (2 if self.pdf is not None and self.pdf.has_alphas else 1) if self.running else 0
        a.alpha_s, a.mz2, a.b0 = self.alpha_s, self.mz2, self.b0
        a.fixed_q2 = self.fixed_q2
        if self.pdf is not None:
            a.d_pdf = self.pdf.table.data_ptr()
            a.nchannels = len(self.channels[0])
            for i, (c1, c2) in enumerate(zip(*self.channels)):
                a.chan_fl1[i], a.chan_fl2[i] = c1, c2
        return a

    def launch(self, divisions, seed, iteration, first_event, nevents, inv_total, partial, nblocks, train,
               skip_accumulate=False):
        a = self._args()
        a.skip_accumulate = int(bool(skip_accumulate))
        a.d_grid = divisions.data_ptr()
        a.seed, a.iteration, a.first_event, a.nevents = int(seed), int(iteration), int(first_event), int(nevents)
        a.inv_total_events = float(inv_total)
        a.d_partial = partial.data_ptr()
        a.nblocks = int(nblocks)
        a.accumulate_hist = int(bool(train))
        ws = self._workspace(nevents)
        a.d_workspace = ws.data_ptr() if ws is not None else None
        a.workspace_bytes = ws.numel() if ws is not None else 0
        self._lib.integrand(a)
        if self.event_sink is not None:
            self.event_sink.consume(*self.events(nevents), first_event)

    def events(self, nevents):
        """The device event buffer of the last launch of `nevents` events as tensors (views of the workspace):
        momenta (capacity, nexternal, 4), weight, |M|^2, alpha_s (capacity,); the integrand value of slot i is
        me[i] * weight[i], empty slots have weight 0 (include/madflow_b200_process.h: mfp_integrand_events)."""
        ws = getattr(self, "_ws", None)
        if ws is None:
            raise rt.MadflowB200Error("no event buffer: the one-event-per-thread flavour keeps the events on chip; "
                                      "select the helicity-parallel flavour (matrix.set_variant('hp'))")
        v = self._lib.integrand_events(ws.data_ptr(), nevents)
        cap, n = int(v.capacity), self.nexternal

        def view(ptr, count):
            off = int(ptr) - ws.data_ptr()
            return ws[off:off + count * 8].view(torch.float64)

        return (view(v.d_mom, cap * n * 4).view(cap, n, 4), view(v.d_weight, cap), view(v.d_me, cap),
                view(v.d_alpha_s, cap))

    # -- the same integrand from the separate API calls (reference structure, madflow_exec.py:422-470)
    def python_integrand(self):
        n = self.nexternal
        psg = PhaseSpaceGenerator(n, self.sqrts, self.masses, com_output=not self.lab_frame)
        for var, particle, lo, hi in self.cuts:
            psg.register_cut(var, particle=particle, min_val=lo, max_val=hi)
        if not self.running and not self.model.frozen:
            self.model.freeze_alpha_s(self.alpha_s)

        pdf = self.pdf
        if pdf is not None:
            ini = self.initial_pairs
            had1, had2 = [a for a, _ in ini], [b for _, b in ini]

        def cross_section(xrand, n_dim=None, weight=None):
            all_ps, wts, x1, x2, idx = psg(xrand)
            q2array = None
            if self.fixed_q2 > 0.0:
                q2array = torch.full_like(x1, self.fixed_q2)
            elif self.running or pdf is not None:
                full_mt = torch.sum(psg.mt(all_ps[:, 2:n, :]), dim=-1)
                q2array = (full_mt / 2.0) ** 2
            if self.running:
                if pdf is not None and pdf.has_alphas:
                    alpha = pdf.alphasQ2(q2array).reshape(-1)
                else:
                    alpha = alpha_s_one_loop(q2array, self.alpha_s, self.mz2, self.b0)
            else:
                alpha = None
            smatrix = self.matrix.smatrix(all_ps, *self.model.evaluate(alpha))
            if pdf is not None:   # madflow_exec.py:410-417, 450-454
                proton_1 = pdf.xfxQ2(had1, x1, q2array).reshape(-1, len(had1))
                proton_2 = pdf.xfxQ2(had2, x2, q2array).reshape(-1, len(had2))
                smatrix = torch.sum(proton_1 * proton_2, dim=1) / x1 / x2 * smatrix
            ret = smatrix * wts
            if self.cuts:
                out = torch.zeros(xrand.shape[0], dtype=torch.float64, device=ret.device)
                out[idx[:, 0].long()] = ret
                return out
            return ret

        return cross_section


class MultiProcessIntegrand:
    """Several subprocesses of one hadronic process on the same events -- `p p > t t~` = g g > t t~ and
    q q~ > t t~ -- summed per event, each weighted by its own parton luminosity (the loop over `matrices` in the
    reference's cross_section, scripts/madflow_exec.py:444-455).

    Every subprocess runs the generation and matrix-element stages of its own library on the same Philox
    counters, grid and cuts, so slot i of every event buffer holds the same event; one accumulation kernel then
    forms t = sum_p |M|^2_p * w_p (w_p = xjac * phase-space weight * luminosity_p) and the VEGAS sums of t."""

    def __init__(self, integrands):
        self.parts = list(integrands)
        if not self.parts:
            raise ValueError("at least one subprocess")
        first = self.parts[0]
        for fi in self.parts:
            if (fi.nexternal, fi.sqrts, fi.masses, fi.cuts, fi.lab_frame) != (first.nexternal, first.sqrts, first.masses,
                                                                               first.cuts, first.lab_frame):
                raise ValueError("the subprocesses must share the phase space (particles, masses, cuts, frame)")
            if fi.matrix.variant != "hp":
                fi.matrix.set_variant("hp")   # the kernel pipeline that keeps the events in device memory
        if len(self.parts) > 16:
            raise ValueError(f"{len(self.parts)} subprocesses: mf_vegas_accumulate_sum takes at most 16 terms per event "
                             "(ACC_MAX_TERMS in csrc/pipeline_kernels.cuh)")
        self.n_dim, self.nexternal = first.n_dim, first.nexternal
        self.max_events_per_launch = min(fi.max_events_per_launch for fi in self.parts)
        self._common_blocks = None
        self.event_sink = None   # madflow_b200.events.EventSink: sees the momenta and the summed integrand value

    def nblocks(self):
        """One grid size for all subprocess libraries: it fixes how the event buffers are cut into segments, and
        slot i must hold the same event in every buffer."""
        if self._common_blocks is None:
            for fi in self.parts:
                fi._lib.set_integrand_blocks(0)
            # the largest: a library whose kernel keeps fewer blocks resident runs them in waves (its blocks walk fewer
            # segments each), while the smallest would leave the SMs of the small kernels half empty
            self._common_blocks = max(fi.nblocks() for fi in self.parts)
        for fi in self.parts:
            fi._lib.set_integrand_blocks(self._common_blocks)
        return self._common_blocks

    def launch(self, divisions, seed, iteration, first_event, nevents, inv_total, partial, nblocks, train):
        views = []
        nblocks = self.nblocks()
        for fi in self.parts:
            fi.launch(divisions, seed, iteration, first_event, nevents, inv_total, partial, nblocks, train,
                      skip_accumulate=True)
            views.append(fi._lib.integrand_events(fi._ws.data_ptr(), nevents))
        cap = int(views[0].capacity)
        if any(int(v.capacity) != cap for v in views):
            raise rt.MadflowB200Error("the subprocess libraries lay out their event buffers differently")
        n = len(views)
        fptr = (ctypes.c_void_p * n)(*[v.d_me for v in views])
        wptr = (ctypes.c_void_p * n)(*[v.d_weight for v in views])
        lib = rt.core()
        rt.check(lib, lib.mf_vegas_accumulate_sum(n, fptr, wptr, ctypes.c_void_p(views[0].d_bins), ctypes.c_int64(cap),
                                                  self.n_dim, int(bool(train)), rt.ptr(partial), int(nblocks),
                                                  rt.stream_ptr()))
        if self.event_sink is not None:
            # what the reference hands its LHE writer: all_ps and the summed ret * weight (madflow_exec.py:462-464)
            evs = [fi.events(nevents) for fi in self.parts]
            total = evs[0][2] * evs[0][1]
            for mom, w, me, _ in evs[1:]:
                total = total + me * w
            self.event_sink.consume(evs[0][0], None, total, None, first_event)

    def release(self):
        """Give the libraries their own grid sizes back (the override is per library, not per integrand)."""
        for fi in self.parts:
            fi._lib.set_integrand_blocks(0)
        self._common_blocks = None

    def python_integrand(self):
        """The same sum from the separate API calls, as the reference assembles it."""
        parts = [fi.python_integrand() for fi in self.parts]

        def cross_section(xrand, n_dim=None, weight=None):
            ret = parts[0](xrand, n_dim=n_dim, weight=weight)
            for f in parts[1:]:
                ret = ret + f(xrand, n_dim=n_dim, weight=weight)
            return ret

        return cross_section
