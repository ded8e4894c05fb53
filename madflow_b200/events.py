"""Event output of the fused integrand: histograms and unweighted events, filled on the device.

The reference produces events by calling back into Python from inside the integrand
(scripts/madflow_exec.py:462-464 -> lhe_writer.LheWriter.lhe_parser, python_package/madflow/lhe_writer.py:
151-208) and histograms the LHE file afterwards (example/compare_mg5_hists.py:16-57).  Here an `EventSink`
attached to a `FusedIntegrand` sees, after every launch, the event buffer of the kernel pipeline in device
memory (include/madflow_b200_process.h: mfp_integrand_events):

    sink = EventSink(integrand, histograms=[Histogram("pt", 2, 0.0, 300.0, 50), Histogram("eta", 2, -4.0, 4.0, 50)],
                     unweight=True, capacity=100_000)
    vegas.run_integration(n)                  # histograms accumulate, unweighted events are kept
    sink.histograms[0].values(n_iterations)   # d(sigma)/bin in pb
    with LheWriter(folder, "run_01", no_unweight=True) as w: sink.write_lhe(w)

Only the events that survive the unweighting leave the GPU.
"""
import ctypes

import numpy as np
import torch

from . import _runtime as rt
from . import config

OBSERVABLES = {"pt": 0, "eta": 1, "pseudorapidity": 1, "rapidity": 2, "y": 2, "energy": 3, "E": 3, "mass": 4}


class Histogram:
    """Weighted 1-d histogram of one particle's observable with under/overflow, accumulated on the device."""

    def __init__(self, observable, particle, lo, hi, nbins=50):
        if observable not in OBSERVABLES:
            raise ValueError(f"observable must be one of {sorted(OBSERVABLES)}")
        self.observable, self.particle = observable, int(particle)
        self.lo, self.hi, self.nbins = float(lo), float(hi), int(nbins)
        self._hist = torch.zeros(self.nbins + 2, dtype=torch.float64, device=config.device())

    @property
    def edges(self):
        return np.linspace(self.lo, self.hi, self.nbins + 1)

    def fill(self, mom, w1, w2=None):
        """mom (nevt, nexternal, 4) and the weight factors w1 (* w2) as CUDA float64 tensors."""
        lib = rt.core()
        nevt, nexternal = int(mom.shape[0]), int(mom.shape[1])
        rt.check(lib, lib.mf_event_histogram(rt.ptr(mom), rt.ptr(w1), rt.ptr(w2), ctypes.c_int64(nevt), nexternal,
                                             self.particle, OBSERVABLES[self.observable], ctypes.c_double(self.lo),
                                             ctypes.c_double(self.hi), self.nbins, rt.ptr(self._hist), rt.stream_ptr()))

    def allreduce(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self._hist, op=dist.ReduceOp.SUM)

    def values(self, n_iterations=1, with_overflow=False):
        """Sum of weights per bin divided by the number of iterations that filled it (each VEGAS iteration
        is an independent estimate of the cross section, so this is d(sigma) per bin)."""
        h = self._hist.cpu().numpy() / float(n_iterations)
        return h if with_overflow else h[1:-1]

    def reset(self):
        self._hist.zero_()


class EventSink:
    """Consumes the device event buffer after every launch of a FusedIntegrand."""

    # weight spectrum for the unweighting threshold: sum |w| per logarithmic bin, 4 bins per octave from 2^-120 to 2^40
    _WBINS, _WLOG0, _WPER = 640, -120.0, 4.0

    def __init__(self, integrand, histograms=(), unweight=False, capacity=1_000_000, seed=1234, wmax=None, wmax_scale=8.0,
                 collect_only=False, tail_share=0.1):
        """unweight: keep slot i with probability min(1, |w_i| / wmax) and weight sign(w_i) max(|w_i|, wmax) -- an
        unbiased sample in which all events but the few above wmax carry the same weight; the events above wmax keep
        their own, larger weight (write_lhe scales it, it is never clipped).  wmax: ONE threshold for the whole
        sample -- given, or frozen by freeze_threshold() from the weight statistics collected so far,
        min(largest weight seen, wmax_scale * sum w^2 / sum |w|); with the heavy-tailed weights of a flat phase space
        the largest weight alone would make the efficiency collapse.  collect_only: only gather the weight statistics
        (no histograms, no selection) until freeze_threshold() -- what madflow_exec does during the last warm-up
        iteration.  Without a frozen threshold the sink falls back to the running estimate, which changes from launch
        to launch (a mixture of thresholds: still unbiased event by event through the weights, but not one sample).
        tail_share: freeze_threshold() puts the threshold where the events above it carry this share of sum |w| (from the
        weight spectrum collected so far, summed over ranks), if that is below the other estimate: with the heavy-tailed
        weights of a flat phase space the yield rises by orders of magnitude, at the price of a sample in which
        `tail_share` of the weight sits in a few events that keep their own larger weight."""
        self.integrand = integrand
        self.histograms = list(histograms)
        self.unweight = bool(unweight)
        self.capacity = int(capacity)
        self.seed = int(seed)
        self.wmax = float(wmax) if wmax else None
        self.wmax_scale = float(wmax_scale)
        self.enabled = True
        self.collect_only = bool(collect_only)
        self.tail_share = float(tail_share)
        dev = config.device()
        n = integrand.nexternal
        self._nstat = int(rt.core().mf_weight_stats_blocks())
        self._stat_partial = torch.zeros((self._nstat, 3), dtype=torch.float64, device=dev)
        self._stats = torch.zeros(3, dtype=torch.float64, device=dev)   # max |w|, sum |w|, sum w^2
        self._wspec = torch.zeros(self._WBINS, dtype=torch.float64, device=dev)
        self._wedges = torch.pow(2.0, self._WLOG0 + torch.arange(self._WBINS + 1, dtype=torch.float64, device=dev) / self._WPER)
        if self.unweight:
            self._mom = torch.empty((self.capacity, n, 4), dtype=torch.float64, device=dev)
            self._w = torch.empty(self.capacity, dtype=torch.float64, device=dev)
            self._idx = torch.empty(self.capacity, dtype=torch.int64, device=dev)
            self._count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.launches = 0
        integrand.event_sink = self

    # called by FusedIntegrand.launch
    def consume(self, mom, weight, me, alpha_s, first_event):
        if not self.enabled:
            return
        lib = rt.core()
        nslots = int(mom.shape[0])
        if not self.collect_only:
            for h in self.histograms:
                h.fill(mom, me, weight)
        if self.unweight:
            wmax = self.wmax
            if wmax is None and self.launches > 0:
                wmax = self.threshold()
            if wmax and not self.collect_only:
                # global slot index: unique per (iteration, rank, launch) as long as capacities do not change
                rt.check(lib, lib.mf_select_events(rt.ptr(mom), rt.ptr(me), rt.ptr(weight), ctypes.c_int64(nslots),
                                                   int(mom.shape[1]), ctypes.c_double(wmax), ctypes.c_uint64(self.seed),
                                                   ctypes.c_uint64(2 * int(first_event) + (self.launches << 40)),
                                                   rt.ptr(self._mom), rt.ptr(self._w), rt.ptr(self._idx),
                                                   rt.ptr(self._count), ctypes.c_int64(self.capacity), rt.stream_ptr()))
            rt.check(lib, lib.mf_weight_stats(rt.ptr(me), rt.ptr(weight), ctypes.c_int64(nslots), rt.ptr(self._stat_partial),
                                              self._nstat, rt.stream_ptr()))
            self._stats[0] = torch.maximum(self._stats[0], self._stat_partial[:, 0].max())
            self._stats[1:] += self._stat_partial[:, 1:].sum(dim=0)
            if self.collect_only and self.wmax is None:
                # weight spectrum, in an order that does not depend on scheduling: sort, prefix sums, differences at the edges
                a = torch.sort((me if weight is None else me * weight).abs()).values
                cs = torch.cat([a.new_zeros(1), torch.cumsum(a, 0)])
                at = torch.searchsorted(a, self._wedges)
                self._wspec += cs[at[1:]] - cs[at[:-1]]
        self.launches += 1

    @property
    def max_weight(self):
        return float(self._stats[0].item())

    def threshold(self):
        """The unweighting threshold the next launch will use (see __init__)."""
        if self.wmax is not None:
            return self.wmax
        mx, s1, s2 = self._stats.tolist()
        return min(mx, self.wmax_scale * s2 / s1) if s1 > 0.0 else 0.0

    def freeze_threshold(self):
        """Fix the unweighting threshold -- one value for all ranks, from the statistics of all of them -- and start filling
        histograms / keeping events."""
        if self.unweight and self.wmax is None:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(self._stats[:1], op=dist.ReduceOp.MAX)
                dist.all_reduce(self._stats[1:], op=dist.ReduceOp.SUM)
                dist.all_reduce(self._wspec, op=dist.ReduceOp.SUM)
            wmax = self.threshold()
            spec = self._wspec.cpu().numpy()
            total = float(spec.sum())
            if total > 0.0 and 0.0 < self.tail_share < 1.0:
                above = total - np.cumsum(spec)                 # weight above the upper edge of every bin
                k = int(np.argmax(above <= self.tail_share * total))
                wmax = min(wmax, float(self._wedges[k + 1].item())) if wmax else float(self._wedges[k + 1].item())
            self.wmax = wmax or None
        self.collect_only = False
        return self.wmax

    def overweight(self):
        """(fraction of the kept events above the threshold, their share of the sample's total |weight|)."""
        _, w = self.events()
        if len(w) == 0 or not self.wmax:
            return 0.0, 0.0
        over = np.abs(w) > self.wmax * (1.0 + 1e-12)
        return float(np.mean(over)), float(np.sum(np.abs(w[over])) / np.sum(np.abs(w)))

    def events(self):
        """(momenta (n, nexternal, 4), weights (n,)) of the kept events as numpy arrays, in global-index order."""
        if not self.unweight:
            raise RuntimeError("EventSink(unweight=True) keeps events")
        n = min(int(self._count.item()), self.capacity)
        order = torch.argsort(self._idx[:n])
        return self._mom[:n][order].cpu().numpy(), self._w[:n][order].cpu().numpy()

    @property
    def overflowed(self):
        return self.unweight and int(self._count.item()) > self.capacity

    def write_lhe(self, writer, cross=None):
        """Write the kept events through an LheWriter.  cross: the integrated cross section -- an event at the threshold
        gets weight +-cross (as the reference assigns after unweighting, lhe_writer.py:339-341), an event above it
        +-cross * |w| / wmax: the excess weight of the tail is kept, so the sample stays unbiased.  None: own weights."""
        mom, w = self.events()
        if cross is not None:
            scale = np.maximum(np.abs(w) / self.wmax, 1.0) if self.wmax else 1.0
            w = np.sign(w) * float(cross) * scale
        writer.lhe_parser(mom, w)
        return len(w)
