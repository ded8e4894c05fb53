"""VEGAS integrator with the vegasflow interface madflow drives
(scripts/madflow_exec.py:487-525, utilities.py:90):

    vegas = VegasFlow(ndim, events_per_iteration, events_limit=...)
    vegas.compile(integrand)              # f(xrand, n_dim=, weight=) -> (nevents,) tensor
    res, err = vegas.run_integration(n_iter)
    vegas.freeze_grid(); vegas.events_per_run = ...
    vegas_wrapper(integrand, ndim, n_iter, n_events)

vegasflow itself is a third-party dependency that is not part of the reference tree; the
algorithm here is its published one (Lepage 1978; 50 bins per dimension, alpha = 1.5, see
DESIGN.md "VEGAS") on a Philox4x32-10 counter stream (csrc/philox.cuh).

Two kinds of integrand:
  * any Python callable on torch CUDA tensors: sample kernel -> callable -> accumulate kernel;
  * a `FusedIntegrand` (madflow_b200.integrand): the whole event pipeline including the matrix
    element runs on the device (one persistent kernel for the one-event-per-thread flavour, the three-kernel
    pipeline generate -> matrix element -> accumulate over an event buffer in HBM for the helicity-parallel
    flavour) and only the per-block sums come back.

Every sum of an iteration is formed in an order that depends on the launch shape only (csrc/vegas.cuh::warp_hist_add,
fixed-order block reduction): two runs with the same seed give bit-identical grids.

Multi-GPU (one process per GPU, torch.distributed): each rank takes a contiguous slice of the
iteration's global event range -- the Philox counter is the global event index, so the sample set
is independent of the number of ranks -- and ONE all-reduce of the 2 + ndim*50 accumulators per
iteration merges the ranks; every rank then refines its (identical) grid copy.
"""
import ctypes
import logging
import math
import time

import torch

from . import _runtime as rt
from . import config

logger = logging.getLogger("madflow")

BINS_MAX = 50
HEADER = 4  # accumulators: sum t, sum t^2, events that reached the matrix element, reserved
ALPHA = 1.5
TECH_CUT = 1e-8


def shard_events(n_events, rank, world):
    """Contiguous split of the global event range [0, n_events) over ranks: (first, count)."""
    base, rem = divmod(int(n_events), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def iteration_sigma(res, res2, n_events):
    """Standard error of one iteration: sqrt(max(n*S2 - S1^2, 1e-30)/(n-1))."""
    err2 = max(res2 * n_events - res * res, 1e-30)
    return math.sqrt(err2 / (n_events - 1.0))


def combine_iterations(results):
    """1/sigma^2-weighted average of the iterations -> (result, error, chi2/dof)."""
    wsum = sum(1.0 / s**2 for _, s in results)
    final = sum(r / s**2 for r, s in results) / wsum
    err = math.sqrt(1.0 / wsum)
    chi2 = sum((r - final) ** 2 / s**2 for r, s in results) / max(len(results) - 1, 1)
    return final, err, chi2


class _Range:
    """NVTX range around a phase of the iteration when MADFLOW_B200_NVTX=1 (shows up in Nsight Systems timelines)."""
    enabled = None

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Range.enabled is None:
            import os

            _Range.enabled = os.environ.get("MADFLOW_B200_NVTX", "0") == "1" and torch.cuda.is_available()
        if _Range.enabled:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Range.enabled:
            torch.cuda.nvtx.range_pop()
        return False


def _dist():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def allreduce_sums(sums):
    """The one collective of the path: sum the accumulators over ranks (NCCL on GPU tensors)."""
    dist, _, world = _dist()
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    return sums


class VegasFlow:
    def __init__(self, n_dim, n_events, train=True, events_limit=None, seed=4, **_):
        if n_dim > 32:
            raise ValueError("at most 32 dimensions")
        self.n_dim = int(n_dim)
        self.n_events = int(n_events)
        self.train = bool(train)
        self.events_limit = int(events_limit) if events_limit else None
        self.seed = int(seed)
        self.iteration = 0
        self.integrand = None
        self.history = []
        self.timings = []
        dev = config.device()
        edges = torch.linspace(0.0, 1.0, BINS_MAX + 1, dtype=torch.float64)
        self.divisions = edges.repeat(self.n_dim, 1).contiguous().to(dev)
        self._sums = torch.zeros(HEADER + self.n_dim * BINS_MAX, dtype=torch.float64, device=dev)
        self._partial = None

    # -- vegasflow surface
    @property
    def events_per_run(self):
        return self.n_events

    @events_per_run.setter
    def events_per_run(self, val):
        self.n_events = int(val)

    def freeze_grid(self):
        self.train = False

    def unfreeze_grid(self):
        self.train = True

    def compile(self, integrand, compilable=True):
        self.integrand = integrand

    def save_grid(self, path=None):
        """Checkpoint: the grid, the Philox position (seed, iteration) and the iteration history.  With `path` the
        state is also written as an .npz file; `load_grid(path)` in a later process resumes the very same sequence of
        samples (the random numbers are counters: key = seed, counter = (event, iteration, dimension))."""
        state = {"divisions": self.divisions.cpu().clone(), "iteration": self.iteration, "seed": self.seed,
                 "history": list(self.history), "train": self.train, "n_events": self.n_events}
        if path is not None:
            import numpy as np

            np.savez(path, divisions=state["divisions"].numpy(), iteration=self.iteration, seed=self.seed,
                     history=np.asarray(self.history, dtype=np.float64).reshape(-1, 2), train=self.train,
                     n_events=self.n_events)
        return state

    def load_grid(self, state):
        """Restore a checkpoint: a dict from save_grid() or the path of its .npz file."""
        if isinstance(state, (str, bytes)) or hasattr(state, "__fspath__"):
            import numpy as np

            with np.load(state) as z:
                state = {"divisions": torch.from_numpy(z["divisions"].copy()), "iteration": int(z["iteration"]),
                         "seed": int(z["seed"]), "history": [tuple(map(float, r)) for r in z["history"]],
                         "train": bool(z["train"]), "n_events": int(z["n_events"])}
        if tuple(state["divisions"].shape) != tuple(self.divisions.shape):
            raise ValueError(f"checkpoint grid has shape {tuple(state['divisions'].shape)}, this integrator {tuple(self.divisions.shape)}")
        self.divisions.copy_(state["divisions"].to(self.divisions.device))
        self.iteration, self.seed, self.history = int(state["iteration"]), int(state["seed"]), list(state["history"])
        if "train" in state:
            self.train = bool(state["train"])
        if "n_events" in state:
            self.n_events = int(state["n_events"])

    # -- one iteration
    def _partial_buf(self, nblocks):
        need = nblocks * (HEADER + self.n_dim * BINS_MAX)
        if self._partial is None or self._partial.numel() < need:
            self._partial = torch.empty(need, dtype=torch.float64, device=self.divisions.device)
        return self._partial

    def _run_chunk_generic(self, first, n):
        lib = rt.core()
        dev = self.divisions.device
        x = torch.empty((n, self.n_dim), dtype=torch.float64, device=dev)
        xjac = torch.empty(n, dtype=torch.float64, device=dev)
        bins = torch.empty((self.n_dim, n), dtype=torch.uint8, device=dev)
        rt.check(lib, lib.mf_vegas_sample(rt.ptr(self.divisions), self.n_dim, ctypes.c_uint64(self.seed),
                                          ctypes.c_uint32(self.iteration), ctypes.c_uint64(first),
                                          ctypes.c_int64(n), ctypes.c_double(1.0 / self.n_events), rt.ptr(x),
                                          rt.ptr(xjac), rt.ptr(bins), rt.stream_ptr()))
        f = self.integrand(x, n_dim=self.n_dim, weight=xjac)
        f = rt.to_device(f).reshape(-1)
        if f.numel() != n:
            raise ValueError("the integrand must return one value per event")
        nblocks = int(lib.mf_vegas_blocks())
        partial = self._partial_buf(nblocks)
        rt.check(lib, lib.mf_vegas_accumulate(rt.ptr(f), rt.ptr(xjac), rt.ptr(bins), ctypes.c_int64(n), self.n_dim,
                                              int(self.train), rt.ptr(partial), nblocks, rt.stream_ptr()))
        rt.check(lib, lib.mf_vegas_reduce(rt.ptr(partial), nblocks, self.n_dim, 1, rt.ptr(self._sums),
                                          rt.stream_ptr()))

    def _run_chunk_fused(self, first, n):
        lib = rt.core()
        nblocks = self.integrand.nblocks()
        partial = self._partial_buf(nblocks)
        self.integrand.launch(self.divisions, self.seed, self.iteration, first, n, 1.0 / self.n_events, partial,
                              nblocks, self.train)
        rt.check(lib, lib.mf_vegas_reduce(rt.ptr(partial), nblocks, self.n_dim, 1, rt.ptr(self._sums),
                                          rt.stream_ptr()))

    def run_iteration(self):
        if self.integrand is None:
            raise RuntimeError("compile(integrand) first")
        fused = hasattr(self.integrand, "launch")
        if fused and self.integrand.n_dim != self.n_dim:
            raise ValueError(f"the integrand needs {self.integrand.n_dim} dimensions, VegasFlow has {self.n_dim}")
        _, rank, world = _dist()
        first, count = shard_events(self.n_events, rank, world)
        limit = self.events_limit or (min(count, getattr(self.integrand, "max_events_per_launch", count)) if fused
                                      else 10_000_000)
        self._sums.zero_()
        done = 0
        while done < count:
            n = min(limit, count - done)
            with _Range(f"vegas chunk {first + done}+{n}"):
                if fused:
                    self._run_chunk_fused(first + done, n)
                else:
                    self._run_chunk_generic(first + done, n)
            done += n
        with _Range("vegas all-reduce"):
            allreduce_sums(self._sums)
        if self.train:
            lib = rt.core()
            with _Range("vegas refine"):
                rt.check(lib, lib.mf_vegas_refine(rt.ptr(self.divisions), rt.ptr(self._sums), self.n_dim,
                                                  rt.stream_ptr()))
        res, res2, n_me, _ = self._sums[:HEADER].tolist()  # the only device->host read of the iteration
        self.last_me_events = int(n_me)
        self.iteration += 1
        out = (res, iteration_sigma(res, res2, self.n_events))
        self.history.append(out)
        return out

    def run_integration(self, n_iter, log_time=True, histograms=None):
        results = []
        for i in range(int(n_iter)):
            t0 = time.time()
            res, sigma = self.run_iteration()
            dt = time.time() - t0
            self.timings.append(dt)
            results.append((res, sigma))
            if log_time:
                logger.info("Result for iteration %d: %.4f +/- %.4f(took %.5f s)", i, res, sigma, dt)
        final, err, chi2 = combine_iterations(results)
        logger.info(" > Final results: %g +/- %g", final, err)
        self.last_chi2 = chi2
        return final, err


def vegas_wrapper(integrand, n_dim, n_iter, total_n_events, **kwargs):
    """utilities.py:90 / vegasflow.vegas_wrapper."""
    v = VegasFlow(n_dim, total_n_events, **kwargs)
    v.compile(integrand)
    return v.run_integration(n_iter)
