"""VEGAS importance sampling as madflow uses it through vegasflow.  TEST INFRASTRUCTURE.

PARITY UNPINNED: vegasflow is a third-party dependency (setup.py:11, unpinned, not under
/root/reference).  This restates the published algorithm (G.P. Lepage, J. Comput. Phys. 27
(1978) 192) with vegasflow 1.x's choices as recalled: 50 bins per dimension, ALPHA = 1.5,
random numbers confined to (1e-8, 1 - 1e-8), per-iteration variance (n*S2 - S1^2)/(n-1),
iterations combined with weights 1/sigma^2.  Anchors in the reference: the call sites
scripts/madflow_exec.py:487-525 (VegasFlow(ndim, n_events, events_limit), compile, run_integration,
freeze_grid, events_per_run), utilities.py:90 (vegas_wrapper) and the integrand signature
f(xrand, n_dim=, weight=) at madflow_exec.py:422.
"""
import numpy as np

from . import philox

BINS_MAX = 50
ALPHA = 1.5
TECH_CUT = 1e-8


def uniform_grid(ndim):
    """Bin edges (ndim, BINS_MAX+1), edges[:,0]=0, edges[:,-1]=1."""
    return np.tile(np.linspace(0.0, 1.0, BINS_MAX + 1), (ndim, 1))


def confine(u):
    """Map [0,1) onto (TECH_CUT, 1-TECH_CUT) as vegasflow's generator does."""
    return TECH_CUT + u * (1.0 - 2.0 * TECH_CUT)


def map_to_grid(rnds, grid):
    """rnds (nevt,ndim) -> x (nevt,ndim), bin index (nevt,ndim), jacobian weight (nevt,).
    vegasflow: xn = BINS*(1-r); k = int(xn); x = lo_k + (hi_k-lo_k)*(xn-k); w = prod BINS*(hi_k-lo_k)."""
    xn = BINS_MAX * (1.0 - rnds)
    k = np.clip(xn.astype(np.int64), 0, BINS_MAX - 1)
    aux = xn - k
    d = np.arange(rnds.shape[1])[None, :]
    lo, hi = grid[d, k], grid[d, k + 1]
    delta = hi - lo
    x = lo + delta * aux
    w = np.prod(delta * BINS_MAX, axis=1)
    return x, k, w


def refine_grid_1d(res2, edges):
    """One dimension of vegasflow's refine_grid: smooth, damp with ALPHA, re-bin to equal weight."""
    padded = np.concatenate([[0.0], res2, [0.0]])
    meaner = np.full(BINS_MAX, 3.0)
    meaner[0] = meaner[-1] = 2.0
    smeared = np.maximum((padded[1:-1] + padded[2:] + padded[:-2]) / meaner, 1e-30)
    sum_t = np.sum(smeared)
    aux = (1.0 - smeared / sum_t) / (np.log(sum_t) - np.log(smeared))
    wei = np.power(aux, ALPHA)
    ave = np.sum(wei) / BINS_MAX
    new = [0.0]
    bin_weight, n_bin, cur, prev = 0.0, -1, 0.0, 0.0
    upper = edges[1:]
    for _ in range(BINS_MAX - 1):
        while bin_weight < ave:
            n_bin += 1
            bin_weight += wei[n_bin]
            prev = cur
            cur = upper[n_bin]
        bin_weight -= ave
        delta = (cur - prev) * bin_weight / wei[n_bin]
        new.append(cur - delta)
    new.append(1.0)
    return np.array(new)


def refine_grid(arr_res2, grid):
    return np.stack([refine_grid_1d(arr_res2[d], grid[d]) for d in range(grid.shape[0])])


def accumulate(values, xjac, k):
    """Per-chunk sums: res = sum(xjac*f), res2 = sum((xjac*f)^2), arr_res2[d,bin] = sum of squares."""
    t = xjac * values
    t2 = t * t
    ndim = k.shape[1]
    arr = np.zeros((ndim, BINS_MAX))
    for d in range(ndim):
        arr[d] = np.bincount(k[:, d], weights=t2, minlength=BINS_MAX)
    return np.sum(t), np.sum(t2), arr


def iteration_result(res, res2, n_events):
    """sigma of one iteration (vegasflow: err2 = max(res2*n - res^2, 1e-30); sigma = sqrt(err2/(n-1)))."""
    err2 = max(res2 * n_events - res * res, 1e-30)
    return res, np.sqrt(err2 / (n_events - 1.0))


def combine(results):
    """Weighted average over iterations -> (result, error, chi2/dof)."""
    r = np.array([x[0] for x in results])
    s = np.array([x[1] for x in results])
    w = 1.0 / s**2
    final = np.sum(r * w) / np.sum(w)
    err = np.sqrt(1.0 / np.sum(w))
    chi2 = np.sum((r - final) ** 2 * w) / max(len(results) - 1, 1)
    return final, err, chi2


class Vegas:
    """vegasflow-style driver over the Philox stream (oracle.philox)."""

    def __init__(self, ndim, n_events, seed=4, chunk=200_000):
        self.ndim, self.n_events, self.seed, self.chunk = ndim, int(n_events), seed, chunk
        self.grid = uniform_grid(ndim)
        self.train = True
        self.iteration = 0
        self.history = []

    def compile(self, integrand):
        self.integrand = integrand

    def freeze_grid(self):
        self.train = False

    def run_iteration(self):
        res = res2 = 0.0
        arr = np.zeros((self.ndim, BINS_MAX))
        for first in range(0, self.n_events, self.chunk):
            n = min(self.chunk, self.n_events - first)
            u = confine(philox.uniforms(self.seed, self.iteration, first, n, self.ndim))
            x, k, w = map_to_grid(u, self.grid)
            xjac = w / self.n_events
            f = self.integrand(x, n_dim=self.ndim, weight=xjac)
            a, b, c = accumulate(f, xjac, k)
            res, res2, arr = res + a, res2 + b, arr + c
        if self.train:
            self.grid = refine_grid(arr, self.grid)
        self.iteration += 1
        out = iteration_result(res, res2, self.n_events)
        self.history.append(out)
        return out

    def run_integration(self, n_iter):
        results = [self.run_iteration() for _ in range(n_iter)]
        final, err, _ = combine(results)
        return final, err


def make_cross_section(ir, params_fn, sqrts, masses, pt_cut=None, const=None, lab_frame=True,
                       alpha_s_fn=None, cuts=(), pdf=None, fixed_q2=None, smatrix_fn=None):
    """The integrand of scripts/madflow_exec.py:422-470: ramboflow -> cuts on COM momenta -> boost ->
    alpha_s(q2=(sum mT/2)^2) or frozen -> luminosity * smatrix * wts, zeros at cut events.
    pdf (oracle.pdf.GridPDF): luminosity = sum over the initial states (+ mirrored) of
    xf_a(x1,q2) xf_b(x2,q2) / x1 / x2 (:410-417, 450-454); None = --no_pdf, luminosity 1 (:437-438)."""
    from . import REFERENCE, matrix as om, phasespace as ps

    const = const or REFERENCE
    n = ir["nexternal"]
    gen = ps.PhaseSpaceGenerator(n, sqrts, masses, com_output=not lab_frame, const=const, xfactor="converged")
    if pt_cut is not None:
        for i in range(2, n):
            gen.register_cut("pt", particle=i, min_val=pt_cut)
    for var, particle, lo, hi in cuts:
        gen.register_cut(var, particle=particle, min_val=lo, max_val=hi)

    def cross_section(xrand, n_dim=None, weight=None):
        all_ps, wts, x1, x2, idx = gen(xrand)
        ret = np.zeros(xrand.shape[0])
        if all_ps.shape[0] == 0:
            return ret
        q2 = None
        if fixed_q2:
            q2 = np.full_like(x1, fixed_q2)
        elif alpha_s_fn is not None or pdf is not None:
            full_mt = np.sum(ps.mt(all_ps[:, 2:n, :]), axis=-1)
            q2 = (full_mt / 2.0) ** 2
        params = params_fn(alpha_s_fn(q2)) if alpha_s_fn is not None else params_fn(None)
        val = (smatrix_fn or om.smatrix)(ir, all_ps, params, const)   # smatrix_fn: om.smatrix_recycled for long lists
        if pdf is not None:
            ini = [tuple(pr) for pr in ir["initial_states"]]
            if ir.get("mirror_initial_states"):
                ini += [(b, a) for a, b in ini]
            p1 = pdf.xfxQ2([a for a, _ in ini], x1, q2)
            p2 = pdf.xfxQ2([b for _, b in ini], x2, q2)
            val = np.sum(p1 * p2, axis=1) / x1 / x2 * val
        val = val * wts
        if pt_cut is not None or cuts:
            ret[idx[:, 0]] = val
        else:
            ret = val
        return ret

    return cross_section
