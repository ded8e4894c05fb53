"""Colour-reduced recursion for  g g > t t~ + k g : the evaluation PLAN of the helicity-parallel kernels.

The reference evaluates MG5's diagram list: one HELAS call per sub-diagram, one amplitude per diagram,
JAMP_f = sum_a c_fa amp_a  (madgraph_plugin/template_files/matrix_method_python.inc:106-138,
PyOut_exporter.py:334-375).  Everything in that list is multilinear, so sub-diagrams over the same legs
that carry linearly DEPENDENT colour factors need not be kept apart:

  currents    the raw currents over one leg set S (one per sub-tree, procgen.Generator.currents) span a
              colour space of dimension r <= their number (Jacobi identity: the three structures of a
              three-gluon current span 2 dimensions; a quark current with two gluons: 2 orderings for 3
              sub-trees).  They are replaced by r BASIS currents  B_j = sum_i m_ij raw_i  -- all terms share
              one propagator -- and everything downstream is built from the basis currents only.
  amplitudes  a diagram closes at its centroid vertex: amp = x . Q(rest of the vertex), x = the heaviest
              branch (for g g > t t~ g g: a 3-leg current against the vertex NUMERATOR over the other three
              legs, or a 2-leg current against a 4-leg numerator).  The numerators over one leg set are
              reduced to a colour basis the same way; a row of the amplitude buffer is (x, Q_j) and its JAMP
              coefficients are the colour of that contraction.

This is the Berends-Giele organisation of the same sum of diagrams; it changes the ORDER of additions only
(the kernel still matches the diagram-by-diagram oracle to 1e-12, tests/test_procgen.py, tests/test_gpu_parity.py).
g g > t t~ g g: 54 -> 41 currents, 159 amplitudes -> 67 rows, 672 -> ~250 JAMP terms; the amplitude, JAMP and
pair-object phases of the kernel shrink accordingly.
"""
import itertools

import numpy as np

from .procgen import Generator, Node


def _round_gauss(z, what):
    """Complex coefficient -> exact small Gaussian rational (the QCD colour algebra only produces those)."""
    from fractions import Fraction

    re = Fraction(float(z.real)).limit_denominator(24)
    im = Fraction(float(z.imag)).limit_denominator(24)
    assert abs(complex(re, im) - z) < 1e-9, f"{what}: coefficient {z} is not a small Gaussian rational"
    return complex(float(re), float(im))


def reduce_colours(cols, what=""):
    """cols: colour vectors as dicts {key: complex}.  Returns (chosen, M): indices of a basis picked among the
    vectors themselves (fewest non-zero entries first: the greedy choice is the minimum-weight basis) and the
    coefficients with  cols[i] = sum_j M[i][j] cols[chosen[j]]."""
    keys = sorted({k for c in cols for k in c}, key=repr)
    A = np.array([[complex(c.get(k, 0)) for c in cols] for k in keys]).reshape(len(keys), len(cols))
    order = sorted(range(len(cols)), key=lambda i: (len(cols[i]), i))
    chosen = []
    for i in order:
        if not cols[i]:
            continue
        trial = chosen + [i]
        if np.linalg.matrix_rank(A[:, trial], tol=1e-9) == len(trial):
            chosen = trial
    chosen.sort()
    if not chosen:
        return [], np.zeros((len(cols), 0), dtype=complex)
    B = A[:, chosen]
    M = np.linalg.lstsq(B, A, rcond=None)[0].T
    M = np.array([[_round_gauss(z, what) for z in row] for row in M]).reshape(len(cols), len(chosen))
    assert np.allclose(B @ M.T, A, atol=1e-9), f"{what}: colour reduction failed"
    return chosen, M


class _TermNode(Node):
    """A basis current: the sum of raw sub-trees `terms` = [(raw Node, coefficient)] (they share their propagator)."""
    __slots__ = ("terms",)

    def __init__(self, kind, legs, color, topo, terms):
        super().__init__(kind, legs, "SUM", (), color, topo)
        self.terms = terms


class RGen(Generator):
    """Generator whose currents over one (leg set, kind) are reduced to a colour basis."""

    def currents(self, legs, kind):
        legs = frozenset(legs)
        key = (legs, kind)
        if key in self._memo:
            return self._memo[key]
        if len(legs) == 1:
            return super().currents(legs, kind)
        raw = super().currents(legs, kind)     # built from the REDUCED currents of the sub-sets (memo)
        chosen, M = reduce_colours([nd.color for nd in raw], f"currents{sorted(legs)}{kind}")
        out = []
        for j, ci in enumerate(chosen):
            members = [i for i in range(len(raw)) if M[i][j] != 0]
            out.append(_TermNode(kind, legs, dict(raw[ci].color), "S[" + "+".join(raw[i].topo for i in members) + "]",
                                 [(raw[i], complex(M[i][j])) for i in members]))
        self._memo[key] = out
        return out


def _choose_x(children):
    """Which input of the closing vertex is `x` (the rest forms the pair object Q): the branch with the most legs,
    ties -> the one holding the lowest leg.  A function of the leg partition only, so that all closing vertices
    over one partition share their x."""
    best = max(range(len(children)), key=lambda q: (len(children[q].legs), -min(children[q].legs)))
    return best


def build_plan(k, ext_calls, root="centroid"):
    """The reduced evaluation plan of  g g > t t~ + k g  (JSON-friendly; attached to the IR as ir["plan"] by
    procgen.generate_ir).  ext_calls: the IR's external-wavefunction calls, by leg.

      objects  externals and basis currents in dependency order:
               {"legs", "ext": call | None, "terms": [{"op", "in": [objects], "coef": [re, im], "coup"}], "mass", "width"}
      pairs    basis vertex numerators: {"legs", "terms": [{"op", "jx", "in", "coef", "coup"}]} -- the closing vertex `op`
               with the line at argument position jx left open, its other lines `in`
      rows     {"x": object, "pair": pair, "jamp": [[colour flow, re, im], ...]}:  amp = x . pair  feeds these JAMPs"""
    gen = RGen(k)
    n = gen.n
    amps = gen.amplitudes(root)
    basis = list(itertools.permutations(gen.gluons))
    bindex = {w: j for j, w in enumerate(basis)}

    # objects in dependency order
    objects, oid = [], {}

    def visit(nd):
        if id(nd) in oid:
            return oid[id(nd)]
        if not getattr(nd, "terms", None):      # external
            (leg,) = nd.legs
            obj = {"legs": sorted(nd.legs), "ext": dict(ext_calls[leg]), "terms": []}
        else:
            terms = []
            for raw, coef in nd.terms:
                ins = [visit(ch) for ch in raw.children]
                terms.append({"op": raw.op, "in": ins, "coef": [coef.real, coef.imag], "coup": _COUP[raw.op[:4]]})
            obj = {"legs": sorted(nd.legs), "ext": None, "terms": terms,
                   "mass": _MASS[nd.kind], "width": _WIDTH[nd.kind]}
        oid[id(nd)] = len(objects)
        objects.append(obj)
        return oid[id(nd)]

    for leg in range(n):
        visit(gen.externals[leg])

    # closing vertices grouped by the leg set of x
    groups = {}
    for op, children, col, topo in amps:
        jx = _choose_x(children)
        x = children[jx]
        rest = tuple(ch for q, ch in enumerate(children) if q != jx)
        pkey = (op, jx, tuple(id(ch) for ch in rest))
        g = groups.setdefault((frozenset(x.legs), x.kind), {"xs": {}, "pairs": {}, "col": {}})
        g["xs"].setdefault(id(x), x)
        g["pairs"].setdefault(pkey, (op, jx, rest))
        assert (id(x), pkey) not in g["col"]
        g["col"][(id(x), pkey)] = col

    pairs, rows = [], []
    for (xlegs, xkind), g in sorted(groups.items(), key=lambda kv: (len(kv[0][0]), sorted(kv[0][0]), kv[0][1])):
        pkeys = list(g["pairs"])
        xids = list(g["xs"])
        # colour vector of a vertex numerator = its JAMP coefficients against every x of the group
        vecs = []
        for pk in pkeys:
            v = {}
            for xi in xids:
                for w, c in g["col"].get((xi, pk), {}).items():
                    v[(xi, bindex[w])] = complex(c)
            vecs.append(v)
        chosen, M = reduce_colours(vecs, f"numerators against {sorted(xlegs)}{xkind}")
        for j, ci in enumerate(chosen):
            terms = []
            for i, pk in enumerate(pkeys):
                if M[i][j] == 0:
                    continue
                op, jx, rest = g["pairs"][pk]
                terms.append({"op": op, "jx": jx, "in": [visit(ch) for ch in rest], "coef": [M[i][j].real, M[i][j].imag],
                              "coup": _COUP[op[:4]]})
            legs = sorted(set(range(n)) - set(xlegs))
            pid = len(pairs)
            pairs.append({"legs": legs, "terms": terms})
            for xi in xids:
                jamp = [(f, c) for (x_, f), c in vecs[ci].items() if x_ == xi and c != 0]
                if not jamp:
                    continue
                # overall sign as in procgen.generate_ir (jamp = -colour), see there
                rows.append({"x": visit(g["xs"][xi]), "pair": pid,
                             "jamp": sorted([f, float(round(-c.real, 12)) + 0.0, float(round(-c.imag, 12)) + 0.0] for f, c in jamp)})
    return {"objects": objects, "pairs": pairs, "rows": rows}


_COUP = {"FFV1": "GC_11", "VVV1": "GC_10", "VVVV": "GC_12"}
_MASS = {"o": "mdl_MT", "i": "mdl_MT", "g": "ZERO"}
_WIDTH = {"o": "mdl_WT", "i": "mdl_WT", "g": "ZERO"}


def plan_stats(plan):
    objs = [o for o in plan["objects"] if o["ext"] is None]
    return {
        "currents": len(objs), "current_terms": sum(len(o["terms"]) for o in objs),
        "pairs": len(plan["pairs"]), "pair_terms": sum(len(p["terms"]) for p in plan["pairs"]),
        "rows": len(plan["rows"]), "jamp_terms": sum(len(r["jamp"]) for r in plan["rows"]),
    }
