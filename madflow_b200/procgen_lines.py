"""Tree-level QCD processes with the top line, up to TWO light quark lines and extra gluons, in any crossing:

    q q~ > t t~ (g (g)),   g q > t t~ q (g),   g q~ > t t~ q~ (g),   g g > t t~ q q~,
    q q' > t t~ q q',   q q > t t~ q q,   q q~ > t t~ q' q~',   q q~ > t t~ q q~,  ...

the light-quark subprocesses of `p p > t t~ + jets` (reference: the subprocess loop of
scripts/madflow_exec.py:444-455 over what MG5_aMC generates).  MG5 is absent, so like madflow_b200.procgen this
module restates the Feynman rules in the conventions of the pinned ALOHA routines -- "parity unpinned" with respect
to MG5's own output, checked by BRST invariance of every colour flow, against the hand-built five-point IRs of
procgen.light_line_ttxg_ir and on the generated CUDA code executed on the host (tests/test_procgen.py).

Differences to procgen.Generator (which stays untouched: the compiled g g > t t~ + n g libraries depend on it):
  * colour is NUMERIC: every current carries the colour tensor of its sub-diagram (axes = the external legs below it
    + its own open index), vertices contract explicit SU(3) generators / structure constants, and the amplitudes'
    tensors are projected on the colour-flow basis  (T..)_{t I}(T..)_{O t~}  and  (T..)_{t t~}(T..)_{O I}  (I / O = the
    fermion-flow-in / -out end of the light line) by least squares; the colour matrix is the basis' Gram matrix;
  * two organisations of the same amplitude: every diagram closed at its centroid vertex (default: currents of at most
    n/2 legs, what the helicity-parallel kernels want) or at the vertex of the external t~ (FFV1_0 amplitudes only,
    currents of up to n-1 legs) -- they must agree, which the tests use.
"""
import itertools
import math
from fractions import Fraction

import numpy as np

from .procgen import _partitions, _su3, integer_rows

QUARTIC = {1: ((1, 2), (3, 4)), 3: ((1, 3), (2, 4)), 4: ((1, 4), (2, 3))}   # f^{e,p,q} f^{r,s,e}, UFO models/sm


class Node:
    __slots__ = ("kind", "legs", "op", "children", "ct", "uid", "topo")

    def __init__(self, kind, legs, op, children, ct, topo):
        self.kind, self.legs, self.op, self.children, self.ct, self.topo = kind, frozenset(legs), op, children, ct, topo
        self.uid = None


class LineGenerator:
    """legs: list of roles per external leg, in MG5's leg order (incoming first):
         "g"   gluon
         "to"  t  (outgoing fermion: oxxxxx, the top string's row index)     "ti"  t~ (ixxxxx, its column index)
         "lo"  fermion-flow-out end of the light line (incoming q~ or outgoing q: oxxxxx)
         "li"  fermion-flow-in end (incoming q or outgoing q~: ixxxxx)"""

    OPEN = 40   # einsum axis labels must stay below 52: legs 0..7, children's open indices 20+q, strings 30+

    def __init__(self, roles, ninitial=2):
        self.roles, self.n, self.ninitial = list(roles), len(roles), ninitial
        quarks = sorted(r for r in roles if r != "g")
        self.lines = [tag for tag in "tlm" if tag + "o" in quarks]   # t = top, l / m = light lines of different flavour
        assert "t" in self.lines and quarks == sorted(tag + end for tag in self.lines for end in "oi"), \
            "the top line and up to two light lines, each with both ends"
        self.has_light = len(self.lines) > 1
        self.T, self.f = _su3()
        self.leg_of = {r: k for k, r in enumerate(roles) if r != "g"}
        self.dim = [8 if r == "g" else 3 for r in roles]
        self._memo = {}
        self.externals = {}
        for leg, r in enumerate(roles):
            kind = "g" if r == "g" else f"{r[1]}_{r[0]}"          # "to" -> "o_t", "li" -> "i_l"
            op = "vxxxxx" if r == "g" else ("oxxxxx" if r[1] == "o" else "ixxxxx")
            self.externals[leg] = Node(kind, [leg], op, (), np.eye(self.dim[leg]), f"x{leg}")

    # ---- colour tensors: axes = sorted external legs + the open index (last)
    def _contract(self, vertex, vaxes, nodes, out_open):
        """einsum of the vertex tensor (axes `vaxes`: internal labels 20+q for child q's open index, OPEN for the new one)
        with the children's colour tensors."""
        ops = [vertex, list(vaxes)]
        legs = set()
        for q, nd in enumerate(nodes):
            ops += [nd.ct, sorted(nd.legs) + [20 + q]]
            legs |= nd.legs
        out = sorted(legs) + ([self.OPEN] if out_open else [])
        return np.einsum(*ops, out)

    def admissible(self, legs, kind):
        """a gluon current holds every quark line entirely or not at all; a quark current of line X holds exactly its own
        end of X, and every other line entirely or not at all"""
        s = set(legs)
        for tag in self.lines:
            has_o, has_i = self.leg_of[tag + "o"] in s, self.leg_of[tag + "i"] in s
            if kind != "g" and kind[2] == tag:
                if (has_o, has_i) != ((True, False) if kind[0] == "o" else (False, True)):
                    return False
            elif has_o != has_i:
                return False
        return True

    def currents(self, legs, kind):
        legs = frozenset(legs)
        key = (legs, kind)
        if key in self._memo:
            return self._memo[key]
        out = []
        if len(legs) == 1:
            (leg,) = legs
            nd = self.externals[leg]
            out = [nd] if nd.kind == kind else []
        elif self.admissible(legs, kind):
            s = tuple(sorted(legs))
            T, f = self.T, self.f
            if kind[0] == "o":               # FFV1_1(o, g): string (.. T^a)_{row, open}
                for a, b in self._splits2(s):
                    for o in self.currents(a, kind):
                        for g in self.currents(b, "g"):
                            ct = self._contract(T, [21, 20, self.OPEN], (o, g), True)
                            out.append(Node(kind, legs, "FFV1_1", (o, g), ct, f"F1({o.topo},{g.topo})"))
            elif kind[0] == "i":             # FFV1_2(i, g): string (T^a ..)_{open, column}
                for a, b in self._splits2(s):
                    for i in self.currents(a, kind):
                        for g in self.currents(b, "g"):
                            ct = self._contract(T, [21, self.OPEN, 20], (i, g), True)
                            out.append(Node(kind, legs, "FFV1_2", (i, g), ct, f"F2({i.topo},{g.topo})"))
            else:
                for a, b in self._splits2(s):
                    for line in self.lines:  # FFV1P0_3(i, o): T^a_{o's open, i's open}
                        for i in self.currents(a, "i_" + line):
                            for o in self.currents(b, "o_" + line):
                                ct = self._contract(T, [self.OPEN, 21, 20], (i, o), True)
                                out.append(Node("g", legs, "FFV1P0_3", (i, o), ct, f"J({i.topo},{o.topo})"))
                for part in _partitions(s, 2):
                    for x in self.currents(part[0], "g"):
                        for y in self.currents(part[1], "g"):   # VVV1P0_1(V2, V3) -> leg 1: f^{1,2,3}
                            ct = self._contract(f, [self.OPEN, 20, 21], (x, y), True)
                            out.append(Node("g", legs, "VVV1P0_1", (x, y), ct, f"V({x.topo},{y.topo})"))
                for part in _partitions(s, 3):
                    for x in self.currents(part[0], "g"):
                        for y in self.currents(part[1], "g"):
                            for z in self.currents(part[2], "g"):
                                for kind4, ((p, q), (r, s_)) in QUARTIC.items():
                                    lab = {1: self.OPEN, 2: 20, 3: 21, 4: 22}
                                    ff = np.einsum("epq,rse->pqrs", f, f)
                                    ct = self._contract(ff, [lab[p], lab[q], lab[r], lab[s_]], (x, y, z), True)
                                    out.append(Node("g", legs, f"VVVV{kind4}P0_1", (x, y, z), ct,
                                                    f"W{kind4}({x.topo},{y.topo},{z.topo})"))
        self._memo[key] = out
        return out

    @staticmethod
    def _splits2(s):
        """ordered splits of s into two non-empty parts (both orders)"""
        for part in _partitions(s, 2):
            yield part[0], part[1]
            yield part[1], part[0]

    def amplitudes(self, root="tbar"):
        """root="tbar": FFV1_0(t~, o_t(A), g(B)) over all splits of the other legs -- every diagram exactly once, currents
        of up to n-1 legs.  root="centroid": every diagram closed at its centroid vertex (no branch with more than n/2
        legs; when the centroid is an edge, the end whose heavy branch holds leg 0), as procgen.Generator does: currents of
        at most n/2 legs, FFV1_0 / VVV1_0 / VVVV_0 amplitudes."""
        if root == "centroid":
            return self._amplitudes_centroid()
        tb = self.leg_of["ti"]
        rest = tuple(l for l in range(self.n) if l != tb)
        amps = []
        for a, b in self._splits2(rest):
            for o in self.currents(a, "o_t"):
                for g in self.currents(b, "g"):
                    # (o string)_{t, m} T^a_{m, m'} delta_{m', t~}
                    ct = self._contract(self.T, [21, 20, 22], (o, g, self.externals[tb]), False)
                    amps.append(("FFV1_0", (self.externals[tb], o, g), ct, f"A({o.topo},{g.topo})"))
        return amps

    def _amplitudes_centroid(self):
        n, half = self.n, self.n // 2
        all_legs = tuple(range(n))
        lines = [(tag, tag + "o", tag + "i") for tag in self.lines]
        amps = []

        def allowed(parts):
            sizes = [len(p_) for p_ in parts]
            if max(sizes) > half:
                return False
            if 2 * max(sizes) == n:   # the centroid is an edge: keep the end whose heavy branch holds leg 0
                return 0 in parts[sizes.index(max(sizes))]
            return True

        ff = {k4: np.einsum("epq,rse->pqrs", self.f, self.f) for k4 in QUARTIC}
        for nparts in (3, 4):
            for parts in _partitions(all_legs, nparts):
                if not allowed(parts):
                    continue
                if all(self.admissible(p_, "g") for p_ in parts):
                    for combo in itertools.product(*[self.currents(p_, "g") for p_ in parts]):
                        if nparts == 3:
                            ct = self._contract(self.f, [20, 21, 22], combo, False)
                            amps.append(("VVV1_0", combo, ct, "A3(" + ",".join(x.topo for x in combo) + ")"))
                        else:
                            for k4, ((p, q), (r, s_)) in QUARTIC.items():
                                lab = {1: 20, 2: 21, 3: 22, 4: 23}
                                ct = self._contract(ff[k4], [lab[p], lab[q], lab[r], lab[s_]], combo, False)
                                amps.append((f"VVVV{k4}_0", combo, ct, "A4(" + ",".join(x.topo for x in combo) + ")"))   # one diagram, 3 structures
                elif nparts == 3:
                    for line, ro, ri in lines:   # the vertex sits on this quark line: branches (i side, o side, gluon)
                        for pi_, po_, pg_ in itertools.permutations(parts):
                            if not (self.admissible(pi_, "i_" + line) and self.admissible(po_, "o_" + line) and self.admissible(pg_, "g")):
                                continue
                            for i in self.currents(pi_, "i_" + line):
                                for o in self.currents(po_, "o_" + line):
                                    for g in self.currents(pg_, "g"):
                                        ct = self._contract(self.T, [22, 21, 20], (i, o, g), False)   # T^a_{o's open, i's open}
                                        amps.append(("FFV1_0", (i, o, g), ct, f"A({i.topo},{o.topo},{g.topo})"))
        return amps

    # ---- colour flows
    def colour_flows(self):
        """Basis tensors over all external legs (sorted): the outgoing-flow end of every quark line is connected to the
        incoming-flow end of some line by a string of generators, (T^s1)_{o_1, i_pi(1)} (T^s2)_{o_2, i_pi(2)} ..., over all
        permutations pi and all ordered distributions of the gluons on the strings.  One line: the n! strings of
        procgen.generate_ir; two lines: (T..)_{t I}(T..)_{O t~} and (T..)_{t t~}(T..)_{O I}."""
        gl = [l for l, r in enumerate(self.roles) if r == "g"]
        rows = [self.leg_of[tag + "o"] for tag in self.lines]
        cols = [self.leg_of[tag + "i"] for tag in self.lines]
        L = len(rows)
        flows, names = [], []
        for pi_ in itertools.permutations(range(L)):
            for perm in itertools.permutations(gl):
                for cuts in itertools.combinations_with_replacement(range(len(gl) + 1), L - 1):
                    bounds = (0,) + cuts + (len(gl),)
                    segs = [perm[bounds[q]:bounds[q + 1]] for q in range(L)]
                    ops = []
                    for q in range(L):
                        t_, ax = self._string(segs[q], rows[q], cols[pi_[q]])
                        ops += [t_, ax]
                    flows.append(np.einsum(*ops, list(range(self.n))).reshape(-1))
                    names.append((pi_, tuple(segs)))
        return np.stack(flows, axis=1), names

    def _string(self, gluons, row, col):
        """(T^{g1} T^{g2} ..)_{row, col} as (tensor, axis labels)"""
        m, axes, cur = np.eye(3), [row, 30], 30
        for q, g in enumerate(gluons):
            m = np.einsum(m, axes, self.T, [g, cur, cur + 1], [a for a in axes if a != cur] + [g, cur + 1])
            axes = [a for a in axes if a != cur] + [g, cur + 1]
            cur += 1
        m = np.einsum(m, axes, np.eye(3), [cur, col], [a for a in axes if a != cur] + [col])
        return m, [a for a in axes if a != cur] + [col]


def _rational(c, what):
    re, im = Fraction(float(c.real)).limit_denominator(5184), Fraction(float(c.imag)).limit_denominator(5184)
    assert abs(complex(re, im) - c) < 1e-10, f"{what}: {c} is not a small rational"
    return re, im


def generate_ir(roles, name, process, pdg, initial_states, mirror=True, ninitial=2, root="centroid"):
    """IR of the process whose external legs have the given roles (see LineGenerator)."""
    gen = LineGenerator(roles, ninitial)
    n = gen.n
    amps = [(op, ch, ct, topo) for op, ch, ct, topo in gen.amplitudes(root)]
    B, flow_names = gen.colour_flows()
    # Two light lines of the SAME flavour: the fermion-flow-in ends are indistinguishable, so the diagrams with the two
    # lines re-paired (l: lo-mi, m: mo-li) contribute with the opposite sign (Fermi statistics).  A second generator
    # with the roles of the two in-ends exchanged enumerates them; colour tensors carry leg labels, so they project on
    # the same flows.
    gens = [gen]
    if "l" in gen.lines and "m" in gen.lines and abs(pdg[gen.leg_of["lo"]]) == abs(pdg[gen.leg_of["mo"]]):
        swap = {"li": "mi", "mi": "li"}
        gen2 = LineGenerator([swap.get(r, r) for r in roles], ninitial)
        gens.append(gen2)
        amps += [(op, ch, -ct, "X" + topo) for op, ch, ct, topo in gen2.amplitudes(root)]
    # For N = 3 the flows become linearly dependent at large multiplicities (three quark lines + a gluon, ...): the
    # decomposition is then not unique -- harmless for |M|^2, but the JAMPs are no longer individually gauge invariant
    independent = np.linalg.matrix_rank(B) == B.shape[1]
    Bpinv = np.linalg.pinv(B)   # one factorisation for all amplitudes: coefficients = pinv(B) tensor (minimum norm)
    # schedule: externals, then per amplitude the currents it needs (each wavefunction keeps its own slot)
    calls, order = [], []
    uid = itertools.count()

    def visit(nd):
        if nd.uid is not None:
            return
        for ch in nd.children:
            visit(ch)
        nd.uid = next(uid)
        order.append(nd)

    def emit(nd):
        line = nd.kind[-1] if nd.kind != "g" else "g"
        mass, width = ("mdl_MT", "mdl_WT") if line == "t" else ("ZERO", "ZERO")
        if not nd.children:
            (leg,) = nd.legs
            role = roles[leg]
            incoming = leg < ninitial
            # gluon: -1 incoming, +1 outgoing; o-type end (oxxxxx): incoming antiquark -1 / outgoing quark +1; i-type end
            # (ixxxxx): incoming quark +1 / outgoing antiquark -1
            nsf = (-1 if incoming else 1) if role == "g" or role[1] == "o" else (1 if incoming else -1)
            calls.append({"op": nd.op, "out": nd.uid, "leg": leg, "mass": mass, "nsf": nsf})
        else:
            coup = {"FFV1": "GC_11", "VVV1": "GC_10", "VVVV": "GC_12"}[nd.op[:4]]
            calls.append({"op": nd.op, "out": nd.uid, "in": [c.uid for c in nd.children], "coup": coup, "mass": mass, "width": width})

    for leg in range(n):
        visit(gen.externals[leg])
        for other in gens[1:]:   # the same external wavefunctions (a leg's call does not depend on the pairing)
            assert other.externals[leg].op == gen.externals[leg].op
            other.externals[leg].uid = gen.externals[leg].uid
    for nd in order:
        emit(nd)
    jamp = [[] for _ in flow_names]
    for a_idx, (op, children, ct, topo) in enumerate(amps):
        before = len(order)
        for ch in children:
            visit(ch)
        for nd in order[before:]:
            emit(nd)
        calls.append({"op": op, "amp": a_idx, "in": [c.uid for c in children],
                      "coup": {"FFV1": "GC_11", "VVV1": "GC_10", "VVVV": "GC_12"}[op[:4]]})
        coef = Bpinv @ ct.reshape(-1)
        assert np.allclose(B @ coef, ct.reshape(-1), atol=1e-10), f"the colour flows do not span {topo}"
        for k_, c in enumerate(coef):
            if independent:
                re, im = _rational(c, topo)
            else:   # any decomposition gives the same |M|^2 = c^+ (B^+ B) c; the minimum-norm one has no nice fractions
                re, im = (0.0 if abs(c.real) < 1e-13 else float(c.real)), (0.0 if abs(c.imag) < 1e-13 else float(c.imag))
            if re or im:
                jamp[k_].append((a_idx, -float(re), -float(im)))
    gram = (B.conj().T @ B).real
    rows = [[_rational(complex(v), "colour matrix")[0] for v in row] for row in gram]
    nums, dens = integer_rows(rows)
    hel_states = []
    for leg, role in enumerate(roles):
        anti_like = (role != "g" and role[1] == "i") == (leg >= ninitial)   # as before for t~ and the light line
        hel_states.append([1, -1] if role != "g" and anti_like else [-1, 1])

    colour_avg = 1
    for leg in range(ninitial):
        colour_avg *= 8 if roles[leg] == "g" else 3
    identical = 1   # identical particles in the final state
    for code in set(pdg[ninitial:]):
        identical *= math.factorial(pdg[ninitial:].count(code))
    topos = {t.replace("W1", "W").replace("W3", "W").replace("W4", "W") for _, _, _, t in amps}
    return {
        "name": name, "process": process, "nexternal": n, "ninitial": ninitial, "ndiags": len(topos), "ncomb": 2**n,
        "nwavefuncs": len(order), "helicities": [list(h) for h in itertools.product(*hel_states)],
        "denominator": 4 * colour_avg * identical,
        "params": ["mdl_MT", "mdl_WT"], "couplings": sorted({c["coup"] for c in calls if "coup" in c}),
        "initial_states": initial_states, "mirror_initial_states": bool(mirror), "pdg": pdg,
        "masses": ["mdl_MT" if r in ("to", "ti") else "ZERO" for r in roles],
        "calls": calls, "jamp": jamp, "color_num": nums, "color_denom": dens,
        "color_basis": [[list(pi_), [list(sg) for sg in segs]] for pi_, segs in flow_names],
        "color_flows_independent": bool(independent),
    }


LIGHT = [2, 4, 1, 3]   # u, c, d, s in MG5's order

PROCESSES = {   # name -> (roles, process string, pdg of the first flavour, initial states of all light flavours)
    "1_uux_ttx": (["li", "lo", "to", "ti"], "u u~ > t t~", [2, -2, 6, -6], [[q, -q] for q in LIGHT]),
    "1_uux_ttxg": (["li", "lo", "to", "ti", "g"], "u u~ > t t~ g", [2, -2, 6, -6, 21], [[q, -q] for q in LIGHT]),
    "1_gu_ttxu": (["g", "li", "to", "ti", "lo"], "g u > t t~ u", [21, 2, 6, -6, 2], [[21, q] for q in LIGHT]),
    "1_gux_ttxux": (["g", "lo", "to", "ti", "li"], "g u~ > t t~ u~", [21, -2, 6, -6, -2], [[21, -q] for q in LIGHT]),
    "1_uux_ttxgg": (["li", "lo", "to", "ti", "g", "g"], "u u~ > t t~ g g", [2, -2, 6, -6, 21, 21], [[q, -q] for q in LIGHT]),
    "1_gu_ttxug": (["g", "li", "to", "ti", "lo", "g"], "g u > t t~ u g", [21, 2, 6, -6, 2, 21], [[21, q] for q in LIGHT]),
    "1_gux_ttxuxg": (["g", "lo", "to", "ti", "li", "g"], "g u~ > t t~ u~ g", [21, -2, 6, -6, -2, 21], [[21, -q] for q in LIGHT]),
    "1_gg_ttxuux": (["g", "g", "to", "ti", "lo", "li"], "g g > t t~ u u~", [21, 21, 6, -6, 2, -2], [[21, 21]]),
    # two light lines (the four-quark subprocesses of p p > t t~ j j); same-flavour lines get the exchange diagrams with
    # the Fermi sign automatically.  initial_states: (parton of hadron 1, parton of hadron 2), mirrored where listed once
    "1_uu_ttxuu": (["li", "mi", "to", "ti", "lo", "mo"], "u u > t t~ u u", [2, 2, 6, -6, 2, 2], [[q, q] for q in LIGHT]),
    "1_ud_ttxud": (["li", "mi", "to", "ti", "lo", "mo"], "u d > t t~ u d", [2, 1, 6, -6, 2, 1],
                   [[a, b] for k_, a in enumerate(LIGHT) for b in LIGHT[k_ + 1:]]),
    "1_uxux_ttxuxux": (["lo", "mo", "to", "ti", "li", "mi"], "u~ u~ > t t~ u~ u~", [-2, -2, 6, -6, -2, -2], [[-q, -q] for q in LIGHT]),
    "1_uxdx_ttxuxdx": (["lo", "mo", "to", "ti", "li", "mi"], "u~ d~ > t t~ u~ d~", [-2, -1, 6, -6, -2, -1],
                       [[-a, -b] for k_, a in enumerate(LIGHT) for b in LIGHT[k_ + 1:]]),
    "1_uux_ttxuux": (["li", "lo", "to", "ti", "mo", "mi"], "u u~ > t t~ u u~", [2, -2, 6, -6, 2, -2], [[q, -q] for q in LIGHT]),
    # the final flavour differs from the initial one: three choices, counted by listing every initial state three times
    "1_uux_ttxddx": (["li", "lo", "to", "ti", "mo", "mi"], "u u~ > t t~ d d~", [2, -2, 6, -6, 1, -1], [[q, -q] for q in LIGHT] * 3),
    "1_udx_ttxudx": (["li", "mo", "to", "ti", "lo", "mi"], "u d~ > t t~ u d~", [2, -1, 6, -6, 2, -1],
                     [[a, -b] for a in LIGHT for b in LIGHT if a != b]),
    # no light line: an independent derivation of the built-in processes (cross-check only, see the tests)
    "1_gg_ttx": (["g", "g", "to", "ti"], "g g > t t~", [21, 21, 6, -6], [[21, 21]]),
    "1_gg_ttxg": (["g", "g", "to", "ti", "g"], "g g > t t~ g", [21, 21, 6, -6, 21], [[21, 21]]),
    "1_gg_ttxgg": (["g", "g", "to", "ti", "g", "g"], "g g > t t~ g g", [21, 21, 6, -6, 21, 21], [[21, 21]]),
}


def process_ir(name, root="centroid"):
    roles, proc, pdg, initial = PROCESSES[name]
    mirror = initial != [[21, 21]] and not all(a == b for a, b in initial)   # identical beams partons need no mirror
    n = len(roles)
    return generate_ir(roles, name, f"{proc} WEIGHTED<={n - 2} @1", pdg, initial, mirror, root=root)
