"""Built-in tree-level process generator for  g g > t t~ + k g  (k = 0..3) and  q q~ > t t~, QCD only.

Why it exists: the IR of a process (process_ir.py) is what MG5_aMC computes and hands to
madflow's exporter -- the diagrams, the HELAS call list, the JAMP coefficients and the colour
matrix (madgraph_plugin/PyOut_exporter.py:174-176, 301-324, 334-375).  MG5_aMC is a third-party
program that is not part of the reference tree and is not available offline, so for the BASELINE
configurations beyond g g > t t~ this module produces the IR itself.  Everything here is restated
from the textbook Feynman rules of QCD in the conventions the pinned ALOHA routines fix
(csrc/aloha_sm.cuh); it is checked by the counts MG5 is known to produce
(diagrams 3/16/123/1240, amplitudes 3/18/159, colour flows 2/6/24/120, denominators
256/256/512/1536 -- SURVEY.md section 8), by gauge invariance, Bose symmetry and closed forms
(tests/test_procgen.py) -- "parity unpinned" with respect to MG5's own call ORDER.

Construction
  legs      0,1 incoming gluons; 2 = t (oxxxxx, +1); 3 = t~ (ixxxxx, -1); 4.. outgoing gluons.
  currents  every distinct sub-tree over a leg subset S with one off-shell line is ONE wavefunction,
            shared by all diagrams containing it (MG5 reuses wavefunctions the same way):
              g(S)  gluon current:  VVV1P0_1(g,g) | VVVV{1,3,4}P0_1(g,g,g) | FFV1P0_3(i,o)
              o(S)  quark current flowing out of t:   FFV1_1(o, g)
              i(S)  quark current flowing into t~:    FFV1_2(i, g)
  diagrams  each diagram is closed at its centroid vertex (every branch has <= n/2 legs), which
            keeps all currents at <= n/2 legs:  FFV1_0(i,o,g) | VVV1_0(g,g,g) | VVVV{1,3,4}_0(g,g,g,g).
            A four-gluon vertex contributes its three colour/Lorentz structures as three
            amplitudes (root) or three wavefunctions (internal) -- MG5's counting.
  colour    basis = strings (T^{a_s1} ... T^{a_sn})_{i jbar} over the n! gluon orderings; the colour
            factor of every current is kept as a linear combination of words (with a hole symbol
            for the open adjoint index of a current containing the quark line), using
            T^a f^{abc} = -i [T^b, T^c].  Colour matrix from the Fierz identity, exact rationals.
"""
import itertools
import math
from fractions import Fraction

HOLE = "*"
I = complex(0, 1)


# ------------------------------------------------------------------------------------------------
# colour word algebra: dict {tuple(word): complex coefficient}
def c_add(a, b, scale=1):
    out = dict(a)
    for w, c in b.items():
        v = out.get(w, 0) + scale * c
        if v == 0:
            out.pop(w, None)
        else:
            out[w] = v
    return out


def c_mul(a, b):
    out = {}
    for wa, ca in a.items():
        for wb, cb in b.items():
            w = wa + wb
            v = out.get(w, 0) + ca * cb
            if v == 0:
                out.pop(w, None)
            else:
                out[w] = v
    return out


def c_scale(a, s):
    return {w: c * s for w, c in a.items()}


def c_comm(a, b):
    return c_add(c_mul(a, b), c_mul(b, a), -1)


def c_sub_hole(q, repl):
    """Replace the hole symbol in every word of q by the colour object repl."""
    out = {}
    for w, c in q.items():
        k = w.index(HOLE)
        left, right = {w[:k]: 1}, {w[k + 1:]: 1}
        out = c_add(out, c_scale(c_mul(c_mul(left, repl), right), c))
    return out


def has_hole(col):
    return any(HOLE in w for w in col)


def f_contract(slots, order):
    """Colour of f^{order[0], order[1], order[2]} given the three slot colour objects.

    Exactly one slot is the 'insertion' slot: the new off-shell index (given as None, all other
    slots pure) or the slot whose colour object contains the hole.  Returns the colour object that
    replaces T^t:  T^t f^{t u v} = -i [T^u, T^v]  ((t,u,v) a cyclic rotation of `order`).  When a
    hole is the insertion slot and another slot is None, None stands for the new hole."""
    a = [slots[k] for k in order]
    ins = [k for k in range(3) if a[k] is not None and has_hole(a[k])]
    if ins:
        t = ins[0]
    else:
        t = [k for k in range(3) if a[k] is None][0]
    u, v = a[(t + 1) % 3], a[(t + 2) % 3]
    u = {(HOLE,): 1} if u is None else u
    v = {(HOLE,): 1} if v is None else v
    return t, c_scale(c_comm(u, v), -I)


# ------------------------------------------------------------------------------------------------
class Node:
    """One wavefunction (external or off-shell current)."""
    __slots__ = ("kind", "legs", "op", "children", "color", "uid", "topo")

    def __init__(self, kind, legs, op, children, color, topo):
        self.kind, self.legs, self.op, self.children, self.color, self.topo = kind, legs, op, children, color, topo
        self.uid = None

    @property
    def quarkful(self):
        return self.kind == "g" and 2 in self.legs


def _partitions(items, nparts):
    """Unordered partitions of a sorted tuple into nparts non-empty blocks (each block sorted, blocks
    ordered by their smallest element)."""
    items = tuple(items)
    if nparts == 1:
        yield (items,)
        return
    first, rest = items[0], items[1:]
    # the block containing `first`
    for r in range(0, len(rest) - (nparts - 1) + 1):
        for extra in itertools.combinations(rest, r):
            block = (first,) + extra
            remaining = tuple(x for x in rest if x not in extra)
            for tail in _partitions(remaining, nparts - 1):
                yield (block,) + tail


class Generator:
    def __init__(self, n_final_gluons):
        self.k = n_final_gluons
        self.n = 4 + n_final_gluons
        self.gluons = (0, 1) + tuple(range(4, self.n))
        self.half = self.n // 2
        self._memo = {}
        self.externals = {}
        for leg in range(self.n):
            if leg == 2:
                nd = Node("o", frozenset([2]), "oxxxxx", (), {(): 1}, "t")
            elif leg == 3:
                nd = Node("i", frozenset([3]), "ixxxxx", (), {(): 1}, "tb")
            else:
                nd = Node("g", frozenset([leg]), "vxxxxx", (), {(leg,): 1}, f"g{leg}")
            self.externals[leg] = nd

    # -- currents
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
            self._memo[key] = out
            return out
        s = tuple(sorted(legs))
        if kind == "o":
            if 2 in legs and 3 not in legs:
                for part in _partitions(s, 2):
                    a, b = (part[0], part[1]) if 2 in part[0] else (part[1], part[0])
                    for o in self.currents(a, "o"):
                        for g in self.currents(b, "g"):
                            out.append(Node("o", legs, "FFV1_1", (o, g), c_mul(o.color, g.color),
                                            f"F1({o.topo},{g.topo})"))
        elif kind == "i":
            if 3 in legs and 2 not in legs:
                for part in _partitions(s, 2):
                    a, b = (part[0], part[1]) if 3 in part[0] else (part[1], part[0])
                    for i in self.currents(a, "i"):
                        for g in self.currents(b, "g"):
                            out.append(Node("i", legs, "FFV1_2", (i, g), c_mul(g.color, i.color),
                                            f"F2({i.topo},{g.topo})"))
        elif kind == "g":
            nq = (2 in legs) + (3 in legs)
            if nq == 1:
                out = []
            else:
                if nq == 2:
                    for part in _partitions(s, 2):
                        for a, b in ((part[0], part[1]), (part[1], part[0])):
                            if 3 in a and 2 in b and 2 not in a and 3 not in b:
                                for i in self.currents(a, "i"):
                                    for o in self.currents(b, "o"):
                                        col = c_mul(c_mul(o.color, {(HOLE,): 1}), i.color)
                                        out.append(Node("g", legs, "FFV1P0_3", (i, o), col, f"J({i.topo},{o.topo})"))
                for part in _partitions(s, 2):
                    for x in self.currents(part[0], "g"):
                        for y in self.currents(part[1], "g"):
                            out.append(self._vvv_current(legs, x, y))
                for part in _partitions(s, 3):
                    for x in self.currents(part[0], "g"):
                        for y in self.currents(part[1], "g"):
                            for z in self.currents(part[2], "g"):
                                out.extend(self._vvvv_currents(legs, x, y, z))
        self._memo[key] = out
        return out

    @staticmethod
    def _ordered(nodes):
        """Canonical argument order: the current containing the quark line first, then by lowest leg."""
        return sorted(nodes, key=lambda nd: (not nd.quarkful, min(nd.legs)))

    def _vvv_current(self, legs, x, y):
        x, y = self._ordered([x, y])
        # VVV1P0_1(V2=x, V3=y) -> leg 1: colour f^{1,2,3} = f^{new, x, y}
        t, repl = f_contract({1: None, 2: x.color, 3: y.color}, (1, 2, 3))
        if x.quarkful:
            col = c_sub_hole(x.color, repl)
        else:
            col = repl
        return Node("g", frozenset(legs), "VVV1P0_1", (x, y), col, f"V({x.topo},{y.topo})")

    # UFO models/sm vertex g g g g: colour f(-1,1,2) f(3,4,-1) <-> VVVV1, f(-1,1,3) f(2,4,-1) <-> VVVV3,
    # f(-1,1,4) f(2,3,-1) <-> VVVV4
    QUARTIC = {1: ((1, 2), (3, 4)), 3: ((1, 3), (2, 4)), 4: ((1, 4), (2, 3))}

    def _quartic_color(self, kind, slots):
        """Colour of f^{e,p,q} f^{r,s,e} (structure `kind`) for slot colour objects {1..4}: None marks
        the new off-shell index.  Returns (quarkful slot or None, colour object)."""
        (p, q), (r, s) = self.QUARTIC[kind]
        holes = [k for k in (1, 2, 3, 4) if slots[k] is not None and has_hole(slots[k])]
        tslot = holes[0] if holes else [k for k in (1, 2, 3, 4) if slots[k] is None][0]

        def obj(k):
            return {(HOLE,): 1} if slots[k] is None else slots[k]

        # write the f containing the insertion slot as sigma * f^{t,u,e} and the other as f^{v,w,e}
        if tslot in (p, q):
            # f^{e,p,q} = f^{p,q,e}
            if tslot == p:
                u, sigma = q, 1
            else:
                u, sigma = p, -1          # f^{p,t,e} = -f^{t,p,e}
            v, w = r, s
        else:
            if tslot == r:
                u, sigma = s, 1
            else:
                u, sigma = r, -1
            v, w = p, q                    # f^{e,p,q} = f^{p,q,e}
        inner = c_scale(c_comm(obj(v), obj(w)), -I)      # E = -i [V, W]
        repl = c_scale(c_comm(obj(u), inner), -I * sigma)  # T^t f^{t,u,e} U E = -i [U, E]
        return (holes[0] if holes else None), repl

    def _vvvv_currents(self, legs, x, y, z):
        x, y, z = self._ordered([x, y, z])
        out = []
        for kind in (1, 3, 4):
            hole_slot, repl = self._quartic_color(kind, {1: None, 2: x.color, 3: y.color, 4: z.color})
            col = c_sub_hole(x.color, repl) if x.quarkful else repl
            out.append(Node("g", frozenset(legs), f"VVVV{kind}P0_1", (x, y, z), col,
                            f"W{kind}({x.topo},{y.topo},{z.topo})"))
        return out

    # -- diagrams (root vertices)
    def amplitudes(self, root="centroid"):
        """List of (op, children, colour dict over full words, diagram key).

        root="centroid" (default) closes every diagram at its centroid vertex; root="tbar" closes
        every diagram at the vertex of the external t~ (only FFV1_0 amplitudes, only o-type and
        pure-gluon currents) -- an independent organisation used as a cross-check."""
        n, half = self.n, self.half
        all_legs = tuple(range(n))
        amps = []
        if root == "tbar":
            rest = tuple(l for l in all_legs if l != 3)
            tb = self.externals[3]
            for parts in _partitions(rest, 2):
                a, b = (parts[0], parts[1]) if 2 in parts[0] else (parts[1], parts[0])
                for o in self.currents(a, "o"):
                    for g in self.currents(b, "g"):
                        col = c_mul(c_mul(o.color, g.color), tb.color)
                        amps.append(("FFV1_0", (tb, o, g), col, f"A({tb.topo},{o.topo},{g.topo})"))
            return amps

        def allowed(parts):
            sizes = [len(p) for p in parts]
            if max(sizes) > half:
                return False
            if 2 * max(sizes) == n:  # the centroid is an edge: keep the end whose heavy branch holds leg 0
                heavy = parts[sizes.index(max(sizes))]
                return 0 in heavy
            return True

        for nparts in (3, 4):
            for parts in _partitions(all_legs, nparts):
                if not allowed(parts):
                    continue
                with2 = [p for p in parts if 2 in p]
                with3 = [p for p in parts if 3 in p]
                same = with2[0] is with3[0] or with2[0] == with3[0]
                if not same:
                    if nparts != 3:
                        continue
                    (pg,) = [p for p in parts if 2 not in p and 3 not in p]
                    for i in self.currents(with3[0], "i"):
                        for o in self.currents(with2[0], "o"):
                            for g in self.currents(pg, "g"):
                                col = c_mul(c_mul(o.color, g.color), i.color)
                                amps.append(("FFV1_0", (i, o, g), col, f"A({i.topo},{o.topo},{g.topo})"))
                else:
                    lists = [self.currents(p, "g") for p in parts]
                    for combo in itertools.product(*lists):
                        nodes = self._ordered(list(combo))
                        q = nodes[0]
                        assert q.quarkful
                        if nparts == 3:
                            t, repl = f_contract({1: q.color, 2: nodes[1].color, 3: nodes[2].color}, (1, 2, 3))
                            col = c_sub_hole(q.color, repl)
                            amps.append(("VVV1_0", tuple(nodes), col, "A3(" + ",".join(x.topo for x in nodes) + ")"))
                        else:
                            for kind in (1, 3, 4):
                                _, repl = self._quartic_color(kind, {1: q.color, 2: nodes[1].color, 3: nodes[2].color,
                                                                     4: nodes[3].color})
                                col = c_sub_hole(q.color, repl)
                                amps.append((f"VVVV{kind}_0", tuple(nodes), col,
                                             "A4(" + ",".join(x.topo for x in nodes) + ")"))
        return amps


# ------------------------------------------------------------------------------------------------
# colour matrix
def _trace_value(traces, N=3, _memo={}):
    """Product of traces of generator words in which every index appears exactly twice -> Fraction."""
    key = tuple(sorted(_canon(t) for t in traces))
    if key in _memo:
        return _memo[key]
    val = None
    # empty traces
    if all(len(t) == 0 for t in traces):
        val = Fraction(N) ** len(traces)
    else:
        if any(len(t) == 1 for t in traces):
            val = Fraction(0)
        else:
            # pick the first index of the first non-empty trace
            ti = next(k for k, t in enumerate(traces) if len(t))
            t = traces[ti]
            a = t[0]
            rest = [x for k, x in enumerate(traces) if k != ti]
            if a in t[1:]:
                j = t.index(a, 1)
                X, Y = t[1:j], t[j + 1:]
                # Tr(T^a X T^a Y) = 1/2 (Tr X Tr Y - 1/N Tr(XY))
                val = Fraction(1, 2) * (_trace_value(rest + [X, Y], N) - Fraction(1, N) * _trace_value(rest + [X + Y], N))
            else:
                oi = next(k for k, x in enumerate(rest) if a in x)
                o = rest[oi]
                j = o.index(a)
                Yc = o[j + 1:] + o[:j]
                X = t[1:]
                others = [x for k, x in enumerate(rest) if k != oi]
                # Tr(T^a X) Tr(T^a Y) = 1/2 (Tr(XY) - 1/N Tr X Tr Y)
                val = Fraction(1, 2) * (_trace_value(others + [X + Yc], N) - Fraction(1, N) * _trace_value(others + [X, Yc], N))
    _memo[key] = val
    return val


def _canon(t):
    """Canonical rotation of a cyclic word with indices renamed by first appearance is overkill here:
    use the lexicographically smallest rotation."""
    if not t:
        return ()
    rots = [t[k:] + t[:k] for k in range(len(t))]
    return min(rots)


def color_matrix(basis):
    """C[s][s'] = sum over colours of (T^s)_{ij} conj((T^s')_{ij}) = Tr(T^{s1}..T^{sn} T^{s'n}..T^{s'1})."""
    memo = {}
    rows = []
    for s in basis:
        row = []
        for sp in basis:
            # relabel so that s = (0,1,2,..): the value depends on the relative permutation only
            pos = {g: k for k, g in enumerate(s)}
            rel = tuple(pos[g] for g in sp)
            if rel not in memo:
                word = tuple(range(len(s))) + tuple(reversed(rel))
                memo[rel] = _trace_value([word])
            row.append(memo[rel])
        rows.append(row)
    return rows


def integer_rows(rows):
    """Per-row common denominator and integer numerators (MG5: get_line_denominators/numerators)."""
    nums, dens = [], []
    for row in rows:
        d = 1
        for v in row:
            d = d * v.denominator // math.gcd(d, v.denominator)
        dens.append(d)
        nums.append([int(v * d) for v in row])
    return nums, dens


# ------------------------------------------------------------------------------------------------
def _coef(c):
    """Complex coefficient -> (re, im) floats; the QCD coefficients are Gaussian integers."""
    c = complex(c)
    return float(round(c.real)), float(round(c.imag))


def generate_ir(n_final_gluons, name=None, root="centroid"):
    """IR of  g g > t t~ + n_final_gluons g.  From two final-state gluons on the IR also carries ir["plan"]: the
    colour-reduced evaluation plan of the same sum of diagrams (madflow_b200/recursion.py), which the helicity-parallel
    kernels evaluate instead of the diagram list; the oracle and the reference-style call list do not know it."""
    k = n_final_gluons
    gen = Generator(k)
    n = gen.n
    amps = gen.amplitudes(root)
    basis = list(itertools.permutations(gen.gluons))
    bindex = {w: j for j, w in enumerate(basis)}

    # schedule: externals, then for every amplitude (in generation order) the currents it needs
    calls, slot_of, order = [], {}, []
    uid = itertools.count()

    def visit(nd):
        if nd.uid is not None:
            return
        for ch in nd.children:
            visit(ch)
        nd.uid = next(uid)
        order.append(nd)

    for leg in range(n):
        visit(gen.externals[leg])
    items = []  # ("wf", node) | ("amp", index)
    for nd in order:
        items.append(("wf", nd))
    for a_idx, (op, children, col, topo) in enumerate(amps):
        before = len(order)
        for ch in children:
            visit(ch)
        for nd in order[before:]:
            items.append(("wf", nd))
        items.append(("amp", a_idx))

    # last use of every wavefunction -> slot reuse (MG5 reuses w[] slots the same way)
    last_use = {}
    for pos, (what, obj) in enumerate(items):
        kids = obj.children if what == "wf" else amps[obj][1]
        for ch in kids:
            last_use[ch.uid] = pos
    free, slot_of, nslots = [], {}, 0
    release_at = {}
    for uid_, pos in last_use.items():
        release_at.setdefault(pos, []).append(uid_)
    mass_of = {"o": "mdl_MT", "i": "mdl_MT", "g": "ZERO"}
    width_of = {"o": "mdl_WT", "i": "mdl_WT", "g": "ZERO"}
    coup_of = {"FFV1": "GC_11", "VVV1": "GC_10", "VVVV": "GC_12"}
    for pos, (what, obj) in enumerate(items):
        if what == "wf":
            nd = obj
            # inputs may be released only after the output is written: allocate first
            if free:
                slot = free.pop(0)
            else:
                slot = nslots
                nslots += 1
            slot_of[nd.uid] = slot
            if not nd.children:
                (leg,) = nd.legs
                nsf = {"vxxxxx": (-1 if leg < 2 else 1), "oxxxxx": 1, "ixxxxx": -1}[nd.op]
                calls.append({"op": nd.op, "out": slot, "leg": leg, "mass": mass_of[nd.kind], "nsf": nsf})
            else:
                calls.append({"op": nd.op, "out": slot, "in": [slot_of[c.uid] for c in nd.children],
                              "coup": coup_of[nd.op[:4]], "mass": mass_of[nd.kind], "width": width_of[nd.kind]})
            if nd.uid not in last_use:  # never used (cannot happen for a valid schedule)
                free.append(slot)
        else:
            op, children, col, topo = amps[obj]
            calls.append({"op": op, "amp": obj, "in": [slot_of[c.uid] for c in children], "coup": coup_of[op[:4]]})
        for uid_ in release_at.get(pos, []):
            free.append(slot_of[uid_])
            free.sort()

    jamp = [[] for _ in basis]
    for a_idx, (op, children, col, topo) in enumerate(amps):
        for w, c in col.items():
            # overall sign chosen so that g g > t t~ reproduces MG5's jamp line
            # (tests/mockup_debug_me.py:530: jamp = [i amp0 - amp1, -i amp0 - amp2]); |M|^2 cannot see it
            re, im = _coef(-c)
            jamp[bindex[w]].append((a_idx, re, im))

    nums, dens = integer_rows(color_matrix(basis))
    hel_states = [[-1, 1]] * n
    hel_states[3] = [1, -1]  # MG5 lists the antiquark's helicities reversed (mockup_debug_me.py:424-441)
    helicities = [list(h) for h in itertools.product(*hel_states)]
    ndiags = len({topo.replace("W1", "W").replace("W3", "W").replace("W4", "W") + ("" if not op.startswith("VVVV") else "")
                  for op, _, _, topo in amps})
    suffix = "g" * k
    ir = {
        "name": name or f"1_gg_ttx{suffix}",
        "process": "g g > t t~" + " g" * k + " WEIGHTED<=%d @1" % (2 + k),
        "nexternal": n, "ninitial": 2, "ndiags": ndiags, "ncomb": 2**n, "nwavefuncs": nslots,
        "helicities": helicities,
        "denominator": 4 * 64 * math.factorial(k),
        "params": ["mdl_MT", "mdl_WT"],
        "couplings": sorted({c["coup"] for c in calls if "coup" in c}),
        "initial_states": [[21, 21]], "mirror_initial_states": False,
        "pdg": [21, 21, 6, -6] + [21] * k,
        "masses": ["ZERO", "ZERO", "mdl_MT", "mdl_MT"] + ["ZERO"] * k,
        "calls": calls, "jamp": jamp, "color_num": nums, "color_denom": dens,
        "color_basis": [list(w) for w in basis],
    }
    if k >= 2 and root == "centroid":
        from . import recursion

        ir["plan"] = recursion.build_plan(k, {c["leg"]: c for c in calls if "leg" in c})
    return ir


def qqbar_ttx_ir(pp=True, name="1_uux_ttx"):
    """IR of  q q~ > t t~  (one s-channel gluon), the second subprocess of madflow's default `p p > t t~`
    (MG5 writes it as matrix_1_uux_ttx with all light flavours in `initial_states`; `mirror_initial_states`
    because either proton may supply the quark: PyOut_exporter.py:186-191, madflow_exec.py:141-155).
    MG5 itself is absent, so like every generated process this is restated from the Feynman rules and checked
    by a closed form (tests/test_procgen.py): parity unpinned with respect to MG5's own output.

      legs    0 = q (incoming fermion, ixxxxx nsf +1), 1 = q~ (oxxxxx nsf -1), 2 = t, 3 = t~
      colour  T^a_{..} T^a_{..} = 1/2 (delta delta - 1/N delta delta): two colour flows with JAMPs
              (1/2, -1/6) x amp and colour matrix [[9, 3], [3, 9]]  (sum = 2 = Tr(T^a T^b)^2 x 8 ... = (N^2-1)/4)
      average 4 spin states x 9 colours"""
    hel_states = [[1, -1], [-1, 1], [-1, 1], [1, -1]]   # antiparticle-like legs listed reversed, as for t~ in g g > t t~
    flavours = [[2, -2], [4, -4], [1, -1], [3, -3]] if pp else [[2, -2]]
    return {
        "name": name,
        "process": "u u~ > t t~ WEIGHTED<=2 @1",
        "nexternal": 4, "ninitial": 2, "ndiags": 1, "ncomb": 16, "nwavefuncs": 5,
        "helicities": [list(h) for h in itertools.product(*hel_states)],
        "denominator": 36,
        "params": ["mdl_MT", "mdl_WT"],
        "couplings": ["GC_11"],
        "initial_states": flavours, "mirror_initial_states": bool(pp),
        "pdg": [2, -2, 6, -6],
        "masses": ["ZERO", "ZERO", "mdl_MT", "mdl_MT"],
        "calls": [
            {"op": "ixxxxx", "out": 0, "leg": 0, "mass": "ZERO", "nsf": +1},
            {"op": "oxxxxx", "out": 1, "leg": 1, "mass": "ZERO", "nsf": -1},
            {"op": "oxxxxx", "out": 2, "leg": 2, "mass": "mdl_MT", "nsf": +1},
            {"op": "ixxxxx", "out": 3, "leg": 3, "mass": "mdl_MT", "nsf": -1},
            {"op": "FFV1P0_3", "out": 4, "in": [0, 1], "coup": "GC_11", "mass": "ZERO", "width": "ZERO"},
            {"op": "FFV1_0", "amp": 0, "in": [3, 2, 4], "coup": "GC_11"},
        ],
        "jamp": [[(0, 0.5, 0.0)], [(0, -1.0 / 6.0, 0.0)]],
        "color_num": [[9, 3], [3, 9]],
        "color_denom": [1, 1],
    }


# ------------------------------------------------------------------------------------------------
# two quark lines: q q~ > t t~ g and its crossings
LIGHT_LINE_KINDS = ("uux_ttxg", "gu_ttxu", "gux_ttxux")


def _su3():
    """Generators T^a = lambda^a / 2 and structure constants f^{abc} of SU(3), numerically."""
    import numpy as np

    lam = np.zeros((8, 3, 3), dtype=complex)
    lam[0][0, 1] = lam[0][1, 0] = 1
    lam[1][0, 1], lam[1][1, 0] = -1j, 1j
    lam[2][0, 0], lam[2][1, 1] = 1, -1
    lam[3][0, 2] = lam[3][2, 0] = 1
    lam[4][0, 2], lam[4][2, 0] = -1j, 1j
    lam[5][1, 2] = lam[5][2, 1] = 1
    lam[6][1, 2], lam[6][2, 1] = -1j, 1j
    lam[7] = np.diag([1, 1, -2]) / math.sqrt(3)
    T = lam / 2
    f = np.zeros((8, 8, 8))
    for a in range(8):
        for b in range(8):
            comm = T[a] @ T[b] - T[b] @ T[a]          # [T^a, T^b] = i f^{abc} T^c,  Tr(T^c T^d) = delta/2
            for c in range(8):
                f[a, b, c] = (-2j * np.trace(comm @ T[c])).real
    return T, f


def light_line_ttxg_ir(kind="uux_ttxg"):
    """IR of the five-point processes with a light quark line, the subprocesses of `p p > t t~ j` besides
    g g > t t~ g:   "uux_ttxg"  q q~ > t t~ g,   "gu_ttxu"  g q > t t~ q,   "gux_ttxux"  g q~ > t t~ q~
    (MG5's names for the first flavour; `initial_states` lists all light flavours, `mirror_initial_states`
    because either proton may supply either parton -- PyOut_exporter.py:186-191, madflow_exec.py:141-155).

    Five diagrams -- the gluon attached to the top line (2), to the light line (2), to the exchanged gluon (1).  The
    three processes are crossings of one amplitude: only the external wavefunctions differ (which leg is the
    fermion-flow-in end `I` of the light line, which the flow-out end `O`, which the gluon `G`).  Colour: the tensor
    of every diagram (generators and structure constants of SU(3) written out numerically) is projected on the four
    colour flows  T^a_{t I} d_{O tb},  d_{t I} T^a_{O tb},  T^a_{t tb} d_{O I},  d_{t tb} T^a_{O I};  the colour matrix
    is the Gram matrix of the flows.  MG5's own output is not in the reference tree: parity unpinned; checked by BRST
    invariance of every colour flow, by the explicit colour sum and on the generated CUDA code (tests/test_procgen.py)."""
    import numpy as np

    roles = {   # leg -> (wavefunction call, nsf); I / O = ends of the light line, G = the gluon
        "uux_ttxg": dict(I=(0, "ixxxxx", +1), O=(1, "oxxxxx", -1), G=(4, "vxxxxx", +1), pdg=[2, -2, 6, -6, 21],
                         initial=[[2, -2], [4, -4], [1, -1], [3, -3]], colour_avg=9, process="u u~ > t t~ g"),
        "gu_ttxu": dict(G=(0, "vxxxxx", -1), I=(1, "ixxxxx", +1), O=(4, "oxxxxx", +1), pdg=[21, 2, 6, -6, 2],
                        initial=[[21, 2], [21, 4], [21, 1], [21, 3]], colour_avg=24, process="g u > t t~ u"),
        "gux_ttxux": dict(G=(0, "vxxxxx", -1), O=(1, "oxxxxx", -1), I=(4, "ixxxxx", -1), pdg=[21, -2, 6, -6, -2],
                          initial=[[21, -2], [21, -4], [21, -1], [21, -3]], colour_avg=24, process="g u~ > t t~ u~"),
    }[kind]
    ext = {2: ("oxxxxx", +1, "mdl_MT"), 3: ("ixxxxx", -1, "mdl_MT")}
    for key in "IOG":
        leg, op, nsf = roles[key]
        ext[leg] = (op, nsf, "ZERO")
    calls = [{"op": ext[leg][0], "out": leg, "leg": leg, "mass": ext[leg][2], "nsf": ext[leg][1]} for leg in range(5)]
    I_, O_, G_, T_, TB_ = roles["I"][0], roles["O"][0], roles["G"][0], 2, 3
    top = dict(coup="GC_11", mass="mdl_MT", width="mdl_WT")
    light = dict(coup="GC_11", mass="ZERO", width="ZERO")
    calls += [
        dict(op="FFV1P0_3", out=5, **{"in": [I_, O_]}, **light),              # gluon from the light line
        dict(op="FFV1_1", out=6, **{"in": [T_, G_]}, **top),
        dict(op="FFV1_0", amp=0, **{"in": [TB_, 6, 5]}, coup="GC_11"),        # gluon radiated off the t
        dict(op="FFV1_2", out=6, **{"in": [TB_, G_]}, **top),
        dict(op="FFV1_0", amp=1, **{"in": [6, T_, 5]}, coup="GC_11"),         # ... off the t~
        dict(op="FFV1P0_3", out=7, **{"in": [TB_, T_]}, **light),             # gluon from the top line
        dict(op="FFV1_2", out=6, **{"in": [I_, G_]}, **light),
        dict(op="FFV1_0", amp=2, **{"in": [6, O_, 7]}, coup="GC_11"),         # ... off the I end of the light line
        dict(op="FFV1_1", out=6, **{"in": [O_, G_]}, **light),
        dict(op="FFV1_0", amp=3, **{"in": [I_, 6, 7]}, coup="GC_11"),         # ... off the O end
        dict(op="VVV1_0", amp=4, **{"in": [7, 5, G_]}, coup="GC_10"),         # ... off the exchanged gluon
    ]
    # colour tensors [t, tb, O, I, a]: a fermion line contributes (T T ..)_{out end, in end}, radiation ordered from the
    # outgoing end; the three-gluon vertex f^{123} in the order of the VVV1_0 arguments
    T, f = _su3()
    d3 = np.eye(3)
    TT = np.einsum("aij,bjk->abik", T, T)
    diagrams = [np.einsum("abik,bol->ikola", TT, T), np.einsum("baik,bol->ikola", TT, T),
                np.einsum("bik,baol->ikola", T, TT), np.einsum("bik,abol->ikola", T, TT),
                np.einsum("bca,bik,col->ikola", f, T, T)]
    flows = [np.einsum("ail,ok->ikola", T, d3), np.einsum("il,aok->ikola", d3, T),
             np.einsum("aik,ol->ikola", T, d3), np.einsum("ik,aol->ikola", d3, T)]
    B = np.stack([b.reshape(-1) for b in flows], axis=1)
    jamp = [[] for _ in flows]
    for d_idx, tensor in enumerate(diagrams):
        coef, *_ = np.linalg.lstsq(B, tensor.reshape(-1), rcond=None)
        assert np.allclose(B @ coef, tensor.reshape(-1), atol=1e-12), "the colour flows do not span this diagram"
        for k_, c in enumerate(coef):
            re = Fraction(float(c.real)).limit_denominator(36)
            im = Fraction(float(c.imag)).limit_denominator(36)
            assert abs(complex(re, im) - c) < 1e-12
            if re or im:
                jamp[k_].append((d_idx, -float(re), -float(im)))   # overall sign as in generate_ir
    gram = (B.conj().T @ B).real
    rows = [[Fraction(float(v)).limit_denominator(36) for v in row] for row in gram]
    assert np.allclose([[float(v) for v in row] for row in rows], gram, atol=1e-12) and np.allclose(gram, gram.T)
    nums, dens = integer_rows(rows)
    hel_states = [[-1, 1]] * 5
    for leg, (op, nsf, _) in ext.items():
        if (op == "ixxxxx") == (leg >= 2):   # antiparticle-like legs (incoming fermion, outgoing antifermion) listed reversed
            hel_states[leg] = [1, -1]
    return {
        "name": "1_" + kind, "process": roles["process"] + " WEIGHTED<=3 @1",
        "nexternal": 5, "ninitial": 2, "ndiags": 5, "ncomb": 32, "nwavefuncs": 8,
        "helicities": [list(h) for h in itertools.product(*hel_states)],
        "denominator": 4 * roles["colour_avg"],
        "params": ["mdl_MT", "mdl_WT"], "couplings": ["GC_10", "GC_11"],
        "initial_states": roles["initial"], "mirror_initial_states": True,
        "pdg": roles["pdg"], "masses": ["ZERO", "ZERO", "mdl_MT", "mdl_MT", "ZERO"],
        "calls": calls, "jamp": jamp, "color_num": nums, "color_denom": dens,
    }


# `p p > ...` processes of the command line: the subprocess libraries whose luminosity-weighted matrix elements
# are summed (madflow_exec.py:444-455)
MULTI_PROCESSES = {"p p > t t~": ["1_gg_ttx", "1_uux_ttx"],
                   "p p > t t~ g": ["1_gg_ttxg", "1_uux_ttxg"],
                   "p p > t t~ j": ["1_gg_ttxg", "1_gu_ttxu", "1_gux_ttxux", "1_uux_ttxg"],
                   # the light-line six-point libraries come from procgen_lines and are built on demand
                   # (python -m madflow_b200.build 1_uux_ttxgg 1_gu_ttxug ...), the four-quark ones included
                   "p p > t t~ g g": ["1_gg_ttxgg", "1_uux_ttxgg"],
                   "p p > t t~ j j": ["1_gg_ttxgg", "1_gg_ttxuux", "1_gu_ttxug", "1_gux_ttxuxg", "1_uux_ttxgg", "1_uu_ttxuu",
                                      "1_ud_ttxud", "1_uxux_ttxuxux", "1_uxdx_ttxuxdx", "1_uux_ttxuux", "1_uux_ttxddx",
                                      "1_udx_ttxudx"]}


def builtin_irs():
    """Processes compiled into the package besides the pinned g g > t t~."""
    return [generate_ir(1), generate_ir(2), generate_ir(3), qqbar_ttx_ir()] + [light_line_ttxg_ir(k) for k in LIGHT_LINE_KINDS]
