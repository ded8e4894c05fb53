"""CUDA emitter: process IR -> one sm_100a translation unit per process.

This is the backend the north star asks the `pyout` plugin to gain.  It does the job of the
reference's regex transpiler (python_package/madflow/custom_op_generator.py:22-118 and
custom_op/*.py, which turn the generated Python `matrix()` into a TensorFlow custom op evaluated
one helicity per launch) -- but from the structured IR instead of from Python text, and into ONE
fused kernel per process: all helicities, JAMP sums and the colour contraction
(see csrc/process_kernels.cuh).

The emitted `Proc::matrix()` is straight-line code: the ordered HELAS call list with MG5-style
wavefunction slot reuse, JAMPs accumulated as soon as each amplitude exists (same left-to-right
order as the reference's jamp line, PyOut_exporter.py:334-375), and the colour quadratic form
Re sum_ij J_i cf_ij conj(J_j)/denom_j (matrix_method_python.inc:137) with the integer matrix as
compile-time constants (symmetric form when all row denominators are equal).
"""
import os
import subprocess
from fractions import Fraction

from . import process_ir

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIBDIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
GENDIR = os.path.join(CSRC, "generated")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-shared",
]

# Algorithmic FP64 flop per call, reference form of each routine (complex*complex = 6,
# complex*real = 2, complex+-complex = 2, complex/complex = 11, *(+-1), *(+-i) free; momenta,
# masses, widths real).  The first four are SURVEY.md section 8(d)'s counts of the frozen MG5
# routines; the others are counted the same way on the expression MG5's ALOHA writes for them.
FLOPS = {
    "vxxxxx": 30, "ixxxxx": 30, "oxxxxx": 30, "sxxxxx": 2,
    "FFV1_0": 108, "FFV1_1": 370, "FFV1_2": 370, "VVV1P0_1": 248,
    "VVV1_0": 208, "FFV1P0_3": 176,
    "VVVV1_0": 140, "VVVV3_0": 140, "VVVV4_0": 140,
    "VVVV1P0_1": 176, "VVVV3P0_1": 176, "VVVV4P0_1": 176,
}

# known alpha_s dependence of the SM QCD couplings: c = (re + i im) * G^power
# GC_10, GC_11 pinned by tests/mockup_debug_me.py:24-25; GC_12 = i G^2 is [EXT] models/sm.
COUPLING_DEFS = {"GC_10": (-1.0, 0.0, 1), "GC_11": (0.0, 1.0, 1), "GC_12": (0.0, 1.0, 2)}


def flops_per_event(ir):
    """F_alg of SURVEY.md section 8(d): ncomb*(sum calls + F_jamp + F_colour) + ncomb + 1."""
    per_hel = sum(FLOPS[c["op"]] for c in ir["calls"])
    fj = 0
    for terms in ir["jamp"]:
        fj += 2 * (len(terms) - 1)
        for _, re, im in terms:
            if (abs(re), abs(im)) not in ((1.0, 0.0), (0.0, 1.0)):
                fj += 2
    ncolor = len(ir["jamp"])
    per_hel += fj + ncolor * ncolor * 10 + 2 * ncolor
    return ir["ncomb"] * per_hel + ir["ncomb"] + 1


def _cname(ir):
    return "".join(ch if ch.isalnum() else "_" for ch in ir["name"])


def _par_expr(ir, name):
    if name == "ZERO":
        return "0.0"
    return f"par[{ir['params'].index(name)}]"


def _coup_expr(ir, c):
    e = f"coup[{ir['couplings'].index(c['coup'])}]"
    return f"(-{e})" if c.get("coup_sign", 1) < 0 else e


def _emit_call(ir, c):
    op = c["op"]
    if op in ("vxxxxx", "ixxxxx", "oxxxxx"):
        extra = "sqh, " if op == "vxxxxx" else ""
        return (f"mf::{op}(p[{c['leg']}], {_par_expr(ir, c['mass'])}, P_hel(icomb, {c['leg']}), {c['nsf']}, "
                f"{extra}w{c['out']});")
    if op == "sxxxxx":
        return f"mf::sxxxxx(p[{c['leg']}], {c['nsf']}, w{c['out']});"
    ins = ", ".join(f"w{i}" for i in c["in"])
    fn = op
    if op.startswith("VVVV"):
        kind = op[4]
        fn = f"VVVV_0<{kind}>" if op.endswith("_0") else f"VVVVP0_1<{kind}>"
    if "amp" in c:
        return f"mf::{fn}({ins}, {_coup_expr(ir, c)})"
    return (f"mf::{fn}({ins}, {_coup_expr(ir, c)}, {_par_expr(ir, c['mass'])}, {_par_expr(ir, c['width'])}, "
            f"w{c['out']});")


def _jamp_update(j, re, im, amp):
    """C++ statement adding (re + i im)*amp to jamp j."""
    if im == 0.0:
        if re == 1.0:
            return f"J{j} += {amp};"
        if re == -1.0:
            return f"J{j} -= {amp};"
        return f"J{j} += {re!r} * {amp};"
    if re == 0.0:
        if im == 1.0:
            return f"J{j} += mul_i({amp});"
        if im == -1.0:
            return f"J{j} += mul_mi({amp});"
        return f"J{j} += {im!r} * mul_i({amp});"
    return f"J{j} += mk({re!r}, {im!r}) * {amp};"


def _emit_colour(ir, J=lambda i: f"J{i}"):
    """Colour quadratic form Re sum_ij J_i cf_ij conj(J_j)/denom_j (matrix_method_python.inc:137) with the
    integer matrix as compile-time constants; symmetric form when all row denominators are equal."""
    lines = []
    ncolor = len(ir["jamp"])
    cf, den = ir["color_num"], ir["color_denom"]
    uniform = len(set(den)) == 1 and all(cf[i][j] == cf[j][i] for i in range(ncolor) for j in range(ncolor))
    if uniform:
        lines.append("    double me = 0.0;")
        for i in range(ncolor):
            terms = []
            for j in range(i + 1, ncolor):
                if cf[i][j] != 0:
                    terms.append((j, 2 * cf[i][j]))
            # row i: J_i . (cf_ii J_i + sum_{j>i} 2 cf_ij J_j)
            lines.append(f"    {{ double tr = {float(cf[i][i])!r} * {J(i)}.re, ti = {float(cf[i][i])!r} * {J(i)}.im;")
            for j, v in terms:
                lines.append(f"      tr += {float(v)!r} * {J(j)}.re; ti += {float(v)!r} * {J(j)}.im;")
            lines.append(f"      me += {J(i)}.re * tr + {J(i)}.im * ti; }}")
        lines.append(f"    return me / {float(den[0])!r};")
    else:
        lines.append("    double me = 0.0;")
        for j in range(ncolor):
            lines.append("    { cxd z = mk(0.0, 0.0);")
            for i in range(ncolor):
                if cf[i][j] != 0:
                    lines.append(f"      z += {float(cf[i][j])!r} * {J(i)};")
            lines.append(f"      me += (z.re * {J(j)}.re + z.im * {J(j)}.im) / {float(den[j])!r}; }}")
        lines.append("    return me;")
    return "\n".join(lines)


def emit_matrix_body(ir):
    lines = []
    slots = sorted({c["out"] for c in ir["calls"] if "out" in c})
    lines.append("    cxd " + ", ".join(f"w{s}[6]" for s in slots) + ";")
    ncolor = len(ir["jamp"])
    lines.append("    cxd " + ", ".join(f"J{j} = mk(0.0, 0.0)" for j in range(ncolor)) + ";")
    by_amp = {}
    for j, terms in enumerate(ir["jamp"]):
        for k, re, im in terms:
            by_amp.setdefault(k, []).append((j, float(re), float(im)))
    for c in ir["calls"]:
        if "amp" in c:
            k = c["amp"]
            uses = by_amp.get(k, [])
            if not uses:
                continue
            lines.append(f"    {{ const cxd amp = {_emit_call(ir, c)};")
            for j, re, im in uses:
                lines.append("      " + _jamp_update(j, re, im, "amp"))
            lines.append("    }")
        else:
            lines.append("    " + _emit_call(ir, c))
    lines.append(_emit_colour(ir))
    return "\n".join(lines)


# ------------------------------------------------------------------------------------------------
# helicity-parallel variant (csrc/process_kernels_hp.cuh)
STRAIGHT_LINE_MAX_CALLS = 400   # longer call lists are only emitted in the helicity-parallel (table) form
THREAD_MAX_CALLS = 64     # call lists up to this length also get the one-event-per-thread kernels
HP_UNROLL_MAX_AMPS = 32   # amplitude lists up to this length are emitted as straight-line code
HP_TYPES = {"vxxxxx": 0, "oxxxxx": 1, "ixxxxx": 2, "FFV1_1": 3, "FFV1_2": 4, "FFV1P0_3": 5, "VVV1P0_1": 6,
            "VVVV1P0_1": 7, "VVVV3P0_1": 8, "VVVV4P0_1": 9}


def hp_analyse(ir):
    """Undo the slot reuse of the call list: every write creates a distinct wavefunction with the set
    of external legs below it.  Returns (wfs, exts, items, amps, cxd per event)."""
    cur = {}      # slot -> wavefunction id
    wfs, exts, items, amps = [], [], [], []
    for c in ir["calls"]:
        if "leg" in c:
            w = len(wfs)
            wfs.append({"legs": (c["leg"],)})
            cur[c["out"]] = w
            exts.append({"call": c, "out": w})
        elif "amp" in c:
            amps.append({"call": c, "in": [cur[s] for s in c["in"]]})
        else:
            ins = [cur[s] for s in c["in"]]
            legs = tuple(sorted(set().union(*[wfs[i]["legs"] for i in ins])))
            assert len(legs) == sum(len(wfs[i]["legs"]) for i in ins), "children must not share legs"
            w = len(wfs)
            wfs.append({"legs": legs})
            cur[c["out"]] = w
            items.append({"call": c, "in": ins, "out": w})
    off = 0
    for w in wfs:
        w["level"] = len(w["legs"])
        w["nv"] = 1 << w["level"]
        w["off"] = off
        w["mask"] = sum(1 << l for l in w["legs"])
        off += 2 + 4 * w["nv"]
    items.sort(key=lambda it: (wfs[it["out"]]["level"], HP_TYPES[it["call"]["op"]]))
    return wfs, exts, items, amps, off


def hp_config(ir, key):
    """Launch shape of the helicity-parallel kernels, MADFLOW_B200_HP_<KEY> overrides (tools/build_variants.py):
      E          events per block
      NCG        colour groups = threads per (event, helicity combination); each keeps NCOLOR/NCG JAMPs in registers
      NB         rows of the amplitude buffer = amplitudes per batch
      SCRATCH    shared-memory scratch for the pair objects of a batch, complex numbers per event
      MINBLOCKS  resident blocks per SM the register allocation aims at
    Defaults from measurements on B200: up to 64 helicity combinations two events per
    block and all JAMPs in one thread; beyond, one event per block (its wavefunctions fill a third of the
    shared memory), 8 colour groups and batches of 64 amplitudes; up to 64 combinations the batch size is what
    still lets two blocks of two events share an SM (21 rows: 14.5e6 events/s for g g > t t~ g g, 16 rows: 13.8e6)."""
    env = os.environ.get("MADFLOW_B200_HP_" + key)
    if env:
        return int(env)
    if ir["ncomb"] > 64:
        return {"E": 1, "NCG": 8, "NB": 64, "SCRATCH": 4096, "MINBLOCKS": 1}[key]
    return {"E": max(1, 128 // ir["ncomb"]), "NCG": 1, "NB": 21, "SCRATCH": 512, "MINBLOCKS": 2}[key]


def hp_chain(ir):
    """Amplitude tiles that accumulate in the tensor-core accumulators: amplitudes of one batch with the same colour
    signature (up to +-1, +-i) and the same split of the legs between pair object and wavefunction are stored ONCE
    (1 = on; MADFLOW_B200_HP_CHAIN overrides).  The batches already keep pair objects of one kind and colour signature
    together, which is what puts chain members into one batch (ordering by legs first was tried: not better).
    Default off until measured on the GPU (plan for round 2)."""
    env = os.environ.get("MADFLOW_B200_HP_CHAIN")
    return int(env) if env else 0


def hp_passes(ir):
    """Helicity passes: with more than 64 helicity combinations the amplitude / JAMP / colour phases run
    once per helicity of the last external leg (64 combinations per pass), so that the JAMPs of a pass
    fit in registers and their exchange area in shared memory."""
    if os.environ.get("MADFLOW_B200_HP_NPASS"):
        return int(os.environ["MADFLOW_B200_HP_NPASS"])
    return max(1, ir["ncomb"] // 64)


def hp_colour_mode(ir, ncg):
    """'thread': each thread owns all JAMPs of its helicity; 'groups': generated code, JAMPs exchanged through
    shared memory; 'mma' / 'loop': table driven over the block-symmetrised colour matrix, on the FP64 tensor
    cores / on the CUDA cores (the A/B pair behind DESIGN.md's tensor-core decision)."""
    mode = os.environ.get("MADFLOW_B200_HP_COLOUR")
    if mode:
        return mode
    if ncg == 1:
        return "thread"
    return "mma" if len(ir["jamp"]) >= 48 else "groups"


def emit_hp(ir):
    """Tables + the generated amplitude/JAMP/colour code of the helicity-parallel kernels."""
    wfs, exts, items, amps, wfsize = hp_analyse(ir)
    n = ir["nexternal"]
    assert ir["ncomb"] == 2**n, "the hp kernels need the full 2^n helicity table"
    maxlevel = max(w["level"] for w in wfs)
    NH = ir["ncomb"]
    NPASS = hp_passes(ir)
    assert NPASS in (1, 2)
    NHP = NH // NPASS
    LSTAR = n - 1 if NPASS > 1 else None   # the leg whose helicity is fixed within a pass (top variant bit)
    big = len(amps) > 400                  # tables beyond the 64 KB of constant memory live in global memory

    def pidx(name):
        return -1 if name == "ZERO" else ir["params"].index(name)

    def both(ctype, name, count, body, const=True):
        space = "__device__ __constant__" if const else "__device__ const"
        return (f"{space} {ctype} d_{name}[{count}] = {{{body}}};\n"
                f"static const {ctype} h_{name}[{count}] = {{{body}}};")

    def vmask(out_legs, in_legs):
        """bits of the output's variant index that make up the input's variant index (both ascending)"""
        return sum(1 << q for q, l in enumerate(out_legs) if l in in_legs)

    L = []
    L.append(both("mf::HpWf", "wf", len(wfs), ", ".join(f"{{{w['off']}u, {w['nv']}, {w['mask']}}}" for w in wfs)))
    ext_by_leg = sorted(exts, key=lambda x: x["call"]["leg"])
    assert [x["call"]["leg"] for x in ext_by_leg] == list(range(n))
    L.append(both("mf::HpExt", "ext", n, ", ".join(
        f"{{{HP_TYPES[x['call']['op']]}, {x['call']['leg']}, {x['call']['nsf']}, {pidx(x['call']['mass'])}, {x['out']}}}"
        for x in ext_by_leg)))
    rows = []
    for it in items:
        c = it["call"]
        ins = it["in"] + [0] * (3 - len(it["in"]))
        vm = [vmask(wfs[it["out"]]["legs"], wfs[i]["legs"]) for i in it["in"]] + [0] * (3 - len(it["in"]))
        ioff = [wfs[i]["off"] for i in ins]
        inv = [wfs[i]["nv"] for i in ins]
        W = wfs[it["out"]]
        rows.append(f"{{{HP_TYPES[c['op']]}, {len(it['in'])}, {pidx(c['mass'])}, {pidx(c['width'])}, "
                    f"{ir['couplings'].index(c['coup'])}, {1 if c.get('coup_sign', 1) < 0 else 0}, {W['off']}, {W['nv']}, "
                    f"{{{ioff[0]}, {ioff[1]}, {ioff[2]}}}, {{{inv[0]}, {inv[1]}, {inv[2]}}}, "
                    f"{{{vm[0]}, {vm[1]}, {vm[2]}}}}}")
    L.append(both("mf::HpItem", "items", max(len(rows), 1), ",\n  ".join(rows) if rows else "{0}"))

    def pext(v, m):
        return sum(((v >> b_) & 1) << q for q, b_ in enumerate(b2 for b2 in range(5) if m >> b2 & 1))

    # work items of the current phases: (current, variant, variants of the inputs), level by level
    crow, begins = [], []
    for lev in range(0, maxlevel + 2):
        begins.append(len(crow))
        for idx, it in enumerate(items):
            W = wfs[it["out"]]
            if W["level"] != lev:
                continue
            masks = [vmask(W["legs"], wfs[i]["legs"]) for i in it["in"]] + [0] * (3 - len(it["in"]))
            for v in range(W["nv"]):
                crow.append(f"{{{idx}, {v}, {{{pext(v, masks[0])}, {pext(v, masks[1])}, {pext(v, masks[2])}}}, 0}}")
    L.append(both("mf::HpPairItem", "cur_items", max(len(crow), 1), ", ".join(crow) if crow else "{0, 0, {0, 0, 0}, 0}", const=False))
    L.append(both("int", "level_begin", len(begins), ", ".join(map(str, begins))))
    tables = "\n".join(L)

    by_amp = {}
    for j, terms in enumerate(ir["jamp"]):
        for k, re, im in terms:
            by_amp.setdefault(k, []).append((j, float(re), float(im)))
    used = [am for am in amps if by_amp.get(am["call"]["amp"])]

    # pair objects: amp = x . Q(rest of the vertex), x = the input with the most legs (with helicity passes:
    # preferably one that does not hold the pass leg, so that the pass halves the rows and not the columns)
    QUARTIC = {"1": ((+1, (1, 4), (2, 3)), (-1, (1, 3), (2, 4))),
               "3": ((+1, (1, 4), (2, 3)), (-1, (1, 2), (3, 4))),
               "4": ((+1, (1, 3), (2, 4)), (-1, (1, 2), (3, 4)))}
    pairs, pair_index, amp_rows = [], {}, []
    for am in used:
        c = am["call"]
        ins = am["in"]
        sizes = [wfs[w]["level"] for w in ins]
        cands = [q for q in range(len(ins)) if sizes[q] == max(sizes)]
        free = [q for q in cands if LSTAR not in wfs[ins[q]]["legs"]]
        jx = (free or cands)[0]
        op = c["op"]
        term = (0, 0)
        if op == "FFV1_0":
            I, O, G = ins
            ptype, rest = [("ROW", (O, G)), ("COL", (I, G)), ("CUR", (I, O))][jx]
        elif op == "VVV1_0":
            ptype, rest = "VVV", (ins[(jx + 1) % 3], ins[(jx + 2) % 3])
        else:
            ptype = "VVVV"
            others = [q for q in range(4) if q != jx]
            rest = tuple(ins[q] for q in others)
            pos = {q + 1: others.index(q) for q in others}   # vertex position -> index into `rest`
            enc = []
            for sign, pa, pb in QUARTIC[op[4]]:
                if jx + 1 in pa:
                    vec, dot = [q for q in pa if q != jx + 1][0], pb
                else:
                    vec, dot = [q for q in pb if q != jx + 1][0], pa
                enc.append((0x40 if sign < 0 else 0) | pos[vec] << 4 | pos[dot[0]] << 2 | pos[dot[1]])
            term = tuple(enc)
        coup, neg = ir["couplings"].index(c["coup"]), 1 if c.get("coup_sign", 1) < 0 else 0
        key = (ptype, rest, term, coup, neg)
        if key not in pair_index:
            legs = tuple(sorted(set().union(*[wfs[w]["legs"] for w in rest])))
            pair_index[key] = len(pairs)
            pairs.append(dict(type=ptype, rest=rest, term=term, coup=coup, neg=neg, legs=legs, nv=1 << len(legs)))
        amp_rows.append(dict(am=am, x=ins[jx], pair=pair_index[key]))
    assert all(pr["nv"] <= 32 for pr in pairs), "pair objects hold up to 32 helicity variants"

    # batches: a batch closes at a pair boundary when the scratch area or the amplitude buffer (HP_NB rows)
    # would overflow; pair objects of one kind together, so that the warps of a batch run the same routine
    scratch = hp_config(ir, "SCRATCH")
    NB = hp_config(ir, "NB")
    NCG = hp_config(ir, "NCG")
    PT = {"ROW": 0, "COL": 1, "CUR": 2, "VVV": 3, "VVVV": 4}
    by_pair = {}
    for k, r in enumerate(amp_rows):
        by_pair.setdefault(r["pair"], []).append(k)
    batches, cur_pairs, cur_amps, fill = [], [], [], 0
    sig_id = {}
    for r in amp_rows:   # colour signature of an amplitude: its JAMP coefficients up to a common phase
        t = by_amp[r["am"]["call"]["amp"]]
        c0 = complex(t[0][1], t[0][2])
        r["sig"] = sig_id.setdefault(tuple((j, complex(re, im) / c0) for j, re, im in sorted(t)), len(sig_id))
    # ... and within a kind, pair objects whose amplitudes share a colour signature next to each other (see the JAMP code)
    chain_rows = hp_chain(ir) and len(used) > HP_UNROLL_MAX_AMPS   # rows of the amplitude buffer = chains, not amplitudes

    def row_key(k):
        r = amp_rows[k]
        return (r["sig"], wfs[r["x"]]["legs"], pairs[r["pair"]]["legs"])

    cur_keys = set()
    for pi in sorted(by_pair, key=lambda q: (PT[pairs[q]["type"]], pairs[q]["nv"], min(amp_rows[k]["sig"] for k in by_pair[q]), q)):
        need = 4 * pairs[pi]["nv"]
        assert need <= scratch and len(by_pair[pi]) <= NB
        new_keys = {row_key(k) for k in by_pair[pi]}
        rows_after = len(cur_keys | new_keys) + 4 if chain_rows else len(cur_amps) + len(by_pair[pi])   # + 4: phases that do not chain
        if fill + need > scratch or rows_after > NB:
            batches.append((cur_pairs, cur_amps))
            cur_pairs, cur_amps, fill = [], [], 0
            cur_keys = set()
        cur_keys |= new_keys
        pairs[pi]["off"] = fill
        fill += need
        cur_pairs.append(pi)
        cur_amps += by_pair[pi]
    if cur_amps:
        batches.append((cur_pairs, cur_amps))
    prow = []
    for pr in pairs:
        rest = list(pr["rest"]) + [0] * (3 - len(pr["rest"]))
        vm = [vmask(pr["legs"], wfs[w]["legs"]) for w in pr["rest"]] + [0] * (3 - len(pr["rest"]))
        ioff = [wfs[w]["off"] for w in rest]
        inv = [wfs[w]["nv"] for w in rest]
        prow.append(f"{{{PT[pr['type']]}, {len(pr['rest'])}, {pr['coup']}, {pr['neg']}, {{{pr['term'][0]}, {pr['term'][1]}}}, "
                    f"{pr['nv']}, {pr.get('off', 0)}, {{{ioff[0]}, {ioff[1]}, {ioff[2]}}}, {{{inv[0]}, {inv[1]}, {inv[2]}}}, "
                    f"{{{vm[0]}, {vm[1]}, {vm[2]}}}}}")

    def spread(legs, v):
        """helicity-combination bits (within the pass) of variant v of an object over `legs` (ascending)"""
        return sum(((v >> q) & 1) << l for q, l in enumerate(legs) if l != LSTAR)

    def vrange(legs, nv, p):
        """variants of an object that pass p needs: the half with the pass leg's helicity = p, or all"""
        if LSTAR is not None and LSTAR in legs:
            return range(p * nv // 2, (p + 1) * nv // 2)   # the pass leg is the highest leg = top variant bit
        return range(nv)

    # tensor-core tiles: amplitude(variant of Q, variant of x) = sum_k Q_k x_k is an (nvq x 4)(4 x nvx)
    # complex product; one work item = 8 variants of Q (rows) x 8 variants of x (columns)
    irow, trow, brow, urow = [], [], [], []
    ncolor = len(ir["jamp"])
    NJ = -(-ncolor // NCG)
    # rows of the amplitude buffer: one per amplitude, or (hp_chain) one per chain of amplitudes that the tile phase
    # adds up: [(index into amp_rows, phase relative to the first member), ...]
    chain_on = bool(hp_chain(ir)) and len(used) > HP_UNROLL_MAX_AMPS
    PHASE_CODE = {1: 0, -1: 1, 1j: 2, -1j: 3}

    def first_coef(k):
        t = by_amp[amp_rows[k]["am"]["call"]["amp"]]
        return complex(t[0][1], t[0][2])

    batch_rows = []
    for cur_pairs, cur_amps in batches:
        rows, index = [], {}
        for k in cur_amps:
            r = amp_rows[k]
            key = (r["sig"], wfs[r["x"]]["legs"], pairs[r["pair"]]["legs"]) if chain_on else ("own", k)
            if key in index:
                ph = first_coef(k) / first_coef(rows[index[key]][0][0])
                if ph in PHASE_CODE:
                    rows[index[key]].append((k, ph))
                    continue
                key = ("own", k)
            index[key] = len(rows)
            rows.append([(k, 1)])
        batch_rows.append(rows)
    for p in range(NPASS):
        for bi, (cur_pairs, cur_amps) in enumerate(batches):
            ib, tb = len(irow), len(trow)
            for pi in sorted(cur_pairs, key=lambda q: (PT[pairs[q]["type"]], pairs[q]["nv"], q)):
                pr = pairs[pi]
                masks = [vmask(pr["legs"], wfs[w]["legs"]) for w in pr["rest"]] + [0] * (3 - len(pr["rest"]))
                for v in vrange(pr["legs"], pr["nv"], p):
                    iv = [pext(v, m) for m in masks]
                    irow.append(f"{{{pi}, {v}, {{{iv[0]}, {iv[1]}, {iv[2]}}}, 0}}")
            ub = len(urow)
            for slot, row in enumerate(batch_rows[bi]):
                r = amp_rows[row[0][0]]
                xw, pr = wfs[r["x"]], pairs[r["pair"]]
                assert not set(xw["legs"]) & set(pr["legs"]) and len(xw["legs"]) + len(pr["legs"]) == n
                qr, xr = vrange(pr["legs"], pr["nv"], p), vrange(xw["legs"], xw["nv"], p)
                for q0 in range(qr.start, qr.stop, 8):
                    for x0 in range(xr.start, xr.stop, 8):
                        qv, xv = min(8, qr.stop - q0), min(8, xr.stop - x0)
                        rowh = [spread(pr["legs"], q0 + i) if i < qv else 0 for i in range(8)]
                        colh = [spread(xw["legs"], x0 + i) if i < xv else 0 for i in range(8)]
                        urow.append(f"{{{len(trow)}u, {len(row)}u}}")
                        for mi, (km, ph) in enumerate(row):   # the members of a chain: same geometry, own objects
                            mx, mp = wfs[amp_rows[km]["x"]], pairs[amp_rows[km]["pair"]]
                            assert (mx["legs"], mp["legs"], mx["nv"], mp["nv"]) == (xw["legs"], pr["legs"], xw["nv"], pr["nv"])
                            # flags: 1 = adds to the tile before it, 2 = the next tile adds to it, phase code << 2
                            flags = (1 if mi else 0) | (2 if mi + 1 < len(row) else 0) | PHASE_CODE[ph] << 2
                            trow.append(f"{{{mp['off']}, {mx['off'] + 2}, {mp['nv']}, {mx['nv']}, {q0}, {x0}, {qv}, {xv}, {slot}, {flags}, "
                                        f"{{{', '.join(map(str, rowh))}}}, {{{', '.join(map(str, colh))}}}}}")
            # with chains the warps take UNITS (chains of tiles, d_units) instead of single tiles
            brow.append(f"{{{ib}, {len(irow)}, {ub if chain_on else tb}, {len(urow) if chain_on else len(trow)}}}")
    # JAMP code per (batch, colour group).  Amplitudes of a batch that feed the same colours of the group with the
    # same coefficients up to a common phase (+-1, +-i) are summed first and the sum is applied once:
    #   J_c += k_c (A_0 + p_1 A_1 + ...)   instead of   J_c += k_c A_0; J_c += k_c p_1 A_1; ...
    def phase_add(dst, ph, src):
        """C++ statement dst += ph * src for ph in {1, -1, i, -i}"""
        if ph == 1:
            return f"{dst} += {src};"
        if ph == -1:
            return f"{dst} -= {src};"
        return f"{dst} += mul_i({src});" if ph == 1j else f"{dst} += mul_mi({src});"

    jamp_cases = [[] for _ in range(NCG)]
    jamp_terms = 0
    for bi, (cur_pairs, cur_amps) in enumerate(batches):
        for cg in range(NCG):
            groups = {}   # signature within the colour group -> [(slot, phase relative to the group's first amplitude)]
            for slot, row in enumerate(batch_rows[bi]):
                k = row[0][0]   # the first member of a chain carries its JAMP coefficients
                terms = [(j - cg * NJ, complex(re, im)) for j, re, im in by_amp[amp_rows[k]["am"]["call"]["amp"]] if j // NJ == cg]
                if not terms:
                    continue
                unit = all(c in (1, -1, 1j, -1j) for _, c in terms)
                c0 = terms[0][1] if unit else 1.0
                sig = tuple((jl, c / c0) for jl, c in sorted(terms)) if unit else ("own", slot)
                groups.setdefault(sig, []).append((slot, c0, terms))
            stm = []
            for sig, members in groups.items():
                slot0, c00, terms0 = members[0]
                if len(members) == 1:
                    upd = " ".join(_jamp_update(jl, c.real, c.imag, "a").replace(f"J{jl} ", f"J[{jl}] ") for jl, c in terms0)
                    stm.append(f"{{ const cxd a = ab[{slot0 * NHP}]; {upd} }}")
                    jamp_terms += len(terms0)
                    continue
                body = [f"cxd a = ab[{slot0 * NHP}];"]
                for slot, c0, _ in members[1:]:
                    body.append(phase_add("a", c0 / c00, f"ab[{slot * NHP}]"))
                body += [_jamp_update(jl, c.real, c.imag, "a").replace(f"J{jl} ", f"J[{jl}] ") for jl, c in terms0]
                stm.append("{ " + " ".join(body) + " }")
                jamp_terms += len(members) - 1 + len(terms0)
            jamp_cases[cg].append(f"      case {bi}: {{ " + "\n        ".join(stm) + " } break;")
    tables += "\n" + both("mf::HpPair", "pairs", max(len(prow), 1), ",\n  ".join(prow) if prow else "{0}", const=not big)
    tables += "\n" + both("mf::HpPairItem", "pair_items", max(len(irow), 1), ", ".join(irow) if irow else "{0, 0, {0, 0, 0}, 0}", const=False)
    tables += "\n" + both("mf::HpTile", "tiles", max(len(trow), 1), ",\n  ".join(trow) if trow else "{0}", const=len(trow) * 32 <= 24576)
    tables += "\n" + both("mf::HpBatch", "batches", max(len(brow), 1), ", ".join(brow) if brow else "{0, 0, 0, 0}")
    if chain_on:
        tables += "\n" + both("uint2", "units", max(len(urow), 1), ", ".join(urow) if urow else "{0u, 0u}", const=False)

    A = ["    switch (cg) {"]
    for cg in range(NCG):
        A.append(f"    case {cg}:")
        A.append("      switch (b) {")
        A += jamp_cases[cg]
        A.append("      default: break;")
        A.append("      }")
        A.append("      break;")
    A.append("    default: break;")
    A.append("    }")
    unroll = len(used) <= HP_UNROLL_MAX_AMPS
    cmode = "thread" if unroll else hp_colour_mode(ir, NCG)
    ncp = -(-ncolor // 8) * 8
    if cmode == "thread":
        assert NCG == 1
        C = _emit_colour(ir, J=lambda i: f"J[{i}]")
    elif cmode == "groups":
        C = _emit_colour_groups(ir, NCG, NJ, NHP)
    else:
        C = "    return 0.0;  // not used: the colour contraction of this process is table driven (d_cfsym)"
    cf, den = ir["color_num"], ir["color_denom"]
    if cmode in ("mma", "loop"):
        assert len(set(den)) == 1 and all(cf[i][j] == cf[j][i] for i in range(ncolor) for j in range(ncolor))
        # block-symmetrised colour matrix: blocks of 8x8 colours; below the block diagonal 0, on it cf, above 2 cf
        vals = []
        for a_ in range(ncp):
            for b_ in range(ncp):
                v = 0
                if a_ < ncolor and b_ < ncolor:
                    v = cf[a_][b_] * (0 if b_ // 8 < a_ // 8 else (1 if b_ // 8 == a_ // 8 else 2))
                vals.append(f"{float(v)!r}")
        tables += "\n" + both("double", "cfsym", ncp * ncp, ", ".join(vals), const=False)
    else:
        tables += "\n" + both("double", "cfsym", 1, "0.0", const=False)
    # straight-line flavour of the same phase for short amplitude lists
    U = ["    cxd " + ", ".join(f"J{j} = mk(0.0, 0.0)" for j in range(len(ir["jamp"]))) + ";",
         "    cxd a[6], b[6], c[6], d[6];"]
    if unroll:
        for am in used:
            c = am["call"]
            for q, w in enumerate(am["in"]):
                U.append(f"    mf::hp_load_amp<Proc>(wf_e, vtab, h, {w}, {'abcd'[q]});")
            op = c["op"]
            fn = f"VVVV_0<{op[4]}>" if op.startswith("VVVV") else op
            args = ", ".join("abcd"[: len(am["in"])])
            U.append(f"    {{ const cxd amp = mf::{fn}({args}, {_coup_expr(ir, c)});")
            for j, re, im in by_amp[c["amp"]]:
                U.append("      " + _jamp_update(j, re, im, "amp"))
            U.append("    }")
        U.append(_emit_colour(ir))
    else:
        U.append("    return 0.0;  // not used: the amplitudes of this process run on the tensor cores")
    return tables, "\n".join(A), C, "\n".join(U), dict(
        wfsize=wfsize, maxlevel=maxlevel, nwf=len(wfs), nitems=len(items), namps=len(used), unroll=unroll,
        nbatch=len(batches), npairs=len(pairs), nitems_pair=len(irow), ntiles=len(trow), ncg=1 if unroll else NCG, jamp_terms=jamp_terms,
        npass=1 if unroll else NPASS, cmode={"thread": 0, "groups": 1, "mma": 2, "loop": 3}[cmode], ncp=ncp,
        colour_denom=float(den[0]),
        nb=max(len(rows) for rows in batch_rows) if batch_rows else 1, chain=1 if chain_on else 0,
        nrows=sum(len(rows) for rows in batch_rows),
        scratch=max(max((pairs[pi]["off"] + 4 * pairs[pi]["nv"] for pi in b_[0]), default=0) for b_ in batches) if batches else 0)


def _emit_colour_groups(ir, ncg, nj, nh):
    """Colour quadratic form with the JAMPs of one helicity spread over `ncg` threads (colour group g owns
    colours [g*nj, (g+1)*nj)): every unordered pair (a, b) is evaluated by exactly one thread -- the owner of
    a when both are in one group, else alternating between the two owners -- which reads the other
    group's JAMP from shared memory (jb[c * NH])."""
    ncolor = len(ir["jamp"])
    cf, den = ir["color_num"], ir["color_denom"]
    assert len(set(den)) == 1 and all(cf[i][j] == cf[j][i] for i in range(ncolor) for j in range(ncolor)), \
        "colour groups need a symmetric colour matrix with one denominator"
    L = ["    double me = 0.0;", "    switch (cg) {"]
    for g in range(ncg):
        L.append(f"    case {g}: {{")
        own = range(g * nj, min((g + 1) * nj, ncolor))
        for a in own:
            terms = []
            for b in range(ncolor):
                if b == a or cf[a][b] == 0:
                    continue
                gb = b // nj
                if gb == g:
                    if b > a:
                        terms.append((b, True))
                elif (a + b) % 2 == (0 if g < gb else 1):
                    terms.append((b, False))
            L.append(f"      {{ double tr = {float(cf[a][a])!r} * J[{a - g * nj}].re, ti = {float(cf[a][a])!r} * J[{a - g * nj}].im;")
            for b, mine in terms:
                src = f"J[{b - g * nj}]" if mine else f"jb[{b * nh}]"
                L.append(f"        {{ const cxd o = {src}; tr += {float(2 * cf[a][b])!r} * o.re; ti += {float(2 * cf[a][b])!r} * o.im; }}")
            L.append(f"        me += J[{a - g * nj}].re * tr + J[{a - g * nj}].im * ti; }}")
        L.append("    } break;")
    L.append("    default: break;")
    L.append("    }")
    L.append(f"    return me / {float(den[0])!r};")
    return "\n".join(L)


def use_hp_default(ir):
    """Default kernel flavour per process: one event per thread while the wavefunctions of one
    helicity fit in registers, helicity-parallel blocks beyond (DESIGN.md "Kernel mapping")."""
    return len(ir["calls"]) > 16


def choose_launch(ir):
    """Block size / minimum resident blocks per SM by process size (one event per thread)."""
    ncalls = len(ir["calls"])
    if ncalls <= 16:
        return 128, 3
    if ncalls <= 64:
        return 128, 2
    return 128, 1


def emit_process_source(ir, block=None, minblocks=None):
    process_ir.validate(ir)
    n = ir["nexternal"]
    ncolor = len(ir["jamp"])
    namps = len({c["amp"] for c in ir["calls"] if "amp" in c})
    b, mb = choose_launch(ir)
    block = block or b
    minblocks = minblocks or mb
    hel_flat = ", ".join(str(int(h)) for row in ir["helicities"] for h in row)
    cdefs = []
    for cname in ir["couplings"]:
        law = ir.get("coupling_defs", {}).get(cname) or COUPLING_DEFS.get(cname)
        if law is None:
            raise ValueError(f"coupling {cname}: alpha_s dependence unknown to the CUDA backend")
        cdefs.append(tuple(law))

    def switch(vals, fmt):
        body = " ".join(f"case {i}: return {fmt(v)};" for i, v in enumerate(vals))
        return f"switch (i) {{ {body} default: return {fmt(0)}; }}"

    hp_tables, hp_jamp, hp_colour, hp_unrolled, hp = emit_hp(ir)
    hp_unroll = 'true' if hp['unroll'] else 'false'
    # only emitted when chains are on, so that the default sources stay as they were measured
    hp_chain_members = ("\n  // chains of tiles accumulate in the tensor-core accumulators (hp_mma_chains); the warps take units = chains\n"
                        "  static constexpr bool HP_CHAIN = true;\n"
                        "  MF_DEV static uint2 unit(int i) { return MF_TAB(units)[i]; }") if hp['chain'] else ""
    hp_scratch_n = 0 if hp['unroll'] else hp['scratch']
    hp_nb = 0 if hp['unroll'] else hp['nb']
    hp_ncg = hp['ncg']
    hp_nj = -(-ncolor // hp_ncg)
    hp_e = hp_config(ir, "E")
    use_hp = "true" if use_hp_default(ir) else "false"
    has_thread = "true" if (len(ir["calls"]) <= THREAD_MAX_CALLS or os.environ.get("MADFLOW_B200_BUILD_THREAD") == "1") else "false"
    hp_minblocks, hp_wfsize, hp_maxlevel, hp_nwf, hp_nitems = hp_config(ir, "MINBLOCKS"), hp["wfsize"], hp["maxlevel"], hp["nwf"], hp["nitems"]
    pnames = ", ".join(f'"{p}"' for p in ir["params"]) or '""'
    cnames = ", ".join(f'"{c}"' for c in ir["couplings"]) or '""'
    src = f"""// GENERATED by madflow_b200.codegen -- do not edit.  Process: {ir.get('process', ir['name'])}
// One fused FP64 kernel per process: HELAS wavefunctions -> ALOHA vertices -> JAMP -> colour matrix.
#include "process_kernels_hp.cuh"

namespace {{
__device__ __constant__ signed char d_hel[{ir['ncomb'] * n}] = {{{hel_flat}}};
static const signed char h_hel[{ir['ncomb'] * n}] = {{{hel_flat}}};

{hp_tables}

#ifdef __CUDA_ARCH__
#define MF_TAB(name) d_##name
#else
#define MF_TAB(name) h_##name
#endif

MF_DEV int P_hel(int icomb, int leg) {{
#ifdef __CUDA_ARCH__
  return d_hel[icomb * {n} + leg];
#else
  return h_hel[icomb * {n} + leg];
#endif
}}

struct Proc {{
  static constexpr int NEXT = {n}, NINIT = {ir['ninitial']}, NCOMB = {ir['ncomb']}, NCOLOR = {ncolor};
  static constexpr int NDIAGS = {ir['ndiags']}, NAMPS = {namps}, NWF = {ir['nwavefuncs']};
  static constexpr int NPAR = {len(ir['params'])}, NCOUP = {len(ir['couplings'])};
  static constexpr int BLOCK = {block}, MINBLOCKS = {minblocks};
  static constexpr double DENOM = {float(ir['denominator'])!r};
  static constexpr double FLOPS = {float(flops_per_event(ir))!r};
  static const char* name() {{ return "{ir['name']}"; }}
  static const char* param_name(int i) {{ static const char* n[] = {{{pnames}}}; return n[i]; }}
  static const char* coupling_name(int i) {{ static const char* n[] = {{{cnames}}}; return n[i]; }}
  MF_DEV static constexpr double coup_re(int i) {{ {switch([c[0] for c in cdefs], lambda v: repr(float(v)))} }}
  MF_DEV static constexpr double coup_im(int i) {{ {switch([c[1] for c in cdefs], lambda v: repr(float(v)))} }}
  MF_DEV static constexpr int coup_power(int i) {{ {switch([c[2] for c in cdefs], lambda v: str(int(v)))} }}
  MF_DEV static int hel(int icomb, int leg) {{ return P_hel(icomb, leg); }}

  // helicity-parallel variant (process_kernels_hp.cuh)
  static constexpr bool USE_HP = {use_hp};
  // the one-event-per-thread kernels are only compiled while a helicity's wavefunctions can stay in
  // registers; beyond that they spill to DRAM (profiles/r01_ttxgg_thread_per_event.summary.txt)
  static constexpr bool HAS_THREAD = {has_thread};
  static constexpr int HP_E = {hp_e}, HP_MINBLOCKS = {hp_minblocks}, HP_WFSIZE = {hp_wfsize}, HP_MAXLEVEL = {hp_maxlevel};
  static constexpr int HP_NWF = {hp_nwf}, HP_NITEMS = {hp_nitems}, HP_NAMPS = {hp['namps']};
  static constexpr int HP_NBATCH = {hp['nbatch']}, HP_NPAIRS = {hp['npairs']}, HP_SCRATCH = {hp_scratch_n};
  // tensor-core amplitude phase: HP_NPASS helicity passes of HP_NHP combinations (the pass = the helicity of
  // the last leg); HP_NB rows of HP_NHP amplitudes per event and batch; the JAMPs of one helicity
  // combination are spread over HP_NCG threads, HP_NJ colours each
  static constexpr int HP_NPASS = {hp['npass']}, HP_NHP = NCOMB / HP_NPASS;
  static constexpr int HP_NB = {hp_nb}, HP_NCG = {hp_ncg}, HP_NJ = {hp_nj}, HP_NTILES = {hp['ntiles']};
  static constexpr int HP_THREADS = HP_E * HP_NHP * HP_NCG;                // threads per block
  // tile descriptors per warp and trip (x HP_E tiles in flight): 2 tiles in flight measured best -- more only adds
  // padded tiles at the end of a batch (g g > t t~ g g: 17.7e6 events/s with 2, 16.5e6 with 4, 14.5e6 with 8)
  static constexpr int HP_TILES_IN_FLIGHT = {max(1, int(os.environ.get("MADFLOW_B200_HP_MT", 2)) // hp_e)};{hp_chain_members}
  // colour contraction: 0 in-thread, 1 generated code over colour groups, 2 tensor cores, 3 CUDA-core loop
  // (2 and 3 read the block-symmetrised matrix d_cfsym and JAMP planes of HP_PLANE doubles per colour)
  static constexpr int HP_COLOUR = {hp['cmode']}, HP_NCP = {hp['ncp']}, HP_PLANE = HP_NHP + 4;
  static constexpr double HP_COLOUR_DENOM = {hp['colour_denom']!r};
  // shared-memory cxd per event: wavefunctions | pair objects | amplitude buffer (the last two double as the
  // JAMP exchange area of the colour groups)
  static constexpr int HP_XCHG = HP_COLOUR >= 2 ? HP_NCP * HP_PLANE : (HP_NCG > 1 ? NCOLOR * HP_NHP : 0);
  static constexpr int HP_EVSIZE = HP_WFSIZE + (HP_SCRATCH + HP_NB * HP_NHP > HP_XCHG ? HP_SCRATCH + HP_NB * HP_NHP : HP_XCHG);
  // the E events of a block sit HP_EVSTRIDE apart; the stride is padded to 8/E (mod 8) elements of 16 bytes so
  // that threads working on the same object of different events hit different banks
  static constexpr int HP_EVSTRIDE = HP_E > 1 ? HP_EVSIZE + ((8 / HP_E) - HP_EVSIZE % 8 + 8) % 8 : HP_EVSIZE;
  MF_DEV static const double* cfsym() {{ return MF_TAB(cfsym); }}
  MF_DEV static mf::HpWf wf(int w) {{ return MF_TAB(wf)[w]; }}
  MF_DEV static mf::HpExt ext(int leg) {{ return MF_TAB(ext)[leg]; }}
  MF_DEV static mf::HpItem item(int i) {{ return mf::hp_fetch32(&MF_TAB(items)[i]); }}
  MF_DEV static int level_begin(int L) {{ return MF_TAB(level_begin)[L]; }}
  MF_DEV static const mf::HpTile* tile(int i) {{ return &MF_TAB(tiles)[i]; }}
  MF_DEV static mf::HpPair pair(int i) {{ return mf::hp_fetch32(&MF_TAB(pairs)[i]); }}
  MF_DEV static mf::HpPairItem pair_item(int i) {{ return MF_TAB(pair_items)[i]; }}  // one 8-byte load
  MF_DEV static mf::HpPairItem cur_item(int i) {{ return MF_TAB(cur_items)[i]; }}
  MF_DEV static mf::HpBatch batch(int i) {{ return MF_TAB(batches)[i]; }}
  // JAMP updates of batch `b` for colour group `cg` (warp-uniform switches): ab = the event's amplitude
  // buffer at this thread's helicity combination, row r at ab[r * HP_NHP]; JAMP registers addressed statically
  MF_DEV static void jamp_batch(int b, int cg, const cxd* ab, cxd (&J)[HP_NJ]) {{
{hp_jamp}
  }}
  // colour quadratic form (HP_COLOUR 0 / 1); the other groups' JAMPs are read from jb[colour * HP_NHP]
  MF_DEV static double colour_sum(int cg, const cxd (&J)[HP_NJ], const cxd* jb) {{
{hp_colour}
  }}
  static constexpr bool HP_UNROLL = {hp_unroll};
  MF_DEV static double hp_amps_unrolled(const cxd* wf_e, const unsigned char* vtab, int h, const cxd* coup);

  // Matrix_{_cname(ir)}.matrix for helicity row `icomb`
  MF_DEV static double matrix(const double (*p)[4], int icomb, const double* par, const cxd* coup, double sqh) {{
{emit_matrix_body(ir) if len(ir["calls"]) <= STRAIGHT_LINE_MAX_CALLS else "    return 0.0 / 0.0;  // not emitted: the straight-line form of this process is too long to be useful"}
  }}
}};

MF_DEV double Proc::hp_amps_unrolled(const cxd* wf_e, const unsigned char* vtab, int h, const cxd* coup) {{
{hp_unrolled}
}}
}}  // namespace

MF_DEFINE_PROCESS(Proc)
"""
    return src


def compile_source(src_path, out, verbose=False, extra_flags=()):
    """nvcc one generated process source into a shared library (sm_100a)."""
    cmd = ["nvcc"] + NVCC_FLAGS + list(extra_flags) + ["-I", CSRC, "-o", out, src_path]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src_path}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stderr)
    return out


def lib_path(ir_or_name):
    name = ir_or_name if isinstance(ir_or_name, str) else ir_or_name["name"]
    return os.path.join(LIBDIR, f"libmfp_{name}.so")


def build_process(ir, out=None, verbose=False, extra_flags=()):
    """Emit and compile one process.  Returns the path of the shared library."""
    os.makedirs(GENDIR, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    src_path = os.path.join(GENDIR, f"proc_{ir['name']}.cu")
    text = emit_process_source(ir)
    old = open(src_path).read() if os.path.exists(src_path) else None
    out = out or lib_path(ir)
    if old == text and os.path.exists(out) and os.path.getmtime(out) >= _newest_header_mtime():
        return out
    with open(src_path, "w") as fh:
        fh.write(text)
    with open(os.path.join(GENDIR, f"proc_{ir['name']}.json"), "w") as fh:
        fh.write(process_ir.dumps(ir))
    return compile_source(src_path, out, verbose, extra_flags)


def _newest_header_mtime():
    t = 0.0
    for f in os.listdir(CSRC):
        if f.endswith((".cuh", ".h", ".cu")):
            t = max(t, os.path.getmtime(os.path.join(CSRC, f)))
    t = max(t, os.path.getmtime(os.path.join(os.path.dirname(CSRC), "..", "include", "madflow_b200_process.h")))
    return t
