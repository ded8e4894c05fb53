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


# hand-written device routines (csrc/helas.cuh, csrc/aloha_sm.cuh); every other vertex of a call list must come with its
# ALOHA routine, written by the pyout plugin's CUDA ALOHA writer (madgraph_plugin/PyOut_create_aloha.py), in
# ir["aloha_routines"] = {routine name: device source text}
BUILTIN_OPS = set(FLOPS)
PLUGIN_ROUTINE_FLOPS = 300   # nominal operation count of a plugin-written routine (F_alg bookkeeping only)


def plugin_ops(ir):
    """Vertex routines of the call list that are not hand-written: they are taken from ir["aloha_routines"]."""
    ops = sorted({c["op"] for c in ir["calls"]} - BUILTIN_OPS)
    missing = [op for op in ops if op not in ir.get("aloha_routines", {})]
    if missing:
        raise ValueError(f"vertex routines {missing} are neither built in nor supplied in ir['aloha_routines'] "
                         "(the pyout plugin's ALOHA writer produces them from the UFO model)")
    return ops


def hp_available(ir):
    """The helicity-parallel kernels evaluate the vertices of the QCD sector through table-driven numerators
    (HP_TYPES); call lists with other Lorentz structures run in the one-event-per-thread flavour."""
    return not plugin_ops(ir)


def _emit_plugin_routines(ir):
    """The ALOHA routines supplied with the IR + adapters to the complex type of the kernels (cxd)."""
    ops = plugin_ops(ir)
    if not ops:
        return ""
    L = ["namespace plg {", "typedef cuda::std::complex<double> cxtype;"]
    for op in ops:
        L.append(ir["aloha_routines"][op].strip())
    L.append("}  // namespace plg")
    for op in ops:
        call = next(c for c in ir["calls"] if c["op"] == op)
        nin = len(call["in"])
        names = [f"W{q}" for q in range(nin)]
        conv = " ".join(f"plg::cxtype x{q}[6]; for (int k = 0; k < 6; ++k) x{q}[k] = plg::cxtype({n}[k].re, {n}[k].im);"
                        for q, n in enumerate(names))
        xs = ", ".join(f"x{q}" for q in range(nin))
        sig = ", ".join(f"const cxd* {n}" for n in names)
        if "amp" in call:
            L.append(f"MF_DEV cxd plg_{op}({sig}, cxd COUP) {{ {conv} plg::cxtype v(0., 0.); "
                     f"plg::{op}({xs}, plg::cxtype(COUP.re, COUP.im), v); return mk(v.real(), v.imag()); }}")
        else:
            L.append(f"MF_DEV void plg_{op}({sig}, cxd COUP, double M, double W, cxd* OUT) {{ {conv} plg::cxtype o[6]; "
                     f"plg::{op}({xs}, plg::cxtype(COUP.re, COUP.im), M, W, o); "
                     "for (int k = 0; k < 6; ++k) OUT[k] = mk(o[k].real(), o[k].imag()); }")
    return "\n".join(L)


def flops_per_event(ir):
    """F_alg of SURVEY.md section 8(d): ncomb*(sum calls + F_jamp + F_colour) + ncomb + 1."""
    per_hel = sum(FLOPS.get(c["op"], PLUGIN_ROUTINE_FLOPS) for c in ir["calls"])
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
    fn = "mf::" + op
    if op not in BUILTIN_OPS:
        fn = "plg_" + op       # written by the plugin's ALOHA writer (_emit_plugin_routines)
    elif op.startswith("VVVV"):
        kind = op[4]
        fn = f"mf::VVVV_0<{kind}>" if op.endswith("_0") else f"mf::VVVVP0_1<{kind}>"
    if "amp" in c:
        return f"{fn}({ins}, {_coup_expr(ir, c)})"
    return (f"{fn}({ins}, {_coup_expr(ir, c)}, {_par_expr(ir, c['mass'])}, {_par_expr(ir, c['width'])}, "
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


def hp_config(ir, key):
    """Launch shape of the helicity-parallel kernels, MADFLOW_B200_HP_<KEY> overrides (tools/build_variants.py):
      E          events per block
      NCG        colour groups = threads per (event, helicity combination); each keeps NCOLOR/NCG JAMPs in registers
      NB         rows of the amplitude buffer = amplitudes per batch
      SCRATCH    shared-memory scratch for the pair objects of a batch, complex numbers per event
      MINBLOCKS  resident blocks per SM the register allocation aims at
      PERSIST    pair objects (vertex numerators) of up to this many helicity variants are evaluated once, together
                 with the currents of their level, and stay in shared memory; the others are evaluated batch by batch
      PERSIST_FREE   with helicity passes: the pair objects that do not hold the pass leg are the same in every pass; up to
                 this many complex numbers of them per event are evaluated once before the passes (a phase of its own
                 after the last level of currents) and kept in shared memory
      TSPLIT     objects with more terms than this are evaluated by 2, 4 or 8 neighbouring lanes (a few terms each, summed
                 with warp shuffles in a fixed order); 0 = one thread per (object, variant) whatever its length.  With few
                 units per phase and up to 12 terms per object (g g > t t~ g g g) the longest unit sets the phase's time
      TMEMJ      1 = the JAMP accumulators live in Tensor Memory between the JAMP phases of the batches instead of in
                 registers (frees NCOLOR/NCG complex registers for the current / pair / tile phases)
      MT         tiles of the amplitude phase in flight per warp (x events per block)
      PREFIN     1 = the momenta / couplings of the next group of events are fetched into registers while the current
                 group is evaluated (smatrix_kernel_hp)
      SLU        1 = packed units: the units of a phase sorted into warp trips of one class of objects (sequence of vertex
                 kinds + propagator kind), one descriptor per trip and one 64-bit word per (unit, term) left to read at run
                 time (process_kernels_hp.cuh, "SLU"); needs TSPLIT = 0
    Defaults from measurements on B200: up to 64 helicity combinations two events per
    block and all JAMPs in one thread (Tensor Memory off: with two blocks per SM it halves the rate, 2.4e7 -> 1.4e7
    events/s for g g > t t~ g g, profiles/r02j_ttxgg_tmem.log); beyond, one event per block, 8 colour groups, batches
    of 64 amplitudes, JAMPs in Tensor Memory."""
    env = os.environ.get("MADFLOW_B200_HP_" + key)
    if env:
        return int(env)
    if ir["ncomb"] > 64:
        # g g > t t~ g g g, measured (profiles/r02k_ttxggg_tmem_tuning.log): JAMPs parked in Tensor Memory 7.9e5 -> 9.4e5
        # events/s (no more spills at 128 registers), + pass-independent pair objects kept 9.9e5; 4 colour groups 6.3e5
        return {"E": 1, "NCG": 8, "NB": 64, "SCRATCH": 2048 if hp_use_plan(ir) else 4096, "MINBLOCKS": 1, "PERSIST": 0,
                "PERSIST_FREE": 3500 if hp_use_plan(ir) else 0, "TSPLIT": 0 if hp_use_plan(ir) else 3, "TMEMJ": 1,
                "SLU": 1 if hp_use_plan(ir) else 0, "MT": 1 if hp_use_plan(ir) else 2, "PREFIN": 1}[key]
    # NB: 22 rows let the 64 rows of the reduced g g > t t~ g g fit in 3 batches (with two blocks per SM still resident)
    # MINBLOCKS: g g > t t~ g (32 combinations, straight-line amplitudes) runs 4.6 % faster with three blocks per SM at 168
    # registers, spills included (1.41e8 -> 1.48e8 events/s; four blocks: 1.03e8), profiles/r02zt_*, r02zu_*
    # the four-quark six-point processes (<= 40 calls, 6 colour flows): four blocks per SM at 128 registers, +22...24 %
    # (q q~ > t t~ q q~ 6.6e7 -> 8.0e7, q q' > t t~ q q' 1.13e8 -> 1.39e8); the 73-call light-line processes stay at two
    # (three: -6 %), profiles/r02zv_*
    small6 = ir["ncomb"] == 64 and not hp_use_plan(ir) and len(ir["calls"]) <= 40
    return {"E": max(1, 128 // ir["ncomb"]), "NCG": 1, "NB": 22 if hp_use_plan(ir) else 21, "SCRATCH": 512,
            "MINBLOCKS": 3 if (ir["ncomb"] == 32 and len(ir["calls"]) > 24) else (4 if small6 else 2),
            "PERSIST": 8 if hp_use_plan(ir) else 0, "PERSIST_FREE": 0, "TSPLIT": 0, "TMEMJ": 0,
            "SLU": 1 if (hp_use_plan(ir) and ir["ncomb"] == 64) else 0, "MT": 2, "PREFIN": 0}[key]


def hp_use_plan(ir):
    """Evaluate the colour-reduced plan attached to the IR (madflow_b200/recursion.py) instead of the diagram list
    (MADFLOW_B200_HP_REDUCE=0: the diagram list, the A/B partner)."""
    return "plan" in ir and os.environ.get("MADFLOW_B200_HP_REDUCE", "1") != "0"


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


# Vertex structures of the helicity-parallel kernels.  Every vertex is evaluated as a NUMERATOR in dual form: with
# one of its lines left open (the output of a current, or the input `x` of a closing vertex),
#   Q[k] = -i COUP (vertex contracted with the other lines)_k ,  metric signs folded in, so that  amp = sum_k x[2+k] Q[k]
# and a current is the same numerator times its propagator (csrc/process_kernels_hp.cuh::hp_unit):
#   ROW  (O, G):    Obar Gslash            <- FFV1_1(O, G)   | FFV1_0(x, O, G)
#   COL  (I, G):    Gslash I               <- FFV1_2(I, G)   | FFV1_0(I, x, G)
#   CUR  (I, O):    Obar gamma^mu I        <- FFV1P0_3(I, O) | FFV1_0(I, O, x)
#   VVV  (V2, V3):  three-gluon vertex     <- VVV1P0_1       | VVV1_0 with the other two lines in cyclic order
#   VVVV (3 lines): contact term           <- VVVVkP0_1      | VVVVk_0
PT = {"ROW": 0, "COL": 1, "CUR": 2, "VVV": 3, "VVVV": 4}
FINISH = {"none": 0, "g": 1, "o": 2, "i": 3}
PHASE_CODE = {1: 0, -1: 1, 1j: 2, -1j: 3}
# UFO four-gluon structures: VVVVk = sum of sign * g(pa) g(pb) over vertex positions 1..4
QUARTIC = {"1": ((+1, (1, 4), (2, 3)), (-1, (1, 3), (2, 4))),
           "3": ((+1, (1, 4), (2, 3)), (-1, (1, 2), (3, 4))),
           "4": ((+1, (1, 3), (2, 4)), (-1, (1, 2), (3, 4)))}


def _vertex_structure(op, jx, ins):
    """(structure, inputs in the structure's order, four-gluon term codes) of vertex `op` with the line at argument
    position jx left open (jx = None: `op` is a current routine and its output is the open line)."""
    if op == "FFV1_1":
        return "ROW", tuple(ins), (0, 0)
    if op == "FFV1_2":
        return "COL", tuple(ins), (0, 0)
    if op == "FFV1P0_3":
        return "CUR", tuple(ins), (0, 0)
    if op == "VVV1P0_1":
        return "VVV", tuple(ins), (0, 0)
    if op == "FFV1_0":
        full = list(ins)
        full.insert(jx, None)
        I, O, G = full
        return [("ROW", (O, G)), ("COL", (I, G)), ("CUR", (I, O))][jx] + ((0, 0),)
    if op == "VVV1_0":
        full = list(ins)
        full.insert(jx, None)
        return "VVV", (full[(jx + 1) % 3], full[(jx + 2) % 3]), (0, 0)
    assert op.startswith("VVVV"), op
    kind = op[4]
    if jx is None:          # VVVVkP0_1(V2, V3, V4): the output is vertex position 1
        xpos, others = 1, [2, 3, 4]
    else:
        xpos, others = jx + 1, [q for q in (1, 2, 3, 4) if q != jx + 1]
    pos = {q: others.index(q) for q in others}   # vertex position -> index into the inputs
    enc = []
    for sign, pa, pb in QUARTIC[kind]:
        if xpos in pa:
            vec, dot = [q for q in pa if q != xpos][0], pb
        else:
            vec, dot = [q for q in pb if q != xpos][0], pa
        enc.append((0x40 if sign < 0 else 0) | pos[vec] << 4 | pos[dot[0]] << 2 | pos[dot[1]])
    return "VVVV", tuple(ins), tuple(enc)


def hp_plan_from_ir(ir):
    """The diagram list of the IR as an evaluation plan (the format of recursion.build_plan): undo the slot reuse of
    the call list -- every write creates a distinct wavefunction -- one single-term object per current, one
    single-term pair object per distinct (closing vertex, inputs but the heaviest), one row per amplitude."""
    cur = {}      # slot -> object id
    objects, amps = [], []
    for c in ir["calls"]:
        if "leg" in c:
            cur[c["out"]] = len(objects)
            objects.append({"legs": [c["leg"]], "ext": c, "terms": []})
        elif "amp" in c:
            amps.append((c, [cur[s] for s in c["in"]]))
        else:
            ins = [cur[s] for s in c["in"]]
            legs = sorted(set().union(*[objects[i]["legs"] for i in ins]))
            assert len(legs) == sum(len(objects[i]["legs"]) for i in ins), "children must not share legs"
            cur[c["out"]] = len(objects)
            sign = -1 if c.get("coup_sign", 1) < 0 else 1
            objects.append({"legs": legs, "ext": None, "mass": c["mass"], "width": c["width"],
                            "terms": [{"op": c["op"], "in": ins, "coef": [sign, 0], "coup": c["coup"]}]})
    by_amp = {}
    for j, terms in enumerate(ir["jamp"]):
        for k, re, im in terms:
            by_amp.setdefault(k, []).append([j, float(re), float(im)])
    LSTAR = ir["nexternal"] - 1 if hp_passes(ir) > 1 else None
    pairs, pair_index, rows = [], {}, []
    for c, ins in amps:
        if not by_amp.get(c["amp"]):
            continue
        # x = the input with the most legs (with helicity passes: preferably one that does not hold the pass leg,
        # so that the pass halves the rows and not the columns of the amplitude tiles)
        sizes = [len(objects[w]["legs"]) for w in ins]
        cands = [q for q in range(len(ins)) if sizes[q] == max(sizes)]
        free = [q for q in cands if LSTAR not in objects[ins[q]]["legs"]]
        jx = (free or cands)[0]
        rest = [w for q, w in enumerate(ins) if q != jx]
        sign = -1 if c.get("coup_sign", 1) < 0 else 1
        key = (c["op"], jx, tuple(rest), c["coup"], sign)
        if key not in pair_index:
            pair_index[key] = len(pairs)
            legs = sorted(set().union(*[objects[w]["legs"] for w in rest]))
            pairs.append({"legs": legs, "terms": [{"op": c["op"], "jx": jx, "in": rest, "coef": [sign, 0], "coup": c["coup"]}]})
        rows.append({"x": ins[jx], "pair": pair_index[key], "jamp": by_amp[c["amp"]], "amp": c["amp"]})
    return {"objects": objects, "pairs": pairs, "rows": rows}


def emit_hp(ir):
    """Tables + the generated amplitude/JAMP/colour code of the helicity-parallel kernels."""
    reduced = hp_use_plan(ir)
    if not hp_available(ir):
        # vertices outside the table-driven set: only the externals are described, the flavour is switched off (HP_AVAILABLE)
        plan = {"objects": [{"legs": [c["leg"]], "ext": c, "terms": []} for c in ir["calls"] if "leg" in c], "pairs": [], "rows": []}
        reduced = False
    else:
        plan = ir["plan"] if reduced else hp_plan_from_ir(ir)
    n = ir["nexternal"]
    assert ir["ncomb"] == 2**n, "the hp kernels need the full 2^n helicity table"
    NH = ir["ncomb"]
    NPASS = hp_passes(ir)
    assert NPASS in (1, 2)
    NHP = NH // NPASS
    LSTAR = n - 1 if NPASS > 1 else None   # the leg whose helicity is fixed within a pass (top variant bit)
    big = len(plan["rows"]) > 400          # tables beyond the 64 KB of constant memory live in global memory
    scratch = hp_config(ir, "SCRATCH")
    NB = hp_config(ir, "NB")
    NCG = hp_config(ir, "NCG")
    PERSIST = hp_config(ir, "PERSIST")

    def pidx(name):
        return -1 if name == "ZERO" else ir["params"].index(name)

    def both(ctype, name, count, body, const=True):
        space = "__device__ __constant__" if const else "__device__ const"
        return (f"{space} {ctype} d_{name}[{count}] = {{{body}}};\n"
                f"static const {ctype} h_{name}[{count}] = {{{body}}};")

    def vmask(out_legs, in_legs):
        """bits of the output's variant index that make up the input's variant index (both ascending)"""
        return sum(1 << q for q, l in enumerate(out_legs) if l in in_legs)

    def pext(v, m):
        return sum(((v >> b_) & 1) << q for q, b_ in enumerate(b2 for b2 in range(8) if m >> b2 & 1))

    def phase_of(t):
        ph = complex(*t["coef"])
        assert ph in PHASE_CODE, f"coefficient {ph} of a plan term is not a unit"
        return PHASE_CODE[ph]

    # ---- wavefunctions (externals + currents) and their layout in the event area
    wfs, exts = [], []
    off = 0
    # currents that no other object is built from (they only close amplitudes) need no momentum slots
    feeds = {i for o in plan["objects"] for t in o["terms"] for i in t["in"]} | {i for p in plan["pairs"] for t in p["terms"] for i in t["in"]}
    for oi, o in enumerate(plan["objects"]):
        legs = tuple(sorted(o["legs"]))
        w = {"legs": legs, "level": len(legs), "nv": 1 << len(legs), "off": off, "mask": sum(1 << l for l in legs),
             "ext": o["ext"], "terms": o["terms"], "mass": o.get("mass", "ZERO"), "width": o.get("width", "ZERO")}
        w["mom"] = 2 if (o["ext"] is not None or oi in feeds or not reduced) else 0
        off += w["mom"] + 4 * w["nv"]
        if o["ext"] is not None:
            exts.append({"call": o["ext"], "out": len(wfs)})
            w["finish"] = "none"
        else:
            kinds = {"FFV1_1": "o", "FFV1_2": "i"}
            w["finish"] = kinds.get(o["terms"][0]["op"], "g")
            assert all(kinds.get(t["op"], "g") == w["finish"] for t in o["terms"])
        wfs.append(w)
    maxlevel = max(w["level"] for w in wfs)
    # ---- pair objects; the small ones (<= PERSIST variants, inputs ready) join the currents' phases and stay
    pairs = []
    for p in plan["pairs"]:
        legs = tuple(sorted(p["legs"]))
        pr = {"legs": legs, "nv": 1 << len(legs), "terms": p["terms"], "finish": "none"}
        pr["ready"] = 1 + max(wfs[w]["level"] for t in p["terms"] for w in t["in"])
        pr["persist"] = pr["nv"] <= PERSIST and pr["ready"] <= maxlevel
        pairs.append(pr)
    # PERSIST_FREE = shared-memory budget (complex numbers per event) for pass-independent pair objects: those with the
    # most terms first (the evaluations saved per complex number stored)
    budget = hp_config(ir, "PERSIST_FREE") if LSTAR is not None else 0
    for pr in sorted(pairs, key=lambda q: -len(q["terms"])):
        if not pr["persist"] and LSTAR not in pr["legs"] and 4 * pr["nv"] <= budget:
            pr["persist"] = True
            budget -= 4 * pr["nv"]
    for pr in pairs:
        if pr["persist"]:
            pr["abs_off"] = off
            off += 4 * pr["nv"]
    assert all(pr["nv"] <= 32 for pr in pairs), "pair objects hold up to 32 helicity variants"
    wfsize = off

    by_amp = {k: r["jamp"] for k, r in enumerate(plan["rows"])}   # row index -> [(colour, re, im)]
    amp_rows = [dict(x=r["x"], pair=r["pair"], row=k) for k, r in enumerate(plan["rows"])]

    def lowered(t):
        """a plan term as a vertex structure: (structure, inputs, four-gluon codes)"""
        return _vertex_structure(t["op"], t.get("jx"), t["in"])

    def type_key(obj):
        return tuple(PT[lowered(t)[0]] for t in obj["terms"])

    # ---- batches: a batch closes at a pair boundary when the scratch area or the amplitude buffer (HP_NB rows)
    # would overflow; pair objects of one kind together, so that the warps of a batch run the same routines
    by_pair = {}
    for k, r in enumerate(amp_rows):
        by_pair.setdefault(r["pair"], []).append(k)
    sig_id = {}
    for r in amp_rows:   # colour signature of a row: its JAMP coefficients up to a common phase
        t = by_amp[r["row"]]
        c0 = complex(t[0][1], t[0][2])
        r["sig"] = sig_id.setdefault(tuple((j, complex(re, im) / c0) for j, re, im in sorted(map(tuple, t))), len(sig_id))
    batches, cur_pairs, cur_amps, fill = [], [], [], 0
    transient = [pi for pi in by_pair if not pairs[pi]["persist"]]
    for pi in sorted(transient, key=lambda q: (type_key(pairs[q]), pairs[q]["nv"], min(amp_rows[k]["sig"] for k in by_pair[q]), q)):
        need = 4 * pairs[pi]["nv"]
        assert need <= scratch and len(by_pair[pi]) <= NB
        if fill + need > scratch or len(cur_amps) + len(by_pair[pi]) > NB:
            batches.append((cur_pairs, cur_amps))
            cur_pairs, cur_amps, fill = [], [], 0
        pairs[pi]["abs_off"] = wfsize + fill
        fill += need
        cur_pairs.append(pi)
        cur_amps += by_pair[pi]
    if cur_amps:
        batches.append((cur_pairs, cur_amps))
    # rows over persistent pair objects need no pair phase: they fill the free rows of the batches (fewer batches,
    # fewer block barriers), the rest forms batches of its own
    resident = [k for pi in sorted(by_pair) if pairs[pi]["persist"] for k in by_pair[pi]]
    resident.sort(key=lambda k: (amp_rows[k]["x"], amp_rows[k]["pair"]))
    nfull = -(-(len(resident) + sum(len(a) for _, a in batches)) // NB)       # batches needed for all rows
    while len(batches) < nfull:
        batches.append(([], []))
    share = -(-(len(resident) + sum(len(a) for _, a in batches)) // max(len(batches), 1))   # even filling
    for bp, ba in batches:
        while resident and len(ba) < min(NB, share):
            ba.append(resident.pop(0))
    for bp, ba in batches:
        while resident and len(ba) < NB:
            ba.append(resident.pop(0))
    assert not resident
    batches = [b_ for b_ in batches if b_[1]]

    def spread(legs, v):
        """helicity-combination bits (within the pass) of variant v of an object over `legs` (ascending)"""
        return sum(((v >> q) & 1) << l for q, l in enumerate(legs) if l != LSTAR)

    def vrange(legs, nv, p):
        """variants of an object that pass p needs: the half with the pass leg's helicity = p, or all"""
        if LSTAR is not None and LSTAR in legs:
            return range(p * nv // 2, (p + 1) * nv // 2)   # the pass leg is the highest leg = top variant bit
        return range(nv)

    # ---- terms, work items, units.  A unit = one (object, helicity variant): its terms are evaluated by one thread
    # and added up; the last one applies the propagator (currents) and stores.
    trows, irow, urow = [], [], []

    def term_row(obj, t, out_off):
        st, ins, q = lowered(t)
        ins3 = list(ins) + [0] * (3 - len(ins))
        vm = [vmask(obj["legs"], wfs[w]["legs"]) for w in ins] + [0] * (3 - len(ins))
        coup = ir["couplings"].index(t["coup"])
        # finish + 4: a current without momentum slots (out_off then points two elements before its components)
        fin = FINISH[obj["finish"]] + (4 if obj.get("mom", 2) == 0 and obj["finish"] != "none" else 0)
        out = out_off - 2 if fin >= 4 else out_off
        return (f"{{{PT[st]}, {len(ins)}, {coup}, {phase_of(t)}, {{{q[0]}, {q[1]}}}, {obj['nv']}, {out}, "
                f"{{{', '.join(str(wfs[w]['off']) for w in ins3)}}}, {{{', '.join(str(wfs[w]['nv']) for w in ins3)}}}, "
                f"{{{vm[0]}, {vm[1]}, {vm[2]}}}, {fin}, {pidx(obj.get('mass', 'ZERO'))}, {pidx(obj.get('width', 'ZERO'))}}}")

    TSPLIT = hp_config(ir, "TSPLIT")

    def add_units(obj, out_off, variants, phase_units):
        """Append the units of `obj` for `variants` to phase_units as groups [(first item, count), ...]: a unit with more
        than TSPLIT terms is split into 2^k parts that neighbouring lanes evaluate and add up with warp shuffles."""
        first = len(trows)
        masks = []
        for t in obj["terms"]:
            st, ins, q = lowered(t)
            masks.append([vmask(obj["legs"], wfs[w]["legs"]) for w in ins] + [0] * (3 - len(ins)))
            trows.append(term_row(obj, t, out_off))
        nt = len(masks)
        parts = 1
        while TSPLIT and parts < 8 and -(-nt // parts) > TSPLIT:
            parts *= 2
        per = -(-nt // parts)
        for v in variants:
            group = []
            for pi in range(parts):
                mine = list(range(pi * per, min((pi + 1) * per, nt)))
                group.append((len(irow), len(mine)))
                for ti in mine:
                    m = masks[ti]
                    irow.append(f"{{{first + ti}, {v}, {{{pext(v, m[0])}, {pext(v, m[1])}, {pext(v, m[2])}}}, 0}}")
            phase_units.append(group)

    def flush_units(phase_units):
        """Lay the groups of one phase out: largest groups first, so that every group of 2^k parts starts at a multiple of
        2^k (its lanes sit in one warp, the phase starts at lane 0).  Unit = first item | items << 24 | log2(parts) << 28."""
        for group in sorted(phase_units, key=lambda g: -len(g)):     # stable: keeps the type order within a size
            glog = len(group).bit_length() - 1
            for item0, cnt in group:
                assert cnt < 16 and item0 < (1 << 24)
                urow.append(f"{item0 | cnt << 24 | glog << 28}u")

    begins = []
    slu_phases = []   # per phase of the kernel, in its order: [(object, offset of its block, variants)]
    maxlevel = max([maxlevel] + [pr["ready"] for pr in pairs if pr["persist"]])   # phases = levels of currents (+ one for pairs)
    for lev in range(0, maxlevel + 2):
        begins.append(len(urow))
        todo = [(w, w["off"]) for w in wfs if w["ext"] is None and w["level"] == lev]
        todo += [(pr, pr["abs_off"]) for pr in pairs if pr["persist"] and pr["ready"] == lev]
        if 2 <= lev <= maxlevel:
            slu_phases.append([(obj, o_, range(obj["nv"])) for obj, o_ in todo])
        phase_units = []
        for obj, o_ in sorted(todo, key=lambda q: (len(q[0]["terms"]), type_key(q[0]), FINISH[q[0]["finish"]])):
            add_units(obj, o_, range(obj["nv"]), phase_units)
        flush_units(phase_units)

    # ---- tensor-core tiles: amplitude(variant of Q, variant of x) = sum_k Q_k x_k is an (nvq x 4)(4 x nvx)
    # complex product; one work item = 8 variants of Q (rows) x 8 variants of x (columns)
    tile_rows, brow = [], []
    ncolor = len(ir["jamp"])
    NJ = -(-ncolor // NCG)
    for p in range(NPASS):
        for bi, (cur_pairs, cur_amps) in enumerate(batches):
            ub, tb = len(urow), len(tile_rows)
            phase_units = []
            slu_phases.append([(pairs[pi], pairs[pi]["abs_off"], vrange(pairs[pi]["legs"], pairs[pi]["nv"], p)) for pi in cur_pairs])
            for pi in cur_pairs:
                pr = pairs[pi]
                add_units(pr, pr["abs_off"], vrange(pr["legs"], pr["nv"], p), phase_units)
            flush_units(phase_units)
            for slot, k in enumerate(cur_amps):
                r = amp_rows[k]
                xw, pr = wfs[r["x"]], pairs[r["pair"]]
                assert not set(xw["legs"]) & set(pr["legs"]) and len(xw["legs"]) + len(pr["legs"]) == n
                qr, xr = vrange(pr["legs"], pr["nv"], p), vrange(xw["legs"], xw["nv"], p)
                for q0 in range(qr.start, qr.stop, 8):
                    for x0 in range(xr.start, xr.stop, 8):
                        qv, xv = min(8, qr.stop - q0), min(8, xr.stop - x0)
                        rowh = [spread(pr["legs"], q0 + i) if i < qv else 0 for i in range(8)]
                        colh = [spread(xw["legs"], x0 + i) if i < xv else 0 for i in range(8)]
                        tile_rows.append(f"{{{pr['abs_off']}, {xw['off'] + xw['mom']}, {pr['nv']}, {xw['nv']}, {q0}, {x0}, {qv}, {xv}, {slot}, 0, "
                                         f"{{{', '.join(map(str, rowh))}}}, {{{', '.join(map(str, colh))}}}}}")
            brow.append(f"{{{ub}, {len(urow)}, {tb}, {len(tile_rows)}}}")

    # ---- JAMP code per (batch, colour group).  Rows of a batch that feed the same colours of the group with the
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
            for slot, k in enumerate(cur_amps):
                terms = [(j - cg * NJ, complex(re, im)) for j, re, im in by_amp[k] if j // NJ == cg]
                if not terms:
                    continue
                unit = all(c in (1, -1, 1j, -1j) for _, c in terms)
                c0 = terms[0][1] if unit else 1.0
                sig = tuple((jl, c / c0) for jl, c in sorted(terms, key=lambda q: q[0])) if unit else ("own", slot)
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

    # ---- class-specialised straight-line units (process_kernels_hp.cuh "SLU")
    SLU = bool(hp_config(ir, "SLU")) and hp_available(ir)
    slu_tables, slu_stats = "", {}
    if SLU:
        assert not TSPLIT, "packed units split long units themselves (TSPLIT belongs to the table-driven routine)"
        HP_E = hp_config(ir, "E")
        LPU, NWARP = 32 // HP_E, HP_E * NHP * NCG // 32
        assert 32 % HP_E == 0 and NWARP >= 1 and len(ir["couplings"]) <= 7
        KCOST = {"ROW": 10, "COL": 10, "CUR": 10, "VVV": 14, "VVVV": 18}
        FCOST = {"none": 1, "g": 6, "o": 10, "i": 10}

        def slu_terms(obj):
            """The terms of an object for the straight-line units: four-gluon terms over the same three lines with the
            same coupling are merged into ONE term with integer weights (ca, cb, cc) of a (b.c), b (a.c), c (a.b); terms
            sorted by vertex kind so that objects of the same make-up share a class."""
            out, first = [], {}
            for t in obj["terms"]:
                st, ins, q = lowered(t)
                coup, ph = ir["couplings"].index(t["coup"]), complex(*t["coef"])
                assert ph in PHASE_CODE
                if st != "VVVV":
                    out.append({"kind": st, "ins": ins, "coup": coup, "ph": ph, "coefs": None})
                    continue
                coefs = [0, 0, 0]
                for code in q:
                    vec, da, db = (code >> 4) & 3, (code >> 2) & 3, code & 3
                    assert {vec, da, db} == {0, 1, 2}
                    coefs[vec] += -1 if code & 0x40 else 1
                key = (tuple(ins), coup)
                if key in first and ph / out[first[key]]["ph"] in (1, -1):
                    o = out[first[key]]
                    sgn = int((ph / o["ph"]).real)
                    o["coefs"] = [a_ + sgn * b_ for a_, b_ in zip(o["coefs"], coefs)]
                    continue
                first.setdefault(key, len(out))
                out.append({"kind": "VVVV", "ins": ins, "coup": coup, "ph": ph, "coefs": coefs})
            out = [t for t in out if t["coefs"] is None or any(t["coefs"])]
            assert out and all(t["coefs"] is None or max(map(abs, t["coefs"])) <= 2 for t in out)
            return sorted(out, key=lambda t: PT[t["kind"]])

        def desc(off, v, nlegs):
            assert 0 <= off < (1 << 14) and 0 <= v < 32 and nlegs <= 5
            return off | v << 14 | nlegs << 19

        NULL_F = 31          # f-index of a null term: the entry of the f-table that is always zero
        SPLIT_MIN = int(os.environ.get("MADFLOW_B200_HP_SLU_SPLIT_MIN", 24))   # never split units cheaper than this
        SPLIT_MAX = int(os.environ.get("MADFLOW_B200_HP_SLU_SPLIT_MAX", 0))    # 0: off
        words, trips, ranges = [], [], []
        classes = set()
        term_evals = null_evals = 0
        for units_of_phase in slu_phases:
            # units of the phase by class = (vertex kinds in order, propagator); a unit = [header word, [words of term 0], ..]
            by_class = {}
            for obj, o_, variants in units_of_phase:
                terms = slu_terms(obj)
                nomom = obj.get("mom", 2) == 0 and obj["finish"] != "none"
                key = (tuple(t["kind"] for t in terms), obj["finish"], nomom, pidx(obj.get("mass", "ZERO")), pidx(obj.get("width", "ZERO")))
                classes.add(key)
                out_off = o_ - 2 if nomom else o_
                for v in variants:
                    tw = []
                    for t in terms:
                        d = [desc(wfs[w]["off"], pext(v, vmask(obj["legs"], wfs[w]["legs"])), len(wfs[w]["legs"])) for w in t["ins"]]
                        one = [(d[0] | (t["coup"] * 4 + PHASE_CODE[t["ph"]]) << 22, d[1])]
                        if t["kind"] == "VVVV":
                            ca, cb, cc = t["coefs"]
                            one.append((d[2], (ca + 2) | (cb + 2) << 3 | (cc + 2) << 6))
                        tw.append(one)
                        term_evals += 1
                    by_class.setdefault(key, []).append(((desc(out_off, v, len(obj["legs"])), 0), tw))

            def cost(kinds, fin):
                return sum(KCOST[k] for k in kinds) + FCOST[fin]

            # parts: split the units of the class with the most expensive trip in two while that shortens the phase
            # (longest-processing-time schedule of the trips on the warps of the block)
            def part_kinds(kinds, g):
                return tuple(k for k in sorted(PT, key=PT.get) for _ in range(-(-kinds.count(k) // g)))

            def trip_cost(key, g):
                return cost(part_kinds(key[0], g), key[1]) + (4 if g > 1 else 0)

            def makespan(gs):
                load = [0.0] * NWARP
                costs = []
                for key, us in by_class.items():
                    costs += [trip_cost(key, gs[key])] * -(-len(us) * gs[key] // LPU)
                for c in sorted(costs, reverse=True):
                    load[load.index(min(load))] += c
                return max(load)

            gs = {key: 1 for key in by_class}
            for key in by_class:   # optional: units costlier than SPLIT_MAX are split whatever the schedule says
                while (SPLIT_MAX and trip_cost(key, gs[key]) > SPLIT_MAX and gs[key] < 8 and gs[key] * 2 <= LPU
                       and len(key[0]) >= 2 * gs[key]):
                    gs[key] *= 2
            while by_class:
                # the split (of any class) that shortens the phase most; stop when none does
                best, best_span = None, makespan(gs)
                for key in sorted(by_class, key=lambda key: (-trip_cost(key, gs[key]), key)):
                    g = gs[key]
                    if g >= 8 or g * 2 > LPU or len(key[0]) < 2 * g or trip_cost(key, g) < SPLIT_MIN:
                        continue
                    trial = dict(gs)
                    trial[key] = 2 * g
                    span = makespan(trial)
                    if span < best_span:
                        best, best_span = trial, span
                if best is None:
                    break
                gs = best
            ptrips = []
            for key, us in by_class.items():
                kinds, fin = key[0], key[1]
                g = gs[key]
                # every part evaluates the same kinds: per kind ceil(n / g) terms, short parts padded with null terms
                per = {k: -(-kinds.count(k) // g) for k in PT}
                pkinds = tuple(k for k in sorted(PT, key=PT.get) for _ in range(per[k]))
                subs = []   # [(header, flat words)] per (unit, part)
                for hdr, tw in us:
                    for part in range(g):
                        flat = []
                        for k in sorted(PT, key=PT.get):
                            mine = [w_ for w_, kk in zip(tw, kinds) if kk == k][part * per[k]:(part + 1) * per[k]]
                            nullw = [w_ for w_, kk in zip(tw, kinds) if kk == k][:1]
                            while len(mine) < per[k]:   # a null term: the inputs of a real one, f-index of the zero entry
                                w0 = nullw[0]
                                mine.append([((w0[0][0] & 0x3fffff) | NULL_F << 22, w0[0][1])] + w0[1:])
                                null_evals += 1
                            for w_ in mine:
                                flat += w_
                        subs.append([hdr] + flat)
                upt = LPU // g   # units per trip
                for i in range(0, len(us), upt):
                    ptrips.append((key, g, pkinds, subs[i * g:(i + upt) * g], min(upt, len(us) - i)))
            # trips of one class each, longest processing time first onto the least loaded warp
            load, mine = [0.0] * NWARP, [[] for _ in range(NWARP)]
            tcost = lambda q: cost(q[2], q[0][1]) + 4 * (q[1] > 1)
            for q in sorted(ptrips, key=lambda q: (-tcost(q), q[0], q[1])):
                wi = min(range(NWARP), key=lambda i: (load[i], i))
                load[wi] += tcost(q)
                mine[wi].append(q)
            if os.environ.get("MADFLOW_B200_HP_SLU_VERBOSE"):
                print(f"phase {len(ranges) // NWARP}: {len(ptrips)} trips, load max {max(load):.0f} mean {sum(load) / NWARP:.1f}; "
                      + ", ".join(f"{'+'.join(k[0] for k in q[2])}/{q[1]}" for q in sorted(ptrips, key=lambda q: -tcost(q))[:6]))
            for wi in range(NWARP):
                ranges.append((len(trips), len(trips) + len(mine[wi])))
                for key, g, pkinds, subs, nunits in mine[wi]:
                    _, fin, nomom, mi, wi_ = key
                    assert nunits < 256 and len(pkinds) <= 10 and mi < 15 and wi_ < 15
                    trips.append((len(words), nunits | (FINISH[fin] + (4 if nomom else 0)) << 8 | (mi + 1) << 12 | (wi_ + 1) << 16 | len(pkinds) << 20,
                                  sum(PT[k] << (3 * q) for q, k in enumerate(pkinds)), g.bit_length() - 1))
                    for j in range(len(subs[0])):
                        # lanes beyond the trip's units repeat its first unit: they take part in the shuffles and store nothing
                        words += [subs[u][j] if u < len(subs) else subs[u % g][j] for u in range(LPU)]
        words += [(0, 0)] * 128    # the kernel reads up to three words ahead of a unit's last one
        slu_tables = "\n".join([
            both("uint2", "slu_words", max(len(words), 1), ", ".join(f"{{{x}u, {y}u}}" for x, y in words) or "{0u, 0u}", const=False),
            both("uint4", "slu_trips", max(len(trips), 1), ", ".join(f"{{{x}u, {y}u, {z}u, {w_}u}}" for x, y, z, w_ in trips) or "{0u, 0u, 0u, 0u}"),
            both("int2", "slu_ranges", max(len(ranges), 1), ", ".join(f"{{{x}, {y}}}" for x, y in ranges) or "{0, 0}")])
        slu_stats = {"classes": len(classes), "trips": len(trips), "words": len(words), "term_evals": term_evals, "null_evals": null_evals,
                     "split": any(t_[3] for t_ in trips)}
    else:
        slu_tables = "\n".join([both("uint2", "slu_words", 1, "{0u, 0u}", const=False), both("uint4", "slu_trips", 1, "{0u, 0u, 0u, 0u}"),
                                both("int2", "slu_ranges", 1, "{0, 0}")])

    # ---- tables
    L = []
    L.append(both("mf::HpWf", "wf", len(wfs), ", ".join(f"{{{w['off']}u, {w['nv']}, {w['mask']}}}" for w in wfs)))
    ext_by_leg = sorted(exts, key=lambda x: x["call"]["leg"])
    assert [x["call"]["leg"] for x in ext_by_leg] == list(range(n))
    L.append(both("mf::HpExt", "ext", n, ", ".join(
        f"{{{HP_TYPES[x['call']['op']]}, {x['call']['leg']}, {x['call']['nsf']}, {pidx(x['call']['mass'])}, {x['out']}}}"
        for x in ext_by_leg)))
    # the term rows are the hottest table: constant memory (its cache is not squeezed by the shared-memory carve-out the
    # way L1 is) while they fit next to the other constant tables
    terms_const = len(trows) * 32 <= int(os.environ.get("MADFLOW_B200_HP_TERMS_CONST", 48 * 1024))
    L.append(both("mf::HpTerm", "terms", max(len(trows), 1), ",\n  ".join(trows) if trows else "{0}", const=terms_const))
    L.append(both("mf::HpWorkItem", "work_items", max(len(irow), 1), ", ".join(irow) if irow else "{0, 0, {0, 0, 0}, 0}", const=False))
    L.append(both("unsigned", "units", max(len(urow), 1), ", ".join(urow) if urow else "0u", const=False))
    L.append(both("int", "level_begin", len(begins), ", ".join(map(str, begins))))
    L.append(both("mf::HpTile", "tiles", max(len(tile_rows), 1), ",\n  ".join(tile_rows) if tile_rows else "{0}", const=len(tile_rows) * 32 <= 24576))
    L.append(both("mf::HpBatch", "batches", max(len(brow), 1), ", ".join(brow) if brow else "{0, 0, 0, 0}"))
    L.append(slu_tables)
    tables = "\n".join(L)

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
    unroll = len(plan["rows"]) <= int(os.environ.get("MADFLOW_B200_HP_UNROLL_MAX", HP_UNROLL_MAX_AMPS)) and not reduced
    if not hp_available(ir):
        unroll = False
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
    have_frag = False
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
        # the same matrix in the order of the tensor-core A fragments: [row block i][pair of k-blocks][lane][2], lane =
        # 4 r + k holds cfsym[8 i + r][4 kk + k] for kk = 2 kk2, 2 kk2 + 1 -- one coalesced 16-byte load per lane and pair
        if cmode == "mma" and (ncp // 4) % 2 == 0 and os.environ.get("MADFLOW_B200_HP_CFFRAG", "1") == "1":
            frag = []
            for i in range(ncp // 8):
                for kk2 in range(ncp // 8):
                    for lane in range(32):
                        r_, k_ = lane >> 2, lane & 3
                        for q in range(2):
                            frag.append(vals[(8 * i + r_) * ncp + 4 * (2 * kk2 + q) + k_])
            # (single precision would hold these small integers exactly and halve the bytes: measured, no gain)
            tables += "\n" + both("double", "cfsym_frag", len(frag), ", ".join(frag), const=False)
            have_frag = True
        else:
            tables += "\n" + both("double", "cfsym_frag", 2, "0.0, 0.0", const=False)
    else:
        tables += "\n" + both("double", "cfsym", 1, "0.0", const=False)
        tables += "\n" + both("double", "cfsym_frag", 2, "0.0, 0.0", const=False)
    # straight-line flavour of the amplitude phase for short amplitude lists (whole vertices per helicity combination)
    U = ["    cxd " + ", ".join(f"J{j} = mk(0.0, 0.0)" for j in range(len(ir["jamp"]))) + ";",
         "    cxd a[6], b[6], c[6], d[6];"]
    if unroll:
        cur = {}
        ids = iter(range(len(wfs)))
        for c in ir["calls"]:      # wavefunction ids = the order of the writes, as in hp_plan_from_ir
            if "amp" not in c:
                cur[c["out"]] = next(ids)
                continue
            if not any(k == c["amp"] for terms in ir["jamp"] for k, _, _ in terms):
                continue
            for q, s in enumerate(c["in"]):
                U.append(f"    mf::hp_load_amp<Proc>(wf_e, vtab, h, {cur[s]}, {'abcd'[q]});")
            op = c["op"]
            fn = f"VVVV_0<{op[4]}>" if op.startswith("VVVV") else op
            args = ", ".join("abcd"[: len(c["in"])])
            U.append(f"    {{ const cxd amp = mf::{fn}({args}, {_coup_expr(ir, c)});")
            for j, terms in enumerate(ir["jamp"]):
                for k, re, im in terms:
                    if k == c["amp"]:
                        U.append("      " + _jamp_update(j, float(re), float(im), "amp"))
            U.append("    }")
        U.append(_emit_colour(ir))
    else:
        U.append("    return 0.0;  // not used: the amplitudes of this process run on the tensor cores")
    nunits_cur = begins[-1]
    return tables, "\n".join(A), C, "\n".join(U), dict(
        wfsize=wfsize, maxlevel=maxlevel, nwf=len(wfs), nitems=nunits_cur, namps=len(plan["rows"]), unroll=unroll,
        nbatch=len(batches), npairs=len(pairs), nitems_pair=len(urow) - nunits_cur, ntiles=len(tile_rows), ncg=1 if unroll else NCG,
        jamp_terms=jamp_terms, npass=1 if unroll else NPASS, cmode={"thread": 0, "groups": 1, "mma": 2, "loop": 3}[cmode], ncp=ncp,
        colour_denom=float(den[0]), reduced=reduced, cf_frag=have_frag, nterms=len(trows), slu=SLU, slu_stats=slu_stats,
        nb=max(len(a_) for _, a_ in batches) if batches else 1,
        scratch=max((pairs[pi]["abs_off"] - wfsize + 4 * pairs[pi]["nv"] for b_ in batches for pi in b_[0]), default=0))


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
    hp_scratch_n = 0 if hp['unroll'] else hp['scratch']
    hp_nb = 0 if hp['unroll'] else hp['nb']
    hp_ncg = hp['ncg']
    hp_nj = -(-ncolor // hp_ncg)
    hp_e = hp_config(ir, "E")
    hp_ok = hp_available(ir)
    use_hp = "true" if (use_hp_default(ir) and hp_ok) else "false"
    has_thread = "true" if (len(ir["calls"]) <= THREAD_MAX_CALLS or not hp_ok or os.environ.get("MADFLOW_B200_BUILD_THREAD") == "1") else "false"
    if not hp_ok and len(ir["calls"]) > STRAIGHT_LINE_MAX_CALLS:
        raise ValueError("call lists with plugin-written vertex routines are limited to the one-event-per-thread flavour "
                         f"({STRAIGHT_LINE_MAX_CALLS} calls)")
    hp_minblocks, hp_wfsize, hp_maxlevel, hp_nwf, hp_nitems = hp_config(ir, "MINBLOCKS"), hp["wfsize"], hp["maxlevel"], hp["nwf"], hp["nitems"]
    pnames = ", ".join(f'"{p}"' for p in ir["params"]) or '""'
    cnames = ", ".join(f'"{c}"' for c in ir["couplings"]) or '""'
    src = f"""// GENERATED by madflow_b200.codegen -- do not edit.  Process: {ir.get('process', ir['name'])}
// One fused FP64 kernel per process: HELAS wavefunctions -> ALOHA vertices -> JAMP -> colour matrix.
#include "process_kernels_hp.cuh"
{'#include <cuda/std/complex>' if plugin_ops(ir) else ''}
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

{_emit_plugin_routines(ir)}

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
  // false: the call list uses vertex routines outside the table-driven set (written by the plugin's ALOHA writer)
  static constexpr bool HP_AVAILABLE = {'true' if hp_ok else 'false'};
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
  // objects with many terms are split over neighbouring lanes (codegen.hp_config TSPLIT)
  static constexpr bool HP_SPLIT = {'true' if hp_config(ir, "TSPLIT") else 'false'};
  // the JAMP accumulators are parked in Tensor Memory between the JAMP phases of the batches (process_kernels_hp.cuh)
  static constexpr bool HP_TMEM_J = {'true' if (hp_config(ir, "TMEMJ") and not hp['unroll']) else 'false'};
  // tile descriptors per warp and trip (x HP_E tiles in flight): 2 tiles in flight measured best -- more only adds
  // padded tiles at the end of a batch (g g > t t~ g g: 17.7e6 events/s with 2, 16.5e6 with 4, 14.5e6 with 8)
  // g g > t t~ g g g with packed units: 1 (no spills at 128 registers, 1.24e6 -> 1.30e6 events/s; 4: 1.18e6)
  static constexpr int HP_TILES_IN_FLIGHT = {max(1, hp_config(ir, "MT") // hp_e)};
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
  static constexpr bool HP_CF_FRAG = {'true' if hp['cf_frag'] else 'false'};   // cfsym also stored in tensor-core fragment order
  MF_DEV static const double* cfsym_frag() {{ return MF_TAB(cfsym_frag); }}
  MF_DEV static mf::HpWf wf(int w) {{ return MF_TAB(wf)[w]; }}
  MF_DEV static mf::HpExt ext(int leg) {{ return MF_TAB(ext)[leg]; }}
  MF_DEV static mf::HpTerm term(int i) {{ return mf::hp_fetch32(&MF_TAB(terms)[i]); }}
  MF_DEV static mf::HpWorkItem work_item(int i) {{ return MF_TAB(work_items)[i]; }}  // one 8-byte load
  MF_DEV static unsigned unit(int i) {{ return MF_TAB(units)[i]; }}
  MF_DEV static int level_begin(int L) {{ return MF_TAB(level_begin)[L]; }}
  MF_DEV static const mf::HpTile* tile(int i) {{ return &MF_TAB(tiles)[i]; }}
  MF_DEV static mf::HpBatch batch(int i) {{ return MF_TAB(batches)[i]; }}
  // packed units (process_kernels_hp.cuh "SLU"): trips per (phase, warp), one class of objects per trip
  static constexpr bool HP_SLU = {'true' if hp['slu'] else 'false'};
  // the words of the next term are fetched while the current one is evaluated (tables beyond L1; measured per process)
  static constexpr bool HP_SLU_PREFETCH = {'true' if hp['slu_stats'].get('words', 0) * 8 > int(os.environ.get("MADFLOW_B200_HP_SLU_PREFETCH_BYTES", 65536)) else 'false'};
  static constexpr bool HP_PREFETCH_INPUTS = {'true' if hp_config(ir, "PREFIN") else 'false'};
  static constexpr bool HP_SLU_SPLIT = {'true' if hp['slu_stats'].get('split') else 'false'};   // some units are split over lanes
  MF_DEV static int2 slu_range(int i) {{ return MF_TAB(slu_ranges)[i]; }}
  MF_DEV static uint4 slu_trip(int i) {{ return MF_TAB(slu_trips)[i]; }}
  MF_DEV static const uint2* slu_words() {{ return MF_TAB(slu_words); }}
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
