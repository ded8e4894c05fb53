"""Matrix_<proc>.matrix / .smatrix interpreter over a process IR.  TEST INFRASTRUCTURE.

Follows madgraph_plugin/template_files/matrix_method_python.inc:80-138 (helicity sum, HELAS
call list, JAMP, colour contraction) and the frozen instance tests/mockup_debug_me.py:453-543.
The IR (a plain dict, see madflow_b200/process_ir.py) is data: the ordered HELAS call list the
reference's generated `matrix()` would execute, the amp->jamp coefficient table, the integer
colour matrix and its per-row denominators, the helicity table and the averaging denominator.
"""
import numpy as np

from . import REFERENCE, aloha, helas


def _param(params, name, sign=1.0):
    if name == "ZERO":
        return 0.0
    return sign * params[name]


def matrix(ir, all_ps, hel, params, const=REFERENCE, return_jamp=False):
    """One helicity configuration: (nevt,) real.  matrix_method_python.inc:106-138."""
    all_ps = np.asarray(all_ps, dtype=np.float64)
    w = {}
    amp = {}
    ext = {"vxxxxx": helas.vxxxxx, "ixxxxx": helas.ixxxxx, "oxxxxx": helas.oxxxxx, "sxxxxx": helas.sxxxxx}
    for c in ir["calls"]:
        op = c["op"]
        if op in ext:
            leg = c["leg"]
            if op == "sxxxxx":
                w[c["out"]] = helas.sxxxxx(all_ps[:, leg], c["nsf"])
            else:
                mass = _param(params, c["mass"])
                w[c["out"]] = ext[op](all_ps[:, leg], mass, hel[leg], c["nsf"], const)
            continue
        fn = aloha.ROUTINES[op]
        ins = [w[i] for i in c["in"]]
        coup = _param(params, c["coup"], c.get("coup_sign", 1.0))
        if "amp" in c:
            amp[c["amp"]] = fn(*ins, coup)
        else:
            w[c["out"]] = fn(*ins, coup, _param(params, c["mass"]), _param(params, c["width"]))
    jamp = []
    for terms in ir["jamp"]:
        acc = None
        for k, re, im in terms:
            t = complex(re, im) * amp[k]
            acc = t if acc is None else acc + t
        jamp.append(acc)
    jamp = np.stack(np.broadcast_arrays(*jamp))
    if return_jamp:
        return jamp
    cf = np.asarray(ir["color_num"], dtype=np.complex128)
    denom = np.asarray(ir["color_denom"], dtype=np.complex128)
    # matrix_method_python.inc:137
    ret = np.einsum("ie,ij,je->e", jamp, cf, np.conj(jamp) / denom.reshape(-1, 1))
    return ret.real


def smatrix(ir, all_ps, params, const=REFERENCE):
    """Sum over ALL helicity rows, then divide by the averaging factor
    (matrix_method_python.inc:99-104)."""
    all_ps = np.asarray(all_ps, dtype=np.float64)
    ans = np.zeros(all_ps.shape[0])
    for hel in ir["helicities"]:
        ans = ans + matrix(ir, all_ps, hel, params, const)
    return ans / ir["denominator"]


def smatrix_recycled(ir, all_ps, params, const=REFERENCE, chunk=256):
    """The same sum as smatrix(), organised for long call lists: a wavefunction over the leg set S depends on the
    helicities of S only, so it is evaluated once per helicity assignment of its OWN legs and looked up for every
    helicity row that contains that assignment (memoisation of the calls of `matrix`; every value is produced by the
    same routine from the same inputs).  Amplitudes, JAMPs (as one matrix product with the amp->jamp coefficient
    table) and the colour contraction are still evaluated for every helicity row.  g g > t t~ g g g: ~0.1 s per
    point instead of ~7 s, which is what makes parity tests on hundreds of points affordable.  Events are processed
    in chunks to bound the memory of the cache."""
    all_ps = np.asarray(all_ps, dtype=np.float64)
    nevt = all_ps.shape[0]
    if nevt > chunk:
        out = np.empty(nevt)
        for lo in range(0, nevt, chunk):
            sl = slice(lo, min(lo + chunk, nevt))
            sub = {k: (v[sl] if np.ndim(v) else v) for k, v in params.items()}
            out[sl] = smatrix_recycled(ir, all_ps[sl], sub, const, chunk)
        return out
    ext = {"vxxxxx": helas.vxxxxx, "ixxxxx": helas.ixxxxx, "oxxxxx": helas.oxxxxx}
    # undo the slot reuse: one wavefunction id per write; legs below every wavefunction
    cur, legs, steps = {}, [], []          # steps: ("ext"|"wf"|"amp", call, input ids, own id)
    for c in ir["calls"]:
        if "leg" in c:
            cur[c["out"]] = len(legs)
            legs.append((c["leg"],))
            steps.append(("ext", c, (), cur[c["out"]]))
        elif "amp" in c:
            steps.append(("amp", c, tuple(cur[s] for s in c["in"]), c["amp"]))
        else:
            ins = tuple(cur[s] for s in c["in"])
            cur[c["out"]] = len(legs)
            legs.append(tuple(sorted(set().union(*[legs[i] for i in ins]))))
            steps.append(("wf", c, ins, cur[c["out"]]))
    namps = 1 + max(c["amp"] for c in ir["calls"] if "amp" in c)
    ncolor = len(ir["jamp"])
    coef = np.zeros((ncolor, namps), dtype=np.complex128)
    for j, terms in enumerate(ir["jamp"]):
        for k, re, im in terms:
            coef[j, k] += complex(re, im)
    cf = np.asarray(ir["color_num"], dtype=np.float64) / np.asarray(ir["color_denom"], dtype=np.float64)[None, :]
    cache = {}
    ans = np.zeros(nevt)
    for hel in ir["helicities"]:
        w = {}
        amps = np.zeros((namps, nevt), dtype=np.complex128)
        for kind, c, ins, own in steps:
            if kind == "amp":
                coup = _param(params, c["coup"], c.get("coup_sign", 1.0))
                amps[own] = aloha.ROUTINES[c["op"]](*[w[i] for i in ins], coup)
                continue
            key = (own, tuple(hel[l] for l in legs[own]))
            if key not in cache:
                if kind == "ext":
                    leg = c["leg"]
                    cache[key] = ext[c["op"]](all_ps[:, leg], _param(params, c["mass"]), hel[leg], c["nsf"], const)
                else:
                    coup = _param(params, c["coup"], c.get("coup_sign", 1.0))
                    cache[key] = aloha.ROUTINES[c["op"]](*[w[i] for i in ins], coup, _param(params, c["mass"]),
                                                         _param(params, c["width"]))
            w[own] = cache[key]
        jamp = coef @ amps
        # matrix_method_python.inc:137  Re sum_ij J_i cf_ij conj(J_j) / denom_j
        ans += np.einsum("ie,ij,je->e", jamp, cf, np.conj(jamp)).real
    return ans / ir["denominator"]
