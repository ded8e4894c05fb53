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
