"""TEST INFRASTRUCTURE: numpy interpreter of a reduced evaluation plan (madflow_b200/recursion.py).

Evaluates the plan the way the kernels do -- basis currents as sums of raw sub-trees, rows = x . (sum of vertex
numerators), JAMPs from the rows -- but with the oracle's HELAS / ALOHA routines, one helicity row at a time, so
that the REDUCTION itself (which sub-trees are added with which coefficients, which JAMPs a row feeds) can be
compared with the diagram-by-diagram oracle (oracle/matrix.py) without any CUDA code."""
import numpy as np

from oracle import REFERENCE, aloha, helas
from oracle.matrix import _param


def matrix(ir, all_ps, hel, params, const=REFERENCE):
    plan = ir["plan"]
    p = np.asarray(all_ps, dtype=np.float64)
    ext = {"vxxxxx": helas.vxxxxx, "ixxxxx": helas.ixxxxx, "oxxxxx": helas.oxxxxx}
    w = []
    for o in plan["objects"]:
        if o["ext"] is not None:
            c = o["ext"]
            w.append(ext[c["op"]](p[:, c["leg"]], _param(params, c["mass"]), hel[c["leg"]], c["nsf"], const))
            continue
        acc = None
        for t in o["terms"]:
            r = aloha.ROUTINES[t["op"]](*[w[i] for i in t["in"]], _param(params, t["coup"]),
                                        _param(params, o["mass"]), _param(params, o["width"]))
            r = np.concatenate([r[:2], complex(*t["coef"]) * r[2:]])
            acc = r if acc is None else np.concatenate([acc[:2], acc[2:] + r[2:]])
        w.append(acc)
    ncolor = len(ir["jamp"])
    jamp = [0.0] * ncolor
    for row in plan["rows"]:
        amp = 0.0
        for t in plan["pairs"][row["pair"]]["terms"]:
            ins = [w[i] for i in t["in"]]
            ins.insert(t["jx"], w[row["x"]])
            amp = amp + complex(*t["coef"]) * aloha.ROUTINES[t["op"]](*ins, _param(params, t["coup"]))
        for f, re, im in row["jamp"]:
            jamp[f] = jamp[f] + complex(re, im) * amp
    jamp = np.stack(np.broadcast_arrays(*[j + np.zeros(p.shape[0], dtype=complex) for j in jamp]))
    cf = np.asarray(ir["color_num"], dtype=np.complex128)
    denom = np.asarray(ir["color_denom"], dtype=np.complex128)
    return np.einsum("ie,ij,je->e", jamp, cf, np.conj(jamp) / denom.reshape(-1, 1)).real


def smatrix(ir, all_ps, params, const=REFERENCE):
    ans = 0.0
    for hel in ir["helicities"]:
        ans = ans + matrix(ir, all_ps, hel, params, const)
    return ans / ir["denominator"]
