#!/usr/bin/env python3
"""Generate the committed golden vectors in tests/golden/*.npz.

Run ONLY in the build container (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's UNMODIFIED Python sources
    /root/reference/python_package/madflow/wavefunctions_flow.py
    /root/reference/python_package/madflow/phasespace.py
    /root/reference/python_package/madflow/parameters.py
    /root/reference/python_package/madflow/tests/mockup_debug_me.py
with `oracle/tfshim` (a numpy stand-in for the TensorFlow ops they call) first on
sys.path, evaluates them on seeded inputs and stores inputs + outputs.  The tests then check
(1) the oracle restatement and (2) the CUDA path against these files; nothing at test time
reads /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "tfshim"))
sys.path.insert(0, "/root/reference/python_package")
np.seterr(all="ignore")

import madflow.wavefunctions_flow as wf  # noqa: E402
import madflow.phasespace as ps  # noqa: E402
import madflow.parameters as par  # noqa: E402
import madflow.tests.mockup_debug_me as mock  # noqa: E402

MT = 173.0
WT = 1.4915000200271606


def f64(x):
    return np.float64(x)


def momenta_set(rng, n, mass):
    """Random on-shell momenta plus the special kinematics every per-event branch needs."""
    pv = rng.normal(size=(n, 3)) * 300.0
    e = np.sqrt(np.sum(pv**2, axis=1) + mass**2)
    p = np.concatenate([e[:, None], pv], axis=1)
    special = []
    for pz in (250.0, -250.0):  # along the beam: pt == 0, and pp+pz == 0 for pz < 0
        special.append([np.sqrt(pz * pz + mass * mass), 0.0, 0.0, pz])
    if mass != 0.0:
        special.append([mass, 0.0, 0.0, 0.0])  # at rest: pp == 0
    special.append([np.sqrt(3.0**2 + 4.0**2 + mass**2), 3.0, 4.0, 0.0])  # pz == 0
    return np.concatenate([np.array(special), p], axis=0)


def gen_wavefunctions(out):
    rng = np.random.default_rng(20261017)
    rec = {}
    for mass in (0.0, MT):
        p = momenta_set(rng, 60, mass)
        rec[f"p_m{int(mass)}"] = p
        for nsf in (-1, 1):
            # sxxxxx is not exercised: wavefunctions_flow.py:49 calls tf.expand_dims(<scalar>, 1),
            # which is an error in TensorFlow as well -- the reference's sxxxxx cannot run.
            for nhel in (-1, 1):
                key = f"m{int(mass)}_h{nhel}_s{nsf}"
                rec["i_" + key] = wf.ixxxxx(p, f64(mass), f64(nhel), f64(nsf))
                rec["o_" + key] = wf.oxxxxx(p, f64(mass), f64(nhel), f64(nsf))
            for nhel in (-1, 1, 4) + ((0,) if mass != 0.0 else ()):
                key = f"m{int(mass)}_h{nhel}_s{nsf}"
                rec["v_" + key] = wf.vxxxxx(p, f64(mass), f64(nhel), f64(nsf))
    np.savez_compressed(out, **rec)


def gen_aloha(out):
    rng = np.random.default_rng(7)
    n = 48

    def cw():
        return rng.normal(size=(6, n)) + 1j * rng.normal(size=(6, n))

    F1, F2, V2, V3 = cw(), cw(), cw(), cw()
    c10, c11 = mock.GC_10, mock.GC_11
    rec = dict(F1=F1, F2=F2, V2=V2, V3=V3, GC_10=np.complex128(c10), GC_11=np.complex128(c11), M=MT, W=WT)
    rec["FFV1_0"] = mock.FFV1_0(F1, F2, V3, c11)
    rec["FFV1_1"] = mock.FFV1_1(F2, V3, c11, f64(MT), f64(WT))
    rec["FFV1_2"] = mock.FFV1_2(F1, V3, c11, f64(MT), f64(WT))
    rec["VVV1P0_1"] = mock.VVV1P0_1(V2, V3, c10, f64(0.0), f64(0.0))
    np.savez_compressed(out, **rec)


def gen_phasespace(out):
    rng = np.random.default_rng(11)
    rec = {}
    # massless rambo, n = 2..7 (reference tests/test_ps.py:19-39), fixed and per-event sqrts
    for n in range(2, 8):
        x = rng.random((16, 4 * n))
        p, w = ps.rambo(x, n, 7e3, masses=None)
        rec[f"rambo{n}_x"], rec[f"rambo{n}_p"], rec[f"rambo{n}_w"] = x, p, w
    x = rng.random((13, 28))
    sq = rng.random(13) * 7e3
    p, w = ps.rambo(x, 7, sq, masses=None)
    rec["rambo7v_x"], rec["rambo7v_s"], rec["rambo7v_p"], rec["rambo7v_w"] = x, sq, p, w
    # ramboflow: the configs of BASELINE.json (sqrts 13 TeV, top masses) + test_fourmomenta's masses
    cases = {
        "tt": (4, 13e3, [MT, MT]),
        "ttg": (5, 13e3, [MT, MT, 0.0]),
        "ttgg": (6, 13e3, [MT, MT, 0.0, 0.0]),
        "ttggg": (7, 13e3, [MT, MT, 0.0, 0.0, 0.0]),
        "m50_125": (4, 7e3, [50.0, 125.0]),
        "massless5": (5, 7e3, None),
        "tt7": (4, 7e3, [MT, MT]),
    }
    for name, (npart, sqrts, masses) in cases.items():
        x = rng.random((64, 4 * (npart - 2) + 2))
        p, w, x1, x2 = ps.ramboflow(x, npart, sqrts, masses=masses)
        rec[f"rf_{name}_x"], rec[f"rf_{name}_p"], rec[f"rf_{name}_w"] = x, p, w
        rec[f"rf_{name}_x1"], rec[f"rf_{name}_x2"] = x1, x2
        rec[f"rf_{name}_lab"] = ps._boost_to_lab(p, x1, x2)
        if masses is not None and npart > 4:
            # the reference's Newton loop stops for the whole batch as soon as one event is done
            # (phasespace.py:90-92): also store single-event batches, which are batch-independent
            ps1 = [ps.ramboflow(x[i : i + 1], npart, sqrts, masses=masses) for i in range(16)]
            rec[f"rf_{name}_p_single"] = np.concatenate([r[0] for r in ps1])
            rec[f"rf_{name}_w_single"] = np.concatenate([r[1] for r in ps1])
    # 2 -> 1
    x = rng.random((8, 2))
    # sqrts as a float64 tensor, as PhaseSpaceGenerator passes it (phasespace.py:387); a bare
    # Python float would be float32-rounded by tf.sqrt at phasespace.py:243
    p, w, x1, x2 = ps.ramboflow(x, 3, f64(13e3), masses=[91.188])
    rec["rf_21_x"], rec["rf_21_p"], rec["rf_21_w"], rec["rf_21_x1"], rec["rf_21_x2"] = x, p, w, x1, x2
    # PhaseSpaceGenerator with cuts (reference tests/test_ps.py:42-65 and madflow_exec.py:389-395)
    gen = ps.PhaseSpaceGenerator(5, 7e3, algorithm="ramboflow")
    gen.register_cut("pt", particle=3, min_val=60, max_val=300.0)
    x = rng.random((200, 14))
    a, w, x1, x2, idx = gen(x)
    rec["psg5_x"], rec["psg5_p"], rec["psg5_w"], rec["psg5_x1"], rec["psg5_x2"], rec["psg5_idx"] = x, a, w, x1, x2, idx
    gen = ps.PhaseSpaceGenerator(5, 13e3, [MT, MT, 0.0], com_output=False)
    for i in range(2, 5):
        gen.register_cut("pt", particle=i, min_val=30.0)
    # single-event batches => batch-independent Newton iteration count
    outs = [gen(x[i : i + 1]) for i in range(64)]
    keep = [o for o in outs if o[0].shape[0] == 1]
    rec["psglab_x"] = x[:64]
    rec["psglab_pass"] = np.array([o[0].shape[0] == 1 for o in outs])
    rec["psglab_p"] = np.concatenate([o[0] for o in keep])
    rec["psglab_w"] = np.concatenate([o[1] for o in keep])
    rec["psglab_mt"] = np.concatenate([ps.PhaseSpaceGenerator.mt(o[0][:, 2:5, :]) for o in keep])
    rec["const_PI"] = np.float64(ps.PI)
    rec["const_ACC"] = np.float64(ps.ACC)
    rec["const_SQH"] = np.float64(wf.SQH)
    np.savez_compressed(out, **rec)


def gen_matrix(out):
    rng = np.random.default_rng(4)
    rec = {}
    m = mock.Matrix_1_gg_ttx()
    rec["helicities"] = np.asarray(m.helicities)
    rec["denominator"] = np.float64(m.denominator)
    rec["params"] = np.array([MT, WT])
    rec["GC_10"] = np.complex128(mock.GC_10)
    rec["GC_11"] = np.complex128(mock.GC_11)
    for name, sqrts in (("13tev", 13e3), ("7tev", 7e3)):
        x = rng.random((96, 10))
        p, w, x1, x2 = ps.ramboflow(x, 4, sqrts, masses=[MT, MT])
        lab = ps._boost_to_lab(p, x1, x2)
        for frame, mom in (("com", p), ("lab", lab)):
            rec[f"{name}_{frame}_p"] = mom
            rec[f"{name}_{frame}_smatrix"] = m.smatrix(mom, *mock.model_params)
            rec[f"{name}_{frame}_matrix"] = np.stack(
                [m.matrix(mom, h, *mock.model_params) for h in np.asarray(m.helicities)]
            )
    # per-event couplings path of the generated code (template: couplings are [None]-shaped)
    a_s = 0.09 + 0.06 * rng.random(96)
    gs = par._alphas_to_gs(a_s)
    rec["run_alpha_s"], rec["run_gs"] = a_s, gs
    rec["run_smatrix"] = m.smatrix(rec["13tev_lab_p"], f64(MT), f64(WT), -gs, 1j * gs)
    np.savez_compressed(out, **rec)


def gen_model(out):
    import collections

    C = collections.namedtuple("constants", ["mdl_MT", "mdl_WT"])
    F = collections.namedtuple("functions", ["GC_10", "GC_11", "GC_12"])
    model = par.Model(
        C(f64(MT), f64(WT)),
        F(lambda G: -G, lambda G: complex(0, 1) * G, lambda G: complex(0, 1) * G**2),
    )
    a_s = np.array([0.118, 0.1, 0.13, 0.0935])
    ev = model.evaluate(a_s)
    rec = dict(alpha_s=a_s, GC_10=ev[2], GC_11=ev[3], GC_12=ev[4], masses=np.array(model.get_masses()))
    model.freeze_alpha_s(0.118)  # goes through float_me([0.118]) => float32-rounded alpha_s
    fr = model.evaluate(None)
    rec.update(frozen_GC_10=fr[2], frozen_GC_11=fr[3], frozen_GC_12=fr[4])
    np.savez_compressed(out, **rec)


if __name__ == "__main__":
    gen_wavefunctions(os.path.join(HERE, "wavefunctions.npz"))
    gen_aloha(os.path.join(HERE, "aloha_mockup.npz"))
    gen_phasespace(os.path.join(HERE, "phasespace.npz"))
    gen_matrix(os.path.join(HERE, "matrix_gg_ttx.npz"))
    gen_model(os.path.join(HERE, "model.npz"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
