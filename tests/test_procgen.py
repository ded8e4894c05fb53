"""The built-in process generator (madflow_b200/procgen.py), CPU only.

MG5_aMC is not available offline, so nothing here can be compared with MG5 output directly except
g g > t t~ (frozen in the reference's tests/mockup_debug_me.py).  The other processes are pinned by
MG5's known counts and by physics: gauge invariance, Bose symmetry, two independent organisations of
the same amplitude, colour-matrix identities."""
import itertools
import shutil

import numpy as np
import pytest

from conftest import G, MT, WT, sm_params
from madflow_b200 import codegen, process_ir, procgen
from oracle import EXACT, SQH_REF
from oracle import matrix as omatrix
from oracle import phasespace as ops


@pytest.fixture(scope="module")
def irs():
    return {k: procgen.generate_ir(k) for k in (0, 1, 2, 3)}


def test_gg_ttx_reproduces_the_references_generated_code(irs):
    pin, gen = process_ir.gg_ttx_pinned(), irs[0]
    assert gen["calls"] == pin["calls"]
    assert [list(map(tuple, t)) for t in gen["jamp"]] == [list(t) for t in pin["jamp"]]
    assert gen["helicities"] == pin["helicities"]
    assert (gen["color_num"], gen["color_denom"], gen["denominator"]) == ([[16, -2], [-2, 16]], [3, 3], 256)


def test_counts_match_mg5():
    """ndiag 3/16/123/1240, amplitudes 3/18/159, ncolor 2/6/24/120,
    denominators 256/256/512/1536."""
    expect = {0: (3, 3, 2, 256), 1: (16, 18, 6, 256), 2: (123, 159, 24, 512), 3: (1240, None, 120, 1536)}
    for k, (ndiag, namp, ncolor, den) in expect.items():
        ir = procgen.generate_ir(k)
        assert process_ir.validate(ir)
        assert ir["ndiags"] == ndiag and len(ir["jamp"]) == ncolor and ir["denominator"] == den
        if namp:
            assert len({c["amp"] for c in ir["calls"] if "amp" in c}) == namp
        assert ir["ncomb"] == 2 ** (4 + k)
    # g g > t t~ g colour matrix, first row as MG5 prints it
    assert procgen.generate_ir(1)["color_num"][0] == [64, -8, -8, 1, 1, 10] and procgen.generate_ir(1)["color_denom"][0] == 9


def test_colour_matrix_identities(irs):
    for k, ir in irs.items():
        c = np.array(ir["color_num"], dtype=float) / np.array(ir["color_denom"], dtype=float)[:, None]
        assert np.allclose(c, c.T)
        assert np.min(np.linalg.eigvalsh(c)) > -1e-10           # a Gram matrix
        ng = 2 + k
        cf = 4.0 / 3.0
        assert np.allclose(np.diag(c), 3 * cf**ng)              # Tr(T^a1..T^an T^an..T^a1) = N CF^n
        assert len({round(v, 9) for v in c.sum(axis=1)}) == 1   # rows are permutations of each other


def _points(k, n=24, seed=0):
    nn = 4 + k
    x = np.random.default_rng(seed).random((n, 4 * (nn - 2) + 2))
    p, _, x1, x2 = ops.ramboflow(x, nn, 13e3, [MT, MT] + [0.0] * k, const=EXACT, xfactor="converged")
    return p


@pytest.mark.parametrize("k", [1, 2])
def test_gauge_invariance_and_two_organisations(irs, k):
    ir = irs[k]
    alt = procgen.generate_ir(k, root="tbar")  # every diagram closed on the t~ leg instead of the centroid
    p = _points(k)
    params = dict(sm_params(), mdl_WT=0.0)
    a = omatrix.smatrix(ir, p, params, EXACT)
    b = omatrix.smatrix(alt, p, params, EXACT)
    np.testing.assert_allclose(a, b, rtol=5e-12)
    hel = [1, -1, 1, -1] + [1, -1, 1][:k]
    phys = np.max(np.abs(omatrix.matrix(ir, p, hel, params, EXACT, return_jamp=True)))
    for gl in [0, 1] + list(range(4, 4 + k)):
        h = list(hel)
        h[gl] = 4  # BRST polarisation (wavefunctions_flow.py:146-152)
        assert np.max(np.abs(omatrix.matrix(ir, p, h, params, EXACT, return_jamp=True))) < 1e-10 * phys


def test_bose_symmetry(irs):
    params = sm_params()
    for k in (1, 2):
        p = _points(k, seed=3)
        a = omatrix.smatrix(irs[k], p, params)
        q = p.copy()
        q[:, [0, 1]] = q[:, [1, 0]]
        np.testing.assert_allclose(omatrix.smatrix(irs[k], q, params), a, rtol=1e-11)
    q = p.copy()
    q[:, [4, 5]] = q[:, [5, 4]]
    np.testing.assert_allclose(omatrix.smatrix(irs[2], q, params), a, rtol=1e-11)


def test_flops_per_event(irs):
    assert codegen.flops_per_event(irs[0]) == 23697
    assert 1.5e5 < codegen.flops_per_event(irs[1]) < 2.5e5
    assert 2e6 < codegen.flops_per_event(irs[2]) < 4e6


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_generated_cuda_source_on_host_ttxg(irs):
    """The emitted straight-line code for g g > t t~ g, executed on the CPU, against the oracle."""
    import hostcheck as hc

    ir = irs[1]
    lib = hc.process(ir)
    p = _points(1, n=100, seed=7)
    a_s = 0.09 + 0.05 * np.random.default_rng(9).random(100)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    out = hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF)
    np.testing.assert_allclose(out, omatrix.smatrix(ir, p, params), rtol=1e-12)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
@pytest.mark.parametrize("k", [1, 2, 3])
def test_helicity_parallel_data_flow_on_host(irs, k):
    """The helicity-parallel kernels' tables and generated code (currents once per helicity variant,
    pair objects, amplitude tiles -> amplitude buffer -> JAMP code per colour group, helicity passes,
    colour groups / block-symmetrised colour matrix) executed phase by phase on the CPU against the
    oracle; also one helicity row on its own.  g g > t t~ g g g: 2 points (the oracle needs ~7 s each)."""
    import hostcheck as hc

    ir = irs[k]
    lib = hc.process(ir)
    npts = 2 if k == 3 else 6
    p = _points(k, n=npts, seed=11)
    a_s = 0.09 + 0.05 * np.random.default_rng(3).random(npts)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    ref = omatrix.smatrix(ir, p, params)
    out = hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, hp=True)
    np.testing.assert_allclose(out, ref, rtol=1e-12)
    row = 5
    one = hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, only_comb=row, hp=True)
    np.testing.assert_allclose(one, omatrix.matrix(ir, p, ir["helicities"][row], params), rtol=1e-11)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_packed_units_equal_the_table_driven_units_on_host(irs, monkeypatch):
    """The two implementations of the unit phases of g g > t t~ g g -- packed units (warp trips of one class, merged
    four-gluon terms, one 64-bit word per term: the default) and the table-driven routine (MADFLOW_B200_HP_SLU=0, the
    path of every other process) -- executed on the CPU on the same points: equal to 1e-13 and both equal to the oracle.
    Also the bookkeeping of the packing: every (object, helicity variant) of the plan is evaluated, no null term without
    a split unit, and the split of g g > t t~ g g g pads less than 10 % of its term evaluations."""
    import copy

    import hostcheck as hc

    ir = irs[2]
    p = _points(2, n=6, seed=21)
    a_s = 0.09 + 0.05 * np.random.default_rng(5).random(6)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    ref = omatrix.smatrix(ir, p, params)
    stats = codegen.emit_hp(ir)[4]
    assert stats["slu"] and stats["slu_stats"]["null_evals"] == 0 and not stats["slu_stats"]["split"]
    packed = hc.smatrix(hc.process(ir), ir, p, [MT, WT], coup, SQH_REF, hp=True)
    monkeypatch.setenv("MADFLOW_B200_HP_SLU", "0")
    tab_ir = copy.deepcopy(ir)
    tab_ir["name"] = ir["name"] + "_tabledriven"     # not a built-in name: its source stays with the test artefacts
    tab_stats = codegen.emit_hp(tab_ir)[4]
    assert not tab_stats["slu"]
    table = hc.smatrix(hc.process(tab_ir), tab_ir, p, [MT, WT], coup, SQH_REF, hp=True)
    monkeypatch.delenv("MADFLOW_B200_HP_SLU")
    np.testing.assert_allclose(packed, table, rtol=1e-13)
    np.testing.assert_allclose(packed, ref, rtol=1e-12)
    # merged four-gluon terms: fewer term evaluations than work items of the table-driven routine
    assert stats["slu_stats"]["term_evals"] == 1196
    big = codegen.emit_hp(irs[3])[4]["slu_stats"]
    assert big["split"] and 0 < big["null_evals"] < 0.1 * big["term_evals"]


def _mandelstam(p):
    dot = lambda a, b: a[:, 0] * b[:, 0] - np.sum(a[:, 1:] * b[:, 1:], axis=1)
    s = dot(p[:, 0] + p[:, 1], p[:, 0] + p[:, 1])
    t = dot(p[:, 0] - p[:, 2], p[:, 0] - p[:, 2])
    u = dot(p[:, 0] - p[:, 3], p[:, 0] - p[:, 3])
    return s, t, u


def test_qqbar_ttx_closed_form():
    """q q~ > t t~, the second subprocess of `p p > t t~`: the textbook result
    sum |M|^2 / 36 = (4 g^4 / 9) [(m^2 - t)^2 + (m^2 - u)^2 + 2 m^2 s] / s^2  (no top propagator: any width)."""
    ir = procgen.qqbar_ttx_ir()
    assert process_ir.validate(ir)
    assert ir["initial_states"] == [[2, -2], [4, -4], [1, -1], [3, -3]] and ir["mirror_initial_states"]
    assert procgen.qqbar_ttx_ir(pp=False)["initial_states"] == [[2, -2]]
    c = np.array(ir["color_num"], dtype=float) / np.array(ir["color_denom"], dtype=float)[:, None]
    j = np.array([t[0][1] for t in ir["jamp"]])
    assert j @ c @ j == pytest.approx(2.0)          # sum over colours of |T^a_ij T^a_kl|^2 = (N^2 - 1) / 4
    p = _points(0, n=500, seed=5)
    s, t, u = _mandelstam(p)
    exact = 4 * G**4 / 9 * ((MT**2 - t) ** 2 + (MT**2 - u) ** 2 + 2 * MT**2 * s) / s**2
    np.testing.assert_allclose(omatrix.smatrix(ir, p, sm_params(), EXACT), exact, rtol=1e-11)
    assert codegen.flops_per_event(ir) == 16 * (4 * 30 + 176 + 108 + 4 + 40 + 4) + 17


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_generated_cuda_source_on_host_qqbar_ttx():
    """The emitted code of q q~ > t t~ (one event per thread and helicity-parallel), executed on the CPU."""
    import hostcheck as hc

    ir = procgen.qqbar_ttx_ir()
    lib = hc.process(ir)
    p = _points(0, n=200, seed=7)
    a_s = 0.09 + 0.05 * np.random.default_rng(9).random(200)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    ref = omatrix.smatrix(ir, p, params)
    np.testing.assert_allclose(hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF), ref, rtol=1e-12)
    np.testing.assert_allclose(hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, hp=True), ref, rtol=1e-12)
    one = hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, only_comb=6, hp=True)
    np.testing.assert_allclose(one, omatrix.matrix(ir, p, ir["helicities"][6], params), rtol=1e-11, atol=1e-300)


@pytest.mark.parametrize("kind", procgen.LIGHT_LINE_KINDS)
def test_light_line_five_point_processes(kind):
    """q q~ > t t~ g, g q > t t~ q, g q~ > t t~ q~ (the subprocesses of `p p > t t~ j` besides g g > t t~ g): BRST
    invariance of every colour flow for every helicity (Gamma_t = 0), the JAMP / colour-matrix form against the
    explicit sum over the colours of all five external partons, crossing-independent colour data."""
    ir = procgen.light_line_ttxg_ir(kind)
    assert process_ir.validate(ir)
    assert ir["ncomb"] == 32 and ir["ndiags"] == 5 and len(ir["jamp"]) == 4 and ir["mirror_initial_states"]
    assert ir["denominator"] == (36 if kind == "uux_ttxg" else 96)
    assert ir["color_num"] == [[12, 0, 4, 4], [0, 12, 4, 4], [4, 4, 12, 0], [4, 4, 0, 12]] and ir["color_denom"] == [1] * 4
    p = _points(1, n=12, seed=2)
    params = dict(sm_params(), mdl_WT=0.0)
    gl = [c["leg"] for c in ir["calls"] if c["op"] == "vxxxxx"][0]
    nonzero = 0
    for hel in ir["helicities"]:
        phys = np.max(np.abs(omatrix.matrix(ir, p, hel, params, EXACT, return_jamp=True)))
        if phys == 0.0:
            continue          # the light quark line conserves helicity
        nonzero += 1
        h = list(hel)
        h[gl] = 4             # BRST polarisation (wavefunctions_flow.py:146-152)
        assert np.max(np.abs(omatrix.matrix(ir, p, h, params, EXACT, return_jamp=True))) < 1e-11 * phys
    assert nonzero == 16
    # explicit colour sum: sum over (t, t~, O, I, a) of |sum_d tensor_d amp_d|^2 with the SU(3) matrices written out
    T, f = procgen._su3()
    TT = np.einsum("aij,bjk->abik", T, T)
    tensors = [np.einsum("abik,bol->ikola", TT, T), np.einsum("baik,bol->ikola", TT, T),
               np.einsum("bik,baol->ikola", T, TT), np.einsum("bik,abol->ikola", T, TT),
               np.einsum("bca,bik,col->ikola", f, T, T)]
    from oracle import aloha, helas

    hel = next(h for h in ir["helicities"] if np.max(np.abs(omatrix.matrix(ir, p, h, sm_params(), EXACT, return_jamp=True))) > 0)
    w, amp = {}, {}
    ext = {"vxxxxx": helas.vxxxxx, "ixxxxx": helas.ixxxxx, "oxxxxx": helas.oxxxxx}
    pr = sm_params()
    val = lambda name: 0.0 if name == "ZERO" else pr[name]
    for c in ir["calls"]:
        if "leg" in c:
            w[c["out"]] = ext[c["op"]](p[:, c["leg"]], val(c["mass"]), hel[c["leg"]], c["nsf"], EXACT)
        elif "amp" in c:
            amp[c["amp"]] = aloha.ROUTINES[c["op"]](*[w[i] for i in c["in"]], pr[c["coup"]])
        else:
            w[c["out"]] = aloha.ROUTINES[c["op"]](*[w[i] for i in c["in"]], pr[c["coup"]], val(c["mass"]), val(c["width"]))
    full = sum(t[..., None] * amp[d] for d, t in enumerate(tensors))          # (3,3,3,3,8,nevt)
    explicit = np.sum(np.abs(full) ** 2, axis=(0, 1, 2, 3, 4))
    np.testing.assert_allclose(omatrix.matrix(ir, p, hel, pr, EXACT), explicit, rtol=1e-12)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
@pytest.mark.parametrize("kind", procgen.LIGHT_LINE_KINDS)
def test_generated_cuda_source_on_host_light_line(kind):
    """The emitted code of the light-line five-point processes (both kernel flavours), executed on the CPU."""
    import hostcheck as hc

    ir = procgen.light_line_ttxg_ir(kind)
    lib = hc.process(ir)
    p = _points(1, n=60, seed=7)
    a_s = 0.09 + 0.05 * np.random.default_rng(9).random(60)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    ref = omatrix.smatrix(ir, p, params)
    np.testing.assert_allclose(hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF), ref, rtol=1e-12)
    np.testing.assert_allclose(hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, hp=True), ref, rtol=1e-12)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_colour_matrix_against_explicit_su3(irs, k):
    """The colour matrix of g g > t t~ + k g (exact Fierz reduction in procgen._trace_value) against the Gram matrix
    of the colour strings (T^a1 .. T^an)_{ij} built from the Gell-Mann matrices: an independent numerical check."""
    T, _ = procgen._su3()
    ir = irs[k]
    ng = 2 + k
    letters = "abcd"[:ng]
    strings = []
    for word in ir["color_basis"]:
        # tensor [i, j, a_leg...] with the adjoint axes in the order of the gluon legs (0, 1, 4, 5, ...)
        legs = sorted(word)
        m = T[:, :, :]   # (a, i, j)
        expr_in = ",".join(f"{letters[legs.index(g)]}{'ijklm'[q]}{'ijklm'[q + 1]}" for q, g in enumerate(word))
        tensor = np.einsum(f"{expr_in}->i{'ijklm'[ng]}{letters}", *([m] * ng))
        strings.append(tensor.reshape(-1))
    B = np.stack(strings, axis=1)
    gram = (B.conj().T @ B).real
    cm = np.array(ir["color_num"], dtype=float) / np.array(ir["color_denom"], dtype=float)[:, None]
    np.testing.assert_allclose(gram, cm, rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------ general light-line generator
def test_line_generator_reproduces_the_hand_built_processes():
    """procgen_lines (numeric colour, every diagram closed at the t~ vertex) against the hand-built q q~ > t t~ and
    five-point IRs: different call lists, same |M|^2."""
    from madflow_b200 import procgen_lines as pl

    p4, p5 = _points(0, n=40, seed=4), _points(1, n=20, seed=4)
    a = omatrix.smatrix(pl.process_ir("1_uux_ttx"), p4, sm_params(), EXACT)
    np.testing.assert_allclose(a, omatrix.smatrix(procgen.qqbar_ttx_ir(), p4, sm_params(), EXACT), rtol=1e-13)
    for kind in procgen.LIGHT_LINE_KINDS:
        ir = pl.process_ir("1_" + kind)
        assert process_ir.validate(ir) and ir["ndiags"] == 5 and len(ir["jamp"]) == 4
        ref = procgen.light_line_ttxg_ir(kind)
        assert ir["denominator"] == ref["denominator"] and ir["initial_states"] == ref["initial_states"]
        np.testing.assert_allclose(omatrix.smatrix(ir, p5, sm_params(), EXACT), omatrix.smatrix(ref, p5, sm_params(), EXACT), rtol=2e-12)


@pytest.mark.parametrize("name", ["1_uux_ttxgg", "1_gu_ttxug", "1_gux_ttxuxg", "1_gg_ttxuux"])
def test_line_generator_six_point_processes(name):
    """q q~ > t t~ g g and its crossings (36 diagrams, 12 colour flows): BRST invariance of every colour flow for every
    gluon (Gamma_t = 0), Bose symmetry of identical gluons, positive |M|^2."""
    from madflow_b200 import procgen_lines as pl

    ir = pl.process_ir(name)
    assert process_ir.validate(ir)
    assert (ir["ndiags"], len(ir["jamp"]), ir["ncomb"]) == (36, 12, 64)
    assert ir["denominator"] == {"1_uux_ttxgg": 72, "1_gu_ttxug": 96, "1_gux_ttxuxg": 96, "1_gg_ttxuux": 256}[name]
    c = np.array(ir["color_num"], dtype=float) / np.array(ir["color_denom"], dtype=float)[:, None]
    assert np.allclose(c, c.T) and np.min(np.linalg.eigvalsh(c)) > -1e-10
    p = _points(2, n=5, seed=6)
    params = dict(sm_params(), mdl_WT=0.0)
    gluons = [cl["leg"] for cl in ir["calls"] if cl["op"] == "vxxxxx"]
    checked = 0
    for hel in ir["helicities"][::5]:
        phys = np.max(np.abs(omatrix.matrix(ir, p, hel, params, EXACT, return_jamp=True)))
        if phys == 0.0:
            continue
        checked += 1
        for gl in gluons:
            h = list(hel)
            h[gl] = 4
            assert np.max(np.abs(omatrix.matrix(ir, p, h, params, EXACT, return_jamp=True))) < 1e-11 * phys
    assert checked >= 5
    me = omatrix.smatrix(ir, p, sm_params())
    assert np.all(me > 0)
    # the other organisation of the same amplitude (every diagram closed at the t~ vertex): same |M|^2, more work
    alt = pl.process_ir(name, root="tbar")
    assert alt["ndiags"] == 36 and codegen.flops_per_event(alt) > codegen.flops_per_event(ir)
    np.testing.assert_allclose(omatrix.smatrix(alt, p, sm_params()), me, rtol=1e-11)
    same = [g_ for g_ in gluons if (g_ < 2) == (gluons[0] < 2)]
    if len(same) == 2:   # two gluons on the same side of the process: exchanging their momenta changes nothing
        q = p.copy()
        q[:, same] = q[:, same[::-1]]
        np.testing.assert_allclose(omatrix.smatrix(ir, q, sm_params()), me, rtol=1e-11)


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
@pytest.mark.parametrize("name", ["1_uux_ttxgg", "1_gg_ttxuux"])
def test_generated_cuda_source_on_host_line_generator(name):
    """The helicity-parallel tables and code emitted for a six-point light-line process, executed on the CPU."""
    import hostcheck as hc
    from madflow_b200 import procgen_lines as pl

    ir = pl.process_ir(name)
    lib = hc.process(ir)
    p = _points(2, n=4, seed=8)
    a_s = 0.09 + 0.05 * np.random.default_rng(3).random(4)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    ref = omatrix.smatrix(ir, p, params)
    np.testing.assert_allclose(hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, hp=True), ref, rtol=1e-12)
    one = hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, only_comb=9, hp=True)
    np.testing.assert_allclose(one, omatrix.matrix(ir, p, ir["helicities"][9], params), rtol=1e-11, atol=1e-300)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_builtin_processes_against_the_independent_generator(irs, k):
    """g g > t t~ + k g (k = 2: the process the headline numbers are quoted on) derived a second time by
    procgen_lines -- numeric SU(3) colour tensors projected on the n! strings instead of the word algebra, diagrams
    closed at the t~ vertex instead of the centroid: same diagram and amplitude counts, identical colour matrix, same
    |M|^2 with the reference's top width."""
    from madflow_b200 import procgen_lines as pl

    a, b = pl.process_ir("1_gg_ttx" + "g" * k, root="tbar"), irs[k]
    assert a["ndiags"] == b["ndiags"] and a["denominator"] == b["denominator"]
    assert len([c for c in a["calls"] if "amp" in c]) == len([c for c in b["calls"] if "amp" in c])
    assert a["color_num"] == b["color_num"] and a["color_denom"] == b["color_denom"]
    p = _points(k, n=6, seed=12)
    np.testing.assert_allclose(omatrix.smatrix(a, p, sm_params(), EXACT), omatrix.smatrix(b, p, sm_params(), EXACT), rtol=1e-12)


@pytest.mark.parametrize("k", [2, 3])
def test_colour_reduced_plan_equals_the_diagram_list(irs, k):
    """madflow_b200/recursion.py: the plan the helicity-parallel kernels evaluate (basis currents = sums of sub-trees
    with linearly dependent colour, rows = current x basis vertex numerator) against the diagram-by-diagram oracle,
    through a numpy interpreter of the plan that uses the oracle's HELAS / ALOHA routines -- no CUDA code involved."""
    import plan_eval
    from madflow_b200 import recursion

    ir = irs[k]
    stats = recursion.plan_stats(ir["plan"])
    namps = len({c["amp"] for c in ir["calls"] if "amp" in c})
    raw_terms = sum(len(t) for t in ir["jamp"])
    assert stats["rows"] < 0.45 * namps and stats["jamp_terms"] < 0.35 * raw_terms
    assert (stats["rows"], stats["jamp_terms"]) == {2: (64, 208), 3: (388, 1768)}[k]
    # every coefficient of the plan is a unit: the kernels fold it into the coupling (+-1, +-i)
    coefs = {tuple(t["coef"]) for o in ir["plan"]["objects"] + ir["plan"]["pairs"] for t in o["terms"]}
    assert coefs <= {(1.0, 0.0), (-1.0, 0.0), (0.0, 1.0), (0.0, -1.0)}
    npts = 2 if k == 3 else 40
    p = _points(k, n=npts, seed=21)
    params = sm_params(alpha_s=0.09 + 0.05 * np.random.default_rng(5).random(npts))
    np.testing.assert_allclose(plan_eval.smatrix(ir, p, params), omatrix.smatrix(ir, p, params), rtol=1e-13)
    row = 9
    one = plan_eval.matrix(ir, p, ir["helicities"][row], params)
    ref = omatrix.matrix(ir, p, ir["helicities"][row], params)
    assert np.max(np.abs(one - ref) / np.maximum(np.abs(ref), 1e-6 * np.max(np.abs(ref)))) < 1e-9


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_diagram_list_flavour_on_host(irs, monkeypatch):
    """MADFLOW_B200_HP_REDUCE=0: the kernels evaluate the IR's diagram list (single-term objects, one row per
    amplitude) -- the A/B partner of the colour-reduced plan and the path of every IR that carries no plan."""
    import hostcheck as hc

    monkeypatch.setenv("MADFLOW_B200_HP_REDUCE", "0")
    ir = process_ir.clone(irs[2])
    ir["name"] += "_diagrams"
    info = codegen.emit_hp(ir)[4]
    monkeypatch.delenv("MADFLOW_B200_HP_REDUCE")
    red = codegen.emit_hp(irs[2])[4]
    assert not info["reduced"] and red["reduced"] and info["namps"] == 159 and red["namps"] == 64
    assert red["ntiles"] < 0.5 * info["ntiles"] and red["jamp_terms"] < 0.5 * info["jamp_terms"]
    monkeypatch.setenv("MADFLOW_B200_HP_REDUCE", "0")
    lib = hc.process(ir)
    p = _points(2, n=5, seed=13)
    a_s = 0.09 + 0.05 * np.random.default_rng(4).random(5)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    ref = omatrix.smatrix(ir, p, params)
    np.testing.assert_allclose(hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, hp=True), ref, rtol=1e-12)
    one = hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, only_comb=11, hp=True)
    # a single (possibly helicity-suppressed) row: accurate relative to the size of the whole sum
    assert np.max(np.abs(one - omatrix.matrix(ir, p, ir["helicities"][11], params)) / (ref * ir["denominator"])) < 1e-13


# ------------------------------------------------------------------------------ two light lines (four-quark processes)
FOUR_QUARK = ["1_uu_ttxuu", "1_ud_ttxud", "1_uxux_ttxuxux", "1_uxdx_ttxuxdx", "1_uux_ttxuux", "1_uux_ttxddx", "1_udx_ttxudx"]


@pytest.mark.parametrize("name", FOUR_QUARK)
def test_four_quark_processes(name):
    """The four-quark subprocesses of p p > t t~ j j: diagram counts (7, or 14 with the exchange diagrams of identical
    flavours), 6 colour flows, the two organisations of the amplitude agree, charge conjugation maps q q(') onto
    q~ q~(') with t <-> t~."""
    from madflow_b200 import procgen_lines as pl

    ir = pl.process_ir(name)
    assert process_ir.validate(ir) and len(ir["jamp"]) == 6 and ir["color_flows_independent"]
    same = name in ("1_uu_ttxuu", "1_uxux_ttxuxux", "1_uux_ttxuux")
    assert ir["ndiags"] == (14 if same else 7)
    assert ir["denominator"] == (72 if name in ("1_uu_ttxuu", "1_uxux_ttxuxux") else 36)
    p = _points(2, n=4, seed=21)
    me = omatrix.smatrix(ir, p, sm_params(), EXACT)
    assert np.all(me > 0)
    np.testing.assert_allclose(omatrix.smatrix(pl.process_ir(name, root="tbar"), p, sm_params(), EXACT), me, rtol=1e-12)
    conj = {"1_uu_ttxuu": "1_uxux_ttxuxux", "1_ud_ttxud": "1_uxdx_ttxuxdx"}.get(name)
    if conj:
        q = p.copy()
        q[:, [2, 3]] = q[:, [3, 2]]
        np.testing.assert_allclose(omatrix.smatrix(pl.process_ir(conj), q, sm_params(), EXACT), me, rtol=1e-12)


def test_identical_quarks_obey_fermi_statistics():
    """u u > t t~ u u with the two outgoing quarks at the same momentum and helicity: the colour-dressed amplitude
    sum_k JAMP_k flow_k(colours) is antisymmetric under the exchange of their colours (a '+' between the direct and the
    exchanged diagrams would make it symmetric)."""
    from madflow_b200 import procgen_lines as pl

    ir = pl.process_ir("1_uu_ttxuu")
    B, _ = pl.LineGenerator(pl.PROCESSES["1_uu_ttxuu"][0]).colour_flows()
    p5 = _points(1, n=3, seed=5)
    p = np.concatenate([p5[:, :4], p5[:, 4:5] / 2, p5[:, 4:5] / 2], axis=1)
    checked = 0
    for hel in ir["helicities"]:
        if hel[4] != hel[5]:
            continue
        J = omatrix.matrix(ir, p, hel, sm_params(), EXACT, return_jamp=True)
        if np.max(np.abs(J)) == 0.0:
            continue
        full = (B @ J).reshape(3, 3, 3, 3, 3, 3, -1)          # colours of the legs 0..5
        assert np.max(np.abs(full + np.swapaxes(full, 4, 5))) < 1e-13 * np.max(np.abs(full))
        checked += 1
    assert checked >= 4


def test_three_lines_and_a_gluon_gauge_invariance():
    """u d > t t~ u d g (64 diagrams; 18 colour flows that are linearly dependent for N = 3, so the decomposition is the
    minimum-norm one): |M|^2 vanishes for a BRST-polarised gluon."""
    from madflow_b200 import procgen_lines as pl

    ir = pl.generate_ir(["li", "mi", "to", "ti", "lo", "mo", "g"], "1_ud_ttxudg", "u d > t t~ u d g",
                        [2, 1, 6, -6, 2, 1, 21], [[2, 1]], True)
    assert ir["ndiags"] == 64 and len(ir["jamp"]) == 18 and not ir["color_flows_independent"]
    p = _points(3, n=2, seed=9)
    params = dict(sm_params(), mdl_WT=0.0)
    checked = 0
    for hel in ir["helicities"][::9]:
        phys = np.max(np.abs(omatrix.matrix(ir, p, hel, params, EXACT)))
        if phys == 0.0:
            continue
        h = list(hel)
        h[6] = 4
        assert np.max(np.abs(omatrix.matrix(ir, p, h, params, EXACT))) < 1e-20 * phys
        checked += 1
    assert checked >= 3


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_generated_cuda_source_on_host_four_quark():
    """The emitted helicity-parallel code of u u~ > t t~ u u~ (direct + exchanged diagrams), executed on the CPU."""
    import hostcheck as hc
    from madflow_b200 import procgen_lines as pl

    ir = pl.process_ir("1_uux_ttxuux")
    lib = hc.process(ir)
    p = _points(2, n=6, seed=8)
    a_s = 0.09 + 0.05 * np.random.default_rng(3).random(6)
    params = sm_params(alpha_s=a_s)
    coup = np.stack([params[c] for c in ir["couplings"]])
    np.testing.assert_allclose(hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF, hp=True), omatrix.smatrix(ir, p, params), rtol=1e-12)
