"""Parity of the CUDA path (through the C ABI) with the oracle and with the golden vectors that the
reference's own sources produced.  Needs a GPU."""
import ctypes
import math

import numpy as np
import pytest

from conftest import G, MT, WT, sm_params
from oracle import EXACT, REFERENCE, SQH_REF, aloha, helas, philox
from oracle import matrix as omatrix
from oracle import phasespace as ops
from oracle import vegas as ovegas

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

REL_ME = 1e-12  # north star: per-event |M|^2 within 1e-12 relative in FP64
# single wavefunction components on random momenta: sqrt(E+pz) and pp+pz cancel for backward-going
# particles, so a last-bit difference (FMA contraction on the GPU) is amplified by E/(E+pz)
RTOL_WF = 2e-11


@pytest.fixture(scope="module")
def mf():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    import madflow_b200  # noqa: F401
    from madflow_b200 import _runtime as rt
    from madflow_b200 import events, integrand, matrix, parameters, phasespace, vegas, wavefunctions_flow

    class NS:
        pass

    ns = NS()
    ns.rt, ns.matrix, ns.phasespace, ns.vegas, ns.wf = rt, matrix, phasespace, vegas, wavefunctions_flow
    ns.integrand, ns.parameters, ns.events = integrand, parameters, events
    return ns


def cpu(t):
    return t.detach().cpu().numpy()


def _select(mf, m, variant):
    """Choose the kernel flavour; the one-event-per-thread kernels are not compiled for long call lists."""
    try:
        m.set_variant(variant)
    except mf.rt.MadflowB200Error:
        pytest.skip(f"{variant} flavour is not compiled for {m}")
    assert m.variant == variant


# ------------------------------------------------------------------------------ HELAS
def test_wavefunctions_vs_reference_golden(mf, golden):
    g = golden("wavefunctions")
    fns = {"i": mf.wf.ixxxxx, "o": mf.wf.oxxxxx, "v": mf.wf.vxxxxx}
    for key in g.files:
        if key.startswith("p_"):
            continue
        kind, m, h, s = key.split("_")
        mass, nhel, ns = float(m[1:]), int(h[1:]), int(s[1:])
        out = cpu(fns[kind](g[f"p_m{int(mass)}"], mass, nhel, ns))
        assert out.shape == g[key].shape
        np.testing.assert_allclose(out, g[key], rtol=1e-14, atol=1e-300, err_msg=key)


def test_wavefunctions_large_random_vs_oracle(mf):
    rng = np.random.default_rng(8)
    pv = rng.normal(size=(20000, 3)) * 500
    for mass in (0.0, MT):
        p = np.concatenate([np.sqrt(np.sum(pv**2, axis=1, keepdims=True) + mass**2), pv], axis=1)
        for nhel in (-1, 1):
            for ns in (-1, 1):
                np.testing.assert_allclose(cpu(mf.wf.ixxxxx(p, mass, nhel, ns)), helas.ixxxxx(p, mass, nhel, ns), rtol=RTOL_WF, atol=1e-300)
                np.testing.assert_allclose(cpu(mf.wf.oxxxxx(p, mass, nhel, ns)), helas.oxxxxx(p, mass, nhel, ns), rtol=RTOL_WF, atol=1e-300)
                np.testing.assert_allclose(cpu(mf.wf.vxxxxx(p, mass, nhel, ns)), helas.vxxxxx(p, mass, nhel, ns), rtol=RTOL_WF, atol=1e-300)
    s = cpu(mf.wf.sxxxxx(p, -1))
    np.testing.assert_allclose(s, helas.sxxxxx(p, -1), rtol=1e-15)
    assert cpu(mf.wf.vxxxxx(np.zeros((0, 4)), 0.0, 1, 1)).shape == (6, 0)


# ------------------------------------------------------------------------------ ALOHA
def _aloha(mf, rid, ins, coup, M=0.0, W=0.0, amp=False):
    n = ins[0].shape[1]
    dev = [mf.rt.to_device(a, torch.complex128) for a in ins] + [None] * (4 - len(ins))
    out = torch.empty(n if amp else (6, n), dtype=torch.complex128, device="cuda")
    lib = mf.rt.core()
    rc = lib.mf_aloha(rid, *[mf.rt.ptr(d) for d in dev], ctypes.c_int64(n), ctypes.c_double(coup.real),
                      ctypes.c_double(coup.imag), ctypes.c_double(M), ctypes.c_double(W), mf.rt.ptr(out), mf.rt.stream_ptr())
    mf.rt.check(lib, rc)
    return cpu(out)


def test_aloha_vs_reference_golden_and_oracle(mf, golden):
    g = golden("aloha_mockup")
    F1, F2, V2, V3 = g["F1"], g["F2"], g["V2"], g["V3"]
    c10, c11, M, W = complex(g["GC_10"]), complex(g["GC_11"]), float(g["M"]), float(g["W"])
    np.testing.assert_allclose(_aloha(mf, 0, [F1, F2, V3], c11, amp=True), g["FFV1_0"], rtol=1e-13)
    np.testing.assert_allclose(_aloha(mf, 1, [F2, V3], c11, M, W), g["FFV1_1"], rtol=1e-13)
    np.testing.assert_allclose(_aloha(mf, 2, [F1, V3], c11, M, W), g["FFV1_2"], rtol=1e-13)
    np.testing.assert_allclose(_aloha(mf, 3, [V2, V3], c10), g["VVV1P0_1"], rtol=1e-13)
    rng = np.random.default_rng(3)
    V4 = rng.normal(size=V2.shape) + 1j * rng.normal(size=V2.shape)
    c12 = 1j * 1.2177**2
    np.testing.assert_allclose(_aloha(mf, 4, [V2, V3, V4], c10, amp=True), aloha.VVV1_0(V2, V3, V4, c10), rtol=1e-13)
    np.testing.assert_allclose(_aloha(mf, 5, [F1, F2], c11), aloha.FFV1P0_3(F1, F2, c11, 0.0, 0.0), rtol=1e-13)
    for rid, k in ((6, 1), (7, 3), (8, 4)):
        np.testing.assert_allclose(_aloha(mf, rid, [F1, V2, V3, V4], c12, amp=True),
                                   aloha._vvvv_0(k, F1, V2, V3, V4, c12), rtol=1e-12)
    for rid, k in ((9, 1), (10, 3), (11, 4)):
        np.testing.assert_allclose(_aloha(mf, rid, [V2, V3, V4], c12), aloha._vvvv_1(k, V2, V3, V4, c12, 0.0, 0.0), rtol=1e-12)


# ------------------------------------------------------------------------------ matrix element
@pytest.mark.parametrize("variant", ["thread", "hp"])
def test_smatrix_gg_ttx_vs_reference_golden(mf, golden, variant):
    g = golden("matrix_gg_ttx")
    m, model = mf.matrix.get_process("1_gg_ttx")
    m.set_variant(variant)
    assert m.variant == variant
    assert str(m) == "1_gg_ttx" and m.nexternal == 4 and m.ncomb == 16 and m.denominator == 256
    np.testing.assert_array_equal(np.array(m.helicities), g["helicities"])
    params = (g["params"][0], g["params"][1], np.array([g["GC_10"]]), np.array([g["GC_11"]]))
    for key in ("13tev_com", "13tev_lab", "7tev_com", "7tev_lab"):
        p = g[key + "_p"]
        out = cpu(m.smatrix(p, *params))
        np.testing.assert_allclose(out, g[key + "_smatrix"], rtol=REL_ME)
        soa = np.ascontiguousarray(np.transpose(p, (1, 2, 0)))
        np.testing.assert_array_equal(cpu(m.smatrix(soa, *params, layout="soa")), out)
        np.testing.assert_array_equal(m.smatrix_host(p, *params), out)
        scale = np.abs(g[key + "_smatrix"]) * 256
        for ic in range(16):
            one = cpu(m.matrix(p, ic, *params))
            assert np.max(np.abs(one - g[key + "_matrix"][ic].real) / scale) < REL_ME
    gs = g["run_gs"]
    out = cpu(m.smatrix(g["13tev_lab_p"], g["params"][0], g["params"][1], -gs, 1j * gs))
    np.testing.assert_allclose(out, g["run_smatrix"], rtol=REL_ME)
    assert cpu(m.smatrix(np.zeros((0, 4, 4)), *params)).shape == (0,)
    # the host-buffer entry point pipelines chunks of 2^17 events over two streams: several chunks, a ragged tail,
    # per-event couplings
    big = np.tile(g["13tev_lab_p"], (-(-300_001 // g["13tev_lab_p"].shape[0]), 1, 1))[:300_001]
    gsb = np.resize(gs, 300_001)
    ref_big = cpu(m.smatrix(big, g["params"][0], g["params"][1], -gsb, 1j * gsb))
    np.testing.assert_array_equal(m.smatrix_host(big, g["params"][0], g["params"][1], -gsb, 1j * gsb), ref_big)
    np.testing.assert_array_equal(m.smatrix_host(big, *params), cpu(m.smatrix(big, *params)))
    with pytest.raises(TypeError):
        m.smatrix(g["13tev_com_p"], 173.0)
    with pytest.raises(ValueError):
        m.smatrix(np.zeros((3, 5, 4)), *params)


@pytest.mark.parametrize("variant", ["thread", "hp"])
def test_smatrix_gg_ttx_1e5_points_and_model(mf, variant):
    """>= 1e5 RAMBO points against the oracle, couplings from Model.evaluate (frozen and running)."""
    from madflow_b200 import process_ir

    ir = process_ir.gg_ttx_pinned()
    m, model = mf.matrix.get_process("1_gg_ttx")
    m.set_variant(variant)
    x = np.random.default_rng(42).random((100_000, 10))
    p, w, x1, x2 = ops.ramboflow(x, 4, 13e3, [MT, MT], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(1).random(x.shape[0])
    ref = omatrix.smatrix(ir, lab, sm_params(alpha_s=a_s))
    out = cpu(m.smatrix(lab, *model.evaluate(a_s)))
    np.testing.assert_allclose(out, ref, rtol=REL_ME)
    model.freeze_alpha_s(0.118)
    a32 = float(np.float32(0.118))
    ref = omatrix.smatrix(ir, p, sm_params(alpha_s=a32))
    out = cpu(m.smatrix(p, *model.evaluate(None)))
    np.testing.assert_allclose(out, ref, rtol=REL_ME)
    # closed form (Gamma_t = 0): the only deviation is the reference's float32 SQH
    from test_oracle import closed_form_gg_ttx

    g = 2 * math.sqrt(math.pi * a32)
    out0 = cpu(m.smatrix(p, MT, 0.0, np.array([-g + 0j]), np.array([1j * g])))
    # the closed form itself loses digits for forward scattering (t1, t2 are differences): 1e-9 here
    np.testing.assert_allclose(out0 / closed_form_gg_ttx(p, g=g), (SQH_REF / math.sqrt(0.5)) ** 4, rtol=1e-9)


# ------------------------------------------------------------------------------ phase space
def test_rambo_and_ramboflow_vs_reference_golden(mf, golden):
    g = golden("phasespace")
    for n in range(2, 8):
        p, w = mf.phasespace.rambo(g[f"rambo{n}_x"], n, 7e3)
        assert tuple(p.shape) == (16, n, 4)
        np.testing.assert_allclose(cpu(p), g[f"rambo{n}_p"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(cpu(w), g[f"rambo{n}_w"], rtol=1e-12)
        np.testing.assert_allclose(cpu(w), ops.massless_volume(n, 7e3), rtol=1e-6)  # reference tests/test_ps.py:9-39
    p, w = mf.phasespace.rambo(g["rambo7v_x"], 7, g["rambo7v_s"])
    np.testing.assert_allclose(cpu(w), g["rambo7v_w"], rtol=1e-12)
    cases = {"tt": (4, 13e3, [MT, MT]), "m50_125": (4, 7e3, [50.0, 125.0]), "massless5": (5, 7e3, None),
             "tt7": (4, 7e3, [MT, MT])}
    for name, (n, s, ms) in cases.items():
        p, w, x1, x2 = mf.phasespace.ramboflow(g[f"rf_{name}_x"], n, s, ms)
        # n=2 massive: the start value is already the root up to rounding, so batch and per-event agree
        np.testing.assert_allclose(cpu(p), g[f"rf_{name}_p"], rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(cpu(w), g[f"rf_{name}_w"], rtol=1e-9)
        np.testing.assert_allclose(cpu(x1), g[f"rf_{name}_x1"], rtol=1e-14)
        np.testing.assert_allclose(cpu(mf.phasespace._boost_to_lab(p, x1, x2)), g[f"rf_{name}_lab"], rtol=1e-9, atol=1e-7)
    for name, (n, ms) in {"ttg": (5, [MT, MT, 0.0]), "ttgg": (6, [MT, MT, 0.0, 0.0]), "ttggg": (7, [MT, MT, 0.0, 0.0, 0.0])}.items():
        # per-event Newton == the reference evaluated on one-event batches
        p, w, x1, x2 = mf.phasespace.ramboflow(g[f"rf_{name}_x"][:16], n, 13e3, ms)
        np.testing.assert_allclose(cpu(p), g[f"rf_{name}_p_single"], rtol=1e-12, atol=1e-8)
        np.testing.assert_allclose(cpu(w), g[f"rf_{name}_w_single"], rtol=1e-12)
    p, w, x1, x2 = mf.phasespace.ramboflow(g["rf_21_x"], 3, 13e3, [91.188])
    np.testing.assert_allclose(cpu(p), g["rf_21_p"], rtol=1e-13)
    np.testing.assert_allclose(cpu(w), g["rf_21_w"], rtol=1e-12)


def test_phasespace_generator_cuts(mf, golden):
    """reference tests/test_ps.py:42-78 on the CUDA path + the golden cut selections."""
    g = golden("phasespace")
    gen = mf.phasespace.PhaseSpaceGenerator(5, 7e3, algorithm="ramboflow")
    gen.register_cut("pt", particle=3, min_val=60, max_val=300.0)
    a, w, x1, x2, idx = gen(g["psg5_x"])
    np.testing.assert_array_equal(cpu(idx), g["psg5_idx"])
    np.testing.assert_allclose(cpu(a), g["psg5_p"], rtol=1e-12, atol=1e-8)
    np.testing.assert_allclose(cpu(w), g["psg5_w"], rtol=1e-12)
    gen.clear_cuts()
    full, fw, _, _, one = gen(g["psg5_x"])
    assert int(one) == 1 and full.shape[0] == 200
    pt = cpu(gen.pt(full[:, 3, :]))
    mask = (pt > 60.0) & (pt < 300.0)
    np.testing.assert_array_equal(cpu(a), cpu(full)[mask])
    gen = mf.phasespace.PhaseSpaceGenerator(5, 13e3, [MT, MT, 0.0], com_output=False)
    for i in range(2, 5):
        gen.register_cut("pt", particle=i, min_val=30.0)
    a, w, x1, x2, idx = gen(g["psglab_x"])
    np.testing.assert_array_equal(cpu(idx)[:, 0], np.flatnonzero(g["psglab_pass"]))
    np.testing.assert_allclose(cpu(a), g["psglab_p"], rtol=1e-12, atol=1e-8)
    np.testing.assert_allclose(cpu(w), g["psglab_w"], rtol=1e-12)
    np.testing.assert_allclose(cpu(gen.mt(a[:, 2:5, :])), g["psglab_mt"], rtol=1e-11)
    gen = mf.phasespace.PhaseSpaceGenerator(4, 7e3, masses=[50.0, 125.0])
    a, *_ = gen(np.random.default_rng(3).random((100, 10)))
    m2 = cpu(mf.phasespace._invariant_mass(a))
    np.testing.assert_allclose(m2[:, 0], 0.0, atol=1e-9)
    np.testing.assert_allclose(m2[:, 2], 50.0**2, atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(m2[:, 3], 125.0**2, atol=1e-4, rtol=1e-4)
    with pytest.raises(ValueError):
        mf.phasespace.PhaseSpaceGenerator(5, 7e3, masses=[1.0])
    with pytest.raises(ValueError):
        gen.register_cut("rapidity", particle=2, min_val=0)
    with pytest.raises(ValueError):
        gen.register_cut("pt", particle=9, min_val=0)


def test_ramboflow_full_size_properties(mf):
    """BASELINE config sizes: momentum conservation and mass shells on 1e6 tt~gg points."""
    n = 1_000_000
    x = torch.rand((n, 18), dtype=torch.float64, device="cuda")
    p, w, x1, x2 = mf.phasespace.ramboflow(x, 6, 13e3, [MT, MT, 0.0, 0.0])
    tot = p[:, 2:].sum(dim=1) - p[:, 0] - p[:, 1]
    roots = p[:, 0, 0] * 2
    assert float((tot.abs().max(dim=1).values / roots).max()) < 1e-9
    m2 = mf.phasespace._invariant_mass(p)
    assert float((m2[:, 2] - MT**2).abs().max()) < 1e-3 and float(m2[:, 4].abs().max() / (13e3**2)) < 1e-12
    assert bool(torch.isfinite(w).all()) and float(w.min()) > 0


# ------------------------------------------------------------------------------ VEGAS pieces
def test_philox_and_sampling_vs_oracle(mf):
    lib = mf.rt.core()
    out = torch.empty((1000, 7), dtype=torch.float64, device="cuda")
    mf.rt.check(lib, lib.mf_philox_uniform(ctypes.c_uint64(4), ctypes.c_uint32(3), ctypes.c_uint64(2**33 + 5),
                                           ctypes.c_int64(1000), 7, mf.rt.ptr(out), mf.rt.stream_ptr()))
    np.testing.assert_array_equal(cpu(out), philox.uniforms(4, 3, 2**33 + 5, 1000, 7))
    ndim, n = 5, 4096
    grid = ovegas.refine_grid(np.random.default_rng(1).random((ndim, 50)) + 0.1, ovegas.uniform_grid(ndim))
    dg = mf.rt.to_device(grid)
    x = torch.empty((n, ndim), dtype=torch.float64, device="cuda")
    xjac = torch.empty(n, dtype=torch.float64, device="cuda")
    bins = torch.empty((ndim, n), dtype=torch.uint8, device="cuda")
    mf.rt.check(lib, lib.mf_vegas_sample(mf.rt.ptr(dg), ndim, ctypes.c_uint64(9), ctypes.c_uint32(2), ctypes.c_uint64(77),
                                         ctypes.c_int64(n), ctypes.c_double(1.0 / 12345), mf.rt.ptr(x), mf.rt.ptr(xjac),
                                         mf.rt.ptr(bins), mf.rt.stream_ptr()))
    xr, kr, wr = ovegas.map_to_grid(ovegas.confine(philox.uniforms(9, 2, 77, n, ndim)), grid)
    np.testing.assert_array_equal(cpu(bins).T, kr)
    np.testing.assert_allclose(cpu(x), xr, rtol=1e-13, atol=1e-15)  # lo + delta*(xn-k): one FMA on the GPU
    np.testing.assert_allclose(cpu(xjac), wr / 12345, rtol=1e-13)
    # accumulate + reduce + refine
    f = np.random.default_rng(3).random(n) * (np.random.default_rng(4).random(n) > 0.3)
    nb = int(lib.mf_vegas_blocks())
    partial = torch.empty(nb * (4 + ndim * 50), dtype=torch.float64, device="cuda")
    sums = torch.zeros(4 + ndim * 50, dtype=torch.float64, device="cuda")
    df = mf.rt.to_device(f)
    mf.rt.check(lib, lib.mf_vegas_accumulate(mf.rt.ptr(df), mf.rt.ptr(xjac), mf.rt.ptr(bins), ctypes.c_int64(n), ndim, 1,
                                             mf.rt.ptr(partial), nb, mf.rt.stream_ptr()))
    mf.rt.check(lib, lib.mf_vegas_reduce(mf.rt.ptr(partial), nb, ndim, 0, mf.rt.ptr(sums), mf.rt.stream_ptr()))
    a, b, c = ovegas.accumulate(f, wr / 12345, kr)
    s = cpu(sums)
    np.testing.assert_allclose(s[0], a, rtol=1e-12)
    np.testing.assert_allclose(s[1], b, rtol=1e-12)
    assert s[2] == np.count_nonzero(f) and s[3] == 0
    np.testing.assert_allclose(s[4:].reshape(ndim, 50), c, rtol=1e-11, atol=1e-300)
    mf.rt.check(lib, lib.mf_vegas_refine(mf.rt.ptr(dg), mf.rt.ptr(sums), ndim, mf.rt.stream_ptr()))
    np.testing.assert_allclose(cpu(dg), ovegas.refine_grid(c, grid), rtol=1e-10, atol=1e-14)


def test_vegasflow_generic_integrand_known_integral(mf):
    ndim, sig = 4, 0.05
    exact = (sig * math.sqrt(2 * math.pi) * math.erf(0.5 / (sig * math.sqrt(2)))) ** ndim

    def f(x, n_dim=None, weight=None):
        return torch.exp(-torch.sum((x - 0.5) ** 2, dim=1) / (2 * sig**2))

    v = mf.vegas.VegasFlow(ndim, 200_000, seed=4)
    v.compile(f)
    res, err = v.run_integration(6)
    assert abs(res - exact) < 4 * err and err / res < 2e-3
    # same stream as the oracle driver => same numbers, iteration by iteration
    ov = ovegas.Vegas(ndim, 20000, seed=11)
    ov.compile(lambda x, **_: np.exp(-np.sum((x - 0.5) ** 2, axis=1) / (2 * sig**2)))
    gv = mf.vegas.VegasFlow(ndim, 20000, seed=11)
    gv.compile(f)
    for _ in range(3):
        r0, s0 = ov.run_iteration()
        r1, s1 = gv.run_iteration()
        assert abs(r1 / r0 - 1) < 1e-9 and abs(s1 / s0 - 1) < 1e-7
    np.testing.assert_allclose(cpu(gv.divisions), ov.grid, rtol=1e-7, atol=1e-12)
    res2 = mf.vegas.vegas_wrapper(f, ndim, 3, 50_000)
    assert abs(res2[0] - exact) < 5 * res2[1]


# ------------------------------------------------------------------------------ fused integrand
@pytest.mark.parametrize("variant", ["thread", "hp"])
@pytest.mark.parametrize("pt_cut,running,lab", [(None, False, False), (30.0, False, True), (30.0, True, True)])
def test_fused_integrand_equals_separate_calls_and_oracle(mf, pt_cut, running, lab, variant):
    """One iteration of the fused kernel vs (a) the same integrand assembled from the separate C-ABI
    calls and (b) the CPU oracle's cross_section, all on the same Philox points."""
    from madflow_b200 import process_ir

    n_events = 40_000
    m, model = mf.matrix.get_process("1_gg_ttx")
    m.set_variant(variant)
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT], pt_cut=pt_cut, lab_frame=lab,
                                     running=running)
    v1 = mf.vegas.VegasFlow(10, n_events, seed=4)
    v1.compile(fi)
    r1 = v1.run_iteration()
    v2 = mf.vegas.VegasFlow(10, n_events, seed=4)
    v2.compile(fi.python_integrand())
    r2 = v2.run_iteration()
    assert abs(r1[0] / r2[0] - 1) < 1e-11 and abs(r1[1] / r2[1] - 1) < 1e-9
    np.testing.assert_allclose(cpu(v1.divisions), cpu(v2.divisions), rtol=1e-8, atol=1e-13)
    ir = process_ir.gg_ttx_pinned()
    if running:
        def a_fn(q2):
            return 0.118 / (1 + 0.118 * fi.b0 * np.log(q2 / fi.mz2))
        pf = lambda a: sm_params(alpha_s=a)
    else:
        a_fn = None
        pf = lambda a: sm_params(alpha_s=float(np.float32(0.118)))
    xs = ovegas.make_cross_section(ir, pf, 13e3, [MT, MT], pt_cut=pt_cut, lab_frame=lab, alpha_s_fn=a_fn)
    ov = ovegas.Vegas(10, n_events, seed=4)
    ov.compile(xs)
    r0 = ov.run_iteration()
    assert abs(r1[0] / r0[0] - 1) < 1e-10 and abs(r1[1] / r0[1] - 1) < 1e-8
    np.testing.assert_allclose(cpu(v1.divisions), ov.grid, rtol=1e-7, atol=1e-12)


def test_cross_section_gg_ttx_integration(mf):
    """Integrated sigma: fused path vs separate-call path within 1 sigma of the combined MC error,
    and chunked launches give the same answer as a single launch."""
    m, model = mf.matrix.get_process("1_gg_ttx")
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT], pt_cut=30.0, lab_frame=True)
    va = mf.vegas.VegasFlow(10, 200_000, seed=4)
    va.compile(fi)
    ra = va.run_integration(5, log_time=False)
    vb = mf.vegas.VegasFlow(10, 200_000, seed=1234)
    vb.compile(fi.python_integrand())
    rb = vb.run_integration(5, log_time=False)
    assert abs(ra[0] - rb[0]) < math.hypot(ra[1], rb[1]) * 3
    assert ra[1] / ra[0] < 5e-3
    vc = mf.vegas.VegasFlow(10, 200_000, seed=4, events_limit=33_333)
    vc.compile(fi)
    rc = vc.run_integration(5, log_time=False)
    assert abs(rc[0] / ra[0] - 1) < 1e-9
    from madflow_b200.utilities import one_matrix_integration

    m2, model2 = mf.matrix.get_process("1_gg_ttx")
    r = one_matrix_integration(m2, model2, out_masses=[MT, MT], n_events=50_000, n_iter=3)
    assert r[0] > 0 and r[1] / r[0] < 0.02


# ------------------------------------------------------------------------------ generated processes
@pytest.mark.parametrize("variant", ["thread", "hp"])
@pytest.mark.parametrize("name,k,npts", [("1_gg_ttxg", 1, 20001), ("1_gg_ttxgg", 2, 3001), ("1_gg_ttxggg", 3, 512)])
def test_smatrix_generated_processes_vs_oracle(mf, name, k, npts, variant):
    """g g > t t~ g, g g > t t~ g g and g g > t t~ g g g (two helicity passes, 120 colour flows contracted on
    the tensor cores): CUDA kernel vs the oracle, lab-frame RAMBO points at 13 TeV, per-event running couplings.
    The kernels evaluate the colour-reduced plan (recursion.py) for k >= 2, the oracle the diagram list.  For
    g g > t t~ g g g the oracle memoises the wavefunctions per helicity of their own legs (smatrix_recycled, equal to
    smatrix to 4e-16: tests/test_oracle.py), which makes 512 points affordable (plain: ~7 s per point)."""
    from madflow_b200 import procgen

    if k == 3 and variant == "thread":
        pytest.skip("g g > t t~ g g g only exists in the helicity-parallel flavour")
    ir = procgen.generate_ir(k)
    n = 4 + k
    m, model = mf.matrix.get_process(name)
    _select(mf, m, variant)
    assert m.nexternal == n and m.ncomb == 2**n and m.ncolor == math.factorial(2 + k)
    x = np.random.default_rng(100 + k).random((npts, 4 * (n - 2) + 2))
    p, w, x1, x2 = ops.ramboflow(x, n, 13e3, [MT, MT] + [0.0] * k, xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(1).random(npts)
    ref = (omatrix.smatrix_recycled if k == 3 else omatrix.smatrix)(ir, lab, sm_params(alpha_s=a_s))
    out = cpu(m.smatrix(lab, *model.evaluate(a_s)))
    np.testing.assert_allclose(out, ref, rtol=REL_ME)
    soa = np.ascontiguousarray(np.transpose(lab, (1, 2, 0)))
    np.testing.assert_array_equal(cpu(m.smatrix(soa, *model.evaluate(a_s), layout="soa")), out)
    # a few single helicities
    params = model.evaluate(a_s)
    op = sm_params(alpha_s=a_s)
    scale = np.abs(ref) * m.denominator
    nh = 40 if k == 3 else 500   # one helicity row of g g > t t~ g g g costs the oracle 0.06 s per point
    for ic in (0, 7, 2**n - 1):
        one = cpu(m.matrix(lab[:nh], ic, params[0], params[1], *[c[:nh] for c in params[2:]]))
        r1 = omatrix.matrix(ir, lab[:nh], ir["helicities"][ic], {kk: (v[:nh] if np.ndim(v) else v) for kk, v in op.items()})
        assert np.max(np.abs(one - r1) / scale[:nh]) < REL_ME


@pytest.mark.parametrize("variant", ["thread", "hp"])
@pytest.mark.parametrize("name,k,nev", [("1_gg_ttxg", 1, 20000), ("1_gg_ttxgg", 2, 4000), ("1_gg_ttxggg", 3, 2000)])
def test_fused_integrand_generated_processes(mf, name, k, nev, variant):
    """Fused kernel == separate C-ABI calls == oracle cross_section on the same Philox points,
    with pt > 30 GeV cuts, lab-frame momenta and the running coupling (BASELINE configs 2-4)."""
    from madflow_b200 import procgen

    if k == 3 and variant == "thread":
        pytest.skip("g g > t t~ g g g only exists in the helicity-parallel flavour")
    ir = procgen.generate_ir(k)
    n = 4 + k
    masses = [MT, MT] + [0.0] * k
    m, model = mf.matrix.get_process(name)
    _select(mf, m, variant)
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, lab_frame=True, running=True)
    v1 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v1.compile(fi)
    r1 = v1.run_iteration()
    v2 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v2.compile(fi.python_integrand())
    r2 = v2.run_iteration()
    assert abs(r1[0] / r2[0] - 1) < 1e-10 and abs(r1[1] / r2[1] - 1) < 1e-8
    assert v1.last_me_events == v2.last_me_events and 0 < v1.last_me_events <= nev
    assert nev < 1000 or v1.last_me_events < nev  # the pt cuts remove events
    xs = ovegas.make_cross_section(ir, lambda a: sm_params(alpha_s=a), 13e3, masses, pt_cut=30.0, lab_frame=True,
                                   alpha_s_fn=lambda q2: 0.118 / (1 + 0.118 * fi.b0 * np.log(q2 / fi.mz2)),
                                   smatrix_fn=omatrix.smatrix_recycled if k == 3 else None)
    ov = ovegas.Vegas(fi.n_dim, nev, seed=4)
    ov.compile(xs)
    r0 = ov.run_iteration()
    assert abs(r1[0] / r0[0] - 1) < 1e-10 and abs(r1[1] / r0[1] - 1) < 1e-8
    np.testing.assert_allclose(cpu(v1.divisions), ov.grid, rtol=1e-6, atol=1e-11)


# ------------------------------------------------------------------------------ event output
def test_event_histogram_and_unweighting_kernels(mf):
    """mf_event_histogram / mf_max_weight / mf_select_events against numpy on random events."""
    import ctypes

    rt, lib = mf.rt, mf.rt.core()
    rng = np.random.default_rng(11)
    n, nx = 200_000, 5
    mom = rng.normal(size=(n, nx, 4)) * 100.0
    mom[:, :, 0] = np.sqrt(np.sum(mom[:, :, 1:] ** 2, axis=-1) + 50.0**2)
    w1, w2 = rng.random(n) * 2.0, rng.normal(size=n)
    w1[::5] = 0.0  # empty slots
    w = w1 * w2
    d_mom, d_w1, d_w2 = (torch.as_tensor(a).cuda().contiguous() for a in (mom, w1, w2))
    p = mom[:, 3]
    pt = np.hypot(p[:, 1], p[:, 2])
    pabs = np.sqrt(pt**2 + p[:, 3] ** 2)
    obs = {"pt": pt, "eta": 0.5 * np.log((pabs + p[:, 3]) / (pabs - p[:, 3])),
           "rapidity": 0.5 * np.log((p[:, 0] + p[:, 3]) / (p[:, 0] - p[:, 3])), "energy": p[:, 0],
           "mass": np.sqrt(np.maximum(p[:, 0] ** 2 - pabs**2, 0.0))}
    ranges = {"pt": (0.0, 300.0), "eta": (-4.0, 4.0), "rapidity": (-2.0, 2.0), "energy": (50.0, 400.0), "mass": (49.013, 51.017)}   # 50 GeV must not sit on a bin edge
    for name, vals in obs.items():
        lo, hi = ranges[name]
        h = mf.events.Histogram(name, 3, lo, hi, 40)
        h.fill(d_mom, d_w1, d_w2)
        h.fill(d_mom, d_w1, d_w2)
        ref, _ = np.histogram(vals, bins=np.linspace(lo, hi, 41), weights=w)
        got = h.values(2, with_overflow=True)
        np.testing.assert_allclose(got[1:-1], ref, rtol=1e-9, atol=1e-9, err_msg=name)
        np.testing.assert_allclose(got[0] + got[-1], np.sum(w[(vals < lo) | (vals >= hi)]), rtol=1e-9, atol=1e-9)
    nb = int(lib.mf_weight_stats_blocks())
    part = torch.zeros((nb, 3), dtype=torch.float64, device="cuda")
    rt.check(lib, lib.mf_weight_stats(rt.ptr(d_w1), rt.ptr(d_w2), ctypes.c_int64(n), rt.ptr(part), nb, rt.stream_ptr()))
    dmax = part[:, 0].max()
    assert dmax.item() == np.max(np.abs(w))
    np.testing.assert_allclose(cpu(part[:, 1:].sum(dim=0)), [np.sum(np.abs(w)), np.sum(w * w)], rtol=1e-12)
    # unweighting: with wmax = max|w| every kept event has |weight| = wmax and the kept fraction is <|w|>/wmax
    cap = n
    o_mom = torch.empty((cap, nx, 4), dtype=torch.float64, device="cuda")
    o_w = torch.empty(cap, dtype=torch.float64, device="cuda")
    o_idx = torch.empty(cap, dtype=torch.int64, device="cuda")
    runs = []
    for _ in range(2):
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        rt.check(lib, lib.mf_select_events(rt.ptr(d_mom), rt.ptr(d_w1), rt.ptr(d_w2), ctypes.c_int64(n), nx,
                                           ctypes.c_double(dmax.item()), ctypes.c_uint64(7), ctypes.c_uint64(1000),
                                           rt.ptr(o_mom), rt.ptr(o_w), rt.ptr(o_idx), rt.ptr(cnt), ctypes.c_int64(cap),
                                           rt.stream_ptr()))
        k = int(cnt.item())
        order = torch.argsort(o_idx[:k])
        runs.append((cpu(o_idx[:k][order]), cpu(o_w[:k][order]), cpu(o_mom[:k][order])))
    idx, ow, om = runs[0]
    np.testing.assert_array_equal(idx, runs[1][0])  # the selection depends on (seed, index) only
    expect = np.sum(np.abs(w)) / dmax.item()
    assert abs(len(idx) - expect) < 5 * np.sqrt(expect)
    np.testing.assert_array_equal(np.abs(ow), np.full(len(idx), dmax.item()))
    np.testing.assert_array_equal(np.sign(ow), np.sign(w[idx - 1000]))
    np.testing.assert_array_equal(om, mom[idx - 1000])
    assert np.all(w[idx - 1000] != 0.0)


def test_event_sink_histograms_and_lhe(mf, tmp_path):
    """The fused integrand with an EventSink: the histograms add up to the iteration's estimate, the kept events are
    physical, and the LHE file written from them reads back."""
    from madflow_b200.lhe_writer import EventFileFlow, LheWriter

    m, model = mf.matrix.get_process("1_gg_ttxg")
    m.set_variant("hp")
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT, 0.0], pt_cut=30.0, lab_frame=True, running=True)
    v = mf.vegas.VegasFlow(fi.n_dim, 200_000, seed=4)
    v.compile(fi)
    v.run_integration(2, log_time=False)
    v.freeze_grid()
    hists = [mf.events.Histogram("pt", 2, 0.0, 300.0, 50), mf.events.Histogram("eta", 2, -4.0, 4.0, 50)]
    wmax = 20.0 * v.history[-1][0] / 200_000   # 20 x the mean weight of an iteration
    sink = mf.events.EventSink(fi, histograms=hists, unweight=True, capacity=50_000, seed=3, wmax=wmax)
    results = [v.run_iteration() for _ in range(3)]
    for h in hists:
        total = np.sum(h.values(1, with_overflow=True))
        np.testing.assert_allclose(total, sum(r[0] for r in results), rtol=1e-10)
    mom, w = sink.events()
    assert 1000 < len(w) <= 50_000 and not sink.overflowed
    assert np.all(w >= wmax) and np.mean(w == wmax) > 0.3 and sink.max_weight >= np.max(w)
    # unbiased: the kept weights add up to the integral of the three iterations, within the sampling error
    assert abs(np.sum(w) - sum(r[0] for r in results)) < 5.0 * np.sqrt(np.sum(w * w))
    np.testing.assert_allclose(np.sum(mom[:, :2], axis=1), np.sum(mom[:, 2:], axis=1), rtol=1e-9, atol=1e-6)
    np.testing.assert_allclose(mom[:, 2, 0] ** 2 - np.sum(mom[:, 2, 1:] ** 2, axis=-1), MT * MT, rtol=1e-7)
    assert np.all(np.hypot(mom[:, 2:, 1], mom[:, 2:, 2]) > 30.0)
    # ... and so does the part of the sample with pt(top) < 300 GeV compared with the histogram's in-range sum
    sel = np.hypot(mom[:, 2, 1], mom[:, 2, 2]) < 300.0
    in_range = np.sum(hists[0].values(1))   # summed over the three iterations, like the kept events
    assert abs(np.sum(w[sel]) - in_range) < 5.0 * np.sqrt(np.sum(w[sel] ** 2)) + 0.02 * in_range
    res, err, _ = mf.vegas.combine_iterations(results)
    with LheWriter(tmp_path, "run_01", no_unweight=True, pdg=m.ir["pdg"]) as lw:
        n = sink.write_lhe(lw, cross=res)
        lw.store_result((res, err))
        lw.dump_result(tmp_path / "cross_err.txt")
    back = list(EventFileFlow(tmp_path / "Events/run_01/weighted_events.lhe.gz"))
    assert len(back) == n == len(w)
    assert [p.pid for p in back[0]] == [21, 21, 6, -6, 21] and [p.status for p in back[0]] == [-1, -1, 1, 1, 1]
    np.testing.assert_allclose([[p.E, p.px, p.py, p.pz] for p in back[5]], mom[5], rtol=1e-10)
    assert back[0].wgt == pytest.approx(res, rel=1e-7)
    np.testing.assert_allclose(np.loadtxt(tmp_path / "cross_err.txt"), [res, err])


def test_pair_cuts_extension(mf):
    """Delta R and invariant-mass cuts on pairs of particles (an extension: they regulate the final-state collinear
    singularity the reference's single-particle cuts leave open): fused kernel == separate API calls == oracle."""
    from madflow_b200 import procgen

    ir = procgen.generate_ir(2)
    masses = [MT, MT, 0.0, 0.0]
    cuts = [("dr", (4, 5), 0.4, None), ("mij", (2, 3), None, 3000.0)]
    m, model = mf.matrix.get_process("1_gg_ttxgg")
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, cuts=cuts, lab_frame=True, running=True)
    nev = 3000
    v1 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v1.compile(fi)
    r1 = v1.run_iteration()
    v2 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v2.compile(fi.python_integrand())
    r2 = v2.run_iteration()
    assert abs(r1[0] / r2[0] - 1) < 1e-10 and v1.last_me_events == v2.last_me_events
    xs = ovegas.make_cross_section(ir, lambda a: sm_params(alpha_s=a), 13e3, masses, pt_cut=30.0, lab_frame=True, cuts=cuts,
                                   alpha_s_fn=lambda q2: 0.118 / (1 + 0.118 * fi.b0 * np.log(q2 / fi.mz2)))
    ov = ovegas.Vegas(fi.n_dim, nev, seed=4)
    ov.compile(xs)
    r0 = ov.run_iteration()
    assert abs(r1[0] / r0[0] - 1) < 1e-10
    # the pair cuts do remove events on top of the pt cuts
    fi0 = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, lab_frame=True, running=True)
    v0 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v0.compile(fi0)
    v0.run_iteration()
    assert v1.last_me_events < v0.last_me_events
    with pytest.raises(ValueError):
        mf.phasespace.PhaseSpaceGenerator(6, 13e3, masses).register_cut("dr", particle=4, min_val=0.4)


# ------------------------------------------------------------------------------ PDFs
@pytest.fixture(scope="module")
def toy_pdf(tmp_path_factory):
    """A synthetic lhagrid1 set (no real grid is available offline), loaded by the product and by the oracle."""
    from madflow_b200 import pdf as mpdf
    from oracle import pdf as opdf

    d = str(tmp_path_factory.mktemp("lhapdf"))
    opdf.write_toy_set(d)
    return mpdf.mkPDF("ToyPDF/0", dirname=d), opdf.GridPDF.from_set("ToyPDF/0", d)


def test_pdf_kernels_vs_oracle(mf, toy_pdf):
    """mf_pdf_xfxq2 / mf_pdf_alphasq2 (pdfflow's xfxQ2 / alphasQ2) against the oracle: cell interiors, knots,
    the subgrid threshold, frozen edges."""
    pd, og = toy_pdf
    rng = np.random.default_rng(3)
    n = 200_000
    x = 10 ** rng.uniform(-7.5, 0.0, n)
    q2 = 10 ** rng.uniform(0.0, 8.5, n)
    sg = og.subgrids[1]
    x[:60], q2[:60] = sg["x"], sg["q2"][3]
    q2[60:70] = 4.75**2
    pids = [21, 2, -1, 5, -5]
    out = cpu(pd.xfxQ2(pids, x, q2))
    assert out.shape == (n, 5)
    ref = og.xfxQ2(pids, x, q2)
    assert np.max(np.abs(out - ref) / np.max(np.abs(ref), axis=0)) < 1e-13
    np.testing.assert_array_equal(out[:60], sg["xf"][:, 3, [pd.column(p) for p in pids]])
    np.testing.assert_allclose(cpu(pd.xfxQ2_allpid(x[:1000], q2[:1000])), og.xfxQ2(og.pids, x[:1000], q2[:1000]), rtol=1e-11, atol=1e-13)
    assert cpu(pd.xfxQ2([21], x[:7], q2[:7])).shape == (7,)          # squeezed like pdfflow's
    qa = 10 ** rng.uniform(-0.5, 9.0, n)
    qa[:13] = og.as_q2
    np.testing.assert_allclose(cpu(pd.alphasQ2(qa)), og.alphasQ2(qa), rtol=1e-13)
    np.testing.assert_allclose(cpu(pd.alphasQ(np.sqrt(qa[:100]))), og.alphasQ2(qa[:100]), rtol=1e-12)


@pytest.mark.parametrize("variant", ["thread", "hp"])
@pytest.mark.parametrize("name,k,nev,fixed", [("1_gg_ttx", 0, 40000, None), ("1_gg_ttx", 0, 40000, 91.46),
                                              ("1_gg_ttxg", 1, 20000, None)])
def test_fused_integrand_with_pdf(mf, toy_pdf, name, k, nev, fixed, variant):
    """The integrand of madflow_exec.py:422-470 WITH the parton luminosity and the set's alpha_s (dynamic scale
    and `-q` fixed scale): fused kernel == separate C-ABI calls == oracle cross_section on the same Philox points."""
    from madflow_b200 import procgen, process_ir

    pd, og = toy_pdf
    ir = process_ir.gg_ttx_pinned() if k == 0 else procgen.generate_ir(k)
    masses = [MT, MT] + [0.0] * k
    m, model = mf.matrix.get_process(name)
    _select(mf, m, variant)
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, lab_frame=True,
                                     running=fixed is None, pdf=pd, fixed_scale=fixed)
    v1 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v1.compile(fi)
    r1 = v1.run_iteration()
    v2 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v2.compile(fi.python_integrand())
    r2 = v2.run_iteration()
    assert abs(r1[0] / r2[0] - 1) < 1e-10 and abs(r1[1] / r2[1] - 1) < 1e-8
    if fixed is None:
        a_fn, pf = og.alphasQ2, (lambda a: sm_params(alpha_s=a))
    else:
        a0 = float(np.float32(og.alphasQ2([fixed**2])[0]))   # Model.freeze_alpha_s rounds through float32 (parameters.py:53)
        assert fi.alpha_s == a0
        a_fn, pf = None, (lambda a: sm_params(alpha_s=a0))
    xs = ovegas.make_cross_section(ir, pf, 13e3, masses, pt_cut=30.0, lab_frame=True, alpha_s_fn=a_fn, pdf=og,
                                   fixed_q2=fixed**2 if fixed else None)
    ov = ovegas.Vegas(fi.n_dim, nev, seed=4)
    ov.compile(xs)
    r0 = ov.run_iteration()
    assert abs(r1[0] / r0[0] - 1) < 1e-10 and abs(r1[1] / r0[1] - 1) < 1e-8
    np.testing.assert_allclose(cpu(v1.divisions), ov.grid, rtol=1e-6, atol=1e-11)
    # the luminosity changes the answer: the same run without the table differs
    fi0 = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, lab_frame=True, running=fixed is None)
    v0 = mf.vegas.VegasFlow(fi.n_dim, nev, seed=4)
    v0.compile(fi0)
    assert abs(v0.run_iteration()[0] / r1[0] - 1) > 0.1


# ------------------------------------------------------------------------------ p p > t t~
def test_qqbar_ttx_smatrix_vs_oracle_and_closed_form(mf):
    """q q~ > t t~, both kernel flavours: per-event |M|^2 vs the oracle (1e-12) and the textbook closed form."""
    from madflow_b200 import procgen

    ir = procgen.qqbar_ttx_ir()
    x = np.random.default_rng(12).random((50_000, 10))
    p, w, x1, x2 = ops.ramboflow(x, 4, 13e3, [MT, MT], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(13).random(lab.shape[0])
    op = sm_params(alpha_s=a_s)
    ref = omatrix.smatrix(ir, lab, op)
    m, model = mf.matrix.get_process("1_uux_ttx")
    assert m.initial_states == [(2, -2), (4, -4), (1, -1), (3, -3)] and m.mirror_initial_states
    for variant in ("thread", "hp"):
        _select(mf, m, variant)
        out = cpu(m.smatrix(lab, *model.evaluate(a_s)))
        assert np.max(np.abs(out / ref - 1)) < REL_ME
    m.set_variant("default")
    dot = lambda a, b: a[:, 0] * b[:, 0] - np.sum(a[:, 1:] * b[:, 1:], axis=1)
    s = dot(p[:, 0] + p[:, 1], p[:, 0] + p[:, 1])
    t = dot(p[:, 0] - p[:, 2], p[:, 0] - p[:, 2])
    u = dot(p[:, 0] - p[:, 3], p[:, 0] - p[:, 3])
    g4 = (4 * np.pi * a_s) ** 2
    exact = 4 * g4 / 9 * ((MT**2 - t) ** 2 + (MT**2 - u) ** 2 + 2 * MT**2 * s) / s**2
    com = cpu(m.smatrix(p, *model.evaluate(a_s)))
    np.testing.assert_allclose(com, exact, rtol=1e-9)   # SQH is float32-rounded in the reference constants: none here (no gluon legs)


@pytest.mark.parametrize("train", [True, False])
def test_multi_process_integrand_pp_ttx(mf, toy_pdf, train):
    """`p p > t t~` = g g > t t~ + q q~ > t t~ on the same events, each with its own parton luminosity
    (madflow_exec.py:444-455): pipeline + mf_vegas_accumulate_sum == separate C-ABI calls == oracle."""
    from madflow_b200 import procgen, process_ir

    pd, og = toy_pdf
    nev = 60_000
    parts = []
    for name in procgen.MULTI_PROCESSES["p p > t t~"]:
        m, model = mf.matrix.get_process(name)
        parts.append(mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT], pt_cut=30.0, lab_frame=True,
                                                 running=True, pdf=pd))
    multi = mf.integrand.MultiProcessIntegrand(parts)
    try:
        v1 = mf.vegas.VegasFlow(10, nev, seed=4, train=train)
        v1.compile(multi)
        r1 = v1.run_iteration()
        v2 = mf.vegas.VegasFlow(10, nev, seed=4, train=train)
        v2.compile(multi.python_integrand())
        r2 = v2.run_iteration()
        assert abs(r1[0] / r2[0] - 1) < 1e-10 and abs(r1[1] / r2[1] - 1) < 1e-8
        assert v1.last_me_events == v2.last_me_events and 0 < v1.last_me_events < nev
        irs = [process_ir.gg_ttx_pinned(), procgen.qqbar_ttx_ir()]
        xss = [ovegas.make_cross_section(ir, lambda a: sm_params(alpha_s=a), 13e3, [MT, MT], pt_cut=30.0, lab_frame=True,
                                         alpha_s_fn=og.alphasQ2, pdf=og) for ir in irs]
        ov = ovegas.Vegas(10, nev, seed=4)
        if not train:
            ov.freeze_grid()
        ov.compile(lambda x, **kw: xss[0](x) + xss[1](x))
        r0 = ov.run_iteration()
        assert abs(r1[0] / r0[0] - 1) < 1e-10 and abs(r1[1] / r0[1] - 1) < 1e-8
        np.testing.assert_allclose(cpu(v1.divisions), ov.grid, rtol=1e-6, atol=1e-11)
        # each subprocess alone, through its own fused pipeline: the two add up to the sum
        singles = []
        for fi in parts:
            fi._lib.set_integrand_blocks(0)
            v = mf.vegas.VegasFlow(10, nev, seed=4, train=False)
            v.compile(fi)
            singles.append(v.run_iteration()[0])
        assert abs(sum(singles) / r1[0] - 1) < 1e-10 and min(singles) / r1[0] > 1e-3
    finally:
        multi.release()
        for fi in parts:
            fi.matrix.set_variant("default")


# ------------------------------------------------------------------------------ command line
def test_madflow_command_line_end_to_end(mf, toy_pdf, tmp_path):
    """`madflow` (scripts/madflow_exec.py:243-530) end to end: g g > t t~ --no_pdf against a direct VegasFlow run, and
    the reference's default kind of run -- p p > t t~ with a PDF set, warm-up + frozen iterations, --histograms."""
    import gzip
    import json

    from madflow_b200.scripts.madflow_exec import madflow_main

    pd, _ = toy_pdf
    try:
        args, (res, err), folder = madflow_main(["--no_pdf", "-c", "-i", "4", "--events_per_iteration", "200000",
                                                 "--madgraph_process", "g g > t t~", "-o", str(tmp_path / "a")])
        assert folder is None and res > 0 and err / res < 0.02
        m, model = mf.matrix.get_process("1_gg_ttx")
        fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT], pt_cut=30.0, lab_frame=True, running=True)
        v = mf.vegas.VegasFlow(10, 200_000, seed=4)
        v.compile(fi)
        v.run_integration(2, log_time=False)
        ref, referr = v.run_integration(2, log_time=False)
        assert abs(res - ref) < 1e-9 * ref           # same seed, same schedule (2 warm-up + 2 final): the same numbers

        args, (res, err), folder = madflow_main(["--pdf", "ToyPDF", "--pdf_dir", pd.dirname, "-c", "-i", "5", "-f", "2",
                                                 "--events_per_iteration", "200000", "--histograms",
                                                 "--madgraph_process", "p p > t t~", "-o", str(tmp_path / "b")])
        assert res > 0 and err / res < 0.02
        np.testing.assert_allclose(np.loadtxt(folder / "cross_err.txt"), [res, err], rtol=1e-6)
        hists = json.loads((folder / "histograms.json").read_text())
        pt = hists["pt_2"]
        total = sum(pt["dsigma_pb"]) + sum(pt["underflow_overflow_pb"])
        assert abs(total / res - 1) < 0.05            # the histogram is filled from the same (summed) event weights
        with gzip.open(folder / "unweighted_events.lhe.gz", "rt") as fh:
            text = fh.read()
        assert text.count("<event>") > 10 and "</LesHouchesEvent>" in text   # the reference's closing tag (lhe_writer.py:227)
    finally:
        for name in ("1_gg_ttx", "1_uux_ttx"):
            lib = mf.rt.process_lib(name)
            lib.set_variant("default")
            lib.set_integrand_blocks(0)


# ------------------------------------------------------------------------------ p p > t t~ j
def test_light_line_processes_and_pp_ttxj(mf, toy_pdf):
    """q q~ > t t~ g, g q > t t~ q, g q~ > t t~ q~: per-event |M|^2 vs the oracle in both kernel flavours (1e-12), and
    `p p > t t~ j` = the four subprocesses on the same events with their luminosities == separate calls == oracle."""
    from madflow_b200 import procgen

    pd, og = toy_pdf
    x = np.random.default_rng(21).random((20_000, 14))
    p, w, x1, x2 = ops.ramboflow(x, 5, 13e3, [MT, MT, 0.0], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(22).random(lab.shape[0])
    irs = {}
    for kind in procgen.LIGHT_LINE_KINDS:
        ir = irs["1_" + kind] = procgen.light_line_ttxg_ir(kind)
        ref = omatrix.smatrix(ir, lab, sm_params(alpha_s=a_s))
        m, model = mf.matrix.get_process("1_" + kind)
        for variant in ("thread", "hp"):
            _select(mf, m, variant)
            out = cpu(m.smatrix(lab, *model.evaluate(a_s)))
            assert np.max(np.abs(out / ref - 1)) < REL_ME, (kind, variant)
        m.set_variant("default")
    irs["1_gg_ttxg"] = procgen.generate_ir(1)
    names = procgen.MULTI_PROCESSES["p p > t t~ j"]
    nev = 30_000
    parts = []
    for name in names:
        m, model = mf.matrix.get_process(name)
        parts.append(mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT, 0.0], pt_cut=30.0, lab_frame=True,
                                                 running=True, pdf=pd))
    multi = mf.integrand.MultiProcessIntegrand(parts)
    try:
        v1 = mf.vegas.VegasFlow(14, nev, seed=4)
        v1.compile(multi)
        r1 = v1.run_iteration()
        v2 = mf.vegas.VegasFlow(14, nev, seed=4)
        v2.compile(multi.python_integrand())
        r2 = v2.run_iteration()
        assert abs(r1[0] / r2[0] - 1) < 1e-10 and abs(r1[1] / r2[1] - 1) < 1e-8
        xss = [ovegas.make_cross_section(irs[nm], lambda a: sm_params(alpha_s=a), 13e3, [MT, MT, 0.0], pt_cut=30.0,
                                         lab_frame=True, alpha_s_fn=og.alphasQ2, pdf=og) for nm in names]
        ov = ovegas.Vegas(14, nev, seed=4)
        ov.compile(lambda xr, **kw: sum(xs(xr) for xs in xss))
        r0 = ov.run_iteration()
        assert abs(r1[0] / r0[0] - 1) < 1e-10 and abs(r1[1] / r0[1] - 1) < 1e-8
        np.testing.assert_allclose(cpu(v1.divisions), ov.grid, rtol=1e-6, atol=1e-11)
    finally:
        multi.release()
        for fi in parts:
            fi.matrix.set_variant("default")


# ------------------------------------------------------------------------------ round 2: independent pins, determinism
def test_kernel_vs_second_generator(mf):
    """g g > t t~ g g: the kernel (colour-reduced plan of procgen.Generator) against the oracle interpreting the IR of the
    SECOND generator (procgen_lines: numeric SU(3) colour tensors projected on the strings, its own diagram
    enumeration) -- the two share no generator code, only the HELAS / ALOHA conventions."""
    from madflow_b200 import procgen_lines

    ir2 = procgen_lines.process_ir("1_gg_ttxgg")
    m, model = mf.matrix.get_process("1_gg_ttxgg")
    npts = 400
    x = np.random.default_rng(77).random((npts, 18))
    p, w, x1, x2 = ops.ramboflow(x, 6, 13e3, [MT, MT, 0.0, 0.0], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(2).random(npts)
    ref = omatrix.smatrix(ir2, lab, sm_params(alpha_s=a_s))
    np.testing.assert_allclose(cpu(m.smatrix(lab, *model.evaluate(a_s))), ref, rtol=REL_ME)


def test_wavefunction_components_well_conditioned(mf):
    """Component-level parity at 1e-13 where the reference's expressions are well conditioned (|pz| < 0.9 |p|: no
    cancellation in E + pz or pp + pz); RTOL_WF above only covers the forward / backward tails."""
    rng = np.random.default_rng(9)
    pv = rng.normal(size=(60000, 3)) * 400
    pv = pv[np.abs(pv[:, 2]) < 0.9 * np.linalg.norm(pv, axis=1)][:20000]
    for mass in (0.0, MT):
        p = np.concatenate([np.sqrt(np.sum(pv**2, axis=1, keepdims=True) + mass**2), pv], axis=1)
        for nhel in (-1, 1):
            for ns in (-1, 1):
                for dev_fn, ref_fn in ((mf.wf.ixxxxx, helas.ixxxxx), (mf.wf.oxxxxx, helas.oxxxxx), (mf.wf.vxxxxx, helas.vxxxxx)):
                    out, ref = cpu(dev_fn(p, mass, nhel, ns)), ref_fn(p, mass, nhel, ns)
                    scale = np.max(np.abs(ref[2:]), axis=0, keepdims=True)   # per event: the largest component
                    assert np.max(np.abs(out[2:] - ref[2:]) / scale) < 1e-13
                    np.testing.assert_array_equal(out[:2], ref[:2])


def _accumulators(mf, fi, n_events, first, count, seed=11):
    """The VEGAS accumulators (S1, S2, count, per-dimension histograms) of the events [first, first + count)."""
    v = mf.vegas.VegasFlow(fi.n_dim, n_events, seed=seed)
    v.compile(fi)
    v._sums.zero_()
    v._run_chunk_fused(first, count)
    torch.cuda.synchronize()
    return cpu(v._sums).copy()


@pytest.mark.parametrize("name,k,n_events", [("1_gg_ttx", 0, 400_000), ("1_gg_ttxgg", 2, 60_000)])
def test_integration_is_bit_reproducible_and_shard_independent(mf, name, k, n_events):
    """(i) Two runs with the same seed give bit-identical grids and results after 5 adaptive iterations (no
    floating-point atomics: csrc/vegas.cuh::warp_hist_add); (ii) the accumulators of two event shards -- what two
    ranks would feed into the all-reduce -- add up to those of the single run to 1e-13: the sample set does not depend
    on the number of GPUs."""
    m, model = mf.matrix.get_process(name)
    masses = [MT, MT] + [0.0] * k
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0 if k else None, lab_frame=True,
                                     running=bool(k))
    runs = []
    for _ in range(2):
        v = mf.vegas.VegasFlow(fi.n_dim, n_events, seed=21)
        v.compile(fi)
        res = v.run_integration(5, log_time=False)
        runs.append((cpu(v.divisions).copy(), res, list(v.history), v.last_me_events))
    np.testing.assert_array_equal(runs[0][0], runs[1][0])
    assert runs[0][1] == runs[1][1] and runs[0][2] == runs[1][2] and runs[0][3] == runs[1][3]
    full = _accumulators(mf, fi, n_events, 0, n_events)
    f0, c0 = mf.vegas.shard_events(n_events, 0, 2)
    f1, c1 = mf.vegas.shard_events(n_events, 1, 2)
    parts = _accumulators(mf, fi, n_events, f0, c0) + _accumulators(mf, fi, n_events, f1, c1)
    assert full[2] == parts[2] > 0                                      # events that reached the matrix element
    np.testing.assert_allclose(parts[:2], full[:2], rtol=1e-13)
    hist_f, hist_p = full[4:].reshape(fi.n_dim, -1), parts[4:].reshape(fi.n_dim, -1)
    np.testing.assert_allclose(hist_p, hist_f, rtol=1e-12, atol=1e-13 * full[1])
    np.testing.assert_array_equal(full, _accumulators(mf, fi, n_events, 0, n_events))   # and bit-reproducible


def _two_rank_worker(rank, world, port, name, n_events, out):
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    from madflow_b200 import integrand, matrix, vegas

    m, model = matrix.get_process(name)
    fi = integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT, 0.0], pt_cut=30.0, lab_frame=True, running=True)
    v = vegas.VegasFlow(fi.n_dim, n_events, seed=21)
    v.compile(fi)
    res = v.run_integration(3, log_time=False)
    if rank == 0:
        np.savez(out, divisions=v.divisions.cpu().numpy(), res=np.array(res), history=np.array(v.history))
    dist.destroy_process_group()


def test_two_rank_nccl_run_equals_single_rank(mf, tmp_path):
    """A real 2-rank NCCL run (one process per GPU, one all-reduce of the accumulators per iteration) reproduces the
    single-rank run: same events, results equal to 1e-12, grids to 1e-9 after 3 adaptive iterations."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    name, n_events = "1_gg_ttxg", 200_000
    out = str(tmp_path / "two_rank.npz")
    mp.spawn(_two_rank_worker, args=(2, 29571, name, n_events, out), nprocs=2, join=True)
    two = np.load(out)
    m, model = mf.matrix.get_process(name)
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT, 0.0], pt_cut=30.0, lab_frame=True, running=True)
    v = mf.vegas.VegasFlow(fi.n_dim, n_events, seed=21)
    v.compile(fi)
    res = v.run_integration(3, log_time=False)
    np.testing.assert_allclose(two["res"], np.array(res), rtol=1e-10)
    np.testing.assert_allclose(two["history"], np.array(v.history), rtol=1e-10)
    np.testing.assert_allclose(two["divisions"], cpu(v.divisions), rtol=1e-8, atol=1e-12)


def test_vegas_checkpoint_resume(mf, tmp_path):
    """save_grid(path) / load_grid(path): an integration resumed from the .npz checkpoint in a fresh integrator continues
    with the very same samples -- the results of the remaining iterations are bit-identical to the uninterrupted run."""
    m, model = mf.matrix.get_process("1_gg_ttx")
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT], lab_frame=True)
    va = mf.vegas.VegasFlow(fi.n_dim, 100_000, seed=5)
    va.compile(fi)
    va.run_integration(3, log_time=False)
    ck = str(tmp_path / "grid.npz")
    va.save_grid(ck)
    ra = [va.run_iteration() for _ in range(2)]
    vb = mf.vegas.VegasFlow(fi.n_dim, 1, seed=999)
    vb.compile(fi)
    vb.load_grid(ck)
    assert vb.iteration == 3 and vb.seed == 5 and vb.n_events == 100_000 and len(vb.history) == 3
    rb = [vb.run_iteration() for _ in range(2)]
    assert ra == rb
    np.testing.assert_array_equal(cpu(va.divisions), cpu(vb.divisions))


# ------------------------------------------------------------------------------ p p > t t~ j j
PPJJ_SIX_POINT = ["1_gg_ttxuux", "1_gu_ttxug", "1_gux_ttxuxg", "1_uux_ttxgg", "1_uu_ttxuu", "1_ud_ttxud", "1_uxux_ttxuxux",
                  "1_uxdx_ttxuxdx", "1_uux_ttxuux", "1_uux_ttxddx", "1_udx_ttxudx"]


@pytest.mark.parametrize("name", PPJJ_SIX_POINT)
def test_six_point_light_quark_processes_vs_oracle(mf, name):
    """The light-line (q q~ > t t~ g g and crossings, g g > t t~ q q~) and four-quark subprocesses of p p > t t~ j j
    (madflow_b200/procgen_lines.py) on the GPU: per-event |M|^2 against the oracle interpreting the same IR (parity
    unpinned with respect to MG5, like every generated process), running couplings, lab-frame RAMBO points."""
    from madflow_b200 import procgen_lines

    ir = procgen_lines.process_ir(name)
    m, model = mf.matrix.get_process(name)
    assert m.nexternal == 6 and m.ncomb == 64
    npts = 1500
    x = np.random.default_rng(31).random((npts, 18))
    p, w, x1, x2 = ops.ramboflow(x, 6, 13e3, [MT, MT, 0.0, 0.0], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(32).random(npts)
    ref = omatrix.smatrix(ir, lab, sm_params(alpha_s=a_s))
    out = cpu(m.smatrix(lab, *model.evaluate(a_s)))
    np.testing.assert_allclose(out, ref, rtol=REL_ME)


def test_pp_ttxjj_sums_twelve_subprocesses(mf, toy_pdf):
    """`p p > t t~ j j`: the twelve subprocess libraries on the same events with their luminosities (the loop of
    madflow_exec.py:141-155, 444-455) == the separate C-ABI calls == the oracle, on identical Philox points."""
    from madflow_b200 import procgen, procgen_lines

    pd, og = toy_pdf
    names = procgen.MULTI_PROCESSES["p p > t t~ j j"]
    assert len(names) == 12
    irs = {nm: (procgen.generate_ir(2) if nm == "1_gg_ttxgg" else procgen_lines.process_ir(nm)) for nm in names}
    masses = [MT, MT, 0.0, 0.0]
    nev = 3000
    parts = []
    for nm in names:
        m, model = mf.matrix.get_process(nm)
        parts.append(mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=masses, pt_cut=30.0, lab_frame=True,
                                                 running=True, pdf=pd))
    multi = mf.integrand.MultiProcessIntegrand(parts)
    try:
        v1 = mf.vegas.VegasFlow(18, nev, seed=4)
        v1.compile(multi)
        r1 = v1.run_iteration()
        v2 = mf.vegas.VegasFlow(18, nev, seed=4)
        v2.compile(multi.python_integrand())
        r2 = v2.run_iteration()
        assert abs(r1[0] / r2[0] - 1) < 1e-10 and abs(r1[1] / r2[1] - 1) < 1e-8
        xss = [ovegas.make_cross_section(irs[nm], lambda a: sm_params(alpha_s=a), 13e3, masses, pt_cut=30.0, lab_frame=True,
                                         alpha_s_fn=og.alphasQ2, pdf=og) for nm in names]
        ov = ovegas.Vegas(18, nev, seed=4)
        ov.compile(lambda xr, **kw: sum(xs(xr) for xs in xss))
        r0 = ov.run_iteration()
        assert abs(r1[0] / r0[0] - 1) < 1e-10 and abs(r1[1] / r0[1] - 1) < 1e-8
        np.testing.assert_allclose(cpu(v1.divisions), ov.grid, rtol=1e-6, atol=1e-11)
    finally:
        multi.release()


def test_plugin_written_vertex_routine_on_the_gpu(mf, tmp_path, monkeypatch):
    """The pyout plugin's ALOHA path on the device: a call list with an FFV2 vertex (text in MG5's C++ ALOHA format ->
    cpp_to_cuda -> codegen -> nvcc) and a coupling outside GC_10/11/12, evaluated through the C ABI against numpy."""
    import shutil

    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    import test_plugin as tp
    from madflow_b200 import codegen, process_ir
    from madgraph_plugin.PyOut_create_aloha import cpp_to_cuda
    from madgraph_plugin.PyOut_exporter import attach_aloha_routines

    ir = process_ir.gg_ttx_pinned()
    ir["name"] = "1_gg_ttx_ffv2"
    for c in ir["calls"]:
        if c["op"] == "FFV1_1":
            c["op"], c["coup"] = "FFV2_1", "GC_100"
        elif c.get("amp") == 1:
            c["op"], c["coup"] = "FFV2_0", "GC_100"
    ir["couplings"] = sorted({c["coup"] for c in ir["calls"] if "coup" in c})
    ir["coupling_defs"] = {"GC_100": [0.0, 0.44, 2]}
    attach_aloha_routines(ir, {"FFV2_0": cpp_to_cuda(tp.FFV2_0_CPP), "FFV2_1": cpp_to_cuda(tp.FFV2_1_CPP)})
    src = str(tmp_path / "proc.cu")
    open(src, "w").write(codegen.emit_process_source(ir))
    so = codegen.compile_source(src, str(tmp_path / "libmfp_1_gg_ttx_ffv2.so"))
    lib = mf.rt.ProcessLib(so)
    assert lib.variant == "thread"
    with pytest.raises(mf.rt.MadflowB200Error):
        lib.set_variant("hp")
    npts = 5000
    x = np.random.default_rng(3).random((npts, 10))
    p, w, x1, x2 = ops.ramboflow(x, 4, 13e3, [MT, MT], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(4).random(npts)
    params = sm_params(alpha_s=a_s)
    params["GC_100"] = 0.44j * (2.0 * np.sqrt(np.pi * a_s)) ** 2
    monkeypatch.setitem(aloha.ROUTINES, "FFV2_0", tp._np_ffv2_0)
    monkeypatch.setitem(aloha.ROUTINES, "FFV2_1", tp._np_ffv2_1)
    ref = omatrix.smatrix(ir, lab, params)
    coup = torch.as_tensor(np.stack([params[c] for c in ir["couplings"]])).cuda().contiguous()
    out = torch.empty(npts, dtype=torch.float64, device="cuda")
    lib.smatrix(torch.as_tensor(lab).cuda().contiguous(), 0, npts, [MT, WT], coup, 1, SQH_REF, out)
    np.testing.assert_allclose(cpu(out), ref, rtol=REL_ME)


def test_packed_units_equal_the_table_driven_units_on_the_gpu(mf, tmp_path, monkeypatch):
    """g g > t t~ g g: the default library (packed units: warp trips of one class, merged four-gluon terms) against a
    library of the same process built with the table-driven unit routine (MADFLOW_B200_HP_SLU=0, the A/B partner and the
    path of every other process) on 4000 points with running couplings: equal to 1e-13, both within 1e-12 of the oracle
    (recycled wavefunctions); one helicity row as well."""
    import shutil

    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    from madflow_b200 import codegen, procgen

    ir = procgen.generate_ir(2)
    monkeypatch.setenv("MADFLOW_B200_HP_SLU", "0")
    src = str(tmp_path / "proc_tab.cu")
    text = codegen.emit_process_source(ir)
    monkeypatch.delenv("MADFLOW_B200_HP_SLU")
    assert "HP_SLU = false" in text and "HP_SLU = true" in codegen.emit_process_source(ir)
    open(src, "w").write(text)
    tab = mf.rt.ProcessLib(codegen.compile_source(src, str(tmp_path / "libmfp_1_gg_ttxgg_tab.so")))
    dflt = mf.rt.ProcessLib(codegen.lib_path(ir))
    npts = 4000
    x = np.random.default_rng(31).random((npts, 18))
    p, w, x1, x2 = ops.ramboflow(x, 6, 13e3, [MT, MT, 0.0, 0.0], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(32).random(npts)
    params = sm_params(alpha_s=a_s)
    coup = torch.as_tensor(np.stack([params[c] for c in ir["couplings"]])).cuda().contiguous()
    mom = torch.as_tensor(lab).cuda().contiguous()
    outs = []
    for lib in (dflt, tab):
        out = torch.empty(npts, dtype=torch.float64, device="cuda")
        lib.smatrix(mom, 0, npts, [MT, WT], coup, 1, SQH_REF, out)
        outs.append(cpu(out))
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-13)
    np.testing.assert_allclose(outs[0], omatrix.smatrix_recycled(ir, lab, params), rtol=REL_ME)
    rows = []
    for lib in (dflt, tab):
        out = torch.empty(npts, dtype=torch.float64, device="cuda")
        lib.smatrix(mom, 0, npts, [MT, WT], coup, 1, SQH_REF, out, only_comb=37)
        rows.append(cpu(out))
    np.testing.assert_allclose(rows[0], rows[1], rtol=1e-11, atol=1e-300)


def test_one_matrix_integration_with_pdf(mf, toy_pdf):
    """utilities.one_matrix_integration(pdf=, flavours=): the reference's regression harness (utilities.py:42-90,
    tests/test_integration.py) -- luminosity of the given flavour at the fixed scale q, frozen couplings, COM momenta, no
    cuts -- fused path == separate calls, and the first iteration == the oracle on the same Philox points.  The
    reference's 103.4 pb needs the NNPDF31 grid, which is not available offline; this runs on the synthetic set."""
    from madflow_b200 import process_ir
    from madflow_b200.utilities import one_matrix_integration

    pd, og = toy_pdf
    m, model = mf.matrix.get_process("1_gg_ttx")
    ra = one_matrix_integration(m, model, pdf=pd, flavours=(0,), out_masses=[MT, MT], n_events=40_000, n_iter=1, seed=4)
    rb = one_matrix_integration(m, model, pdf=pd, flavours=(0,), out_masses=[MT, MT], n_events=40_000, n_iter=1, seed=4, fused=False)
    assert abs(ra[0] / rb[0] - 1) < 1e-10
    ir = process_ir.gg_ttx_pinned()
    a32 = float(np.float32(0.118))
    xs = ovegas.make_cross_section(ir, lambda a: sm_params(alpha_s=a32), 7e3, [MT, MT], lab_frame=False, pdf=og, fixed_q2=91.46**2)
    ov = ovegas.Vegas(10, 40_000, seed=4)
    ov.compile(xs)
    r0 = ov.run_integration(1)
    assert abs(ra[0] / r0[0] - 1) < 1e-10
    with pytest.raises(ValueError):
        one_matrix_integration(m, model, pdf=pd, out_masses=[MT, MT])


def test_unweighting_threshold_frozen_from_the_weight_spectrum(mf):
    """EventSink(collect_only=True) gathers the weight statistics of one iteration, freeze_threshold() puts ONE threshold
    where the events above it carry `tail_share` of sum |w|; afterwards the kept sample reproduces the iteration's
    estimate (sum of kept weights, each standing for max(|w|, threshold)) within the sampling error and the events above
    the threshold keep their own weight."""
    m, model = mf.matrix.get_process("1_gg_ttx")
    fi = mf.integrand.FusedIntegrand(m, model, sqrts=13e3, masses=[MT, MT], lab_frame=True)
    m.set_variant("hp")
    try:
        v = mf.vegas.VegasFlow(fi.n_dim, 300_000, seed=9)
        v.compile(fi)
        v.run_integration(4, log_time=False)
        sink = mf.events.EventSink(fi, unweight=True, capacity=300_000, seed=3, collect_only=True, tail_share=0.1)
        v.run_iteration()
        wmax = sink.freeze_threshold()
        assert 0.0 < wmax <= sink.max_weight
        spec, edges = cpu(sink._wspec), cpu(sink._wedges)
        share = spec[edges[:-1] >= wmax * (1 - 1e-12)].sum() / spec.sum()
        assert share <= 0.1 + 1e-12                              # the threshold sits at the 10 % tail, to bin resolution
        assert spec[edges[:-1] >= wmax / 2 ** 0.25 * (1 - 1e-12)].sum() / spec.sum() > 0.1
        v.freeze_grid()
        res, sigma = v.run_iteration()
        assert sink.wmax == wmax                                 # one threshold for the whole sample
        mom, w = sink.events()
        assert len(w) > 2000
        # unbiased: every kept event stands for the threshold (or its own larger weight): the sum is the iteration's estimate
        est = np.sum(np.sign(w) * np.maximum(np.abs(w), wmax))
        assert abs(est / res - 1) < 6 * sigma / res + 4 / np.sqrt(len(w))
        frac, wshare = sink.overweight()
        assert 0.0 < frac < 0.3 and 0.0 < wshare < 0.3
    finally:
        m.set_variant("default")
        fi.event_sink = None


def test_smatrix_pinned_host_pipeline(mf):
    """Matrix.smatrix_pinned (pinned host buffers, chunks on two streams) returns exactly what smatrix returns."""
    m, model = mf.matrix.get_process("1_gg_ttxg")
    n = 70_001
    x = np.random.default_rng(12).random((n, 14))
    p, w, x1, x2 = ops.ramboflow(x, 5, 13e3, [MT, MT, 0.0], xfactor="converged")
    lab = ops.boost_to_lab(p, x1, x2)
    a_s = 0.09 + 0.05 * np.random.default_rng(13).random(n)
    params = model.evaluate(a_s)
    ref = cpu(m.smatrix(lab, *params))
    npar = len(m.param_names)
    h_ps = torch.as_tensor(lab).pin_memory()
    h_c = [c.cpu().pin_memory() for c in params[npar:]]
    out = m.smatrix_pinned(h_ps, *params[:npar], *h_c, chunk=1 << 14)
    assert out.is_pinned() and out.device.type == "cpu"
    np.testing.assert_array_equal(out.numpy(), ref)
    with pytest.raises(ValueError):
        m.smatrix_pinned(torch.as_tensor(lab), *params[:npar], *h_c)      # not pinned
