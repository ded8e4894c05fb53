"""The oracle against (1) golden vectors produced by the reference's own sources
(tests/golden/make_golden.py) and (2) the reference's known-answer tests.  CPU only."""
import math

import numpy as np
import pytest

from conftest import G, MT, WT, sm_params
from madflow_b200 import process_ir
from oracle import EXACT, PI_REF, REFERENCE, SQH_REF, aloha, helas, matrix as omatrix, model, philox
from oracle import phasespace as ops
from oracle import vegas as ovegas


def test_constants_are_the_references(golden):
    g = golden("phasespace")
    assert REFERENCE.SQH == float(g["const_SQH"]) == 0.7071067690849304
    assert REFERENCE.PI == float(g["const_PI"])
    assert REFERENCE.ACC == float(g["const_ACC"])
    assert REFERENCE.GEV2PB == 389379360.0


def test_helas_bitwise_vs_reference(golden):
    g = golden("wavefunctions")
    fns = {"i": helas.ixxxxx, "o": helas.oxxxxx, "v": helas.vxxxxx}
    n = 0
    for key in g.files:
        if key.startswith("p_"):
            continue
        kind, m, h, s = key.split("_")
        mass, nhel, ns = float(m[1:]), int(h[1:]), int(s[1:])
        out = fns[kind](g[f"p_m{int(mass)}"], mass, nhel, ns)
        np.testing.assert_array_equal(out, g[key], err_msg=key)
        n += 1
    assert n == 30


def test_aloha_pinned_bitwise(golden):
    g = golden("aloha_mockup")
    c10, c11, M, W = complex(g["GC_10"]), complex(g["GC_11"]), float(g["M"]), float(g["W"])
    np.testing.assert_array_equal(aloha.FFV1_0(g["F1"], g["F2"], g["V3"], c11), g["FFV1_0"])
    np.testing.assert_array_equal(aloha.FFV1_1(g["F2"], g["V3"], c11, M, W), g["FFV1_1"])
    np.testing.assert_array_equal(aloha.FFV1_2(g["F1"], g["V3"], c11, M, W), g["FFV1_2"])
    np.testing.assert_array_equal(aloha.VVV1P0_1(g["V2"], g["V3"], c10, 0.0, 0.0), g["VVV1P0_1"])


def test_aloha_derived_routines_are_consistent(golden):
    """Unpinned routines: amplitude = (off-shell current without propagator) . (third leg)."""
    g = golden("aloha_mockup")
    F1, F2, V2, V3 = g["F1"], g["F2"], g["V2"], g["V3"]
    rng = np.random.default_rng(5)
    V4 = rng.normal(size=V2.shape) + 1j * rng.normal(size=V2.shape)
    c = 0.3 - 0.7j
    # FFV1P0_3 contracted with V3 reproduces FFV1_0
    cur = aloha.FFV1P0_3(F1, F2, c, 0.0, 0.0)
    P = aloha._mom(cur, -1.0)
    p2 = P[0] ** 2 - P[1] ** 2 - P[2] ** 2 - P[3] ** 2
    np.testing.assert_allclose(aloha._vdot(cur, V3) * p2, aloha.FFV1_0(F1, F2, V3, c), rtol=1e-12)
    # VVV1P0_1 contracted with a third vector whose momentum closes the vertex reproduces VVV1_0
    V1 = V4.copy()
    V1[0], V1[1] = -(V2[0] + V3[0]), -(V2[1] + V3[1])
    cur = aloha.VVV1P0_1(V2, V3, c, 0.0, 0.0)
    P = aloha._mom(cur, -1.0)
    p2 = P[0] ** 2 - P[1] ** 2 - P[2] ** 2 - P[3] ** 2
    np.testing.assert_allclose(aloha._vdot(cur, V1) * p2, aloha.VVV1_0(V1, V2, V3, c), rtol=1e-11)
    for k in (1, 3, 4):
        cur = aloha._vvvv_1(k, V2, V3, V4, c, 0.0, 0.0)
        P = aloha._mom(cur, -1.0)
        p2 = P[0] ** 2 - P[1] ** 2 - P[2] ** 2 - P[3] ** 2
        np.testing.assert_allclose(aloha._vdot(cur, F1) * p2, aloha._vvvv_0(k, F1, V2, V3, V4, c), rtol=1e-11)


def test_matrix_gg_ttx_vs_reference(golden):
    g = golden("matrix_gg_ttx")
    ir = process_ir.gg_ttx_pinned()
    assert process_ir.validate(ir)
    np.testing.assert_array_equal(np.array(ir["helicities"], dtype=float), g["helicities"])
    assert ir["denominator"] == float(g["denominator"])
    params = {"mdl_MT": float(g["params"][0]), "mdl_WT": float(g["params"][1]),
              "GC_10": complex(g["GC_10"]), "GC_11": complex(g["GC_11"])}
    for key in ("13tev_com", "13tev_lab", "7tev_com", "7tev_lab"):
        p = g[key + "_p"]
        np.testing.assert_allclose(omatrix.smatrix(ir, p, params), g[key + "_smatrix"], rtol=2e-15)
        for ic in (0, 3, 9, 15):
            np.testing.assert_allclose(omatrix.matrix(ir, p, ir["helicities"][ic], params),
                                       g[key + "_matrix"][ic].real, rtol=1e-13, atol=1e-16)
    gs = g["run_gs"]
    np.testing.assert_allclose(omatrix.smatrix(ir, g["13tev_lab_p"], dict(params, GC_10=-gs, GC_11=1j * gs)),
                               g["run_smatrix"], rtol=2e-15)


def closed_form_gg_ttx(p, g=G, mt=MT):
    """spin/colour averaged |M|^2(gg -> tt~), Gamma_t = 0."""
    def dot(a, b):
        return a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2] - a[:, 3] * b[:, 3]
    s = dot(p[:, 0] + p[:, 1], p[:, 0] + p[:, 1])
    t = dot(p[:, 0] - p[:, 2], p[:, 0] - p[:, 2])
    u = dot(p[:, 0] - p[:, 3], p[:, 0] - p[:, 3])
    t1, t2, rho = (mt**2 - t) / s, (mt**2 - u) / s, 4 * mt**2 / s
    return g**4 * (1 / (6 * t1 * t2) - 3 / 8) * (t1**2 + t2**2 + rho - rho**2 / (4 * t1 * t2))


def test_gg_ttx_closed_form():
    ir = process_ir.gg_ttx_pinned()
    x = np.random.default_rng(1).random((500, 10))
    p, _, _, _ = ops.ramboflow(x, 4, 13e3, [MT, MT], const=EXACT, xfactor="converged")
    params = dict(sm_params(), mdl_WT=0.0)
    exact = omatrix.smatrix(ir, p, params, const=EXACT)
    np.testing.assert_allclose(exact, closed_form_gg_ttx(p), rtol=5e-12)
    # the reference's float32 SQH shifts every gluon polarisation: (SQH32/SQH64)^(2*ngluon)
    ref = omatrix.smatrix(ir, p, params, const=REFERENCE)
    np.testing.assert_allclose(ref / exact, (SQH_REF / math.sqrt(0.5)) ** 4, rtol=1e-13)


def test_gg_ttx_gauge_invariance():
    """BRST check (wavefunctions_flow.py:146-152): a gluon polarisation replaced by its momentum
    (nhel = 4) makes every JAMP vanish relative to the size of the individual amplitudes."""
    ir = process_ir.gg_ttx_pinned()
    x = np.random.default_rng(2).random((64, 10))
    p, _, _, _ = ops.ramboflow(x, 4, 13e3, [MT, MT], const=EXACT, xfactor="converged")
    params = dict(sm_params(), mdl_WT=0.0)
    phys = omatrix.matrix(ir, p, [1, -1, 1, -1], params, return_jamp=True)
    brst = omatrix.matrix(ir, p, [4, -1, 1, -1], params, return_jamp=True)
    assert np.max(np.abs(brst)) < 1e-9 * np.max(np.abs(phys))


# ------------------------------------------------------------------------------- phase space
def test_phasespace_vs_reference(golden):
    g = golden("phasespace")
    for n in range(2, 8):
        p, w = ops.rambo(g[f"rambo{n}_x"], n, 7e3)
        np.testing.assert_array_equal(p, g[f"rambo{n}_p"])
        np.testing.assert_allclose(w, g[f"rambo{n}_w"], rtol=1e-13)
    p, w = ops.rambo(g["rambo7v_x"], 7, g["rambo7v_s"])
    np.testing.assert_array_equal(p, g["rambo7v_p"])
    cases = {"tt": (4, 13e3, [MT, MT]), "ttg": (5, 13e3, [MT, MT, 0.0]), "ttgg": (6, 13e3, [MT, MT, 0.0, 0.0]),
             "ttggg": (7, 13e3, [MT, MT, 0.0, 0.0, 0.0]), "m50_125": (4, 7e3, [50.0, 125.0]),
             "massless5": (5, 7e3, None), "tt7": (4, 7e3, [MT, MT])}
    for name, (n, s, m) in cases.items():
        p, w, x1, x2 = ops.ramboflow(g[f"rf_{name}_x"], n, s, m)  # xfactor="reference": batch semantics
        np.testing.assert_array_equal(p, g[f"rf_{name}_p"], err_msg=name)
        np.testing.assert_allclose(w, g[f"rf_{name}_w"], rtol=1e-13)
        np.testing.assert_array_equal(ops.boost_to_lab(p, x1, x2), g[f"rf_{name}_lab"])
        if m is not None and n > 4:
            # per-event convergence == the reference run on one-event batches
            pc, wc, _, _ = ops.ramboflow(g[f"rf_{name}_x"][:16], n, s, m, xfactor="converged")
            np.testing.assert_array_equal(pc, g[f"rf_{name}_p_single"])
            np.testing.assert_allclose(wc, g[f"rf_{name}_w_single"], rtol=1e-13)
    p, w, x1, x2 = ops.ramboflow(g["rf_21_x"], 3, 13e3, [91.188])
    np.testing.assert_array_equal(p, g["rf_21_p"])
    np.testing.assert_allclose(w, g["rf_21_w"], rtol=1e-14)


def test_phasespace_generator_cuts_vs_reference(golden):
    g = golden("phasespace")
    gen = ops.PhaseSpaceGenerator(5, 7e3)
    gen.register_cut("pt", particle=3, min_val=60, max_val=300.0)
    a, w, x1, x2, idx = gen(g["psg5_x"])
    np.testing.assert_array_equal(a, g["psg5_p"])
    np.testing.assert_array_equal(idx, g["psg5_idx"])
    gen = ops.PhaseSpaceGenerator(5, 13e3, [MT, MT, 0.0], com_output=False, xfactor="converged")
    for i in range(2, 5):
        gen.register_cut("pt", particle=i, min_val=30.0)
    a, w, x1, x2, idx = gen(g["psglab_x"])
    np.testing.assert_array_equal(idx[:, 0], np.flatnonzero(g["psglab_pass"]))
    np.testing.assert_allclose(a, g["psglab_p"], rtol=1e-13, atol=1e-9)
    np.testing.assert_allclose(w, g["psglab_w"], rtol=1e-13)
    np.testing.assert_allclose(ops.mt(a[:, 2:5, :]), g["psglab_mt"], rtol=1e-12)


def test_rambo_volume_reference_kat():
    """reference tests/test_ps.py:9-39: massless weight == analytic volume to 1e-6."""
    rng = np.random.default_rng(0)
    for n in range(2, 8):
        p, w = ops.rambo(rng.random((3, 4 * n)), n, 7e3)
        assert p.shape == (3, n, 4)
        np.testing.assert_allclose(w, ops.massless_volume(n, 7e3), rtol=1e-6)
        p, w = ops.rambo(rng.random((3, 4 * n)), n, 7e3, const=EXACT)
        np.testing.assert_allclose(w, ops.massless_volume(n, 7e3), rtol=1e-13)
    sq = rng.random(13) * 7e3
    _, w = ops.rambo(rng.random((13, 28)), 7, sq)
    np.testing.assert_allclose(w, ops.massless_volume(7, sq), rtol=1e-6)


def test_fourmomenta_reference_kat():
    """reference tests/test_ps.py:68-78."""
    gen = ops.PhaseSpaceGenerator(4, 7e3, masses=[50.0, 125.0])
    a, w, x1, x2, idx = gen(np.random.default_rng(3).random((100, 10)))
    np.testing.assert_allclose(ops.invariant_mass2(a[:, 0:1, :]), 0.0, atol=1e-9)
    np.testing.assert_allclose(ops.invariant_mass2(a[:, 2:3, :]), 50.0**2, atol=1e-4, rtol=1e-4)
    np.testing.assert_allclose(ops.invariant_mass2(a[:, 3:4, :]), 125.0**2, atol=1e-4, rtol=1e-4)
    # momentum conservation in the converged mode
    a, *_ = ops.ramboflow(np.random.default_rng(3).random((100, 14)), 5, 13e3, [MT, MT, 0.0], xfactor="converged")
    np.testing.assert_allclose(np.sum(a[:, 2:], axis=1), a[:, 0] + a[:, 1], atol=1e-6)


def test_model_vs_reference(golden):
    m = golden("model")
    c = model.sm_qcd_couplings(m["alpha_s"])
    for k in ("GC_10", "GC_11", "GC_12"):
        np.testing.assert_array_equal(c[k], m[k])
    c = model.sm_qcd_couplings([model.frozen_alpha_s(0.118)])
    for k in ("GC_10", "GC_11", "GC_12"):
        np.testing.assert_array_equal(c[k], m["frozen_" + k])


# ------------------------------------------------------------------------------- RNG / VEGAS
def test_philox_known_answers():
    """Random123 known-answer vectors for philox4x32-10."""
    def one(c, k):
        r = philox.philox4x32_10([np.array([v], dtype=np.uint32) for v in c], k)
        return [int(x[0]) for x in r]
    assert one((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert one((f, f, f, f), (f, f)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert one((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    u = philox.uniforms(4, 0, 0, 1000, 10)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    # chunking / sharding independence
    np.testing.assert_array_equal(philox.uniforms(4, 2, 600, 400, 10), philox.uniforms(4, 2, 0, 1000, 10)[600:])


def test_vegas_gaussian_known_integral():
    ndim, sig = 4, 0.05
    exact = (sig * math.sqrt(2 * math.pi) * math.erf(0.5 / (sig * math.sqrt(2)))) ** ndim

    def f(x, **_):
        return np.exp(-np.sum((x - 0.5) ** 2, axis=1) / (2 * sig**2))

    v = ovegas.Vegas(ndim, 20000, seed=4)
    v.compile(f)
    res, err = v.run_integration(6)
    assert abs(res - exact) < 4 * err and err / res < 0.01
    # refinement keeps edges monotone with fixed ends
    assert np.all(np.diff(v.grid, axis=1) > 0) and np.all(v.grid[:, 0] == 0) and np.all(v.grid[:, -1] == 1)
    # the adapted grid concentrates bins near the peak
    assert np.min(np.diff(v.grid, axis=1)) < 0.2 / 50


def test_vegas_map_weight_is_jacobian():
    grid = ovegas.refine_grid(np.random.default_rng(1).random((3, 50)) + 0.1, ovegas.uniform_grid(3))
    u = ovegas.confine(np.random.default_rng(2).random((200000, 3)))
    x, k, w = ovegas.map_to_grid(u, grid)
    assert abs(np.mean(w) - 1.0) < 0.02          # E[jacobian] = volume of the unit cube
    assert np.all((x >= 0) & (x <= 1)) and k.min() >= 0 and k.max() <= 49


def test_pdf_interpolation_properties(tmp_path):
    """oracle/pdf.py (LHAPDF log-bicubic + AlphaS_Ipol; pdfflow and every real grid are absent, so parity is
    unpinned): the knots are reproduced exactly, a quadratic in (log x, log Q2) on uniform log grids is
    reproduced to rounding, the interpolant is continuous across cells, a threshold value belongs to the upper
    subgrid, the synthetic shapes are recovered to the accuracy of the coarse grid."""
    from oracle import pdf as opdf

    opdf.write_toy_set(str(tmp_path))
    g = opdf.GridPDF.from_set("ToyPDF/0", str(tmp_path))
    assert g.pids == [-5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 21] and len(g.subgrids) == 2
    for sg in g.subgrids:
        X, Q2 = np.meshgrid(sg["x"], sg["q2"][:-1], indexing="ij")
        got = g.xfxQ2([21, -3], X.ravel(), Q2.ravel())
        ref = sg["xf"][:, :-1][:, :, [10, 2]].reshape(-1, 2)
        np.testing.assert_array_equal(got, ref)
    # the threshold Q = 4.75 is the first knot of the upper subgrid
    lo, hi = g.subgrids
    np.testing.assert_array_equal(g.xfxQ2([21], hi["x"], np.full(60, 4.75**2))[:, 0], hi["xf"][:, 0, 10])
    # quadratic polynomial in (log x, log Q2) on uniform log grids: exact
    lx, lq = np.linspace(-9.0, 0.0, 19), np.linspace(1.0, 12.0, 12)
    poly = lambda a, b: 1.0 + 0.3 * a - 0.05 * a * a + 0.2 * b + 0.01 * b * b + 0.02 * a * b
    sub = dict(x=np.exp(lx), q2=np.exp(lq), pids=[21], xf=poly(lx[:, None], lq[None, :])[:, :, None])
    gq = opdf.GridPDF({"AlphaS_Qs": [], "AlphaS_Vals": []}, [sub])
    rng = np.random.default_rng(0)
    a, b = rng.uniform(-8.5, -0.5, 2000), rng.uniform(2.0, 11.0, 2000)   # away from the one-sided edge cells
    np.testing.assert_allclose(gq.xfxQ2([21], np.exp(a), np.exp(b))[:, 0], poly(a, b), rtol=1e-12)
    # continuity across a cell boundary in x and in Q2
    xk, qk = hi["x"][20], hi["q2"][4]
    eps = 1e-9
    v = g.xfxQ2([21], [xk * (1 - eps), xk * (1 + eps), 1e-3, 1e-3], [1e4, 1e4, qk * (1 - eps), qk * (1 + eps)])[:, 0]
    assert abs(v[0] / v[1] - 1) < 1e-7 and abs(v[2] / v[3] - 1) < 1e-7
    x = 10 ** rng.uniform(-4, -0.3, 1000)
    q2 = 10 ** rng.uniform(np.log10(30.0), 7, 1000)
    shape = 3.0 * x**-0.25 * (1 - x) ** 5 * (1 + 0.3 * np.log(np.sqrt(q2) / 1.65))
    np.testing.assert_allclose(g.xfxQ2([0], x, q2)[:, 0], shape, rtol=5e-4)   # pid 0 = gluon
    # alpha_s: knots exact (the upper subgrid owns the threshold), monotonic, frozen above, power law below
    np.testing.assert_allclose(g.alphasQ2(g.as_q2[[0, 1, 2, 4, 8, 12]]), g.as_vals[[0, 1, 2, 4, 8, 12]], rtol=1e-15)
    aa = g.alphasQ2(10 ** np.linspace(0.5, 7.9, 400))
    assert np.all(np.diff(aa) < 0)
    assert g.alphasQ2([1e9])[0] == g.as_vals[-1] and g.alphasQ2([1.0])[0] > g.as_vals[0]
    assert abs(g.alphasQ2([91.1876**2])[0] - 0.118) < 1e-12


def test_recycled_helicity_sum_equals_the_plain_one():
    """oracle.matrix.smatrix_recycled (wavefunctions memoised per helicity of their own legs, JAMPs as one matrix
    product) is the same sum as smatrix: every value comes from the same routine with the same inputs."""
    from conftest import sm_params
    from madflow_b200 import procgen
    from oracle import matrix as om
    from oracle import phasespace as ops

    for k, npts in ((1, 60), (2, 12)):
        ir = procgen.generate_ir(k)
        n = 4 + k
        x = np.random.default_rng(40 + k).random((npts, 4 * (n - 2) + 2))
        p, _, x1, x2 = ops.ramboflow(x, n, 13e3, [173.0, 173.0] + [0.0] * k, xfactor="converged")
        p = ops.boost_to_lab(p, x1, x2)
        params = sm_params(alpha_s=0.09 + 0.05 * np.random.default_rng(1).random(npts))
        a, b = om.smatrix(ir, p, params), om.smatrix_recycled(ir, p, params, chunk=7)
        np.testing.assert_allclose(b, a, rtol=1e-14)
