"""The CUDA device code executed on the CPU (tests/hostcheck) against the oracle and the golden
vectors: the functions are __host__ __device__, so this checks the very source the GPU kernels are
compiled from, in the container that has no GPU.  Needs nvcc; CPU only."""
import ctypes
import shutil

import numpy as np
import pytest

from conftest import MT, WT, sm_params
from madflow_b200 import process_ir
from oracle import ACC_REF, GEV2PB_REF, PI_REF, SQH_REF, aloha, philox
from oracle import matrix as omatrix
from oracle import phasespace as ops
from oracle import vegas as ovegas

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")


@pytest.fixture(scope="module")
def hc():
    import hostcheck

    return hostcheck


def test_device_helas_on_host(hc, golden):
    lib, g = hc.core(), golden("wavefunctions")
    for key in g.files:
        if key.startswith("p_"):
            continue
        kind, m, h, s = key.split("_")
        mass, nhel, ns = float(m[1:]), int(h[1:]), int(s[1:])
        p = np.ascontiguousarray(g[f"p_m{int(mass)}"])
        out = np.empty((6, p.shape[0]), dtype=np.complex128)
        lib.hc_wavefunction({"i": 0, "o": 1, "v": 2}[kind], hc._dp(p), ctypes.c_longlong(p.shape[0]),
                            ctypes.c_double(mass), nhel, ns, ctypes.c_double(SQH_REF), hc._dp(out.view(np.float64)))
        np.testing.assert_allclose(out, g[key], rtol=1e-14, atol=1e-300, err_msg=key)


def test_device_process_on_host(hc, golden):
    g = golden("matrix_gg_ttx")
    ir = process_ir.gg_ttx_pinned()
    lib = hc.process(ir)
    coup = np.array([g["GC_10"], g["GC_11"]])
    for key in ("13tev_com", "13tev_lab", "7tev_com", "7tev_lab"):
        out = hc.smatrix(lib, ir, g[key + "_p"], g["params"], coup, SQH_REF)
        np.testing.assert_allclose(out, g[key + "_smatrix"], rtol=1e-13)
    gs = g["run_gs"]
    out = hc.smatrix(lib, ir, g["13tev_lab_p"], g["params"], np.stack([-gs, 1j * gs]), SQH_REF)
    np.testing.assert_allclose(out, g["run_smatrix"], rtol=1e-13)


def test_device_ramboflow_on_host(hc, golden):
    lib, g = hc.core(), golden("phasespace")
    cases = {"tt": (4, 13e3, [MT, MT]), "ttg": (5, 13e3, [MT, MT, 0.0]), "ttgg": (6, 13e3, [MT, MT, 0.0, 0.0]),
             "ttggg": (7, 13e3, [MT, MT, 0.0, 0.0, 0.0]), "m50_125": (4, 7e3, [50.0, 125.0])}
    for name, (nx, s, ms) in cases.items():
        x = np.ascontiguousarray(g[f"rf_{name}_x"])
        ne = x.shape[0]
        p, w, x1, x2 = np.empty((ne, nx, 4)), np.empty(ne), np.empty(ne), np.empty(ne)
        ma = np.array(ms + [0.0] * (8 - len(ms)))
        lib.hc_ramboflow(nx, hc._dp(x), ctypes.c_longlong(ne), ctypes.c_double(s), hc._dp(ma), ctypes.c_double(PI_REF),
                         ctypes.c_double(ACC_REF), ctypes.c_double(GEV2PB_REF), 0, hc._dp(p), hc._dp(w), hc._dp(x1),
                         hc._dp(x2))
        pr, wr, _, _ = ops.ramboflow(x, nx, s, ms, xfactor="converged")
        np.testing.assert_allclose(p, pr, rtol=1e-12, atol=1e-8)
        np.testing.assert_allclose(w, wr, rtol=1e-12)


def test_device_philox_on_host(hc):
    lib = hc.core()
    u = np.empty((100, 7))
    lib.hc_philox(ctypes.c_ulonglong(4), 3, ctypes.c_ulonglong(2**33 + 5), ctypes.c_longlong(100), 7, hc._dp(u))
    np.testing.assert_array_equal(u, philox.uniforms(4, 3, 2**33 + 5, 100, 7))


def test_device_vegas_map_on_host(hc):
    lib = hc.core()
    grid = ovegas.refine_grid(np.random.default_rng(1).random((5, 50)) + 0.1, ovegas.uniform_grid(5))
    r = ovegas.confine(np.random.default_rng(2).random((300, 5)))
    x, w = np.empty((300, 5)), np.empty(300)
    bins = np.empty((300, 5), dtype=np.int32)
    lib.hc_vegas_map(hc._dp(np.ascontiguousarray(grid)), hc._dp(r), ctypes.c_longlong(300), 5, hc._dp(x),
                     bins.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), hc._dp(w))
    xr, kr, wr = ovegas.map_to_grid(r, grid)
    np.testing.assert_array_equal(bins, kr)
    np.testing.assert_allclose(x, xr, rtol=1e-15)
    np.testing.assert_allclose(w, wr, rtol=1e-14)


def test_device_pdf_on_host(hc, tmp_path):
    """csrc/pdf.cuh (log-bicubic x f(x, Q2), alpha_s table, luminosity + scale of an event) executed on the CPU
    against oracle/pdf.py on a synthetic lhagrid1 set: knots, cell interiors, subgrid thresholds, frozen edges."""
    from madflow_b200 import pdf as mpdf
    from oracle import pdf as opdf

    opdf.write_toy_set(str(tmp_path))
    og = opdf.GridPDF.from_set("ToyPDF/0", str(tmp_path))
    pd = mpdf.mkPDF("ToyPDF/0", dirname=str(tmp_path))
    T = np.ascontiguousarray(pd._host_table)
    lib = hc.core()
    rng = np.random.default_rng(3)
    n = 4000
    x = 10 ** rng.uniform(-7.5, 0.0, n)            # below xmin: frozen
    q2 = 10 ** rng.uniform(0.0, 8.5, n)            # below q2min / above q2max: frozen
    sg = og.subgrids[1]
    x[:60], q2[:60] = sg["x"], sg["q2"][3]         # on knots
    q2[60:70] = 4.75 ** 2                          # on the subgrid threshold
    x[70:80], q2[70:80] = 1.0, 1e4                 # upper corner
    pids = [21, 2, -1, 5]
    cols = np.array([pd.column(p) for p in pids], dtype=np.int32)
    out = np.empty((n, len(pids)))
    lib.hc_pdf_xfx(hc._dp(T), cols.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(pids), hc._dp(x), hc._dp(q2),
                   ctypes.c_longlong(n), hc._dp(out))
    ref = og.xfxQ2(pids, x, q2)
    scale = np.max(np.abs(ref), axis=0)
    assert np.max(np.abs(out - ref) / scale) < 1e-13
    np.testing.assert_array_equal(out[:60], sg["xf"][:, 3, cols])   # knots exactly
    qa = 10 ** rng.uniform(-0.5, 9.0, n)
    qa[:13] = og.as_q2
    a = np.empty(n)
    lib.hc_pdf_alphas(hc._dp(T), hc._dp(qa), ctypes.c_longlong(n), hc._dp(a))
    np.testing.assert_allclose(a, og.alphasQ2(qa), rtol=1e-13)
    # scale + alpha_s + luminosity of an event, g g and q q~ (+ mirrored) channels
    xr = rng.random((500, 10))
    p, w, x1, x2 = ops.ramboflow(xr, 4, 13e3, [MT, MT], xfactor="converged")
    lab = np.ascontiguousarray(ops.boost_to_lab(p, x1, x2))
    q2e = (np.sum(ops.mt(lab[:, 2:4]), axis=-1) / 2.0) ** 2
    ini = [(1, -1), (2, -2), (-1, 1), (-2, 2)]
    f1 = np.array([pd.column(a_) for a_, _ in ini], dtype=np.int8)
    f2 = np.array([pd.column(b_) for _, b_ in ini], dtype=np.int8)
    as_, lumi = np.empty(500), np.empty(500)
    for fixed in (0.0, 91.46 ** 2):
        lib.hc_event_scale(hc._dp(T), 2, ctypes.c_double(0.118), ctypes.c_double(1.0), ctypes.c_double(0.0),
                           ctypes.c_double(fixed), len(ini), f1.ctypes.data_as(ctypes.POINTER(ctypes.c_byte)),
                           f2.ctypes.data_as(ctypes.POINTER(ctypes.c_byte)), hc._dp(lab), hc._dp(x1), hc._dp(x2),
                           ctypes.c_longlong(500), hc._dp(as_), hc._dp(lumi))
        qq = np.full(500, fixed) if fixed else q2e
        p1, p2 = og.xfxQ2([a_ for a_, _ in ini], x1, qq), og.xfxQ2([b_ for _, b_ in ini], x2, qq)
        np.testing.assert_allclose(lumi, np.sum(p1 * p2, axis=1) / x1 / x2, rtol=1e-12)
        np.testing.assert_allclose(as_, og.alphasQ2(qq), rtol=1e-12)


def test_device_pdf_single_subgrid_on_host(hc, tmp_path):
    """csrc/pdf.cuh on a set with ONE subgrid (no thresholds) and a 5-knot alpha_s table without repeated Q."""
    from madflow_b200 import pdf as mpdf
    from oracle import pdf as opdf

    opdf.write_toy_set(str(tmp_path), name="OneGrid", q_knots=((1.65, 3.0, 10.0, 100.0, 1000.0),))
    info = tmp_path / "OneGrid" / "OneGrid.info"
    lines = [ln for ln in info.read_text().splitlines() if not ln.startswith("AlphaS_")]
    lines += ["AlphaS_Qs: [2.0, 5.0, 20.0, 91.1876, 1000.0]", "AlphaS_Vals: [0.30, 0.21, 0.15, 0.118, 0.088]"]
    info.write_text("\n".join(lines) + "\n")
    og = opdf.GridPDF.from_set("OneGrid/0", str(tmp_path))
    T = np.ascontiguousarray(mpdf.mkPDF("OneGrid/0", dirname=str(tmp_path))._host_table)
    lib = hc.core()
    rng = np.random.default_rng(5)
    n = 2000
    x, q2 = 10 ** rng.uniform(-7.0, 0.0, n), 10 ** rng.uniform(0.3, 6.2, n)
    cols = np.arange(11, dtype=np.int32)
    out = np.empty((n, 11))
    lib.hc_pdf_xfx(hc._dp(T), cols.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 11, hc._dp(x), hc._dp(q2), ctypes.c_longlong(n),
                   hc._dp(out))
    ref = og.xfxQ2(og.pids, x, q2)
    assert np.max(np.abs(out - ref) / np.max(np.abs(ref), axis=0)) < 1e-13
    qa = 10 ** rng.uniform(0.0, 7.0, n)
    a = np.empty(n)
    lib.hc_pdf_alphas(hc._dp(T), hc._dp(qa), ctypes.c_longlong(n), hc._dp(a))
    np.testing.assert_allclose(a, og.alphasQ2(qa), rtol=1e-13)
