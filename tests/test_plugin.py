"""The MG5_aMC `pyout` plugin with the CUDA backend, exercised with stand-ins of the MG5 objects
(MG5_aMC is not installed here; SURVEY.md Appendix G lists the calls the exporter makes).  CPU only."""
import fractions
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

from conftest import ROOT

# the HELAS call lines of the reference's frozen generated code, tests/mockup_debug_me.py:516-528
MOCKUP_CALLS = """w0 = vxxxxx(all_ps[:,0],ZERO,hel[0],float_me(-1))
w1 = vxxxxx(all_ps[:,1],ZERO,hel[1],float_me(-1))
w2 = oxxxxx(all_ps[:,2],mdl_MT,hel[2],float_me(+1))
w3 = ixxxxx(all_ps[:,3],mdl_MT,hel[3],float_me(-1))
w4= VVV1P0_1(w0,w1,GC_10,ZERO,ZERO)
# Amplitude(s) for diagram number 1
amp0= FFV1_0(w3,w2,w4,GC_11)
w4= FFV1_1(w2,w0,GC_11,mdl_MT,mdl_WT)
# Amplitude(s) for diagram number 2
amp1= FFV1_0(w3,w4,w1,GC_11)
w4= FFV1_2(w3,w0,GC_11,mdl_MT,mdl_WT)
# Amplitude(s) for diagram number 3
amp2= FFV1_0(w4,w2,w1,GC_11)""".split("\n")


class FakeProcess(dict):
    def shell_string(self):
        return "1_gg_ttx"

    def nice_string(self):
        return "Process: g g > t t~ WEIGHTED<=2 @1"

    def get_initial_ids(self):
        return [21, 21]


class FakeColorMatrix:
    def __bool__(self):
        return True

    def get_line_denominators(self):
        return [3, 3]

    def get_line_numerators(self, index, denominator):
        return [[16, -2], [-2, 16]][index]


class FakeMatrixElement:
    """What MG5's HelasMatrixElement answers for g g > t t~ (values: mockup_debug_me.py:415-543)."""

    def get(self, key):
        return {"processes": [FakeProcess(legs=[], model=None)], "diagrams": [1, 2, 3],
                "color_basis": {0: 1, 1: 1}, "color_matrix": FakeColorMatrix()}[key]

    def get_nexternal_ninitial(self):
        return (4, 2)

    def get_helicity_combinations(self):
        return 16

    def get_helicity_matrix(self):
        return [[a, b, c, d] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1) for d in (1, -1)]

    def get_denominator_factor(self):
        return 256

    def get_number_of_amplitudes(self):
        return 3

    def get_number_of_wavefunctions(self):
        return 5

    def get_mirror_processes(self):
        return []

    def get_color_amplitudes(self):
        one = fractions.Fraction(1)
        # ((fermion factor, fraction, is_imaginary, Nc power), amplitude number)
        return [[((1, one, True, 0), 1), ((-1, one, False, 0), 2)],
                [((-1, one, True, 0), 1), ((-1, one, False, 0), 3)]]


def test_plugin_registration_surface():
    plugin = importlib.import_module("madgraph_plugin")
    assert set(plugin.new_output) == {"pyout"}
    assert plugin.new_cluster == {} and plugin.new_interface is None
    assert plugin.minimal_mg5amcnlo_version == (2, 5, 0)
    exp = plugin.new_output["pyout"]
    assert (exp.check, exp.exporter, exp.output, exp.grouped_mode, exp.sa_symmetry) == (True, 'v4', 'dir', False, False)
    for method in ("pass_information_from_cmd", "generate_subprocess_directory", "convert_model", "finalize"):
        assert callable(getattr(exp, method))


def test_helas_call_lines_to_ir_reproduces_the_pinned_process():
    from madflow_b200 import process_ir
    from madgraph_plugin import PyOut_exporter

    ir = PyOut_exporter.matrix_element_to_ir(FakeMatrixElement(), MOCKUP_CALLS)
    pin = process_ir.gg_ttx_pinned()
    assert process_ir.validate(ir)
    strip = lambda calls: [{k: v for k, v in c.items() if k != "coup_sign"} for c in calls]
    assert strip(ir["calls"]) == pin["calls"]
    assert [list(t) for t in ir["jamp"]] == [list(t) for t in pin["jamp"]]
    for key in ("name", "nexternal", "ninitial", "ndiags", "ncomb", "nwavefuncs", "helicities", "denominator",
                "params", "couplings", "color_num", "color_denom", "initial_states", "mirror_initial_states"):
        assert ir[key] == pin[key], key


def test_helas_call_parser_rejects_garbage():
    from madgraph_plugin.PyOut_helas_call_writer import parse_helas_calls

    with pytest.raises(ValueError):
        parse_helas_calls(["w1 = something strange"])
    calls = parse_helas_calls(["w5= FFV1_1(w2,w0,-GC_11,mdl_MT,mdl_WT)", "w0 = vxxxxx(all_ps[:,0],ZERO, 4,float_me(-1))"])
    assert calls[0]["coup"] == "GC_11" and calls[0]["coup_sign"] == -1
    assert calls[1] == {"op": "vxxxxx", "out": 0, "leg": 0, "mass": "ZERO", "nsf": -1}


def test_coupling_power_law():
    from madgraph_plugin.PyOut_exporter import coupling_power_law, jamp_coefficient

    assert coupling_power_law("-G") == (-1.0, 0.0, 1)
    assert coupling_power_law("complex(0,1)*G") == (0.0, 1.0, 1)
    assert coupling_power_law("complex(0,1)*G**2") == (0.0, 1.0, 2)
    assert coupling_power_law("cmath.sqrt(G)") is None
    assert jamp_coefficient(1, fractions.Fraction(1, 3), True, 1) == (0.0, 1.0)
    assert jamp_coefficient(-1, fractions.Fraction(1, 3), False, 0) == (-1.0 / 3.0, 0.0)


def test_write_process_files(tmp_path):
    """The exporter's products for one subprocess: IR, CUDA source, generated Python module."""
    from madflow_b200 import process_ir
    from madgraph_plugin import PyOut_exporter

    ir = PyOut_exporter.matrix_element_to_ir(FakeMatrixElement(), MOCKUP_CALLS)
    lines = ("    mdl_MT = param_card['MASS'].get(6).value\n    mdl_WT = param_card['DECAY'].get(6).value\n"
             "    GC_10 = lambda G: complex_me(-G)\n    GC_11 = lambda G: complex_me(complex(0,1)*G)\n")
    path = PyOut_exporter.write_process_files(ir, str(tmp_path), lines, build=False)
    assert os.path.exists(tmp_path / "1_gg_ttx.json") and os.path.exists(tmp_path / "1_gg_ttx.cu")
    src = open(tmp_path / "1_gg_ttx.cu").read()
    assert "MF_DEFINE_PROCESS(Proc)" in src and "mf::VVV1P0_1(w0, w1, coup[0]" in src
    assert process_ir.loads(open(tmp_path / "1_gg_ttx.json").read())["calls"] == ir["calls"]
    # the generated module: get_model_param works from a param_card without MG5
    (tmp_path / "param_card.dat").write_text("Block MASS\n  6 1.730000e+02 # MT\nDECAY 6 1.491500e+00 # WT\n")
    sys.path.insert(0, str(tmp_path))
    try:
        mod = importlib.import_module("matrix_1_gg_ttx")
        text = open(path).read()
        assert "class Matrix_1_gg_ttx(Matrix)" in text and "def get_model_param(model, param_card_path)" in text
        import torch

        if torch.cuda.is_available():
            model = mod.get_model_param(None, str(tmp_path / "param_card.dat"))
            assert [float(m) for m in model.get_masses()] == [173.0]
    finally:
        sys.path.remove(str(tmp_path))
        sys.modules.pop("matrix_1_gg_ttx", None)


def test_param_card_reader(tmp_path):
    from madflow_b200.param_card import ParamCard

    (tmp_path / "card.dat").write_text("Block SMINPUTS\n 1 1.325070e+02 # aEWM1\n 3 1.180000e-01 # aS\n"
                                       "Block MASS\n 6 1.730000e+02\nDECAY 6 1.491500e+00\n")
    card = ParamCard(str(tmp_path / "card.dat"))
    assert card["SMINPUTS"].get(3).value == 0.118 and card["MASS"].get(6).value == 173.0
    assert card["DECAY"].get(6).value == 1.4915


def test_aloha_cpp_to_cuda_retargeting():
    from madgraph_plugin.PyOut_create_aloha import cpp_to_cuda

    cpp = """void VVV1P0_1(std::complex<double> V2[], std::complex<double> V3[], std::complex<double> COUP, double M1, double W1, std::complex<double> V1[])
{
  static std::complex<double> cI = std::complex<double>(0.,1.);
  std::complex<double> TMP1;
  V1[0] = +V2[0]+V3[0];
  TMP1 = (V3[2]*V2[2]);
  V1[2] = COUP * TMP1 * cI;
}"""
    cu = cpp_to_cuda(cpp)
    assert cu.startswith("__host__ __device__ __forceinline__ void VVV1P0_1(const cxtype V2[], const cxtype V3[], cxtype COUP")
    assert "cxtype V1[])" in cu and "const cxtype V1[]" not in cu
    assert "const cxtype cI(0., 1.);" in cu and "static" not in cu and "std::complex" not in cu


def test_leading_order_wrapper(tmp_path):
    """finalize() writes leading_order.py like the reference's exporter (PyOut_exporter.py:442-468, 544-549); the
    script is valid Python that imports the generated matrix module and builds the fused integrand."""
    import ast
    import io

    from madgraph_plugin import PyOut_exporter

    exp = PyOut_exporter.PyOutExporter.__new__(PyOut_exporter.PyOutExporter)
    exp.dir_path = str(tmp_path)
    exp.me_names, exp.proc_names, exp.mass_lists = ["matrix_1_gg_ttx"], ["1_gg_ttx"], [["ZERO", "ZERO", "mdl_MT", "mdl_MT"]]
    out = io.StringIO()
    exp.write_leading_order_wrapper(out, ["generate g g > t t~", "output pyout out"])
    text = out.getvalue()
    ast.parse(text)
    assert "from matrix_1_gg_ttx import Matrix_1_gg_ttx, get_model_param as model_1_gg_ttx" in text
    assert '"1_gg_ttx": (Matrix_1_gg_ttx, model_1_gg_ttx)' in text
    assert '"1_gg_ttx": ["ZERO", "ZERO", "mdl_MT", "mdl_MT"]' in text
    assert "generate g g > t t~" in text and "FusedIntegrand(matrix, model, sqrts=SQRTS" in text


# the text MG5's C++ ALOHA writer produces for two routines of the electroweak sector (UFO FFV2 = Gamma(3,2,-1) ProjM(-1,1))
FFV2_0_CPP = """void FFV2_0(std::complex<double> F1[], std::complex<double> F2[], std::complex<double> V3[], std::complex<double> COUP, std::complex<double> & vertex)
{
  static std::complex<double> cI = std::complex<double> (0., 1.);
  std::complex<double> TMP0;
  TMP0 = (F1[2] * (F2[4] * (V3[2] + V3[5]) + F2[5] * (V3[3] + cI * (V3[4]))) + F1[3] * (F2[4] * (V3[3] - cI * (V3[4])) + F2[5] * (V3[2] - V3[5])));
  vertex = COUP * - cI * TMP0;
}"""
FFV2_1_CPP = """void FFV2_1(std::complex<double> F2[], std::complex<double> V3[], std::complex<double> COUP, double M1, double W1, std::complex<double> F1[])
{
  static std::complex<double> cI = std::complex<double> (0., 1.);
  double P1[4];
  std::complex<double> denom;
  F1[0] = +F2[0] + V3[0];
  F1[1] = +F2[1] + V3[1];
  P1[0] = -F1[0].real();
  P1[1] = -F1[1].real();
  P1[2] = -F1[1].imag();
  P1[3] = -F1[0].imag();
  denom = COUP/((P1[0] * P1[0]) - (P1[1] * P1[1]) - (P1[2] * P1[2]) - (P1[3] * P1[3]) - M1 * (M1 - cI * W1));
  F1[2] = denom * cI * M1 * (F2[4] * (V3[2] + V3[5]) + F2[5] * (V3[3] + cI * (V3[4])));
  F1[3] = denom * - cI * M1 * (F2[4] * (+cI * (V3[4]) - V3[3]) + F2[5] * (V3[5] - V3[2]));
  F1[4] = denom * (-cI) * (F2[4] * (P1[0] * (V3[2] + V3[5]) + (P1[1] * (+cI * (V3[4]) - V3[3]) + (P1[2] * (-1.) * (V3[4] + cI * (V3[3])) - P1[3] * (V3[2] + V3[5])))) + F2[5] * (P1[0] * (V3[3] + cI * (V3[4])) + (P1[1] * (V3[5] - V3[2]) + (P1[2] * (-cI * (V3[2]) + cI * (V3[5])) - P1[3] * (V3[3] + cI * (V3[4]))))));
  F1[5] = denom * cI * (F2[4] * (P1[0] * (+cI * (V3[4]) - V3[3]) + (P1[1] * (V3[2] + V3[5]) + (P1[2] * (-1.) * (+cI * (V3[2] + V3[5])) + P1[3] * (+cI * (V3[4]) - V3[3])))) + F2[5] * (P1[0] * (V3[5] - V3[2]) + (P1[1] * (V3[3] + cI * (V3[4])) + (P1[2] * (V3[4] - cI * (V3[3])) + P1[3] * (V3[5] - V3[2])))));
}"""


def _np_ffv2_0(F1, F2, V3, COUP):
    TMP0 = (F1[2] * (F2[4] * (V3[2] + V3[5]) + F2[5] * (V3[3] + 1j * V3[4]))
            + F1[3] * (F2[4] * (V3[3] - 1j * V3[4]) + F2[5] * (V3[2] - V3[5])))
    return COUP * -1j * TMP0


def _np_ffv2_1(F2, V3, COUP, M1, W1):
    cI = 1j
    F1 = [None] * 6
    F1[0], F1[1] = F2[0] + V3[0], F2[1] + V3[1]
    P1 = [-F1[0].real, -F1[1].real, -F1[1].imag, -F1[0].imag]
    denom = COUP / (P1[0] ** 2 - P1[1] ** 2 - P1[2] ** 2 - P1[3] ** 2 - M1 * (M1 - cI * W1))
    F1[2] = denom * cI * M1 * (F2[4] * (V3[2] + V3[5]) + F2[5] * (V3[3] + cI * V3[4]))
    F1[3] = denom * -cI * M1 * (F2[4] * (cI * V3[4] - V3[3]) + F2[5] * (V3[5] - V3[2]))
    F1[4] = denom * (-cI) * (F2[4] * (P1[0] * (V3[2] + V3[5]) + (P1[1] * (cI * V3[4] - V3[3]) + (P1[2] * (-1.) * (V3[4] + cI * V3[3]) - P1[3] * (V3[2] + V3[5]))))
                             + F2[5] * (P1[0] * (V3[3] + cI * V3[4]) + (P1[1] * (V3[5] - V3[2]) + (P1[2] * (-cI * V3[2] + cI * V3[5]) - P1[3] * (V3[3] + cI * V3[4])))))
    F1[5] = denom * cI * (F2[4] * (P1[0] * (cI * V3[4] - V3[3]) + (P1[1] * (V3[2] + V3[5]) + (P1[2] * (-1.) * (cI * (V3[2] + V3[5])) + P1[3] * (cI * V3[4] - V3[3]))))
                          + F2[5] * (P1[0] * (V3[5] - V3[2]) + (P1[1] * (V3[3] + cI * V3[4]) + (P1[2] * (V3[4] - cI * V3[3]) + P1[3] * (V3[5] - V3[2])))))
    return np.stack(np.broadcast_arrays(*F1))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")
def test_plugin_written_aloha_routines_reach_the_kernel(monkeypatch):
    """A vertex outside the hand-written QCD set (FFV2, the left-handed current of the electroweak sector) and a coupling
    outside GC_10/11/12: the routine text in MG5's C++ ALOHA format goes through cpp_to_cuda -> attach_aloha_routines ->
    codegen (pasted into the translation unit, called from the generated matrix()) -> nvcc, and the generated code,
    executed on the CPU, equals a numpy evaluation of the same call list (reference: PyOut_create_aloha.py:124-197,
    PyOut_exporter.py:378-407, 507-540)."""
    import hostcheck as hc
    from madflow_b200 import codegen, process_ir
    from madgraph_plugin.PyOut_create_aloha import cpp_to_cuda
    from madgraph_plugin.PyOut_exporter import PyOutExporterError, attach_aloha_routines, coupling_power_law
    from oracle import SQH_REF, aloha
    from oracle import matrix as omatrix

    ir = process_ir.gg_ttx_pinned()
    ir["name"] = "1_gg_ttx_ffv2"
    # the t-channel diagram with FFV2 vertices and their own coupling: a different (unphysical) process, the same plumbing
    for c in ir["calls"]:
        if c["op"] == "FFV1_1":
            c["op"], c["coup"] = "FFV2_1", "GC_100"
        elif c.get("amp") == 1:
            c["op"], c["coup"] = "FFV2_0", "GC_100"
    ir["couplings"] = sorted({c["coup"] for c in ir["calls"] if "coup" in c})
    law = coupling_power_law("(ee*complex(0,1))/(sw*cmath.sqrt(2))*G**2", {"ee": 0.3, "sw": 0.48})
    assert law is not None and law[2] == 2
    ir["coupling_defs"] = {"GC_100": list(law)}
    with pytest.raises(PyOutExporterError):
        attach_aloha_routines(process_ir.clone(ir), {})
    with pytest.raises(ValueError):
        codegen.emit_process_source(ir)          # routines neither built in nor supplied
    attach_aloha_routines(ir, {"FFV2_0": cpp_to_cuda(FFV2_0_CPP), "FFV2_1": cpp_to_cuda(FFV2_1_CPP), "FFV9_9": "unused"})
    assert sorted(ir["aloha_routines"]) == ["FFV2_0", "FFV2_1"]
    src = codegen.emit_process_source(ir)
    assert "HP_AVAILABLE = false" in src and "plg_FFV2_1(" in src and "namespace plg" in src
    process_ir.loads(process_ir.dumps(ir))
    lib = hc.process(ir)
    from test_procgen import _points
    from conftest import MT, WT, sm_params

    p = _points(0, n=50, seed=3)
    a_s = 0.09 + 0.05 * np.random.default_rng(3).random(50)
    params = sm_params(alpha_s=a_s)
    G = 2.0 * np.sqrt(np.pi * a_s)
    params["GC_100"] = complex(law[0], law[1]) * G ** law[2]
    coup = np.stack([params[c] for c in ir["couplings"]])
    monkeypatch.setitem(aloha.ROUTINES, "FFV2_0", _np_ffv2_0)
    monkeypatch.setitem(aloha.ROUTINES, "FFV2_1", _np_ffv2_1)
    ref = omatrix.smatrix(ir, p, params)
    out = hc.smatrix(lib, ir, p, [MT, WT], coup, SQH_REF)
    np.testing.assert_allclose(out, ref, rtol=1e-12)
    assert np.all(ref > 0)
