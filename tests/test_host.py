"""Host logic and the C-ABI surface, CPU only (no kernel is launched)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from madflow_b200 import codegen, process_ir
from madflow_b200.vegas import combine_iterations, iteration_sigma, shard_events


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfp?_[a-z0-9_]+)\s*\(", text)))


def test_core_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(os.path.join(ROOT, "madflow_b200", "lib", "libmadflow_b200.so"))
    names = _declared("madflow_b200.h")
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    lib.mf_version.restype = ctypes.c_int
    assert lib.mf_version() == 1


def test_process_library_exports_every_declared_symbol():
    from madflow_b200 import _runtime as rt

    names = _declared("madflow_b200_process.h")
    assert len(names) >= 10
    lib = rt.process_lib("1_gg_ttx")
    for n in names:
        assert hasattr(lib.lib, n), n
    # metadata calls are host-only
    assert lib.name == "1_gg_ttx"
    assert (lib.info.nexternal, lib.info.ncomb, lib.info.ncolor, lib.info.ndiags, lib.info.ndim) == (4, 16, 2, 3, 10)
    assert lib.info.denominator == 256.0
    assert lib.info.flops_per_event == 23697.0
    assert lib.param_names == ["mdl_MT", "mdl_WT"] and lib.coupling_names == ["GC_10", "GC_11"]
    assert lib.coupling_defs == [(-1.0, 0.0, 1), (0.0, 1.0, 1)]
    assert lib.helicities == process_ir.gg_ttx_pinned()["helicities"]


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from madflow_b200 import _runtime as rt
    from madflow_b200.matrix import Matrix

    m = Matrix("1_gg_ttx")
    with pytest.raises(RuntimeError):
        m.smatrix(np.zeros((2, 4, 4)), 173.0, 1.5, np.array([-1.2 + 0j]), np.array([1.2j]))
    with pytest.raises(rt.MadflowB200Error):
        rt.process_lib("no_such_process")


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "madflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_flop_count_matches_survey():
    assert codegen.flops_per_event(process_ir.gg_ttx_pinned()) == 23697


def test_shard_events_partitions_the_range():
    for n in (1, 7, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            parts = [shard_events(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_iteration_statistics():
    rng = np.random.default_rng(0)
    f = rng.random(10000)
    n = f.size
    t = f / n
    res, res2 = t.sum(), (t * t).sum()
    assert abs(iteration_sigma(res, res2, n) - f.std(ddof=1) / np.sqrt(n)) < 1e-12
    final, err, chi2 = combine_iterations([(1.0, 0.1), (1.2, 0.2)])
    assert abs(final - (1.0 / 0.01 + 1.2 / 0.04) / (1 / 0.01 + 1 / 0.04)) < 1e-14
    assert abs(err - (1 / (1 / 0.01 + 1 / 0.04)) ** 0.5) < 1e-14


_GLOO = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from madflow_b200.vegas import allreduce_sums, shard_events
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1001
first, count = shard_events(n, rank, world)
f = torch.arange(n, dtype=torch.float64)[first:first + count]
sums = torch.zeros(4 + 3 * 50, dtype=torch.float64)
sums[0], sums[1] = f.sum(), (f * f).sum()
sums[4 + rank] = 1.0
allreduce_sums(sums)
full = torch.arange(n, dtype=torch.float64)
assert sums[0] == full.sum() and sums[1] == (full * full).sum(), sums[:2]
assert sums[4:4 + world].tolist() == [1.0] * world
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gloo_allreduce(tmp_path):
    script = tmp_path / "gloo_check.py"
    script.write_text(_GLOO)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29671", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_lhe_writer_round_trip(tmp_path):
    """LheWriter with the reference's protocol (lhe_writer.py:94-356): weighted file, unweighting, cross_err.txt."""
    from madflow_b200.lhe_writer import EventFileFlow, FourMomentumFlow, LheWriter

    rng = np.random.default_rng(0)
    ps = rng.normal(size=(60, 6, 4)) * 50.0
    ps[:, :, 0] = np.sqrt(np.sum(ps[:, :, 1:] ** 2, axis=-1) + np.array([0, 0, 173.0, 173.0, 0, 0]) ** 2)
    w = rng.random(60)
    w[::6] = 0.0   # events cut away are not written
    with LheWriter(tmp_path, "run_01") as lw:
        lw.lhe_parser(ps[:30], w[:30])
        lw.lhe_parser(ps[30:], w[30:])
        lw.store_result((12.5, 0.1))
        lw.dump_result(tmp_path / "cross_err.txt")
    kept = np.nonzero(w)[0]
    events = list(EventFileFlow(tmp_path / "Events/run_01/weighted_events.lhe.gz"))
    assert len(events) == len(kept)
    for ev, i in zip(events, kept):
        assert ev.nexternal == 6 and ev.wgt == pytest.approx(w[i], rel=1e-7)
        assert [p.pid for p in ev] == [21, 21, 6, -6, 21, 21] and [p.status for p in ev] == [-1, -1, 1, 1, 1, 1]
        np.testing.assert_allclose([[p.E, p.px, p.py, p.pz] for p in ev], ps[i], rtol=1e-10)
    top = FourMomentumFlow(events[0][2])
    assert top.pt == pytest.approx(np.hypot(ps[kept[0], 2, 1], ps[kept[0], 2, 2]), rel=1e-10)
    assert top.mass == pytest.approx(173.0, rel=1e-6)
    unw = list(EventFileFlow(tmp_path / "Events/run_01/unweighted_events.lhe.gz"))
    assert 0 < len(unw) <= len(events) and all(e.wgt == pytest.approx(12.5) for e in unw)
    np.testing.assert_allclose(np.loadtxt(tmp_path / "cross_err.txt"), [12.5, 0.1])


def test_madflow_cli_arguments_and_process_names():
    """The `madflow` command line of scripts/madflow_exec.py:243-308 (names, defaults, optional values)."""
    from madflow_b200.scripts.madflow_exec import madflow_main, process_library_name

    args, _, _ = madflow_main(["--no_pdf", "-c", "-q", "-i", "6", "-f", "3", "--histograms",
                               "--madgraph_process", "g g > t t~ g g"], quick_return=True)
    assert args.pt_cut == 30.0 and args.fixed_scale == 91.46 and args.iterations == 6 and args.frozen_iter == 3
    assert args.events_per_iteration == int(1e6) and args.massive_particles == 2 and args.histograms and args.no_pdf
    assert process_library_name(args.madgraph_process) == "1_gg_ttxgg"
    assert process_library_name("g g > t t~") == "1_gg_ttx"
    from madflow_b200.scripts.madflow_exec import subprocess_libraries

    assert subprocess_libraries("p p  > t t~") == ["1_gg_ttx", "1_uux_ttx"] and subprocess_libraries("g g > t t~ g") == ["1_gg_ttxg"]
    assert madflow_main(["--dry_run", "--madgraph_process", "p p > t t~"]) == (None, None, None)
    assert subprocess_libraries("p p > t t~ j") == ["1_gg_ttxg", "1_gu_ttxu", "1_gux_ttxux", "1_uux_ttxg"]
    assert process_library_name("g u~ > t t~ u~") == "1_gux_ttxux"
    assert madflow_main(["--dry_run", "--madgraph_process", "p p > t t~ j"]) == (None, None, None)
    # the six-point light-quark libraries are part of the default build (madflow_b200.build.builtin_irs)
    assert madflow_main(["--dry_run", "--madgraph_process", "p p > t t~ g g"]) == (None, None, None)
    assert len(subprocess_libraries("p p > t t~ j j")) == 12
    assert madflow_main(["--dry_run", "--madgraph_process", "p p > t t~ j j"]) == (None, None, None)
    assert madflow_main(["--dry_run"]) == (None, None, None)   # like the reference, a dry run stops before the PDF
    with pytest.raises(SystemExit, match="NNPDF31_nnlo_as_0118"):
        madflow_main(["-i", "2"])             # a missing PDF set is refused, not approximated
    with pytest.raises(SystemExit):
        madflow_main(["--no_pdf", "--dry_run", "--madgraph_process", "e+ e- > t t~"])   # no such library
    assert madflow_main(["--no_pdf", "--dry_run", "--madgraph_process", "g g > t t~ g"]) == (None, None, None)


def test_pdf_set_reader_and_table(tmp_path):
    """madflow_b200.pdf: lhagrid1 reader, the packed table of csrc/pdf.cuh, pdfflow's mkPDF conventions, errors."""
    from madflow_b200 import pdf as mpdf
    from oracle import pdf as opdf

    opdf.write_toy_set(str(tmp_path))
    p = mpdf.mkPDF("ToyPDF/0", dirname=str(tmp_path))
    assert p.flavor_scheme == [-5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 21] and p.column(0) == p.column(21) == 10
    assert p.has_alphas and p.q2min == pytest.approx(1.65**2) and p.q2max == pytest.approx(1e8) and p.nmembers == 1
    T = p._host_table
    nsub, nfl, nas = int(T[0]), int(T[1]), int(T[2])
    assert (nsub, nfl, nas) == (2, 11, 2)
    og = opdf.GridPDF.from_set("ToyPDF/0", str(tmp_path))
    for s_, sg in enumerate(og.subgrids):
        nx, nq, ox, olx, oq, olq, oxf, q2min = T[8 + 8 * s_: 16 + 8 * s_]
        nx, nq, ox, olx, oq, olq, oxf = int(nx), int(nq), int(ox), int(olx), int(oq), int(olq), int(oxf)
        np.testing.assert_array_equal(T[ox:ox + nx], sg["x"])
        np.testing.assert_array_equal(T[olx:olx + nx], np.log(sg["x"]))
        np.testing.assert_array_equal(T[oq:oq + nq], sg["q2"])
        np.testing.assert_array_equal(T[oxf:oxf + nx * nq * nfl].reshape(nx, nq, nfl), sg["xf"])
        assert q2min == sg["q2"][0]
    assert T[3] == og.as_q2[0] and T[6] == og.as_q2[-1] and T[7] == og.as_vals[-1]
    os.environ["LHAPDF_DATA_PATH"] = str(tmp_path)
    try:
        assert mpdf.mkPDF("ToyPDF").member == 0
    finally:
        del os.environ["LHAPDF_DATA_PATH"]
    with pytest.raises(mpdf.PDFError, match="not found"):
        mpdf.mkPDF("NoSuchSet/0", dirname=str(tmp_path))
    with pytest.raises(mpdf.PDFError, match="not found"):
        mpdf.mkPDF("ToyPDF/3", dirname=str(tmp_path))
    with pytest.raises(mpdf.PDFError, match="flavour 6"):
        p.column(6)

    class M:
        initial_states, mirror_initial_states = [(1, -1), (2, -2)], True

    assert mpdf.initial_state_channels(M, p) == ([5, 6, 4, 3], [4, 3, 5, 6])   # madflow_exec.py:141-155
    import torch

    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU implementation"):
            p.xfxQ2([21], [0.1], [1e4])    # nothing but the CUDA kernels behind the product API


def test_ctypes_structs_match_the_c_headers(tmp_path):
    """The ctypes mirrors of mfp_integrand_args / mfp_event_view / mfp_info have the C compiler's layout."""
    import shutil

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from madflow_b200 import _runtime as rt

    fields = {"mfp_integrand_args": ["d_grid", "nevents", "masses", "cuts", "par", "alpha_mode", "sqh", "d_partial",
                                     "d_workspace", "d_pdf", "nchannels", "chan_fl1", "chan_fl2", "fixed_q2", "skip_accumulate"],
              "mfp_event_view": ["d_mom", "capacity", "d_bins"], "mfp_info": ["name", "ndim", "denominator", "flops_per_event"]}
    src = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/madflow_b200_process.h"', "int main(void) {"]
    for st, fl in fields.items():
        src.append(f'  printf("%zu", sizeof({st}));')
        src += [f'  printf(" %zu", offsetof({st}, {f}));' for f in fl]
        src.append('  printf("\\n");')
    src += ["  return 0;", "}"]
    (tmp_path / "layout.c").write_text("\n".join(src))
    subprocess.run(["gcc", str(tmp_path / "layout.c"), "-o", str(tmp_path / "layout")], check=True)
    lines = subprocess.run([str(tmp_path / "layout")], check=True, capture_output=True, text=True).stdout.split("\n")
    for (st, fl), line in zip(fields.items(), lines):
        cls = getattr(rt, st)
        assert [int(v) for v in line.split()] == [ctypes.sizeof(cls)] + [getattr(cls, f).offset for f in fl], st


def test_pdf_set_variants(tmp_path):
    """Sets with one subgrid, without an alpha_s table, malformed files: reader, table header, loud errors."""
    from madflow_b200 import pdf as mpdf
    from oracle import pdf as opdf

    opdf.write_toy_set(str(tmp_path), name="OneGrid", q_knots=((1.65, 3.0, 10.0, 100.0, 1000.0),))
    p = mpdf.mkPDF("OneGrid/0", dirname=str(tmp_path))
    assert len(p.subgrids) == 1 and int(p._host_table[0]) == 1 and p.q2max == pytest.approx(1e6)
    # strip the alpha_s table: the set still gives PDFs, alphasQ2 is refused, the table header says "no alpha_s"
    info = tmp_path / "OneGrid" / "OneGrid.info"
    info.write_text("\n".join(ln for ln in info.read_text().splitlines() if not ln.startswith("AlphaS_Qs") and not ln.startswith("AlphaS_Vals")))
    q = mpdf.mkPDF("OneGrid/0", dirname=str(tmp_path))
    assert not q.has_alphas and int(q._host_table[2]) == 0
    with pytest.raises(mpdf.PDFError, match="AlphaS"):
        q.alphasQ2([100.0])
    # too few knots for the bicubic interpolation, and a file that is not lhagrid1
    opdf.write_toy_set(str(tmp_path), name="Coarse", q_knots=((2.0, 10.0, 100.0),))
    with pytest.raises(mpdf.PDFError, match="4 knots"):
        mpdf.mkPDF("Coarse/0", dirname=str(tmp_path))
    bad = tmp_path / "Coarse" / "Coarse_0000.dat"
    bad.write_text("PdfType: central\nFormat: lhagrid2\n---\n")
    with pytest.raises(mpdf.PDFError, match="lhagrid1"):
        mpdf.mkPDF("Coarse/0", dirname=str(tmp_path))


def test_on_demand_process_library(tmp_path):
    """A light-line six-point process of procgen_lines compiles for sm_100a into a loadable C-ABI library with the
    right metadata (no kernel is launched here; the library is built outside the package tree)."""
    import shutil

    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not available")
    from madflow_b200 import _runtime as rt
    from madflow_b200 import procgen_lines

    ir = procgen_lines.process_ir("1_gu_ttxug")
    src = tmp_path / "proc.cu"
    src.write_text(codegen.emit_process_source(ir))
    out = codegen.compile_source(str(src), str(tmp_path / "libmfp_1_gu_ttxug.so"))
    lib = rt.ProcessLib(out)
    assert lib.name == "1_gu_ttxug"
    assert (lib.info.nexternal, lib.info.ncomb, lib.info.ncolor, lib.info.ndiags, lib.info.ndim) == (6, 64, 12, 36, 18)
    assert lib.info.denominator == 96.0 and lib.info.flops_per_event == codegen.flops_per_event(ir)
    assert lib.coupling_names == ["GC_10", "GC_11", "GC_12"] and lib.helicities == ir["helicities"]
    assert lib.variant == "hp"      # 96 calls: only the helicity-parallel flavour is compiled
    for n in _declared("madflow_b200_process.h"):
        assert hasattr(lib.lib, n), n
