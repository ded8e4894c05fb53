#!/usr/bin/env python3
"""Leading-order cross sections and event generation on the GPU -- the `madflow` command
(python_package/madflow/scripts/madflow_exec.py:243-530).

    python -m madflow_b200.scripts.madflow_exec --madgraph_process "g g > t t~ g g" --no_pdf -c -i 10 -f 5 \\
        --events_per_iteration 10000000 --histograms -o out/

Same arguments and the same warm-up / frozen-grid schedule as the reference (madflow_exec.py:491-510).  Without
`--no_pdf` the parton luminosity and alpha_s come from the LHAPDF set `--pdf` (member 0, madflow_exec.py:342),
interpolated on the GPU (madflow_b200.pdf); the set has to be on disk (`--pdf_dir`, PDFFLOW_DATA_PATH or
LHAPDF_DATA_PATH) -- none ships with this package and a missing set is refused loudly, not approximated.  MG5
process generation is not available: the process must be one of the compiled process libraries -- the built-in
g g > t t~ + n g, or anything exported through the `pyout` plugin's CUDA backend.
Extensions for long runs: --target_error stops the final iterations once the combined relative error is below
it, --unweighted_events sets the capacity of the on-device unweighting buffer.
"""
import argparse
import json
import logging
import sys
import tempfile
import time
from pathlib import Path

logger = logging.getLogger("madflow")

DEFAULT_PDF = "NNPDF31_nnlo_as_0118"


def process_library_name(madgraph_process):
    """'g g > t t~ g' -> '1_gg_ttxg' (MG5's shell_string for process number 1, PyOut_exporter.py:138); the light
    q q~ > t t~ processes share the library MG5 names after the first flavour, 1_uux_ttx."""
    ini, fin = madgraph_process.split(">")
    if "".join(ini.split()) in ("dd~", "ss~", "cc~") and "".join(fin.split()) == "tt~":
        return "1_uux_ttx"
    shell = lambda side: "".join(side.split()).replace("~", "x")
    return f"1_{shell(ini)}_{shell(fin)}"


def subprocess_libraries(madgraph_process):
    """The process libraries behind a process string: one, or for the hadronic processes the generator knows
    (`p p > t t~`) the subprocesses whose luminosity-weighted matrix elements are summed (madflow_exec.py:444-455)."""
    from madflow_b200.procgen import MULTI_PROCESSES

    key = " ".join(madgraph_process.split())
    return list(MULTI_PROCESSES.get(key, [process_library_name(madgraph_process)]))


def build_parser():
    arger = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    arger.add_argument("-v", "--verbose", help="Print extra info", action="store_true")
    arger.add_argument("-p", "--pdf", help="PDF set", type=str, default=DEFAULT_PDF)
    arger.add_argument("--pdf_dir", help="Directory holding the LHAPDF sets (default: PDFFLOW_DATA_PATH / LHAPDF_DATA_PATH)",
                       type=str, default=None)
    arger.add_argument("--no_pdf", help="Don't use a PDF for the initial state", action="store_true")
    arger.add_argument("--madgraph_process", help="Set the madgraph process to be run", type=str, default="g g > t t~")
    arger.add_argument("-m", "--massive_particles", help="Number of massive particles", type=int, default=2)
    arger.add_argument("-q", "--fixed_scale", help="Fix value of scale muR=muF (and alphas(q)), if this flag is not provided "
                       "take dynamical scale q2 = sum(mT)/2", type=float, nargs="?", const=91.46)
    arger.add_argument("-c", "--pt_cut", help="Enable a pt cut for the outgoing particles", type=float, nargs="?", const=30.0)
    arger.add_argument("--histograms", help="Generate LHE files/histograms", action="store_true")
    arger.add_argument("-i", "--iterations", help="Iterations of vegas to run", type=int, default=10)
    arger.add_argument("-f", "--frozen_iter", help="Iterations with frozen grid", type=int, default=0)
    arger.add_argument("--events_per_device", help="How many events to send to each device per launch", type=int)
    arger.add_argument("-o", "--output", help="Output folder", type=Path)
    arger.add_argument("--dry_run", help="Resolve the process and its library but don't run anything", action="store_true")
    arger.add_argument("--events_per_iteration", help="How many events to run per iteration", type=int, default=int(1e6))
    arger.add_argument("--custom_op", help="Accepted for compatibility: the CUDA kernels are the only implementation",
                       action="store_true")
    arger.add_argument("--target_error", help="Stop the final iterations at this relative error of the combined result",
                       type=float, default=None)
    arger.add_argument("--unweighted_events", help="Capacity of the unweighted-event buffer (--histograms)", type=int,
                       default=100_000)
    arger.add_argument("--dr_cut", help="Extension: Delta R > value between every pair of outgoing massless particles (the "
                       "reference's pt cuts leave their collinear singularity open)", type=float, nargs="?", const=0.4)
    arger.add_argument("--seed", type=int, default=4)
    arger.add_argument("--param_card", help="SLHA param_card with the masses and widths of the model (the reference reads "
                       "Cards/param_card.dat of the MG5 output folder; default: that file under --output if it exists, else "
                       "the SM values m_t = 173, Gamma_t = 1.4915)", type=Path, default=None)
    return arger


def madflow_main(args=None, quick_return=False):
    args = build_parser().parse_args(args)
    if quick_return:
        return args, None, None
    from madflow_b200 import config  # noqa: F401  (installs the "madflow" log handler, config.py:22-28)

    if args.verbose:
        logger.setLevel(logging.DEBUG)

    import torch

    from madflow_b200 import events as mfe
    from madflow_b200 import integrand as mfi
    from madflow_b200 import matrix as mfm
    from madflow_b200 import vegas as mfv
    from madflow_b200.lhe_writer import LheWriter

    names = subprocess_libraries(args.madgraph_process)
    name = names[0]
    for nm in names:
        if nm not in mfm.available_processes():
            from madflow_b200 import procgen_lines

            hint = (f"; it can be generated: python -m madflow_b200.build {nm}" if nm in procgen_lines.PROCESSES else "")
            raise SystemExit(f"process '{args.madgraph_process}' ({nm}) has no compiled process library; available: "
                             f"{mfm.available_processes()} (export it through the pyout plugin's CUDA backend{hint})")
    output_path = args.output if args.output is not None else Path(tempfile.mkdtemp(prefix="mad_"))
    output_path.mkdir(parents=True, exist_ok=True)
    if args.dry_run:
        logger.info("Process %s -> libraries %s; dry run, nothing executed", args.madgraph_process, names)
        return None, None, None

    pdf = None
    if not args.no_pdf:   # madflow_exec.py:339-342
        from madflow_b200 import pdf as mfpdf

        try:
            pdf = mfpdf.mkPDF(args.pdf + "/0", dirname=args.pdf_dir)
        except mfpdf.PDFError as e:
            raise SystemExit(str(e))
        logger.info("PDF set %s", pdf)

    dist = None
    rank, world = 0, 1
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank))))

    card = args.param_card
    if card is None and (output_path / "Cards" / "param_card.dat").exists():
        card = output_path / "Cards" / "param_card.dat"   # madflow_exec.py:130-136
    if card is not None:
        logger.info("Model parameters from %s", card)
    matrix, model = mfm.get_process(name, param_card=card)
    if args.histograms:
        matrix.set_variant("hp")   # the kernel pipeline that keeps the events in device memory
    nparticles = int(matrix.nexternal)
    non_massive = nparticles - args.massive_particles - 2
    param_masses = [float(m) for m in model.get_masses()]
    if len(param_masses) < args.massive_particles:
        param_masses *= args.massive_particles
    masses = param_masses[: args.massive_particles] + [0.0] * non_massive   # madflow_exec.py:362-370
    if args.fixed_scale is None:
        logger.info("Set variable muF=muR=sum(mT)/2")
    else:
        logger.info("Setting fixed muF=muR=%.2f GeV, alpha_s = %s", args.fixed_scale,   # madflow_exec.py:376-386
                    "0.118" if pdf is None else "alphasQ2 of the PDF set")
    cuts = []
    if args.dr_cut is not None:
        light = [i for i in range(2, nparticles) if masses[i - 2] == 0.0]
        cuts = [("dr", (i, j), args.dr_cut, None) for a, i in enumerate(light) for j in light[a + 1:]]
        logger.info("Applying Delta R > %.2f to the pairs %s", args.dr_cut, [c[1] for c in cuts])
    # a single-flavour process string against the all-flavour q q~ library of `p p`: only that flavour's luminosity
    one_flavour = {"u u~ > t t~": 2, "d d~ > t t~": 1, "s s~ > t t~": 3, "c c~ > t t~": 4}.get(" ".join(args.madgraph_process.split()))

    def integrand_of(mat, mod):
        return mfi.FusedIntegrand(mat, mod, sqrts=13e3, masses=masses, pt_cut=args.pt_cut, cuts=cuts, lab_frame=True,
                                  running=args.fixed_scale is None, alpha_s=0.118 if pdf is None else None, pdf=pdf,
                                  fixed_scale=args.fixed_scale if pdf is not None else None,
                                  initial_states=[(one_flavour, -one_flavour)] if one_flavour else None,
                                  mirror_initial_states=False if one_flavour else None)

    fi = integrand_of(matrix, model)
    if len(names) > 1:   # one event sample, the subprocesses summed per event
        fi = mfi.MultiProcessIntegrand([fi] + [integrand_of(*mfm.get_process(nm, param_card=card)) for nm in names[1:]])
    if args.events_per_device:
        fi.max_events_per_launch = args.events_per_device
    if nparticles >= 5 and args.frozen_iter == 0:
        logger.warning("With this many particles (> 5) it is recommended to run with frozen iterations")

    vegas = mfv.VegasFlow(fi.n_dim, args.events_per_iteration, seed=args.seed)
    vegas.compile(fi)
    if args.frozen_iter == 0:   # madflow_exec.py:491-510
        warmup_iterations, final_iterations = args.iterations // 2, args.iterations // 2
    else:
        warmup_iterations, final_iterations = max(args.iterations - args.frozen_iter, 2), args.frozen_iter
    t0 = time.time()
    logger.info("Running %d warm-up iterations of %d events each", warmup_iterations, args.events_per_iteration)
    sink = None
    if args.histograms:
        # the unweighting threshold is fixed ONCE, from the weights of the last warm-up iteration (adapted grid), and
        # holds for the whole final run: one sample, one threshold
        vegas.run_integration(warmup_iterations - 1)
        top = 2  # the first outgoing particle: the top quark in the built-in processes (compare_mg5_hists.py:31-32)
        sink = mfe.EventSink(fi, histograms=[mfe.Histogram("pt", top, 0.0, 300.0, 50), mfe.Histogram("eta", top, -4.0, 4.0, 50)],
                             unweight=True, capacity=args.unweighted_events, seed=args.seed, collect_only=True)
        vegas.run_integration(1)
        logger.info("Unweighting threshold frozen at %.6g (largest warm-up weight %.6g)", sink.freeze_threshold() or 0.0,
                    sink.max_weight)
    else:
        vegas.run_integration(warmup_iterations)
    if args.frozen_iter > 0:
        vegas.freeze_grid()
    logger.info("Running %d iterations of %d events each%s", final_iterations, vegas.events_per_run,
                " with the grid frozen" if args.frozen_iter > 0 else "")

    results = []
    n_final = 0
    for _ in range(final_iterations):
        results.append(vegas.run_iteration())
        n_final += 1
        res, err, chi2 = mfv.combine_iterations(results)
        logger.info("Result for final iteration %d: %.6g +/- %.3g -> combined %.6g +/- %.3g (%.3g %%)", n_final,
                    results[-1][0], results[-1][1], res, err, 100 * err / abs(res))
        if args.target_error and err / abs(res) < args.target_error:
            break
    res, err, chi2 = mfv.combine_iterations(results)
    wall = time.time() - t0
    logger.info(" > Final results: %g +/- %g pb  (chi2/dof %.2f, %d events, %.1f s)", res, err, chi2,
                (warmup_iterations + n_final) * args.events_per_iteration, wall)

    proc_folder = None
    if args.histograms:
        proc_name = args.madgraph_process.replace(" ", "_").replace(">", "to").replace("~", "b")
        run = proc_name if world == 1 else f"{proc_name}_rank{rank}"
        for h in sink.histograms:
            h.allreduce()
        pdg = list((matrix.ir or {}).get("pdg") or [])
        if len(names) > 1 and pdg:
            # several subprocesses were summed per event and no channel is recorded: the beams are labelled as protons
            # (the reference writes the hadron-level process the same way), the final state as in the first subprocess
            pdg[:2] = [2212, 2212]
        with LheWriter(output_path, run, no_unweight=True, pdg=pdg or None) as lhe_writer:
            nkept = sink.write_lhe(lhe_writer, cross=res)
            frac, share = sink.overweight()
            if frac > 0.0:
                logger.info("%.2f %% of the kept events lie above the unweighting threshold and carry %.1f %% of the total "
                            "weight: they keep their own (larger) weight in the file", 100 * frac, 100 * share)
            lhe_writer.store_result((res, err))
            proc_folder = output_path / f"Events/{run}"
            lhe_writer.dump_result(proc_folder / "cross_err.txt")
        # the events of the sink are already unweighted on the device: the file is the unweighted sample
        (proc_folder / "weighted_events.lhe.gz").rename(proc_folder / "unweighted_events.lhe.gz")
        if rank == 0:
            hists = {f"{h.observable}_{h.particle}": {"edges": h.edges.tolist(), "dsigma_pb": h.values(n_final).tolist(),
                                                     "underflow_overflow_pb": [float(h.values(n_final, True)[0]),
                                                                               float(h.values(n_final, True)[-1])]}
                     for h in sink.histograms}
            (proc_folder / "histograms.json").write_text(json.dumps(hists, indent=1))
        logger.info("Written %d unweighted events%s, histograms and cross_err.txt to %s", nkept,
                    " (buffer full)" if sink.overflowed else "", proc_folder)
    if hasattr(fi, "release"):
        fi.release()   # the subprocess libraries share one grid size while summed: give it back
    if dist is not None:
        dist.destroy_process_group()
    return args, (res, err), proc_folder


def main():
    madflow_main()


if __name__ == "__main__":
    main()
