"""Build every CUDA library of the package for sm_100a, in-tree (madflow_b200/lib/*.so)."""
import os
import subprocess
import sys

from . import codegen, process_ir

HERE = os.path.dirname(os.path.abspath(__file__))


def builtin_irs():
    """Every process compiled into the package: the pinned g g > t t~, g g > t t~ + 1..3 g, q q~ > t t~, the
    light-line five-point processes (`p p > t t~ j`) and the six-point subprocesses of `p p > t t~ j j` (light-line
    and four-quark, from procgen_lines)."""
    irs = [process_ir.gg_ttx_pinned()]
    try:
        from . import procgen, procgen_lines

        irs += procgen.builtin_irs()
        have = {ir["name"] for ir in irs}
        irs += [procgen_lines.process_ir(n) for n in procgen.MULTI_PROCESSES["p p > t t~ j j"] if n not in have]
    except ImportError:
        pass
    return irs


def build_core(verbose=False):
    os.makedirs(codegen.LIBDIR, exist_ok=True)
    out = os.path.join(codegen.LIBDIR, "libmadflow_b200.so")
    src = os.path.join(codegen.CSRC, "core.cu")
    if os.path.exists(out) and os.path.getmtime(out) >= codegen._newest_header_mtime():
        return out
    cmd = ["nvcc"] + codegen.NVCC_FLAGS + ["-I", codegen.CSRC, "-o", out, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for core.cu:\n{res.stdout}\n{res.stderr}")
    return out


def build_named(name, verbose=False):
    """Compile one process of the light-line generator (procgen_lines.PROCESSES) that is not built by default,
    e.g. build_named("1_uux_ttxgg"); afterwards matrix.get_process(name) and the `madflow` command find it."""
    from . import procgen_lines

    return codegen.build_process(procgen_lines.process_ir(name), verbose=verbose)


def build_all(verbose=False, jobs=None):
    """Compile the core library and every built-in process, `jobs` nvcc processes at a time."""
    from concurrent.futures import ThreadPoolExecutor

    jobs = jobs or max(1, min(8, os.cpu_count() or 1))
    irs = builtin_irs()
    with ThreadPoolExecutor(max_workers=jobs) as pool:
        core = pool.submit(build_core, verbose)
        procs = [pool.submit(codegen.build_process, ir, None, verbose) for ir in irs]
        return [core.result()] + [f.result() for f in procs]


if __name__ == "__main__":
    names = [a for a in sys.argv[1:] if not a.startswith("-")]
    for lib in ([build_named(n, "-v" in sys.argv) for n in names] if names else build_all("-v" in sys.argv)):
        print(lib)
