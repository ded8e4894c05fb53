"""Matrix-element exporter of the `pyout` plugin with the CUDA (sm_100a) backend.

Keeps the call protocol MG5_aMC drives (reference: madgraph_plugin/PyOut_exporter.py):
    PyOutExporter(dir_path)                         :112   (creates the directory)
    .pass_information_from_cmd(cmd)                 :117
    .generate_subprocess_directory(group, model)    :411   -> one matrix_<proc>.py per subprocess
    .convert_model(model, lorentz, couplings)       :486
    .finalize(matrix_elements, history, opts, flags):544   -> Cards/param_card.dat, proc_card_mg5.dat
and the class attributes MG5 inspects (:84-94).  What changes is the product: instead of TensorFlow
code, each subprocess yields
    <dir>/<proc>.json        the process IR (madflow_b200/process_ir.py)
    <dir>/<proc>.cu          one fused FP64 kernel (madflow_b200/codegen.py)
    <dir>/libmfp_<proc>.so   compiled with nvcc for sm_100a (skipped when MADFLOW_B200_NO_BUILD=1)
    <dir>/matrix_<proc>.py   `Matrix_<proc>` with the reference's attributes and smatrix signature,
                             calling the library through the C ABI; `get_model_param` as in the
                             reference's template (template_files/matrix_method_python.inc:36-41).

`matrix_element_to_ir` holds everything that touches the MG5 matrix-element object; it only uses
the calls listed in SURVEY.md Appendix G, so it is testable with stand-in objects.
"""
import fractions
import logging
import os
import shutil

from . import PyOut_helas_call_writer as pyout_helas_call_writer
from ._mg5 import HAVE_MG5, MG5DIR, MadGraph5Error, export_python, export_v4, misc

plugin_path = os.path.dirname(os.path.realpath(__file__))
logger = logging.getLogger('PyOut_plugin.MEExporter')
pjoin = os.path.join


class PyOutExporterError(MadGraph5Error):
    """Error from the PyOut exporter."""


def jamp_coefficient(ff_number, frac, is_imaginary, Nc_power, Nc_value=3):
    """(re, im) of one JAMP coefficient: ff_number * frac * Nc^Nc_power [* i]
    (reference `coeff`, PyOut_exporter.py:44-75)."""
    total = ff_number * fractions.Fraction(frac) * fractions.Fraction(Nc_value) ** Nc_power
    val = float(total)
    return (0.0, val) if is_imaginary else (val, 0.0)


def matrix_element_to_ir(matrix_element, helas_calls, coupling_defs=None):
    """MG5 HelasMatrixElement (or a stand-in) + its HELAS call lines -> process IR."""
    process = matrix_element.get('processes')[0]
    name = process.shell_string()
    nexternal, ninitial = matrix_element.get_nexternal_ninitial()
    calls = pyout_helas_call_writer.parse_helas_calls(helas_calls)
    # JAMP table (PyOut_exporter.py:334-375)
    jamp = []
    for coeff_list in matrix_element.get_color_amplitudes():
        terms = []
        for (coefficient, amp_number) in coeff_list:
            re, im = jamp_coefficient(coefficient[0], coefficient[1], coefficient[2], coefficient[3])
            terms.append((amp_number - 1, re, im))
        jamp.append(terms)
    # colour matrix (PyOut_exporter.py:301-324)
    cm = matrix_element.get('color_matrix')
    if not cm:
        color_num, color_denom = [[1]], [1]
    else:
        color_denom = [int(d) for d in cm.get_line_denominators()]
        color_num = [[int(v) for v in cm.get_line_numerators(i, d)] for i, d in enumerate(color_denom)]
    # masses/widths and couplings, sorted (PyOut_exporter.py:378-407)
    params = sorted({c[k] for c in calls for k in ("mass", "width") if k in c} - {"ZERO"})
    couplings = sorted({c["coup"] for c in calls if "coup" in c})
    legs = process['legs'] if hasattr(process, '__getitem__') else []
    ir = {
        "name": name,
        "process": getattr(process, 'nice_string', lambda: name)(),
        "nexternal": nexternal, "ninitial": ninitial,
        "ndiags": len(matrix_element.get('diagrams')),
        "ncomb": matrix_element.get_helicity_combinations(),
        "nwavefuncs": matrix_element.get_number_of_wavefunctions(),
        "helicities": [list(h) for h in matrix_element.get_helicity_matrix()],
        "denominator": matrix_element.get_denominator_factor(),
        "params": params, "couplings": couplings,
        "initial_states": [list(p.get_initial_ids()) for p in matrix_element.get('processes')],
        "mirror_initial_states": bool(matrix_element.get_mirror_processes()),
        "calls": calls, "jamp": jamp, "color_num": color_num, "color_denom": color_denom,
    }
    if coupling_defs:
        ir["coupling_defs"] = {k: list(v) for k, v in coupling_defs.items() if k in couplings}
    return ir


def coupling_power_law(expr, namespace=None):
    """If a UFO coupling expression is c * G^k (all the QCD ones are) return (Re c, Im c, k), else None."""
    import cmath

    ns = {"cmath": cmath, "complex": complex, "complexconjugate": lambda z: z.conjugate()}
    ns.update(namespace or {})
    try:
        f1 = complex(eval(expr, dict(ns, G=1.0)))
        f2 = complex(eval(expr, dict(ns, G=2.0)))
        f3 = complex(eval(expr, dict(ns, G=3.0)))
    except Exception:
        return None
    if f1 == 0:
        return None
    for k in range(0, 7):
        if abs(f2 - f1 * 2**k) < 1e-12 * abs(f1) * 2**k and abs(f3 - f1 * 3**k) < 1e-12 * abs(f1) * 3**k:
            return (f1.real, f1.imag, k)
    return None


def attach_aloha_routines(ir, routines):
    """ir["aloha_routines"] = the device source of every vertex routine of the call list that the CUDA backend does not
    have hand-written (madflow_b200.codegen.BUILTIN_OPS).  `routines`: {name: text} or a callable producing it (the
    ALOHA computation is only started when something is missing)."""
    from madflow_b200 import codegen

    missing = sorted({c["op"] for c in ir["calls"]} - codegen.BUILTIN_OPS)
    if not missing:
        return ir
    if callable(routines):
        routines = routines()
    absent = [op for op in missing if op not in routines]
    if absent:
        raise PyOutExporterError("ALOHA did not produce the routines %s" % absent)
    ir["aloha_routines"] = {op: routines[op] for op in missing}
    return ir


MATRIX_TEMPLATE = open(pjoin(plugin_path, "template_files", "matrix_method_cuda.inc")).read()


def write_process_files(ir, dir_path, model_parameter_lines, build=None):
    """Write <proc>.json, <proc>.cu, matrix_<proc>.py (and build the library) into dir_path."""
    from madflow_b200 import codegen, process_ir

    proc = ir["name"]
    with open(pjoin(dir_path, f"{proc}.json"), "w") as fh:
        fh.write(process_ir.dumps(ir))
    with open(pjoin(dir_path, f"{proc}.cu"), "w") as fh:
        fh.write(codegen.emit_process_source(ir))
    if build is None:
        build = os.environ.get("MADFLOW_B200_NO_BUILD", "0") != "1"
    if build:
        codegen.compile_source(pjoin(dir_path, f"{proc}.cu"), pjoin(dir_path, f"libmfp_{proc}.so"))
    text = MATRIX_TEMPLATE % {
        "process_string": proc,
        "root_path": MG5DIR,
        "paramnames_const": ",".join('"%s"' % p for p in ir["params"]),
        "paramnames_func": ",".join('"%s"' % c for c in ir["couplings"]),
        "paramtuple_const": ",".join("float(%s)" % p for p in ir["params"]),
        "paramtuple_func": ",".join(ir["couplings"]),
        "model_parameters": model_parameter_lines,
        "nexternal": ir["nexternal"], "ndiags": ir["ndiags"], "ncomb": ir["ncomb"],
    }
    with open(pjoin(dir_path, f"matrix_{proc}.py"), "w") as fh:
        fh.write(text)
    return pjoin(dir_path, f"matrix_{proc}.py")


class PyOutExporter(export_python.ProcessExporterPython):
    """Built on MG5's Python exporter exactly like the reference's (PyOut_exporter.py:78)."""

    check = True
    exporter = 'v4'
    output = 'dir'
    grouped_mode = False
    sa_symmetry = False

    PS_dependent_key = ['aS', 'MU_R']

    def __init__(self, dir_path, *args, **opts):
        os.mkdir(dir_path)
        self.dir_path = dir_path
        self.params_ext, self.params_dep, self.params_indep = [], [], []
        self.coups_dep, self.coups_indep = [], []
        self.me_names, self.proc_names, self.mass_lists = [], [], []
        self.refactorized = False

    def pass_information_from_cmd(self, cmd):
        self.proc_defs = cmd._curr_proc_defs
        self.model = cmd._curr_model

    # -- model parameters: same generated text as the reference (PyOut_exporter.py:198-226, 563-582)
    def get_model_parameter_lines(self, ir):
        lines = '    # External (param_card) parameters\n    '
        lines += "\n    ".join("%(param)s = param_card['%(block)s'].get(%(id)s).value" %
                               {"param": p.name, 'block': p.lhablock, 'id': p.lhacode[0]} for p in self.params_ext)
        lines += '\n\n    #PS-independent parameters\n'
        for p in self.params_indep:
            lines += '    %s = %s\n' % (p.name, p.expr)
        lines += '\n    #PS-dependent parameters\n'
        for p in self.params_dep:
            if p.name == "mdl_sqrt__aS":
                lines += '    %s = %s\n' % (p.name, p.expr)
            else:
                lines += '    %s = lambda G: complex_me(%s)\n' % (p.name, p.expr)
        dep = '\n    # PS-dependent couplings\n'
        for c in self.coups_dep:
            if c.name in ir["couplings"]:
                dep += '    %s = lambda G: complex_me(%s)\n' % (c.name, c.expr)
        for p in self.params_dep:
            if p.name != "mdl_sqrt__aS":
                dep = dep.replace(p.name, '%s(G)' % p.name)
        for c in self.coups_indep:
            if c.name in ir["couplings"]:
                lines += '    %s = lambda G: complex_me(%s) + 0*G\n' % (c.name, c.expr)
        return (lines + dep).replace('cmath', 'np')

    def coupling_defs(self, ir):
        defs = {}
        for c in list(self.coups_dep) + list(self.coups_indep):
            if c.name in ir["couplings"]:
                law = coupling_power_law(c.expr)
                if law is not None:
                    defs[c.name] = law
        return defs

    def write_alohas(self, matrix_element):
        """The ALOHA routines of the matrix elements as CUDA device source, by routine name (reference:
        PyOut_exporter.py:507-540 writes aloha_<proc>.py with TensorFlow code).  Only the routines that are not
        hand-written in csrc/aloha_sm.cuh end up in the process's translation unit (attach_aloha_routines)."""
        from . import PyOut_create_aloha as pyout_create_aloha
        from ._mg5 import aloha_writers, create_aloha

        aloha_model = create_aloha.AbstractALOHAModel(os.path.basename(self.model.get('modelpath')))
        aloha_model.add_Lorentz_object(self.model.get('lorentz'))
        wanted_lorentz = set(sum([me.get_used_lorentz() for me in self.matrix_elements], []))
        if wanted_lorentz:
            aloha_model.compute_subset(list(wanted_lorentz))
        else:
            aloha_model.compute_all(save=False)
        routines = {}
        for k, v in aloha_model.items():
            routine = pyout_create_aloha.PyOutAbstractRoutine(v)
            routines[aloha_writers.get_routine_name(abstract=routine)] = routine.write(output_dir='')
        proc = matrix_element.get('processes')[0].shell_string()
        with open(pjoin(self.dir_path, 'aloha_%s.cuh' % proc), 'w') as fout:
            fout.write(pyout_create_aloha.CUDA_PROLOGUE + "\n".join(routines.values()))
        return routines

    def generate_subprocess_directory(self, subproc_group, fortran_model, me=None):
        self.helas_writer = pyout_helas_call_writer.PyOutUFOHelasCallWriter(self.model)
        super(PyOutExporter, self).__init__(subproc_group, self.helas_writer)
        self.refactorize()
        for matrix_element in self.matrix_elements:
            calls = self.helas_call_writer.get_matrix_element_calls(matrix_element, False)
            ir = matrix_element_to_ir(matrix_element, calls)
            ir["coupling_defs"] = {k: list(v) for k, v in self.coupling_defs(ir).items()}
            attach_aloha_routines(ir, lambda: self.write_alohas(matrix_element))
            write_process_files(ir, self.dir_path, self.get_model_parameter_lines(ir))
            proc = ir["name"]
            self.me_names.append('matrix_%s' % proc)
            self.proc_names.append(proc)
            model = matrix_element.get('processes')[0]['model']
            self.mass_lists.append([model.get_particle(l['id'])['mass']
                                    for l in matrix_element.get('processes')[0]['legs']])

    def convert_model(self, model, wanted_lorentz=[], wanted_couplings=[]):
        """Copy the UFO model next to the output (PyOut_exporter.py:486-504)."""
        target = pjoin(self.dir_path, 'bin', 'internal', 'ufomodel')
        shutil.rmtree(target, ignore_errors=True)
        shutil.copytree(model.get('modelpath'), target, ignore=shutil.ignore_patterns('*.pyc', '*.dat', '*.py~'))

    def write_leading_order_wrapper(self, outfile, history):
        """leading_order.py: a runnable integration script for the exported process
        (PyOut_exporter.py:442-468, template_files/leading_order.inc)."""
        imports, tree_level, masses = '', '\n', '\n'
        for name, proc, mass in zip(self.me_names, self.proc_names, self.mass_lists):
            info = {'me': name, 'proc': proc, 'me_class': name.capitalize()}
            imports += 'from %(me)s import %(me_class)s, get_model_param as model_%(proc)s\n' % info
            tree_level += '    "%(proc)s": (%(me_class)s, model_%(proc)s),\n' % info
            masses += ('    "%s": [' % proc) + ", ".join('"%s"' % m for m in mass) + '],\n'
        try:
            info_lines = "Generated with the madflow B200 backend and MG5_aMC v%s" % misc.get_pkg_info().get('version', '?')
        except Exception:
            info_lines = "Generated with the madflow B200 backend"
        hist = "\n".join(str(h) for h in history) if history else ""
        template = open(pjoin(plugin_path, 'template_files', 'leading_order_cuda.inc')).read()
        outfile.write(template % {'info_lines': info_lines, 'history': hist, 'matrix_element_imports': imports,
                                  'tree_level_keys': tree_level, 'masses': masses})

    def finalize(self, matrix_elements, history, mg5options, flaglist):
        """leading_order.py, Cards/param_card.dat and the command history (PyOut_exporter.py:544-559)."""
        with open(pjoin(self.dir_path, 'leading_order.py'), 'w') as fout:
            self.write_leading_order_wrapper(fout, history)
        cardpath = pjoin(self.dir_path, 'Cards')
        if not os.path.isdir(cardpath):
            os.mkdir(cardpath)
            export_v4.UFO_model_to_mg4.create_param_card_static(self.model, pjoin(cardpath, 'param_card.dat'))
        if history and os.path.isdir(cardpath):
            history.write(pjoin(cardpath, 'proc_card_mg5.dat'))

    def refactorize(self, wanted_couplings=[]):
        """Split parameters/couplings by alpha_s dependence (PyOut_exporter.py:586-632)."""
        if self.refactorized:
            return
        self.refactorized = True
        keys = sorted(self.model['parameters'].keys(), key=len)
        for key in keys:
            to_add = [o for o in self.model['parameters'][key] if o.name]
            if key == ('external',):
                self.params_ext += to_add
            elif any(k in key for k in self.PS_dependent_key):
                self.params_dep += to_add
            else:
                self.params_indep += to_add
        for key, coup_list in self.model['couplings'].items():
            sel = [c for c in coup_list if (not wanted_couplings or c.name in wanted_couplings)]
            if any(k in key for k in self.PS_dependent_key):
                self.coups_dep += sel
            else:
                self.coups_indep += sel
