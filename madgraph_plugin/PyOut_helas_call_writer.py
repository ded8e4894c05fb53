"""HELAS call writer of the `pyout` plugin.

The reference's writer (madgraph_plugin/PyOut_helas_call_writer.py:19-136) formats one line per
wavefunction/amplitude of the generated `matrix()`:

    w3 = ixxxxx(all_ps[:,3],mdl_MT,hel[3],float_me(-1))        external   (:40-54)
    w4= FFV1_1(w2,w0,GC_11,mdl_MT,mdl_WT)                       off-shell  (:101-120)
    amp0= FFV1_0(w3,w2,w4,GC_11)                                amplitude

The CUDA backend keeps exactly these lines (so MG5's `get_matrix_element_calls` machinery is used
unchanged) and `parse_helas_calls` turns them into the structured call list of the process IR
(madflow_b200/process_ir.py) from which the kernel is emitted.
"""
import re

from ._mg5 import HAVE_MG5, helas_call_writers, helas_objects, aloha_writers

_EXT = re.compile(r"^w(\d+)\s*=\s*([ivos]xxxxx)\(all_ps\[:,(\d+)\],(?:\s*([\w]+),\s*(hel\[(\d+)\]|4),)?\s*float_me\(([+-]?\d+)\)\)$")
_CALL = re.compile(r"^(w|amp)(\d+)\s*=\s*(\w+)\((.*)\)$")


def parse_helas_calls(lines):
    """Lines of the generated matrix() -> IR calls (see madflow_b200.process_ir)."""
    calls = []
    for raw in lines:
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        m = _EXT.match(line)
        if m:
            slot, op, leg, mass, _, _, nsf = m.groups()
            calls.append({"op": op, "out": int(slot), "leg": int(leg), "mass": mass or "ZERO", "nsf": int(nsf)})
            continue
        m = _CALL.match(line)
        if not m:
            raise ValueError(f"cannot parse HELAS call line: {raw!r}")
        kind, idx, name, args = m.groups()
        args = [a.strip() for a in args.split(",") if a.strip()]
        wfs = [int(a[1:]) for a in args if re.fullmatch(r"w\d+", a)]
        rest = [a for a in args if not re.fullmatch(r"w\d+", a)]
        if kind == "amp":
            if len(rest) != 1:
                raise ValueError(f"amplitude with {len(rest)} couplings is not supported: {raw!r}")
            coup = rest[0]
            calls.append({"op": name, "amp": int(idx), "in": wfs, "coup": coup.lstrip("-"),
                          "coup_sign": -1 if coup.startswith("-") else 1})
        else:
            if len(rest) != 3:
                raise ValueError(f"off-shell call needs (coupling, mass, width): {raw!r}")
            coup, mass, width = rest
            calls.append({"op": name, "out": int(idx), "in": wfs, "coup": coup.lstrip("-"),
                          "coup_sign": -1 if coup.startswith("-") else 1, "mass": mass, "width": width})
    return calls


class PyOutUFOHelasCallWriter(helas_call_writers.PythonUFOHelasCallWriter):
    """Same call formats as the reference's writer (PyOut_helas_call_writer.py:19-136)."""

    def generate_helas_call(self, argument, gauge_check=False):
        if not isinstance(argument, helas_objects.HelasWavefunction) and \
           not isinstance(argument, helas_objects.HelasAmplitude):
            raise self.PhysicsObjectError("get_helas_call must be called with wavefunction or amplitude")

        if isinstance(argument, helas_objects.HelasAmplitude) and argument.get('interaction_id') == 0:
            self.add_amplitude(argument.get_call_key(), lambda amp: "#")
            return

        if isinstance(argument, helas_objects.HelasWavefunction) and not argument.get('mothers'):
            name = helas_call_writers.HelasCallWriter.mother_dict[argument.get_spin_state_number()].lower()
            call = "w%d = " + name + 'x' * (6 - len(name)) + "(all_ps[:,%d],"
            scalar = argument.get('spin') == 1
            brst = gauge_check and argument.get('spin') == 3 and argument.get('mass') == 'ZERO'
            if not scalar:
                call += "%s, 4," if brst else "%s,hel[%d],"
            call += "float_me(%+d))"

            def sign(wf):
                if wf.is_boson():
                    return (-1) ** (wf.get('state') == 'initial')
                return -(-1) ** wf.get_with_flow('is_part')

            if scalar:
                call_function = lambda wf: call % (wf.get('me_id') - 1, wf.get('number_external') - 1, sign(wf))
            elif brst:
                call_function = lambda wf: call % (wf.get('me_id') - 1, wf.get('number_external') - 1, 'ZERO', sign(wf))
            else:
                call_function = lambda wf: call % (wf.get('me_id') - 1, wf.get('number_external') - 1,
                                                   wf.get('mass'), wf.get('number_external') - 1, sign(wf))
        else:
            outgoing = argument.find_outgoing_number() if isinstance(argument, helas_objects.HelasWavefunction) else 0
            lor = [str(l) for l in argument.get('lorentz')]
            flag = []
            if argument.needs_hermitian_conjugate():
                flag = ['C%d' % i for i in argument.get_conjugate_index()]
            arg = {'routine_name': aloha_writers.combine_name('%s' % lor[0], lor[1:], outgoing, flag, True),
                   'wf': ("w%%(%d)d," * len(argument.get('mothers'))) % tuple(range(len(argument.get('mothers')))),
                   'coup': ("%%(coup%d)s," * len(argument.get('coupling'))) % tuple(range(len(argument.get('coupling'))))}
            if isinstance(argument, helas_objects.HelasWavefunction):
                arg['out'], arg['mass'] = 'w%(out)d', "%(M)s,%(W)s"
            else:
                arg['coup'] = arg['coup'][:-1]
                arg['out'], arg['mass'] = 'amp%(out)d', ''
            call = '%(out)s= %(routine_name)s(%(wf)s%(coup)s%(mass)s)' % arg
            call_function = lambda wf: call % wf.get_helas_call_dict(index=0)

        if isinstance(argument, helas_objects.HelasWavefunction):
            if not gauge_check:
                self.add_wavefunction(argument.get_call_key(), call_function)
        else:
            self.add_amplitude(argument.get_call_key(), call_function)
        return call_function
