"""Access to MG5_aMC from the plugin, with importable stand-ins when MG5 is absent.

Inside MG5 (`PLUGIN/pyout` on its path) the real classes are used.  Outside -- in this
repository's tests, where MG5_aMC is not installed -- the base classes are plain `object`
subclasses so that the MG5-independent logic of the plugin (IR construction, file writing) can be
exercised with duck-typed stand-ins of the MG5 objects (SURVEY.md Appendix G lists the calls made).
"""
HAVE_MG5 = True
try:
    from madgraph import MadGraph5Error, MG5DIR
    import madgraph.iolibs.export_python as export_python
    import madgraph.iolibs.helas_call_writers as helas_call_writers
    import madgraph.iolibs.export_v4 as export_v4
    import madgraph.core.helas_objects as helas_objects
    import madgraph.various.misc as misc
    import aloha
    import aloha.create_aloha as create_aloha
    import aloha.aloha_writers as aloha_writers
except ImportError:  # pragma: no cover - exercised only without MG5
    HAVE_MG5 = False
    MG5DIR = ""

    class MadGraph5Error(Exception):
        pass

    class _NS:
        pass

    export_python = _NS()
    export_python.ProcessExporterPython = object
    helas_call_writers = _NS()
    helas_call_writers.PythonUFOHelasCallWriter = object
    export_v4 = helas_objects = misc = aloha = create_aloha = None
    aloha_writers = _NS()
    aloha_writers.ALOHAWriterForCPP = object
