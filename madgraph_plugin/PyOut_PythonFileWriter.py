"""File writer that does not split lines (reference: madgraph_plugin/PyOut_PythonFileWriter.py:11-20)."""
from ._mg5 import HAVE_MG5

if HAVE_MG5:
    import madgraph.iolibs.file_writers as file_writers

    class PyOutPythonWriter(file_writers.FileWriter):
        def write_line(self, line):
            return ["%s\n" % line]
else:

    class PyOutPythonWriter(object):
        def write_line(self, line):
            return ["%s\n" % line]
