"""`pyout` output plugin for MG5_aMC with a CUDA (sm_100a) backend.

Same registration surface as the reference's madgraph_plugin/__init__.py:29-49: MG5 reads
`new_output`, `new_cluster`, `new_interface` and the version gates from this module, and
`output pyout <dir>` instantiates `PyOutExporter`.
"""
import os
import sys

root_path = os.path.split(os.path.dirname(os.path.realpath(__file__)))[0]
if root_path not in sys.path:
    sys.path.insert(0, root_path)

from . import PyOut_exporter  # noqa: E402

# 1. new output mode: "output pyout PATH"
new_output = {"pyout": PyOut_exporter.PyOutExporter}
# 2. no new cluster support
new_cluster = {}
# 3. no new interface
new_interface = None

__author__ = "madflow_b200"
__version__ = (0, 1, 0)
minimal_mg5amcnlo_version = (2, 5, 0)
maximal_mg5amcnlo_version = (1000, 1000, 1000)
latest_validated_version = (2, 5, 0)
