"""Coupling model: numpy restatement of python_package/madflow/parameters.py.  TEST INFRASTRUCTURE."""
import numpy as np


def alphas_to_gs(alpha_s):
    """parameters.py:13-15: G = 2 sqrt(pi alpha_s) as complex."""
    return (2.0 * np.sqrt(np.pi * np.asarray(alpha_s, dtype=np.float64))).astype(np.complex128)


def sm_qcd_couplings(alpha_s):
    """The alpha_s-dependent couplings of models/sm the QCD processes use.  GC_10 = -G and
    GC_11 = iG are pinned by tests/mockup_debug_me.py:24-25; GC_12 = iG^2 is [EXT] models/sm."""
    G = alphas_to_gs(alpha_s)
    return {"GC_10": -G, "GC_11": 1j * G, "GC_12": 1j * G**2}


def frozen_alpha_s(alpha_s):
    """Model.freeze_alpha_s (parameters.py:44-53) evaluates at float_me([alpha_s]) -- a bare
    Python float in a list goes through float32."""
    return float(np.float32(alpha_s))
