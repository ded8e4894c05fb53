"""Settings: dtypes, logging and numerical-constant mode.

Mirrors python_package/madflow/config.py of the reference (env-driven dtypes :16-34, logger
:49-61, complex helpers :66-79) with torch tensors in place of TensorFlow ones.  The hot path is
FP64/complex128 only: MADFLOW_FLOAT=32 is refused rather than silently honoured.
"""
import logging
import os

import torch

_log_level_idx = os.environ.get("MADFLOW_LOG_LEVEL", "2")
_float_env = os.environ.get("MADFLOW_FLOAT", "64")
_int_env = os.environ.get("MADFLOW_INT", "32")
if _float_env != "64":
    raise ValueError("madflow_b200 evaluates matrix elements in FP64 only (MADFLOW_FLOAT must be 64)")

DTYPE = torch.float64
DTYPEINT = torch.int32 if _int_env == "32" else torch.int64
DTYPECOMPLEX = torch.complex128

LOG_DICT = {"0": logging.ERROR, "1": logging.WARNING, "2": logging.INFO, "3": logging.DEBUG}
logger = logging.getLogger("madflow")
logger.setLevel(LOG_DICT.get(_log_level_idx, logging.INFO))
if not logger.handlers:
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("[%(levelname)s] (<madflow>) %(message)s"))
    logger.addHandler(_h)


def device():
    """The CUDA device of this process (LOCAL_RANK under torchrun)."""
    if not torch.cuda.is_available():
        raise RuntimeError("madflow_b200 needs a CUDA device: there is no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def float_me(x):
    return torch.as_tensor(x, dtype=DTYPE, device=device())


def int_me(x):
    return torch.as_tensor(x, dtype=DTYPEINT, device=device())


def complex_me(x):
    return torch.as_tensor(x, dtype=DTYPECOMPLEX, device=device())


def complex_tf(real, imag):
    return torch.complex(float_me(real), float_me(imag))


def run_eager(flag=True):
    """Kept for API compatibility (reference: pdfflow.configflow.run_eager); nothing is traced."""
    return None


class Constants:
    """Numerical constants of the path.

    mode "reference" (default) uses the values the reference effectively computes with -- it
    passes bare Python floats through tf.cast, which rounds them to float32 first:
      SQH     wavefunctions_flow.py:10   0.7071067690849304
      PI      phasespace.py:16           3.1415927410125732
      ACC     phasespace.py:17           1.000000013351432e-10
      GEV2PB  phasespace.py:316          389379360.0
    mode "exact" uses the true double-precision values.  Select with MADFLOW_B200_CONSTANTS.
    """

    def __init__(self, mode=None):
        import numpy as np

        mode = mode or os.environ.get("MADFLOW_B200_CONSTANTS", "reference")
        if mode not in ("reference", "exact"):
            raise ValueError("constants mode must be 'reference' or 'exact'")
        self.mode = mode
        if mode == "reference":
            self.SQH = float(np.float32(np.sqrt(np.float32(0.5))))
            self.PI = float(np.float32(np.pi))
            self.ACC = float(np.float32(1e-10))
            self.GEV2PB = float(np.float32(389379365.6))
        else:
            self.SQH = float(np.sqrt(0.5))
            self.PI = float(np.pi)
            self.ACC = 1e-10
            self.GEV2PB = 389379365.6


CONSTANTS = Constants()


def set_constants(mode):
    global CONSTANTS
    CONSTANTS = Constants(mode)
    return CONSTANTS


def get_constants():
    return CONSTANTS
