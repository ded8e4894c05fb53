"""Coupling model -- same interface as python_package/madflow/parameters.py (`Model` :18-98).

`constants` is a namedtuple of masses/widths (and alpha_s-independent couplings); `functions` a
namedtuple of callables G -> coupling.  evaluate(alpha_s) returns [constants..., couplings...]
with the couplings as complex128 CUDA tensors of shape (nevt,) -- or shape (1,) when frozen
(parameters.py:44-53) -- which is what Matrix.smatrix(all_ps, *params) takes.
"""
from itertools import chain

import numpy as np
import torch

from . import config


def _alphas_to_gs(alpha_s):
    """parameters.py:13-15."""
    return (2.0 * torch.sqrt(np.pi * alpha_s)).to(torch.complex128)


class Model:
    def __init__(self, constants, functions):
        self._tuple_constants = constants
        self._tuple_functions = functions
        self._constants = list(constants)
        self._to_evaluate = list(functions)
        self._frozen = []

    @property
    def frozen(self):
        """Whether the model is frozen for a given value of alpha_s or not"""
        return bool(self._frozen)

    def freeze_alpha_s(self, alpha_s):
        """parameters.py:44-53.  The reference evaluates at float_me([alpha_s]); a bare Python float
        in a list is float32-rounded by that cast, reproduced in constants mode "reference"."""
        if self.frozen:
            raise ValueError("The model is already frozen")
        if config.get_constants().mode == "reference":
            alpha_s = float(np.float32(alpha_s))
        self._frozen = self._evaluate(config.float_me([alpha_s]))

    def unfreeze(self):
        """Remove the frozen status"""
        self._frozen = []

    def _evaluate(self, alpha_s):
        alpha_s = config.float_me(alpha_s)
        gs = _alphas_to_gs(alpha_s)
        results = [torch.as_tensor(fun(gs), dtype=torch.complex128, device=gs.device) for fun in self._to_evaluate]
        if not results:
            return self._constants
        if not self._constants:
            return results
        return list(chain.from_iterable([self._constants, results]))

    def get_masses(self):
        """Get the masses that entered the model as constants (parameters.py:74-80)."""
        return [val for key, val in self._tuple_constants._asdict().items() if key.startswith("mdl_M")]

    def parse_parameter(self, parameter_name):
        """Parse a (constant) parameter given its string name (parameters.py:82-91)."""
        if parameter_name == "ZERO":
            return 0.0
        if hasattr(self._tuple_constants, parameter_name):
            return getattr(self._tuple_constants, parameter_name)
        if hasattr(self._tuple_functions, parameter_name):
            return getattr(self._tuple_functions, parameter_name)
        raise AttributeError(f"The model class does not contain parameter {parameter_name}")

    def evaluate(self, alpha_s=None):
        """Evaluate alpha_s; if the model is frozen returns the frozen values (parameters.py:93-98)."""
        if self.frozen:
            return self._frozen
        return self._evaluate(alpha_s)
