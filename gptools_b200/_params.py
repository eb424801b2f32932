"""Parameter container shared by kernels and mean functions: values, fixed flags, names, hyperprior,
and the free-parameter views the GaussianProcess uses.  Semantics follow the constructor blocks of the
reference's Kernel (kernel/core.py:154-218) and MeanFunction (mean.py:88-170) and their accessor
properties (kernel/core.py:259-354)."""
import inspect
import warnings

import numpy as np

from .error_handling import GPArgumentError
from .utils import IndependentJointPrior, MaskedBounds, UniformJointPrior


def count_fun_params(fun, skip):
    """Number of positional parameters of ``fun`` after its first ``skip`` arguments (keyword arguments
    with defaults excluded); None when ``fun`` takes *args (mean.py:92-104, kernel/gibbs.py:276-283)."""
    target = fun
    extra = 0
    try:
        spec = inspect.getfullargspec(target)
    except TypeError:
        spec = inspect.getfullargspec(target.__call__)
        extra = 1
    if inspect.ismethod(target) or (not inspect.isfunction(target) and not inspect.isbuiltin(target) and extra == 0
                                    and hasattr(target, "__call__") and not inspect.isroutine(target)):
        extra = 1
    if spec.varargs is not None:
        return None
    n = len(spec.args) - skip - extra
    if spec.defaults is not None:
        n -= len(spec.defaults)
    return n


class ParamHolder(object):
    """Mixin: ``params``, ``fixed_params``, ``param_names``, ``hyperprior`` and the free-* views."""

    def _init_params(self, num_params, initial_params, fixed_params, param_bounds, param_names, enforce_bounds,
                     hyperprior, warn_default_bounds=False, arg_error=GPArgumentError):
        if num_params < 0 or not isinstance(num_params, (int, np.integer)):
            raise ValueError("num_params must be an integer >= 0!")
        self.num_params = int(num_params)
        if param_names is None:
            param_names = [''] * self.num_params
        elif len(param_names) != self.num_params:
            raise ValueError("param_names must be a list of length num_params!")
        self.param_names = np.asarray(param_names, dtype=str)
        self.enforce_bounds = enforce_bounds
        if initial_params is None:
            if fixed_params is not None:
                raise arg_error("Must pass explicit parameter values if fixing parameters!")
            initial_params = np.ones(self.num_params, dtype=float)
            fixed_params = np.zeros(self.num_params, dtype=float)
        else:
            if len(initial_params) != self.num_params:
                raise ValueError("Length of initial_params must be equal to num_params!")
            if fixed_params is None:
                fixed_params = np.zeros(self.num_params, dtype=float)
            elif len(fixed_params) != self.num_params:
                raise ValueError("Length of fixed_params must be equal to num_params!")
        self.fixed_params = np.asarray(fixed_params, dtype=bool)
        if param_bounds is None and hyperprior is None:
            if warn_default_bounds and (~self.fixed_params).any():
                warnings.warn("Neither param_bounds nor hyperprior were specified when creating the kernel, "
                              "defaults may not be appropriate for your data.")
            param_bounds = self.num_params * [(0.0, 1e16)]
        elif param_bounds is not None and len(param_bounds) != self.num_params:
            raise ValueError("Length of param_bounds must be equal to num_params!")
        if hyperprior is None:
            hyperprior = UniformJointPrior(param_bounds)
        else:
            try:
                iter(hyperprior)
            except TypeError:
                pass
            else:
                if len(hyperprior) != self.num_params:
                    raise ValueError("If hyperprior is a list its length must be equal to num_params!")
                hyperprior = IndependentJointPrior(hyperprior)
        self.params = np.array(initial_params, dtype=float)
        self.hyperprior = hyperprior

    # -- bounds live in the hyperprior (kernel/core.py:212-218) --
    @property
    def param_bounds(self):
        return self.hyperprior.bounds

    @param_bounds.setter
    def param_bounds(self, value):
        self.hyperprior.bounds = value

    def set_hyperparams(self, new_params):
        """Set the FREE hyperparameters (kernel/core.py:259-287)."""
        new_params = np.array(new_params, dtype=float)
        if len(new_params) != len(self.free_params):
            raise ValueError("Length of new_params must be {:d}!".format(len(self.free_params)))
        if self.enforce_bounds:
            for idx, (val, bound) in enumerate(zip(new_params, self.free_param_bounds)):
                if bound[0] is not None and val < bound[0]:
                    new_params[idx] = bound[0]
                elif bound[1] is not None and val > bound[1]:
                    new_params[idx] = bound[1]
        self.params[~self.fixed_params] = new_params

    @property
    def num_free_params(self):
        return int(np.sum(~self.fixed_params))

    @property
    def free_param_idxs(self):
        return np.arange(0, self.num_params)[~self.fixed_params]

    @property
    def free_params(self):
        return MaskedBounds(self.params, self.free_param_idxs)

    @free_params.setter
    def free_params(self, value):
        self.params[self.free_param_idxs] = np.asarray(value, dtype=float)

    @property
    def free_param_bounds(self):
        return MaskedBounds(self.hyperprior.bounds, self.free_param_idxs)

    @free_param_bounds.setter
    def free_param_bounds(self, value):
        for i, v in zip(self.free_param_idxs, value):
            self.hyperprior.bounds[i] = v

    @property
    def free_param_names(self):
        return MaskedBounds(self.param_names, self.free_param_idxs)

    @free_param_names.setter
    def free_param_names(self, value):
        self.param_names = np.asarray(self.param_names, dtype=str)
        self.param_names[~self.fixed_params] = value
