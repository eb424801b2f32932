"""Covariance kernels: the reference's Kernel(Xi, Xj, ni, nj, hyper_deriv) plugin surface
(gptools/kernel/), with the closed forms evaluated by the CUDA library."""
from .core import *  # noqa: F401,F403
from .noise import *  # noqa: F401,F403
from .squared_exponential import *  # noqa: F401,F403
from .matern import *  # noqa: F401,F403
from .gibbs import *  # noqa: F401,F403
from .warping import *  # noqa: F401,F403
