"""gptools_b200 -- B200-native (sm_100a) implementation of the gptools GP likelihood / prediction hot path
behind the reference's GaussianProcess / Kernel API.  See DESIGN.md."""
__version__ = "0.1.0"

from .error_handling import GPArgumentError, GPImpossibleParamsError  # noqa: F401
from .utils import *  # noqa: F401,F403
from .kernel import *  # noqa: F401,F403
from .mean import *  # noqa: F401,F403
from .gaussian_process import GaussianProcess, Constraint  # noqa: F401
