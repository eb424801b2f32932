"""gptools_b200 (bootstrap import; full API wired below once the host modules exist)."""
__version__ = "0.1.0"
from .error_handling import GPArgumentError, GPImpossibleParamsError  # noqa: F401
