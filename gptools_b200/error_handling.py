"""Exception classes (same names as the reference's gptools/error_handling.py)."""


class GPArgumentError(Exception):
    """Raised when an incorrect combination of keyword arguments is given."""


class GPImpossibleParamsError(Exception):
    """Raised when the hyperparameters have zero prior probability."""
