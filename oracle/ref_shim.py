"""TEST INFRASTRUCTURE -- not product code.

Import the UNMODIFIED reference package (``/root/reference/gptools``) in this
container under a small compatibility shim, so that golden vectors can be
generated from the reference itself (``tests/golden/make_golden.py``) and the
numpy restatement in ``oracle/gp_oracle.py`` can be pinned against it.

The reference was written for Python 2.7 / scipy 0.14 (README.rst:8); on
Python 3.12 / scipy 1.18 it needs:

1. the numpy names it reaches through the ``scipy`` namespace
   (``scipy.array``, ``scipy.zeros`` ... used on nearly every line of
   gaussian_process.py / kernel/*.py);
2. ``scipy.misc.factorial`` (kernel/rational_quadratic.py:28, utils.py:1146);
3. ``inspect.getargspec`` (kernel/gibbs.py:279, kernel/core.py:866, mean.py:94);
4. stub ``matplotlib`` modules (unguarded import, gaussian_process.py:37-38);
5. the Cython module ``gptools.kernel._matern`` (kernel/_matern.pyx): provided
   here by the reference's own ``matern.c`` compiled into
   ``oracle/_ref/libmatern52_ref.so`` (see oracle/Makefile) and driven through
   ctypes -- same per-pair C function, same loop order.

Nothing here is importable on the GPU box (``/root/reference`` does not exist
there); only ``tests/golden/make_golden.py`` and the ``-m "not gpu"`` pinning
tests (which skip when the reference is absent) use it.
"""
import ctypes
import inspect
import os
import sys
import types
import warnings

import numpy as np

REF_ROOT = os.environ.get("GPTOOLS_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "gptools"))


def _install_scipy_numpy_aliases():
    import scipy
    import scipy.special
    for name in dir(np):
        if name.startswith("_"):
            continue
        if not hasattr(scipy, name):
            try:
                setattr(scipy, name, getattr(np, name))
            except Exception:
                pass
    # removed numpy aliases the reference still spells out
    for name, val in (("float", float), ("int", int), ("bool", bool), ("complex", complex)):
        if not hasattr(scipy, name):
            setattr(scipy, name, val)
    if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "factorial"):
        misc = types.ModuleType("scipy.misc")
        misc.factorial = scipy.special.factorial
        sys.modules["scipy.misc"] = misc
        scipy.misc = misc


def _install_py2_builtins():
    """``long`` (kernel/warping.py:127) is a Python-2 builtin; under Python 3 it is ``int``."""
    import builtins
    if not hasattr(builtins, "long"):
        builtins.long = int


def _install_getargspec():
    if not hasattr(inspect, "getargspec"):
        def getargspec(func):
            fs = inspect.getfullargspec(func)
            return (fs.args, fs.varargs, fs.varkw, fs.defaults)
        inspect.getargspec = getargspec


class _Anything(types.ModuleType):
    """Module stub whose every attribute is another stub (matplotlib stand-in)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return self


def _install_matplotlib_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.widgets", "matplotlib.gridspec",
                 "matplotlib.patches", "matplotlib.cm", "matplotlib.colors", "matplotlib.ticker",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.axes_grid1"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)


def _install_matern_module():
    """``gptools.kernel._matern`` backed by the reference's own C source."""
    so = os.path.join(_HERE, "_ref", "libmatern52_ref.so")
    if not os.path.exists(so):
        raise RuntimeError("oracle/_ref/libmatern52_ref.so missing: run `make -C oracle ref`")
    lib = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int32)
    lib.matern52_pairs.argtypes = [dp, dp, ip, ip, ctypes.c_int64, ctypes.c_int32, dp, dp]
    lib.matern52_pairs.restype = None

    def _matern52(Xi, Xj, ni, nj, var):
        Xi = np.ascontiguousarray(Xi, dtype=np.float64)
        Xj = np.ascontiguousarray(Xj, dtype=np.float64)
        ni = np.ascontiguousarray(ni, dtype=np.int32)
        nj = np.ascontiguousarray(nj, dtype=np.int32)
        var = np.ascontiguousarray(var, dtype=np.float64)
        n, d = Xi.shape
        out = np.zeros(n, dtype=np.float64)
        lib.matern52_pairs(Xi.ctypes.data_as(dp), Xj.ctypes.data_as(dp), ni.ctypes.data_as(ip),
                           nj.ctypes.data_as(ip), n, d, var.ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out

    mod = types.ModuleType("gptools.kernel._matern")
    mod._matern52 = _matern52
    sys.modules["gptools.kernel._matern"] = mod
    return mod


_ref_module = None


def load_reference():
    """Return the reference ``gptools`` package (cached)."""
    global _ref_module
    if _ref_module is not None:
        return _ref_module
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_scipy_numpy_aliases()
    _install_getargspec()
    _install_py2_builtins()
    _install_matplotlib_stubs()
    matern_mod = _install_matern_module()
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import gptools  # noqa: the reference package
        import gptools.kernel
    gptools.kernel._matern = matern_mod
    _ref_module = gptools
    return gptools


if __name__ == "__main__":
    g = load_reference()
    print("reference gptools", g.__version__, "loaded from", os.path.dirname(g.__file__))
