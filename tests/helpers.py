"""Shared test helpers: golden-file loading and the case -> (kernel id, inputs) mapping."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

KERNEL_SE, KERNEL_MATERN52, KERNEL_MATERN, KERNEL_GIBBS_TANH = 0, 1, 2, 3

# golden case -> kernel id
CASE_KERNEL = {
    "se2d_kat1": KERNEL_SE,
    "matern52_kat2": KERNEL_MATERN52,
    "matern52_2d_testshape": KERNEL_MATERN52,
    "matern_generic_nu2p5": KERNEL_MATERN,
    "matern_generic_nu3p5": KERNEL_MATERN,
    "matern_generic_nu1p5": KERNEL_MATERN,
    "matern_generic_2d": KERNEL_MATERN,
    "matern_generic_nu2p2": KERNEL_MATERN,       # real order: K_nu by Temme's method on the device
    "matern_generic_nu3p0": KERNEL_MATERN,       # integer order: series zone averaged over nu -+ 0.001
    "matern_generic_2d_nu2p2": KERNEL_MATERN,
    "gibbs_kat3": KERNEL_GIBBS_TANH,
    "gibbs_c5_small": KERNEL_GIBBS_TANH,
    "demo_c1_kat4": KERNEL_SE,
    "c1_synth200": KERNEL_SE,
    "c2_small_matern52": KERNEL_MATERN52,
    "c2_small_matern_generic": KERNEL_MATERN,
    "se_diagnoise": KERNEL_SE,
}


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) if a.size else 0.0


def assert_close(a, b, rtol=1e-9, atol=0.0, what=""):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    bad = ~(np.abs(a - b) <= atol + rtol * np.abs(b))
    bad &= ~(np.isnan(a) & np.isnan(b))
    if bad.any():
        i = np.argmax(np.where(bad, np.abs(a - b), 0))
        raise AssertionError("%s: %d/%d entries differ; worst at %s: got %r want %r (rtol %g atol %g)" % (
            what, bad.sum(), bad.size, np.unravel_index(i, a.shape), a.flat[i], b.flat[i], rtol, atol))


def var_tol(cov_diag_prior, rtol=1e-9):
    """predictive variance is a cancellation K** - |v|^2; agreement is bounded by
    rtol * K** (the prior variance), not by rtol * var."""
    return rtol * np.maximum(np.abs(cov_diag_prior), 1e-300)


# hyper-derivative goldens: Richardson finite differences of the reference's ll and K (make_golden.py: case_hyperfd)
HYPERFD_KERNEL = {"hyperfd_matern52_1d": KERNEL_MATERN52, "hyperfd_matern52_2d": KERNEL_MATERN52,
                  "hyperfd_matern_generic_nu2p5": KERNEL_MATERN, "hyperfd_matern_generic_nu3p5": KERNEL_MATERN,
                  "hyperfd_matern_generic_nu1p5": KERNEL_MATERN, "hyperfd_matern_generic_2d": KERNEL_MATERN,
                  "hyperfd_matern_generic_nu2p2": KERNEL_MATERN, "hyperfd_matern_generic_nu3p0": KERNEL_MATERN,
                  "hyperfd_gibbs_direct": KERNEL_GIBBS_TANH, "hyperfd_gibbs_T": KERNEL_GIBBS_TANH}


def richardson_fd(f, theta, i, h):
    """(4 D(h/2) - D(h)) / 3 with D the central difference of f along theta[i]."""
    def D(hh):
        tp, tm = np.array(theta, dtype=float), np.array(theta, dtype=float)
        tp[i] += hh
        tm[i] -= hh
        return (f(tp) - f(tm)) / (2.0 * hh)
    return (4.0 * D(h / 2.0) - D(h)) / 3.0
