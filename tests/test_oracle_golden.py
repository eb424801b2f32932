"""Pin the numpy restatement (oracle/gp_oracle.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py) and the KAT values recorded in SURVEY.md 8c.

CPU only. These tests are what makes the oracle a trustworthy checker for the CUDA path.
"""
import numpy as np
import pytest

from helpers import CASE_KERNEL, KERNEL_SE, assert_close, load_golden
from oracle import gp_oracle as orc


def _T(gd):
    return gd["T"] if "T" in gd else None


@pytest.mark.parametrize("D", [1, 2, 3])
def test_se_pair_lists(D):
    gd = load_golden("se_pairs_D%d" % D)
    args = (gd["Xi"], gd["Xj"], gd["ni"], gd["nj"], gd["params"])
    assert_close(orc.se_pairs(*args), gd["val"], rtol=1e-13, atol=1e-300, what="value")
    z = np.zeros_like(gd["ni"])
    assert_close(orc.se_pairs(gd["Xi"], gd["Xj"], z, z, gd["params"]), gd["val0"], rtol=1e-14, what="val0")
    for p in range(D + 1):
        got = orc.se_pairs(*args, hyper_deriv=p)
        assert_close(got, gd["hd%d" % p], rtol=1e-12, what="hyper_deriv %d" % p)
        got0 = orc.se_pairs(gd["Xi"], gd["Xj"], z, z, gd["params"], hyper_deriv=p)
        assert_close(got0, gd["hd0_%d" % p], rtol=1e-14, what="hyper_deriv0 %d" % p)


@pytest.mark.parametrize("case", [c for c in CASE_KERNEL if c not in ("c1_synth200", "c2_small_matern52",
                                                                     "c2_small_matern_generic", "gibbs_c5_small")])
def test_K_matrix(case):
    gd = load_golden(case)
    K = orc.compute_Kij(CASE_KERNEL[case], gd["params"], gd["X"], None, gd["n"], None)
    # the reference's own tolerance for this quantity is 1e-8 abs (tests/test_matern.py:31)
    assert_close(K, gd["K"], rtol=1e-12, atol=1e-14, what=case + " K")


def test_matern_generic_matches_reference_generic_K():
    gd = load_golden("matern52_2d_testshape")
    K = orc.compute_Kij(orc.KERNEL_MATERN, gd["params_generic"], gd["X"], None, gd["n"], None)
    assert_close(K, gd["K_generic"], rtol=1e-12, atol=1e-14, what="generic K")
    # and the reference's own property (tests/test_matern.py:31): C path == generic path to 1e-8
    np.testing.assert_array_almost_equal(gd["K"], gd["K_generic"], decimal=8)


@pytest.mark.parametrize("case", sorted(CASE_KERNEL))
def test_ll_alpha(case):
    gd = load_golden(case)
    ns = float(gd["noise_sigma"][0]) if "noise_sigma" in gd else 0.0
    grad_idx = None
    if "ll_deriv" in gd and CASE_KERNEL[case] == KERNEL_SE:
        grad_idx = np.arange(len(gd["params"]))
    r = orc.compute_K_L_alpha_ll(CASE_KERNEL[case], gd["params"], gd["X"], gd["n"], gd["y"], gd["err_y"],
                                 T=_T(gd), noise_sigma=ns, grad_idx=grad_idx)
    assert_close(r["ll"] + gd["log_prior"], gd["ll"], rtol=1e-12, what=case + " ll")
    assert_close(r["alpha"].ravel(), gd["alpha"], rtol=1e-9, atol=1e-9 * np.abs(gd["alpha"]).max(), what=case + " alpha")
    if grad_idx is not None:
        nk = len(gd["params"])
        # uniform / gamma prior derivative is not part of the oracle; compare the likelihood part
        if case in ("se2d_kat1", "se_diagnoise"):
            assert_close(r["ll_deriv"], gd["ll_deriv"][:nk], rtol=1e-10, what=case + " ll_deriv")


def test_kat1_values_from_survey():
    """SURVEY.md 8c KAT-1 numbers, typed in from the survey (independent of the npz)."""
    gd = load_golden("se2d_kat1")
    r = orc.compute_K_L_alpha_ll(KERNEL_SE, [1.3, 0.7, 1.1], gd["X"], gd["n"], gd["y"], gd["err_y"], grad_idx=[0, 1, 2])
    assert_close(r["ll"] + (-43.74911676688687), -20.205363765586593, rtol=1e-12)
    assert_close(r["ll_deriv"], [-7.084597593024648, 20.863151889933373, 9.298697347173814], rtol=1e-10)
    assert_close(r["K"][0, :3], [1.6900000000000002, 1.6649137752633176, 1.659904236194991], rtol=1e-14)
    assert_close(r["K"][0, 6:9], [0, -0.18330997001381968, 0.42398257947550183], rtol=1e-13)
    assert_close(r["K"][6, 6:9], [3.4489795918367343, 3.3776004608536585, 3.279263521999551], rtol=1e-13)


@pytest.mark.parametrize("case", ["se2d_kat1", "matern52_kat2", "gibbs_kat3", "gibbs_c5_small", "se_diagnoise",
                                  "matern_generic_nu2p5", "matern_generic_nu3p5", "matern_generic_nu2p2",
                                  "matern_generic_nu3p0"])
def test_predict_full(case):
    gd = load_golden(case)
    kid = CASE_KERNEL[case]
    ns = float(gd["noise_sigma"][0]) if "noise_sigma" in gd else 0.0
    r = orc.compute_K_L_alpha_ll(kid, gd["params"], gd["X"], gd["n"], gd["y"], gd["err_y"], T=_T(gd), noise_sigma=ns)
    Xs = np.atleast_2d(gd["Xs"])
    if Xs.shape[0] == 1 and gd["X"].shape[1] == 1:
        Xs = Xs.T
    mean, std, cov = orc.predict(kid, gd["params"], gd["X"], gd["n"], r["L"], r["alpha"], Xs, np.zeros(Xs.shape, dtype=int), T=_T(gd))
    assert_close(mean, gd["mean"], rtol=1e-9, atol=1e-12, what=case + " mean")
    prior = np.diag(orc.compute_Kij(kid, gd["params"], Xs, None, np.zeros(Xs.shape, dtype=int), None))
    assert np.all(np.abs(cov - gd["cov"]) <= 1e-9 * prior.max())
    assert np.all(np.abs(std ** 2 - gd["std"] ** 2) <= 1e-9 * prior)
    if "draw" in gd:
        samp = orc.draw_sample(mean, cov, gd["rand_vars"])
        assert_close(samp, gd["draw"], rtol=1e-7, atol=1e-7, what=case + " draw")


@pytest.mark.parametrize("case", ["demo_c1_kat4", "c1_synth200", "c2_small_matern52", "c2_small_matern_generic"])
def test_predict_mean_std_blocked(case):
    gd = load_golden(case)
    kid = CASE_KERNEL[case]
    r = orc.compute_K_L_alpha_ll(kid, gd["params"], gd["X"], gd["n"], gd["y"], gd["err_y"])
    Xs = gd["Xs"][:, None]
    for order, suffix in ((0, ""), (1, "_d1")):
        ns = np.full(Xs.shape, order, dtype=int)
        mean, std = orc.predict_blocked(kid, gd["params"], gd["X"], gd["n"], r["L"], r["alpha"], Xs, ns, block=100)
        scale = np.abs(gd["mean" + suffix]).max()
        assert_close(mean, gd["mean" + suffix], rtol=1e-9, atol=1e-9 * scale, what=case + " mean" + suffix)
        prior = np.diag(orc.compute_Kij(kid, gd["params"], Xs[:1], None, ns[:1], None))[0]
        ok = np.abs(std ** 2 - gd["std" + suffix] ** 2) <= 1e-9 * prior
        ok |= np.isnan(std) & np.isnan(gd["std" + suffix])
        assert ok.all(), case + " var" + suffix


def test_c3_kat5():
    gd = load_golden("c3_kat5")
    for b in range(len(gd["theta_idx"])):
        r = orc.compute_K_L_alpha_ll(KERNEL_SE, gd["theta"][b], gd["X"], gd["n"], gd["y"], gd["err_y"],
                                     grad_idx=[0, 1, 2] if b in (0, 7) else None)
        assert_close(r["ll"] + gd["log_prior"], gd["ll"][b], rtol=1e-11, what="ll[%d]" % b)
        if b in (0, 7):
            assert_close(r["ll_deriv"], gd["ll_deriv"][b], rtol=1e-9, what="grad[%d]" % b)
    # SURVEY KAT-5 (typed from the survey)
    assert_close(gd["ll"][0], 622.1177149587926, rtol=1e-12)
    assert_close(gd["ll_deriv"][0], [-43.61422487, 417.65577873, 271.8373869], rtol=1e-8)
    assert_close(gd["ll"][7], 648.2002071542964, rtol=1e-12)
