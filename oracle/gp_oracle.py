"""TEST INFRASTRUCTURE -- not product code.

CPU (numpy/scipy, float64) restatement of the gptools hot path: pairwise
covariance closed forms -> compute_Kij -> compute_K_L_alpha_ll (+ gradient)
-> predict -> draw_sample.  Every function cites the reference file:line it
follows (paths relative to /root/reference/gptools).  It is the checker for the
CUDA path (tests/, __graft_entry__.smoke()) and the ``cpu_baseline`` /
``--impl reference`` leg of bench.py.  The product (gptools_b200/) never
imports it.

Parity status: PINNED.  tests/test_oracle_golden.py checks this module against
golden vectors produced by the unmodified reference run under
oracle/ref_shim.py (tests/golden/make_golden.py, tests/golden/*.npz) and against
the known-answer values recorded in SURVEY.md section 8c (KAT-1..5).

Numerics come from third-party libraries exactly as in the reference:
scipy.linalg.cholesky / cho_solve / solve_triangular (LAPACK dpotrf / dpotrs /
dtrtrs), scipy.special.eval_hermite / kv / kvp / gamma.  Versions are recorded
next to every golden file.
"""
import sys

import numpy as np
import scipy.linalg
import scipy.special

KERNEL_SE = 0
KERNEL_MATERN52 = 1
KERNEL_MATERN = 2
KERNEL_GIBBS_TANH = 3

EPS = sys.float_info.epsilon

# --------------------------------------------------------------------------
# pairwise closed forms (the Kernel.__call__ bodies)
# --------------------------------------------------------------------------


def _r2l2(tau, ls):
    """kernel/core.py:384-421 (_compute_r2l2): sum_d tau_d^2/l_d^2, 0/0 -> 0."""
    l_mat = np.tile(np.asarray(ls, dtype=float), (tau.shape[0], 1))
    with np.errstate(divide="ignore", invalid="ignore"):
        tol = tau / l_mat
    tol[(tau == 0) & (l_mat == 0)] = 0.0
    return np.sum(tol ** 2, axis=1), l_mat


def se_pairs(Xi, Xj, ni, nj, params, hyper_deriv=None):
    """kernel/squared_exponential.py:110-174.  params = [sigma_f, l_1..l_D]."""
    params = np.asarray(params, dtype=float)
    D = Xi.shape[1]
    ni = np.asarray(ni)
    nj = np.asarray(nj)
    value_only = (ni.astype(np.intc) == 0).all() and (nj.astype(np.intc) == 0).all()
    tau = np.asarray(Xi - Xj, dtype=float)
    r2l2, l_mat = _r2l2(tau, params[-D:])
    k = params[0] ** 2 * np.exp(-r2l2 / 2.0)
    if not value_only:
        n_tot_j = np.asarray(np.sum(nj, axis=1), dtype=np.intc).flatten()
        m = np.asarray(ni + nj, dtype=np.intc)
        sign_j = (-1.0) ** n_tot_j
        herm = (-1.0 / (np.sqrt(2.0) * l_mat)) ** m * scipy.special.eval_hermite(m, tau / (np.sqrt(2.0) * l_mat))
        if hyper_deriv is not None and hyper_deriv > 0:
            d = hyper_deriv - 1
            lp = params[hyper_deriv]
            t = tau[:, d] ** 2.0 / lp ** 3.0
            mask = m[:, d] > 0
            t[mask] -= m[mask, d] / lp
            mask = mask & (tau[:, d] != 0.0)
            x = tau[mask, d] / (np.sqrt(2.0) * lp)
            t[mask] -= (np.sqrt(2.0) * m[mask, d] * tau[mask, d] / lp ** 2.0 *
                        scipy.special.eval_hermite(m[mask, d] - 1, x) / scipy.special.eval_hermite(m[mask, d], x))
            herm[:, d] *= t
        k = sign_j * np.prod(herm, axis=1) * k
    if hyper_deriv is None:
        return k
    if hyper_deriv == 0:
        return 2.0 * k / params[0] if params[0] != 0.0 else np.zeros_like(k)
    if value_only:
        return tau[:, hyper_deriv - 1] ** 2.0 / params[hyper_deriv] ** 3.0 * k
    return k


def matern52_pairs(Xi, Xj, ni, nj, params, hyper_deriv=None):
    """kernel/matern.py:543-555 + kernel/src/matern.c:61-186 (vectorised).

    params = [sigma_f, l_1..l_D]; only sum(ni)<=1 and sum(nj)<=1.
    """
    if hyper_deriv is not None:
        raise NotImplementedError("Hyperparameter derivatives have not been implemented!")
    ni = np.asarray(ni, dtype=np.int32)
    nj = np.asarray(nj, dtype=np.int32)
    if np.any(ni.sum(axis=1) > 1) or np.any(nj.sum(axis=1) > 1):
        raise ValueError("Matern52Kernel only supports 0th and 1st order derivatives")
    params = np.asarray(params, dtype=float)
    Xi = np.asarray(Xi, dtype=float)
    Xj = np.asarray(Xj, dtype=float)
    D = Xi.shape[1]
    var = np.square(params[-D:])
    SQRT_5 = 2.2360679774997898
    FIVE_THIRDS = 1.6666666666666667
    npair = Xi.shape[0]
    # matern.c:61-71 -- accumulation order d = 0..D-1
    r2 = np.zeros(npair)
    for d in range(D):
        disp = Xi[:, d] - Xj[:, d]
        r2 = r2 + disp * disp / var[d]
    zero = r2 == 0
    r = np.sqrt(r2)
    s5r = SQRT_5 * r
    e = np.exp(-s5r)
    # index of the first '1' in ni / nj (matern.c:32-39), -1 if none
    has_i = (ni == 1).any(axis=1)
    has_j = (nj == 1).any(axis=1)
    idx_i = np.argmax(ni == 1, axis=1)
    idx_j = np.argmax(nj == 1, axis=1)
    rows = np.arange(npair)
    out = np.zeros(npair)
    # value (matern.c:77-88)
    c = ~has_i & ~has_j
    val = (1.0 + s5r + FIVE_THIRDS * r2) * e
    val[zero] = 1.0
    out[c] = val[c]
    dk_over_r = -FIVE_THIRDS * (1 + s5r) * e
    # d/dXi_n (matern.c:95-106)
    c = has_i & ~has_j
    g = dk_over_r * ((Xi[rows, idx_i] - Xj[rows, idx_i]) / var[idx_i])
    g[zero] = 0.0
    out[c] = g[c]
    # d/dXj_m = dkernel_dXn(xj, xi, m) (matern.c:182-184)
    c = ~has_i & has_j
    g = dk_over_r * ((Xj[rows, idx_j] - Xi[rows, idx_j]) / var[idx_j])
    g[zero] = 0.0
    out[c] = g[c]
    # mixed second derivative (matern.c:112-147)
    c = has_i & has_j
    if c.any():
        same = idx_i == idx_j
        dn = Xi[rows, idx_i] - Xj[rows, idx_i]
        dm = Xi[rows, idx_j] - Xj[rows, idx_j]
        with np.errstate(divide="ignore", invalid="ignore"):
            dr_dXn = dn / (r * var[idx_i])
            dr_dYm = -dm / (r * var[idx_j])
            d2r_r3 = dn * dm / (var[idx_i] * var[idx_j])
            d2r_r3 = np.where(same, d2r_r3 - r * r / var[idx_i], d2r_r3)
            d2k = FIVE_THIRDS * (5 * r2 - s5r - 1) * e
            h = dk_over_r * d2r_r3 / r2 + d2k * dr_dXn * dr_dYm
        h0 = np.where(same, FIVE_THIRDS / var[idx_i], 0.0)
        h = np.where(zero, h0, h)
        out[c] = h[c]
    return params[0] ** 2 * out


# ---- generic Matern (ChainRuleKernel + utils) ---------------------------------


def _poch(a, n):
    """utils.py:1369-1395 (fixed_poch): prod_{k<n} (a+k)."""
    p = 1.0
    for k in range(int(n)):
        p *= (a + k)
    return p


def _bell(n, k, x):
    """utils.py:1520-1570 (incomplete_bell_poly) recurrence."""
    if n == 0 and k == 0:
        return np.ones(x.shape[0])
    if k == 0 or n == 0:
        return np.zeros(x.shape[0])
    res = np.zeros(x.shape[0])
    for m in range(0, n - k + 1):
        res += x[:, m] * scipy.special.binom(n - 1, m) * _bell(n - (m + 1), k - 1, x)
    return res


def _Kn2Der(nu, y, n=0):
    """utils.py:1397-1427: d^n/dy^n K_nu(sqrt(y))."""
    y = np.asarray(y, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore"):
        sq = np.sqrt(y)
        if n == 0:
            return scipy.special.kv(nu, sq)
        K = np.zeros_like(y)
        x = np.asarray([_poch(1.5 - j, j) * y ** (0.5 - j) for j in np.arange(1.0, n + 1.0)]).T
        for k in range(1, n + 1):
            K += scipy.special.kvp(nu, sq, n=k) * _bell(n, k, x)
    return K


def _yn2Kn2Der(nu, y, n=0, tol=5e-4, nterms=1, nu_step=0.001):
    """utils.py:1429-1518: d^n/dy^n [y^{nu/2} K_nu(sqrt y)] incl. the series zone 0<y<=tol."""
    n = int(n)
    y = np.asarray(y, dtype=float)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if n == 0:
            K = y ** (nu / 2.0) * scipy.special.kv(nu, np.sqrt(y))
            K[y == 0.0] = scipy.special.gamma(nu) / 2.0 ** (1.0 - nu)
        else:
            K = np.zeros_like(y)
            for k in range(0, n + 1):
                K += (scipy.special.binom(n, k) * _poch(1.0 + nu / 2.0 - k, k) *
                      y ** (nu / 2.0 - k) * _Kn2Der(nu, y, n=n - k))
            mask = (y == 0.0)
            if mask.any():
                if int(nu) == nu:
                    K[mask] = 0.5 * (_yn2Kn2Der(nu - nu_step, y[mask], n, tol, nterms, nu_step) +
                                     _yn2Kn2Der(nu + nu_step, y[mask], n, tol, nterms, nu_step))
                elif n > nu:
                    K[mask] = scipy.special.gamma(-nu) * _poch(1 + nu - n, n) * np.inf
                else:
                    K[mask] = (scipy.special.gamma(nu) * scipy.special.gamma(n + 1.0) /
                               (2.0 ** (1.0 - nu + 2.0 * n) * _poch(1.0 - nu, n) * scipy.special.factorial(n)))
        if tol > 0.0:
            mask = (y <= tol) & (y > 0.0)
            K[mask] = 0.0
            if int(nu) == nu:
                K[mask] = 0.5 * (_yn2Kn2Der(nu - nu_step, y[mask], n, tol, nterms, nu_step) +
                                 _yn2Kn2Der(nu + nu_step, y[mask], n, tol, nterms, nu_step))
            else:
                for k in np.arange(n, n + nterms, dtype=float):
                    K[mask] += (scipy.special.gamma(nu) * _poch(1.0 + k - n, n) * y[mask] ** (k - n) /
                                (2.0 ** (1.0 - nu + 2 * k) * _poch(1.0 - nu, k) * scipy.special.factorial(k)))
                for k in np.arange(0, nterms, dtype=float):
                    K[mask] += (scipy.special.gamma(-nu) * _poch(1.0 + nu + k - n, n) * y[mask] ** (nu + k - n) /
                                (2.0 ** (1.0 + nu + 2.0 * k) * _poch(1.0 + nu, k) * scipy.special.factorial(k)))
    return K


def _set_partitions(items):
    """All set partitions of a list (utils.py:1572-1656 enumerates the same set via
    restricted-growth strings; the sum over partitions is order independent up to
    round-off -- we keep the lexicographic RGS order)."""
    n = len(items)
    if n == 0:
        return []
    out = []

    def rec(pos, rgs, nblocks):
        if pos == n:
            blocks = [[] for _ in range(nblocks)]
            for it, b in zip(items, rgs):
                blocks[b].append(it)
            out.append(blocks)
            return
        for b in range(nblocks + 1):
            rec(pos + 1, rgs + [b], max(nblocks, b + 1))

    rec(1, [0], 1)
    return out


def matern_pairs(Xi, Xj, ni, nj, params, hyper_deriv=None):
    """kernel/core.py:691-816 (ChainRuleKernel) + kernel/matern.py:296-459.

    params = [sigma_f, nu, l_1..l_D].
    """
    if hyper_deriv is not None:
        raise NotImplementedError("Hyperparameter derivatives have not been implemented!")
    params = np.asarray(params, dtype=float)
    D = Xi.shape[1]
    nu = params[1]
    ls = params[2:2 + D]
    tau = np.asarray(Xi - Xj, dtype=float)
    n_tot_j = np.asarray(np.sum(nj, axis=1), dtype=np.intc).flatten()
    m = np.asarray(ni + nj, dtype=np.intc)
    k = np.zeros(Xi.shape[0])
    pref = 2.0 ** (1.0 - nu) / scipy.special.gamma(nu)
    for state in np.unique(m, axis=0):
        idxs = (m == state).all(axis=1)
        t = tau[idxs]
        r2l2, _ = _r2l2(t, ls)
        y = 2.0 * nu * r2l2
        pattern = []
        for d in range(D):
            pattern.extend(int(state[d]) * [d])
        if len(pattern) == 0:
            with np.errstate(invalid="ignore", divide="ignore"):
                kk = pref * y ** (nu / 2.0) * scipy.special.kv(nu, np.sqrt(y))
            kk[r2l2 == 0] = 1.0
            k[idxs] = kk
            continue
        acc = np.zeros(t.shape[0])
        for part in _set_partitions(pattern):
            # matern.py:420-459 (_compute_dk_dtau_on_partition)
            order = len(part)
            n1 = 0
            fac = np.ones_like(y)
            dead = False
            for b in part:
                if len(b) > 2 or (len(b) == 2 and b[0] != b[1]):
                    dead = True
                    break
                if len(b) == 1:
                    fac = fac * (4.0 * nu * t[:, b[0]] / ls[b[0]] ** 2.0)
                    n1 += 1
                else:
                    fac = fac * (4.0 * nu / ls[b[0]] ** 2.0)
            if dead:
                continue
            dk = pref * _yn2Kn2Der(nu, y, n=order)
            if n1 > 0:
                mask = (y == 0.0)
                tau_pow = 2 * (nu - order) + n1
                if tau_pow == 0:
                    dk[mask] = np.nan
                elif tau_pow > 0:
                    dk[mask] = 0.0
            acc += dk * fac
        k[idxs] = acc
    return params[0] ** 2.0 * (-1.0) ** n_tot_j * k


def tanh_warp(x, n, l1, l2, lw, x0):
    """kernel/gibbs.py:458-461."""
    if n == 0:
        return (l1 + l2) / 2.0 - (l1 - l2) / 2.0 * np.tanh((x - x0) / lw)
    return -(l1 - l2) / (2.0 * lw) * (np.cosh((x - x0) / lw)) ** (-2.0)


def gibbs_tanh_pairs(Xi, Xj, ni, nj, params, hyper_deriv=None):
    """kernel/gibbs.py:288-423 with l(x) = tanh_warp (gibbs.py:426-465).

    params = [sigma_f, l1, l2, lw, x0]; dispatch on the (ni, nj) pair, orders <= 1.
    """
    if hyper_deriv is not None:
        raise NotImplementedError("Hyperparameter derivatives have not been implemented!")
    params = np.asarray(params, dtype=float)
    x = np.asarray(Xi, dtype=float)[:, 0]
    y = np.asarray(Xj, dtype=float)[:, 0]
    a = np.asarray(ni, dtype=int)[:, 0]
    b = np.asarray(nj, dtype=int)[:, 0]
    if (a > 1).any() or (b > 1).any():
        raise NotImplementedError("Derivatives greater than [1, 1] are not supported!")
    lx = tanh_warp(x, 0, *params[1:])
    ly = tanh_warp(y, 0, *params[1:])
    lx1 = tanh_warp(x, 1, *params[1:])
    ly1 = tanh_warp(y, 1, *params[1:])
    d = x - y
    S = lx ** 2 + ly ** 2
    e = np.exp(-d ** 2 / S)
    k = np.zeros(x.shape[0])
    c = (a == 0) & (b == 0)
    k[c] = (np.sqrt(2.0 * lx * ly / S) * e)[c]
    c = (a == 1) & (b == 0)
    if c.any():
        v = (e * ly * (-4 * d * lx ** 3 - 4 * d * lx * ly ** 2 + 4 * d ** 2 * lx ** 2 * lx1 - lx ** 4 * lx1 + ly ** 4 * lx1)
             ) / (np.sqrt(2 * lx * ly) * S ** 2.5)
        k[c] = v[c]
    c = (a == 0) & (b == 1)
    if c.any():
        v = (e * lx * (4 * d * ly ** 3 + 4 * d * ly * lx ** 2 + 4 * d ** 2 * ly ** 2 * ly1 - ly ** 4 * ly1 + lx ** 4 * ly1)
             ) / (np.sqrt(2 * lx * ly) * S ** 2.5)
        k[c] = v[c]
    c = (a == 1) & (b == 1)
    if c.any():
        v = (e * (
            -lx ** 8 * lx1 * ly1 +
            4 * lx ** 7 * (2 * ly - d * ly1) -
            4 * lx ** 5 * ly * (4 * d ** 2 - 6 * ly ** 2 - 3 * d * ly * ly1) +
            ly ** 6 * lx1 * (4 * d * ly + 4 * d ** 2 * ly1 - ly ** 2 * ly1) +
            4 * lx ** 6 * lx1 * (-5 * d * ly + d ** 2 * ly1 + 2 * ly ** 2 * ly1) -
            4 * lx * ly ** 4 * (4 * d ** 2 * ly - 2 * ly ** 3 + 4 * d ** 3 * ly1 - 5 * d * ly ** 2 * ly1) -
            4 * lx ** 3 * ly ** 2 * (8 * d ** 2 * ly - 6 * ly ** 3 + 4 * d ** 3 * ly1 - 9 * d * ly ** 2 * ly1) +
            2 * lx ** 4 * ly * lx1 * (8 * d ** 3 - 18 * d * ly ** 2 - 18 * d ** 2 * ly * ly1 + 9 * ly ** 3 * ly1) +
            4 * lx ** 2 * ly ** 2 * lx1 * (4 * d ** 3 * ly - 3 * d * ly ** 3 + 4 * d ** 4 * ly1 -
                                           9 * d ** 2 * ly ** 2 * ly1 + 2 * ly ** 4 * ly1)
        )) / (2 * np.sqrt(2 * lx * ly) * S ** 4.5)
        k[c] = v[c]
    return params[0] ** 2 * k


_PAIRS = {
    KERNEL_SE: se_pairs,
    KERNEL_MATERN52: matern52_pairs,
    KERNEL_MATERN: matern_pairs,
    KERNEL_GIBBS_TANH: gibbs_tanh_pairs,
}


def _powerset(items):
    """utils.py: powerset() as used by ProductKernel (kernel/core.py:645): every sub-multiset by position."""
    from itertools import chain, combinations
    return chain.from_iterable(combinations(items, r) for r in range(len(items) + 1))


def _product_pairs(fa, fb, Xi, Xj, ni, nj):
    """ProductKernel.__call__ (kernel/core.py:628-668): for every distinct row of [ni, nj] the multiset of unit
    derivatives is split between the two factors in every possible way (the power set BY POSITION, so equal splits
    are visited repeatedly -- that repetition is the binomial weight)."""
    D = Xi.shape[1]
    nij = np.hstack((ni, nj))
    result = np.zeros(Xi.shape[0])
    for row in np.unique(nij, axis=0):
        pattern = []
        for idx in range(len(row)):
            pattern.extend(int(row[idx]) * [idx])
        sel = (nij == row).all(axis=1)
        cnt = int(sel.sum())
        for sub in _powerset(list(range(len(pattern)))):
            n1 = np.zeros((cnt, 2 * D), dtype=int)
            n2 = np.zeros((cnt, 2 * D), dtype=int)
            for pos in range(len(pattern)):
                (n1 if pos in sub else n2)[:, pattern[pos]] += 1
            result[sel] += fa(Xi[sel], Xj[sel], n1[:, :D], n1[:, D:]) * fb(Xi[sel], Xj[sel], n2[:, :D], n2[:, D:])
    return result


def composite_pairs(structure, params, Xi, Xj, ni, nj, hyper_deriv=None):
    """Sum of products of leaf kernels: the flattened form of a SumKernel / ProductKernel tree
    (kernel/core.py:549-670).  structure = (leaf_kids, leaf_nparams, term_masks); params is the reference's
    concatenation of the operands' parameter vectors (kernel/core.py:452-459).  hyper_deriv follows SumKernel
    (kernel/core.py:576-582: only the operand owning the parameter contributes); inside a product the owning factor is
    replaced by its own hyper-derivative (the reference raises NotImplementedError there, kernel/core.py:626-627)."""
    kids, nps, masks = structure
    Xi = np.atleast_2d(np.asarray(Xi, dtype=float))
    Xj = np.atleast_2d(np.asarray(Xj, dtype=float))
    ni = np.atleast_2d(np.asarray(ni, dtype=int))
    nj = np.atleast_2d(np.asarray(nj, dtype=int))
    params = np.asarray(params, dtype=float)
    offs = np.concatenate([[0], np.cumsum(nps)])
    owner, hl = None, None
    if hyper_deriv is not None:
        owner = int(np.searchsorted(offs, hyper_deriv, side="right") - 1)
        hl = int(hyper_deriv - offs[owner])

    def leaf(q):
        def f(xi, xj, mi, mj):
            return _PAIRS[kids[q]](xi, xj, mi, mj, params[offs[q]:offs[q + 1]], hyper_deriv=hl if q == owner else None)
        return f

    total = np.zeros(Xi.shape[0])
    for mask in masks:
        members = [q for q in range(len(kids)) if (mask >> q) & 1]
        if owner is not None and owner not in members:
            continue
        f = leaf(members[-1])
        for q in reversed(members[:-1]):   # right-nested products: a * (b * (c ...))
            f = (lambda fa, fb: (lambda xi, xj, mi, mj: _product_pairs(fa, fb, xi, xj, mi, mj)))(leaf(q), f)
        total += f(Xi, Xj, ni, nj)
    return total


def kernel_pairs(kid, params, Xi, Xj, ni, nj, hyper_deriv=None):
    if hasattr(kid, "structure"):  # composite descriptor (anything with .structure = (kids, nparams, masks))
        return composite_pairs(kid.structure, params, Xi, Xj, ni, nj, hyper_deriv=hyper_deriv)
    return _PAIRS[kid](Xi, Xj, ni, nj, params, hyper_deriv=hyper_deriv)


# --------------------------------------------------------------------------
# GaussianProcess numerics
# --------------------------------------------------------------------------


def compute_Kij(kid, params, Xi, Xj, ni, nj, hyper_deriv=None, block_rows=None):
    """gaussian_process.py:1535-1605: flattened pair lists, pair index i*Mj + j.

    ``block_rows`` evaluates the same pair lists in row blocks of Xi (identical
    values; bounds the size of the (Mi*Mj, D) temporaries for large problems).
    """
    Xi = np.atleast_2d(np.asarray(Xi, dtype=float))
    ni = np.atleast_2d(np.asarray(ni, dtype=int))
    if Xj is None:
        Xj, nj = Xi, ni
    Xj = np.atleast_2d(np.asarray(Xj, dtype=float))
    nj = np.atleast_2d(np.asarray(nj, dtype=int))
    Mi, Mj = Xi.shape[0], Xj.shape[0]
    if block_rows is None or block_rows >= Mi:
        out = kernel_pairs(kid, params, np.repeat(Xi, Mj, axis=0), np.tile(Xj, (Mi, 1)),
                           np.repeat(ni, Mj, axis=0), np.tile(nj, (Mi, 1)), hyper_deriv=hyper_deriv)
        return np.reshape(out, (Mi, -1))
    K = np.empty((Mi, Mj))
    for r0 in range(0, Mi, block_rows):
        r1 = min(Mi, r0 + block_rows)
        K[r0:r1] = compute_Kij(kid, params, Xi[r0:r1], Xj, ni[r0:r1], nj, hyper_deriv=hyper_deriv)
    return K


def compute_K_L_alpha_ll(kid, params, X, n, y, err_y, T=None, noise_sigma=0.0, diag_factor=1e2,
                         mu_y=None, grad_idx=None, block_rows=None):
    """gaussian_process.py:1418-1522 without the hyperprior terms (added by the caller).

    noise_sigma : DiagonalNoiseKernel sigma_n (gaussian_process.py:1436-1437); 0 = ZeroKernel.
    mu_y        : T.mu(X, n) already evaluated (gaussian_process.py:1455-1459), or None.
    grad_idx    : indices into ``params`` of the free kernel hyperparameters; when given,
                  ll_deriv[i] = 0.5 (alpha' dK alpha - tr(K_tot^{-1} dK)) (gaussian_process.py:1495-1504).
    """
    X = np.atleast_2d(np.asarray(X, dtype=float))
    n = np.atleast_2d(np.asarray(n, dtype=int))
    y = np.asarray(y, dtype=float)
    err_y = np.asarray(err_y, dtype=float)
    K = compute_Kij(kid, params, X, None, n, None, block_rows=block_rows)
    noise_K = noise_sigma ** 2.0 * np.eye(X.shape[0]) if noise_sigma != 0.0 else np.zeros((X.shape[0],) * 2)
    if T is not None:
        KnK = T.dot(K + noise_K).dot(T.T)
    else:
        KnK = K + noise_K
    K_tot = KnK + np.diag(err_y ** 2.0) + diag_factor * EPS * np.eye(len(y))
    L = scipy.linalg.cholesky(K_tot, lower=True)
    y_alph = y - mu_y if mu_y is not None else y
    alpha = scipy.linalg.cho_solve((L, True), np.atleast_2d(y_alph).T)
    ll = (-0.5 * np.atleast_2d(y_alph).dot(alpha) - np.log(np.diag(L)).sum() -
          0.5 * len(y) * np.log(2.0 * np.pi))[0, 0]
    out = {"K": K, "K_tot": K_tot, "L": L, "alpha": alpha, "ll": ll}
    if grad_idx is not None:
        g = np.zeros(len(grad_idx))
        for i, pi in enumerate(grad_idx):
            dK = compute_Kij(kid, params, X, None, n, None, hyper_deriv=int(pi), block_rows=block_rows)
            if T is not None:
                dK = T.dot(dK).dot(T.T)
            g[i] = 0.5 * (alpha.T.dot(dK.dot(alpha))[0, 0] - np.trace(scipy.linalg.cho_solve((L, True), dK)))
        out["ll_deriv"] = g
    return out


def predict(kid, params, X, n, L, alpha, Xstar, nstar, T=None, return_cov=True, mu_star=None,
            output_transform=None):
    """gaussian_process.py:965-1006 (noise=False)."""
    Xstar = np.atleast_2d(np.asarray(Xstar, dtype=float))
    nstar = np.atleast_2d(np.asarray(nstar, dtype=int))
    Kstar = compute_Kij(kid, params, X, Xstar, n, nstar)
    if T is not None:
        Kstar = T.dot(Kstar)
    mean = Kstar.T.dot(alpha)
    if mu_star is not None:
        mean = mean + np.atleast_2d(mu_star).T
    if output_transform is not None:
        mean = output_transform.dot(mean)
    mean = mean.ravel()
    if not return_cov:
        return mean
    v = scipy.linalg.solve_triangular(L, Kstar, lower=True)
    Kss = compute_Kij(kid, params, Xstar, None, nstar, None)
    cov = Kss - v.T.dot(v)
    if output_transform is not None:
        cov = output_transform.dot(cov.dot(output_transform.T))
    with np.errstate(invalid="ignore"):
        std = np.sqrt(np.diagonal(cov))
    return mean, std, cov


def predict_blocked(kid, params, X, n, L, alpha, Xstar, nstar, T=None, block=2000):
    """mean/std of :func:`predict` evaluated in blocks of test points -- per-point
    quantities are block independent, which is how the
    reference has to be driven for M* >~ 1e4 (it forms the full M* x M* covariance,
    gaussian_process.py:984-987)."""
    Xstar = np.atleast_2d(np.asarray(Xstar, dtype=float))
    nstar = np.atleast_2d(np.asarray(nstar, dtype=int))
    means, stds = [], []
    for s0 in range(0, Xstar.shape[0], block):
        m, s, _ = predict(kid, params, X, n, L, alpha, Xstar[s0:s0 + block], nstar[s0:s0 + block], T=T)
        means.append(m)
        stds.append(s)
    return np.concatenate(means), np.concatenate(stds)


def draw_sample(mean, cov, rand_vars, diag_factor=1e3):
    """gaussian_process.py:1295-1300, 1330 (method='cholesky', explicit rand_vars)."""
    Lc = scipy.linalg.cholesky(cov + diag_factor * EPS * np.eye(cov.shape[0]), lower=True, check_finite=False)
    return np.atleast_2d(mean).T + Lc.dot(rand_vars)
