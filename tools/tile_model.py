"""Numpy model of the batched kernel's tile algorithm (gptools_b200/csrc/batched4.cu), used to validate the
slot / in-place conventions before writing CUDA.  Phase 1: left-looking Cholesky with inverted diagonal
tiles; phase 2: XT = L^{-T} by block substitution, in place; phase 3: K^{-1} tiles = sum XT XT^T."""
import numpy as np

TB = 8


def nt(A, B):
    return A @ B.T


def run(K, y):
    M = K.shape[0]
    nT = M // TB
    blk = lambda A, I, J: A[I * TB:(I + 1) * TB, J * TB:(J + 1) * TB]
    slot = {}
    D, DT = {}, {}
    z = np.zeros(M)
    logdet = 0.0
    # ---- phase 1
    for k in range(nT):
        C = {}
        for I in range(k, nT):
            acc = np.zeros((TB, TB))
            for j in range(k):
                acc += nt(slot[(I, j)], slot[(k, j)])
            C[I] = blk(K, I, k) - acc
        Lkk = np.linalg.cholesky(C[k])
        logdet += np.log(np.diag(Lkk)).sum()
        D[k] = np.linalg.inv(Lkk)
        DT[k] = D[k].T.copy()
        rk = y[k * TB:(k + 1) * TB].copy()
        for j in range(k):
            rk -= slot[(k, j)] @ z[j * TB:(j + 1) * TB]
        z[k * TB:(k + 1) * TB] = D[k] @ rk
        for I in range(k + 1, nT):
            slot[(I, k)] = nt(C[I], D[k])
    L = np.zeros((M, M))
    for (I, J), v in slot.items():
        L[I * TB:(I + 1) * TB, J * TB:(J + 1) * TB] = v
    for k in range(nT):
        L[k * TB:(k + 1) * TB, k * TB:(k + 1) * TB] = np.linalg.inv(D[k])
    assert np.allclose(L @ L.T, K)
    # ---- phase 2 (in place): slot(I,J) <- XT(J,I)
    for I in range(1, nT):
        for J in range(0, I):          # ascending J
            ST = np.zeros((TB, TB))
            for m in range(J, I):
                a = DT[J] if m == J else slot[(m, J)]
                b = slot[(I, m)]       # still L(I,m): m >= J, not yet overwritten
                ST += nt(a, b)
            slot[(I, J)] = -nt(ST, D[I])
    X = np.linalg.inv(L)
    for (I, J), v in slot.items():
        assert np.allclose(v, X[I * TB:(I + 1) * TB, J * TB:(J + 1) * TB].T), (I, J)
    # alpha = X^T z
    alpha = np.zeros(M)
    for J in range(nT):
        a = DT[J] @ z[J * TB:(J + 1) * TB]
        for m in range(J + 1, nT):
            a += slot[(m, J)] @ z[m * TB:(m + 1) * TB]
        alpha[J * TB:(J + 1) * TB] = a
    assert np.allclose(alpha, np.linalg.solve(K, y))
    # ---- phase 3
    Kinv = np.linalg.inv(K)
    for J in range(nT):
        for I in range(J, nT):
            acc = np.zeros((TB, TB))
            for m in range(I, nT):
                a = DT[I] if m == I else slot[(m, I)]
                if m == J:
                    b = DT[J]
                else:
                    b = slot[(m, J)]
                acc += nt(a, b)
            assert np.allclose(acc, blk(Kinv, I, J)), (I, J)
    ll = -0.5 * z @ z - logdet - 0.5 * M * np.log(2 * np.pi)
    s, ld = np.linalg.slogdet(K)
    assert np.allclose(ll, -0.5 * y @ np.linalg.solve(K, y) - 0.5 * ld - 0.5 * M * np.log(2 * np.pi))
    return True


if __name__ == "__main__":
    rs = np.random.RandomState(0)
    for nT in (1, 2, 3, 5, 8):
        M = nT * TB
        A = rs.randn(M, M)
        K = A @ A.T + M * np.eye(M)
        assert run(K, rs.randn(M))
    print("tile model OK")


def gj_inverse_factor(A):
    """Model of potrf_inv_tile (batched4.cu): in-place sweep returning X = chol(A)^{-1} and the pivots d.
    Position (r, c): for c > j still holds the Schur complement a_rc, for c <= j holds e_rc = (L_unit^{-1})_rc."""
    n = A.shape[0]
    V = np.tril(A).copy()
    d = np.zeros(n)
    for j in range(n):
        w = np.empty(n)
        w[j:] = V[j:, j]          # column j of the current Schur complement (incl. pivot)
        w[:j] = V[j, :j]          # row j of L_unit^{-1}
        d[j] = w[j]
        invd = 1.0 / d[j]
        for r in range(j + 1, n):
            mult = w[r] * invd
            for c in range(r + 1):
                if c == j:
                    V[r, c] = -mult
                else:
                    V[r, c] -= mult * w[c]
    rs = 1.0 / np.sqrt(d)
    X = np.tril(V, -1) * rs[:, None] + np.diag(rs)
    return X, d


if __name__ == "__main__":
    rs_ = np.random.RandomState(1)
    for n in (1, 5, 64):
        B = rs_.randn(n, n)
        A = B @ B.T + n * np.eye(n)
        X, d = gj_inverse_factor(A)
        Lc = np.linalg.cholesky(A)
        assert np.allclose(X, np.linalg.inv(Lc), rtol=1e-10, atol=1e-12)
        assert np.allclose(0.5 * np.log(d).sum(), np.log(np.diag(Lc)).sum())
    print("gj_inverse_factor OK")


def blocked_gj_inverse_factor(A, b=8):
    """Model of the BLOCKED sweep (batched4.cu potrf_inv_tile): 8x8 pivot blocks, every rank-8 update a pair of
    DMMAs, in place.  Position (I,K): K > J Schur complement; (I,J) panel L_IJ then Y_IJ; K < J: Y_IK / X_JK."""
    n = A.shape[0]
    nb = n // b
    V = np.tril(A).copy()
    for i in range(nb):           # keep the diagonal blocks symmetric-full like the tile in shared memory
        V[i*b:(i+1)*b, i*b:(i+1)*b] = A[i*b:(i+1)*b, i*b:(i+1)*b]
    B = lambda I, K: V[I*b:(I+1)*b, K*b:(K+1)*b]
    logdet = 0.0
    def diag8(P):
        X, d = gj_inverse_factor(P)
        return X, 0.5 * np.log(np.prod(d))
    X0, ld = diag8(B(0, 0)); B(0, 0)[:] = X0; logdet += ld
    for J in range(nb):
        Xp = B(J, J).copy()
        for K in range(J):                      # (b) X_JK = Xp Y_JK
            B(J, K)[:] = Xp @ B(J, K)
        for I in range(J + 1, nb):              # (c) L_IJ = V_IJ Xp^T
            B(I, J)[:] = B(I, J) @ Xp.T
        for I in range(J + 1, nb):              # (d) trailing, (e) inverse part
            for K in range(J + 1, I + 1):
                B(I, K)[:] -= B(I, J) @ B(K, J).T
            for K in range(J):
                B(I, K)[:] -= B(I, J) @ B(J, K)
        for I in range(J + 1, nb):              # (e') Y_IJ = -L_IJ Xp
            B(I, J)[:] = -B(I, J) @ Xp
        if J + 1 < nb:                          # (a) next pivot block
            Xn, ld = diag8(B(J + 1, J + 1)); B(J + 1, J + 1)[:] = Xn; logdet += ld
    return np.tril(V), logdet


if __name__ == "__main__":
    rs_ = np.random.RandomState(2)
    for n in (8, 16, 64):
        Bm = rs_.randn(n, n)
        A = Bm @ Bm.T + n * np.eye(n)
        X, ld = blocked_gj_inverse_factor(A)
        Lc = np.linalg.cholesky(A)
        assert np.allclose(X, np.linalg.inv(Lc), rtol=1e-10, atol=1e-12), n
        assert np.allclose(ld, np.log(np.diag(Lc)).sum())
    print("blocked_gj_inverse_factor OK")
