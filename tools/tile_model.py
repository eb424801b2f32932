"""Numpy model of the batched kernel's tile algorithm (gptools_b200/csrc/batched.cu), used to validate the
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
