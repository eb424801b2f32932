"""Timings of the non-headline configurations of BASELINE.json through the C-ABI (development / profiles aid).
usage: bench_configs.py [c4|c2|c5|all] [M_locations]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from numpy.random import RandomState
from gptools_b200._lib import Device


def timed(f, n=1):
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        r = f()
    return (time.perf_counter() - t0) / n, r


def c4(nloc):
    rs = RandomState(0)
    X0 = rs.rand(nloc, 2)
    f = lambda x: np.sin(3 * x[:, 0]) * np.cos(2 * x[:, 1])
    X = np.vstack([X0, X0, X0])
    n = np.vstack([np.zeros((nloc, 2), int), np.tile([1, 0], (nloc, 1)), np.tile([0, 1], (nloc, 1))])
    y = np.concatenate([f(X0), 3 * np.cos(3 * X0[:, 0]) * np.cos(2 * X0[:, 1]),
                        -2 * np.sin(3 * X0[:, 0]) * np.sin(2 * X0[:, 1])]) + 0.05 * rs.randn(3 * nloc)
    M = 3 * nloc
    d = Device(0)
    d.set_data(X, n, y, np.full(M, 0.05))
    d.set_kernel(0, 3, 1e2)
    th = np.array([1.0, 0.05, 0.05]) if nloc >= 8192 else np.array([1.0, 0.1, 0.1])
    t, (ll, _, st) = timed(lambda: d.ll(th, 0.0))
    print("C4-shape M=%d: assemble+potrf+solve %.3f s  -> %.2f TFLOP/s (M^3/3)  status %d ll %.6g" % (
        M, t, M ** 3 / 3.0 / t * 1e-12, st, ll))
    if M <= 12288:
        t, (ll, g, st) = timed(lambda: d.ll(th, 0.0, grad_idx=[0, 1, 2]))
        print("   ll+grad (inverse + fused trace) %.3f s -> %.2f TFLOP/s (M^3)" % (t, M ** 3 / t * 1e-12))
        d.ll(th, 0.0)
    for Ms in (100000, 1000000):
        Xs = RandomState(2).rand(Ms, 2)
        ns = np.zeros((Ms, 2), dtype=int)
        t, _ = timed(lambda: d.predict(Xs, ns, want_var=False))
        print("   predict mean only  M*=%d: %.3f s  %.3g points/s  (equivalent K* bytes %.0f GB/s)" % (
            Ms, t, Ms / t, 8.0 * M * Ms / t * 1e-9))
    Ms = 100000 if M <= 24576 else 28416   # one whole wave of the solve GEMMs (148 SMs x 3 CTAs x 64 rows)
    Xs = RandomState(2).rand(Ms, 2)
    ns = np.zeros((Ms, 2), dtype=int)
    t, _ = timed(lambda: d.predict(Xs, ns, want_var=True))
    print("   predict mean+var   M*=%d: %.3f s  %.3g points/s  trsm %.2f TFLOP/s (M^2 M*)" % (
        Ms, t, Ms / t, float(M) ** 2 * Ms / t * 1e-12))


def c2():
    rs = RandomState(0)
    X = np.sort(rs.rand(2000)) * 10
    Xa = np.concatenate([X, X])[:, None]
    n = np.concatenate([np.zeros(2000, int), np.ones(2000, int)])[:, None]
    y = np.concatenate([np.sin(X), np.cos(X)]) + 0.05 * rs.randn(4000)
    d = Device(0)
    d.set_data(Xa, n, y, np.full(4000, 0.05))
    for name, kid, th in (("Matern52", 1, [1.0, 0.8]), ("Matern(nu=5/2)", 2, [1.0, 2.5, 0.8])):
        d.set_kernel(kid, len(th), 1e2)
        t, (ll, _, st) = timed(lambda: d.ll(np.array(th), 0.0), 3)
        gi = [0, 1] if kid == 1 else [0, 2]
        tg, (_, g, _) = timed(lambda: d.ll(np.array(th), 0.0, grad_idx=gi), 3)
        print("C2 %s M=4000: ll+grad %.4f s (grad %s)" % (name, tg, np.array2string(g, precision=6)))
        d.ll(np.array(th), 0.0)
        Xs = np.linspace(0, 10, 100000)[:, None]
        tp, _ = timed(lambda: d.predict(Xs, np.zeros((100000, 1), int), want_var=True))
        tm, _ = timed(lambda: d.predict(Xs, np.zeros((100000, 1), int), want_var=False))
        print("C2 %s M=4000: ll %.4f s (status %d, ll %.6f); predict 1e5 mean+std %.3f s (%.3g pts/s), mean only %.3f s" % (
            name, t, st, ll, tp, 1e5 / tp, tm))


def c5():
    rs = RandomState(0)
    Nq, Mo, W = 4000, 500, 400
    Xq = np.linspace(0, 1.1, Nq)
    T = np.zeros((Mo + 1, Nq + 1))
    for i, s in enumerate(rs.randint(0, Nq - W, size=Mo)):
        T[i, s:s + W] = 1.1 / Nq
    T[Mo, Nq] = 1.0
    X = np.concatenate([Xq, [0.0]])[:, None]
    n = np.concatenate([np.zeros(Nq, int), [1]])[:, None]
    y = np.concatenate([rs.rand(Mo) * 0.3 + 0.1, [0.0]])
    err = np.concatenate([np.full(Mo, 0.02), [0.0]])
    d = Device(0)
    d.set_data(X, n, y, err, T)
    d.set_kernel(3, 5, 1e2)
    th = np.array([1.5, 0.6, 0.1, 0.05, 0.9])
    t, (ll, _, st) = timed(lambda: d.ll(th, 0.0), 3)
    Xs = np.linspace(0, 1.1, 400)[:, None]
    tp, (mean, var, cov) = timed(lambda: d.predict(Xs, np.zeros((400, 1), int), want_cov=True), 3)
    rv = rs.randn(400, 1000)
    td, (samp, sst) = timed(lambda: d.draw_sample(mean, cov, rv, 1e3 * 2.220446049250313e-16), 3)
    print("C5 Gibbs+T N=4001 M=501: ll %.4f s (status %d ll %.6f); predict full cov 400 pts %.4f s; draw 1000 samples %.4f s (status %d)" % (
        t, st, ll, tp, td, sst))
    # 64 walkers: one gpt_ll_batched call (thetas back to back on the device) against 64 gpt_ll calls
    B = 64
    ths = np.hstack([th * np.exp(0.03 * rs.randn(B, 5)), np.zeros((B, 1))])
    gi = [0, 1, 2, 3, 4]
    tb, (llb, gb, stb) = timed(lambda: d.ll_batched(ths, grad_idx=gi), 1)
    tl, _ = timed(lambda: [d.ll(ths[b, :5], 0.0, grad_idx=gi) for b in range(B)], 1)
    tbv, _ = timed(lambda: d.ll_batched(ths), 1)
    tlv, _ = timed(lambda: [d.ll(ths[b, :5], 0.0) for b in range(B)], 1)
    print("C5 64 thetas, ll+grad: gpt_ll_batched %.4f s (%.2f ms / theta), 64 x gpt_ll %.4f s; ll only: %.4f s against %.4f s; all ok %s" % (
        tb, 1e3 * tb / B, tl, tbv, tlv, (stb == 0).all()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("c2", "all"):
        c2()
    if what in ("c5", "all"):
        c5()
    if what in ("c4", "all"):
        for nloc in ([int(sys.argv[2])] if len(sys.argv) > 2 else [4096, 8192]):
            c4(nloc)
