"""Library baselines on the B200 (not the product): cuBLAS Dgemm = FP64 roofline
denominator; torch.linalg.cholesky (cuSOLVER/MAGMA) batched 4096x512 and single 16384."""
import json, time, torch
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
out = {"gpu": torch.cuda.get_device_name(0)}
def ev(f, n=3):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = ev(lambda: torch.matmul(a, b))
    out["dgemm_%d_tflops" % n] = 2.0 * n ** 3 / ms * 1e-9
    del a, b
# sustained dgemm (2 s)
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
torch.cuda.synchronize(); t0 = time.time(); cnt = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    for _ in range(5): torch.matmul(a, b); cnt += 1
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["dgemm_8192_sustained_tflops"] = cnt * 2.0 * n ** 3 / e0.elapsed_time(e1) * 1e-9
del a, b
# batched cholesky 4096 x 512
B, M = 4096, 512
x = torch.randn(B, M, M // 4, dtype=torch.float64, device=dev)
A = x @ x.transpose(1, 2) + M * torch.eye(M, dtype=torch.float64, device=dev)
del x
ms = ev(lambda: torch.linalg.cholesky(A), n=2)
out["torch_batched_cholesky_4096x512_ms"] = ms
out["torch_batched_cholesky_tflops_M3over3"] = B * M ** 3 / 3.0 / ms * 1e-9
ms = ev(lambda: torch.linalg.inv(A[:1024]), n=2)
out["torch_batched_inv_1024x512_ms"] = ms
del A
n = 16384
x = torch.randn(n, n, dtype=torch.float64, device=dev)
A = x @ x.T + n * torch.eye(n, dtype=torch.float64, device=dev); del x
ms = ev(lambda: torch.linalg.cholesky(A), n=2)
out["torch_cholesky_16384_ms"] = ms
out["torch_cholesky_16384_tflops"] = n ** 3 / 3.0 / ms * 1e-9
print(json.dumps(out, indent=1))
