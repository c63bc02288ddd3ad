"""K8 evidence: the dense FP64 tensor-core Cholesky of the explicit Schur path next to cuSOLVER's potrf
(torch.linalg.cholesky) and to the FP64 GEMM rate cuBLAS reaches on the same GPU (the practical DMMA peak)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from apex_solver_b200.context import GpuContext

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {"gpu": torch.cuda.get_device_name(0), "rows": []}
m = 8192
a = torch.randn(m, m, dtype=torch.float64, device="cuda"); b = torch.randn(m, m, dtype=torch.float64, device="cuda")
ms = timed(lambda: torch.matmul(a, b), 5)
out["dgemm_8192_tflops"] = 2 * m ** 3 / ms / 1e9
del a, b
g = GpuContext()
for n in (4032, 8064, 14016, 24000):
    flops = n ** 3 / 3
    ours = g.dense_cholesky_bench(n, 3)
    x = torch.rand(n, n, dtype=torch.float64, device="cuda") * 2 - 1
    A = (x + x.T) * 0.5
    A.diagonal().copy_(torch.full((n,), float(n), dtype=torch.float64, device="cuda"))
    del x
    cus = timed(lambda: torch.linalg.cholesky(A), 3)
    del A
    torch.cuda.empty_cache()
    out["rows"].append({"n": n, "ours_ms": ours, "ours_tflops": flops / ours / 1e9, "cusolver_ms": cus, "cusolver_tflops": flops / cus / 1e9,
                        "ours_over_cusolver": cus / ours, "ours_frac_of_dgemm": flops / ours / 1e9 / out["dgemm_8192_tflops"]})
print(json.dumps(out))
