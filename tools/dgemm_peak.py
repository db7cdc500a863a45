"""Measure the cuBLAS DGEMM peak (roofline denominator for fp64 tensor-pipe stages). Library call, not product."""
import json, torch
assert torch.cuda.is_available()
res = {}
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2): torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{n}_tflops"] = 2 * n**3 / best / 1e9
    # sustained 2 s
    import time
    t0 = time.time(); cnt = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 2.0:
        torch.matmul(a, b); cnt += 1
        if cnt % 8 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    res[f"dgemm_{n}_tflops_sustained"] = 2 * n**3 * cnt / e0.elapsed_time(e1) / 1e9
print(json.dumps(res))
