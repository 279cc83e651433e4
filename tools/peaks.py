"""Measure library GEMM throughput on this GPU (roofline denominators the driver did not write): cuBLASLt int8
(torch._int_mm), bf16 (torch.matmul) and fp8 (torch._scaled_mm), best of 10 at 8192^3 (16384 for fp8/int8 too)."""
import json, sys, torch
dev = torch.device("cuda", 0)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
out = {}
for n in (8192, 16384):
    a = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev).t().contiguous().t()
    try:
        ms = timeit(lambda: torch._int_mm(a, b))
        out[f"int8_tops_{n}"] = round(2.0 * n ** 3 / ms / 1e9, 1)
    except Exception as e:  # noqa: BLE001
        out[f"int8_{n}_error"] = str(e)[:200]
    x = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    y = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    ms = timeit(lambda: torch.matmul(x, y))
    out[f"bf16_tflops_{n}"] = round(2.0 * n ** 3 / ms / 1e9, 1)
    try:
        xf = x.to(torch.float8_e4m3fn)
        yf = y.to(torch.float8_e4m3fn).t().contiguous().t()
        one = torch.ones((), device=dev)
        ms = timeit(lambda: torch._scaled_mm(xf, yf, scale_a=one, scale_b=one, out_dtype=torch.bfloat16))
        out[f"fp8_tflops_{n}"] = round(2.0 * n ** 3 / ms / 1e9, 1)
    except Exception as e:  # noqa: BLE001
        out[f"fp8_{n}_error"] = str(e)[:200]
    del a, b, x, y
print(json.dumps(out))
