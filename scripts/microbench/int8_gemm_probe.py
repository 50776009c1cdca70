#!/usr/bin/env python
"""What does the library int8 GEMM (cuBLASLt through torch._int_mm) reach on this GPU, next to bf16?  Context for K2's roofline:
K2's tcgen05 kind::i8 pipeline alone (UAVM_K2_DBG=2) runs at 2.38 Pop/s."""
import json, torch
dev = "cuda"
out = {}
for n in (4096, 8192):
    a = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev)
    b = torch.randint(-100, 100, (n, n), dtype=torch.int8, device=dev).t().contiguous().t()
    for _ in range(3): torch._int_mm(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): torch._int_mm(a, b)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[f"int8_{n}"] = {"ms": ms, "tops": 2 * n ** 3 / ms / 1e9}
    x = torch.randn(n, n, dtype=torch.bfloat16, device=dev); y = torch.randn(n, n, dtype=torch.bfloat16, device=dev)
    for _ in range(3): x @ y
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20): x @ y
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out[f"bf16_{n}"] = {"ms": ms, "tflops": 2 * n ** 3 / ms / 1e9}
print(json.dumps(out))
