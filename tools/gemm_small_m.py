"""Small-M (text stream) FP8 GEMMs: time per tile width (FDM_GEMM_BN) against cuBLASLt."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import os, sys, torch
sys.path.insert(0, %r)
from fastdm_b200 import ops
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (m, k, n) in ((512, 3072, 9216), (512, 3072, 3072), (512, 3072, 12288), (512, 12288, 3072), (1024, 5120, 5120), (4352, 3072, 9216)):
    x = torch.randn(m, k, device="cuda", dtype=torch.bfloat16); w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * 0.02
    xq, xs = ops.quantize_to_fp8(x); wq, ws = ops.quantize_to_fp8(w)
    us = t(lambda: ops.fp8_matmul(xq, wq.t(), xs, ws.view(-1), torch.bfloat16, None))
    ref = t(lambda: torch._scaled_mm(xq, wq.t(), xs, ws.view(1, -1), out_dtype=torch.bfloat16))
    print(f"M{m} K{k} N{n}: {us:7.1f} us | cublasLt {ref:7.1f} us")
''' % ROOT
for bn in ("auto", "64", "128", "256"):
    env = dict(os.environ)
    if bn != "auto":
        env["FDM_GEMM_BN"] = bn
    print("--- BN", bn)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-1500:])
