"""One plain and one `residual + gate * out` FP8 GEMM launch (FLUX to_out shape) for ncu captures."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
m, k, n = 8192, 3072, 3072
x = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * 0.02
xq, xs = ops.quantize_to_fp8(x)
wq, ws = ops.quantize_to_fp8(w)
bias = torch.randn(n, device="cuda", dtype=torch.bfloat16)
res = torch.randn(m, n, device="cuda", dtype=torch.bfloat16)
gate = torch.randn(1, n, device="cuda")
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.fp8_matmul(xq, wq.t(), xs, ws.view(-1), torch.bfloat16, bias, out=out)
    ops.fp8_matmul(xq, wq.t(), xs, ws.view(-1), torch.bfloat16, bias, out=out, gate=gate, residual=res, rows_per_batch=m)
torch.cuda.synchronize()
print(float(out.float().abs().mean()))
