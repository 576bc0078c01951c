"""One W8A8 GEMM launch set for `ncu --set full -k regex:gemm_w8a8` captures (FLUX ff.proj_in)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
m, k, n = 8704, 3072, 12288
x = torch.randn(m, k, device="cuda", dtype=torch.bfloat16)
w = torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * 0.02
xq, xs = ops.quantize_to_fp8(x)
wq, ws = ops.quantize_to_fp8(w)
bias = torch.randn(n, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    y = ops.fp8_matmul(xq, wq.t(), xs, ws.view(-1), torch.bfloat16, bias)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
