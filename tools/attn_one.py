import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
b, s, h, hd = 1, 8704, 24, 128
q = torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16)
k = torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16)
v = torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    y = ops.scaled_dot_product_attention(q, k, v, h, h, hd)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
