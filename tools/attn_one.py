"""One attention launch for `ncu --set full -k regex:attn_fwd` captures.

    python tools/attn_one.py [flux|wan]     flux: 8704 tokens x 24 heads, wan: 80640 tokens x 40 heads
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "flux"
b, s, h, hd = (1, 80640, 40, 128) if which == "wan" else (1, 8704, 24, 128)
q = torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16)
k = torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16)
v = torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    y = ops.scaled_dot_product_attention(q, k, v, h, h, hd)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
