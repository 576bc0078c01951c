"""One launch of every memory-bound kernel at a Wan2.2 / FLUX shape, for `ncu --set full` (achieved HBM GB/s)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
dev, bf = "cuda", torch.bfloat16
M, K = 80640, 5120
x = torch.randn(M, K, device=dev, dtype=bf)
ops.quantize_to_fp8(x)                                              # quant_row_kernel fp8
ops.quantize_to_int8(x, False)                                      # quant_row_kernel int8 asym
a, c = torch.rand(1, K, device=dev) + 0.5, torch.rand(1, K, device=dev)
ops.layernorm_modulate_quant(x, a, c, M, torch.float8_e4m3fn)       # ln_mod_quant_kernel
S, H, hd = 8704, 24, 128
xr = torch.randn(S * H, hd, device=dev, dtype=bf)
ops.rms_norm(xr, torch.randn(hd, device=dev, dtype=bf), 1e-6)       # rmsnorm
q = torch.randn(1, S, H * hd, device=dev, dtype=bf)
k = torch.randn(1, S, H * hd, device=dev, dtype=bf)
cs = torch.rand(S, hd, device=dev, dtype=bf)
ops.rotary_pos_embedding(q, k, hd, cs, False)                       # rope_kernel
g = torch.randn(8192, 2 * 12288, device=dev, dtype=bf)
ops.gelu_and_mul(g)                                                 # gelu_mul_kernel
qkv = torch.randn(S, 3 * H * hd, device=dev, dtype=bf)
w = torch.randn(hd, device=dev, dtype=bf)
ops.qk_norm_rope_(qkv, w, w, cs, H, H, hd, 0, H * hd, 0, 1e-6)      # qk_norm_rope_head_kernel (FLUX)
Hw = 40
qkvw = torch.randn(20160, 3 * Hw * hd, device=dev, dtype=bf)
ww = torch.randn(Hw * hd, device=dev, dtype=bf)
csw = torch.rand(20160, hd, device=dev, dtype=bf)
ops.qk_norm_rope_(qkvw, ww, ww, csw, Hw, Hw, hd, 0, Hw * hd, 0, 1e-6, across_heads=True)  # qk_norm_rope_row_kernel (Wan)
torch.cuda.synchronize()
print("done")
