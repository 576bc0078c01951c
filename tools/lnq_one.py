"""One launch each of the quant kernel and the LayerNorm-modulate-quant variants on [80640, 5120] for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
dev, bf = "cuda", torch.bfloat16
M, K = 80640, 5120
x = torch.randn(M, K, device=dev, dtype=bf)
a, c = (torch.rand(1, K, device=dev) + 0.5).to(bf), torch.rand(1, K, device=dev).to(bf)
for _ in range(2):
    ops.quantize_to_fp8(x)
    ops.layernorm_modulate_quant(x, a, c, M, torch.float8_e4m3fn)                         # bf16 modulation vectors
    ops.layernorm_modulate_quant(x, a.float(), c.float(), M, torch.float8_e4m3fn)         # fp32 vectors, bf16 chain
    ops.layernorm_modulate_quant(x, a.float() + 1e-4, c.float(), M, torch.float8_e4m3fn, round_steps=False)  # Wan fp32 chain
torch.cuda.synchronize()
print("done")
