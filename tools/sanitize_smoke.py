"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
dev, bf = "cuda", torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
# attention: CTA-pair path (hd 128, Sk >= 1024), legacy path (hd 64, short Sk, masked), fp8
# attention: persistent CTA-pair kernel (hd 128, 9 KV tiles), one-item CTA-pair kernel with rotating issuers (18 tiles), hd 64 (dense: P next to
# O in TMEM, rotating issuers), short Sk
for (b, sq, sk, h, hd) in ((1, 700, 1100, 2, 128), (1, 700, 2200, 1, 128), (2, 300, 200, 2, 64), (1, 300, 2300, 1, 64), (1, 520, 512, 1, 128)):
    q = torch.randn(b, sq, h * hd, device=dev, generator=g).to(bf)
    k = torch.randn(b, sk, h * hd, device=dev, generator=g).to(bf)
    v = torch.randn(b, sk, h * hd, device=dev, generator=g).to(bf)
    y = ops.scaled_dot_product_attention(q, k, v, h, h, hd)
    assert torch.isfinite(y.float()).all()
q = torch.randn(1, 640, 2 * 128, device=dev, generator=g).to(bf); k = torch.randn(1, 1300, 2 * 128, device=dev, generator=g).to(bf)
v = torch.randn(1, 1300, 2 * 128, device=dev, generator=g).to(bf)
mask = (torch.rand(1, 2, 5, 21, device=dev) < 0.5).to(torch.int8); mask[:, :, :, 0] = 1
ops.sparse_scaled_dot_product_attention(q, k, v, 2, 2, 128, sparse_mask=mask, block_q=128, block_k=64)
# GEMMs: plain, gelu, gated residual, int8, ragged N
x = torch.randn(300, 256, device=dev, generator=g).to(bf); w = (torch.randn(208, 256, device=dev, generator=g) * 0.05).to(bf)
xq, xs = ops.quantize_to_fp8(x); wq, ws = ops.quantize_to_fp8(w)
bias = torch.randn(208, device=dev).to(bf)
ops.fp8_matmul(xq, wq.t(), xs, ws.view(-1), bf, bias)
ops.fp8_matmul(xq, wq.t(), xs, ws.view(-1), bf, bias, act="gelu_tanh")
res = torch.randn(300, 208, device=dev).to(bf); gate = torch.randn(1, 208, device=dev); out = torch.empty_like(res)
ops.fp8_matmul(xq, wq.t(), xs, ws.view(-1), bf, bias, out=out, gate=gate, residual=res, rows_per_batch=300)
xi, si, zi = ops.quantize_to_int8(x, False); wi, wsi = ops.quantize_to_int8(w, True)[:2]
adj = wi.to(torch.int32).sum(dim=1, dtype=torch.int32).view(1, -1).contiguous()
ops.int8_matmul(xi, wi.t(), si, wsi.view(-1), bf, adj, zi, bias)
# CTA-pair GEMM (>= 74 tiles of 256 x 256): plain, GELU, gated residual, int8 with zero points; ragged M (4870 rows)
xl = torch.randn(4870, 256, device=dev, generator=g).to(bf); wl = (torch.randn(1024, 256, device=dev, generator=g) * 0.05).to(bf)
xlq, xls = ops.quantize_to_fp8(xl); wlq, wls = ops.quantize_to_fp8(wl)
bl = torch.randn(1024, device=dev).to(bf)
yl = ops.fp8_matmul(xlq, wlq.t(), xls, wls.view(-1), bf, bl)
ref = (xlq.float() * xls) @ (wlq.float() * wls).t() + bl.float()
assert (yl.float() - ref).abs().max() <= 0.02 * ref.abs().max() + 0.05, "pair GEMM mismatch"
ops.fp8_matmul(xlq, wlq.t(), xls, wls.view(-1), bf, bl, act="gelu_tanh")
resl = torch.randn(4870, 1024, device=dev).to(bf); gl = torch.randn(1, 1024, device=dev).to(bf).float(); outl = torch.empty_like(resl)
ops.fp8_matmul(xlq, wlq.t(), xls, wls.view(-1), bf, bl, out=outl, gate=gl, residual=resl, rows_per_batch=4870)
xli, sli, zli = ops.quantize_to_int8(xl, False); wli, wlsi = ops.quantize_to_int8(wl, True)[:2]
adjl = wli.to(torch.int32).sum(dim=1, dtype=torch.int32).view(1, -1).contiguous()
ops.int8_matmul(xli, wli.t(), sli, wlsi.view(-1), bf, adjl, zli, bl)
# elementwise
ops.rms_norm(torch.randn(77, 4, 128, device=dev).to(bf), torch.randn(128, device=dev).to(bf), 1e-6)
qq = torch.randn(1, 77, 4 * 128, device=dev).to(bf); kk = torch.randn(1, 77, 4 * 128, device=dev).to(bf)
ops.rotary_pos_embedding(qq, kk, 128, torch.rand(77, 128, device=dev).to(bf), False)
ops.gelu_and_mul(torch.randn(50, 512, device=dev).to(bf))
ops.layernorm_modulate_quant(torch.randn(90, 256, device=dev).to(bf), torch.rand(1, 256, device=dev).to(bf).float(),
                             torch.rand(1, 256, device=dev).to(bf).float(), 90, torch.float8_e4m3fn)
qkv = torch.randn(77, 3 * 4 * 128, device=dev).to(bf)
ops.qk_norm_rope_(qkv, torch.randn(128, device=dev).to(bf), torch.randn(128, device=dev).to(bf),
                  torch.rand(77, 128, device=dev).to(bf), 4, 4, 128, 0, 512, 0, 1e-6)
torch.cuda.synchronize()
print("sanitize smoke ok")
