"""Block-sparse attention with the reference's radial mask at the Wan2.2 shape (4 heads): time vs dense."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
from fastdm_b200.sparse import radial_block_mask, sparge_mask_convert
dev, bf = "cuda", torch.bfloat16
frames, tpf, h, hd = 21, 3840, 4, 128
s = frames * tpf
q, k, v = (torch.randn(1, s, h * hd, device=dev, dtype=bf) for _ in range(3))
conv = sparge_mask_convert(radial_block_mask(frames, tpf, 64, 0.3, "wan", device=dev), 64)
mask = conv.to(torch.int8)[None, None].expand(1, h, -1, -1).contiguous()


def timeit(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


dense = timeit(lambda: ops.scaled_dot_product_attention(q, k, v, h, h, hd))
sparse = timeit(lambda: ops.sparse_scaled_dot_product_attention(q, k, v, h, h, hd, sparse_mask=mask, block_q=128, block_k=64))
dens = conv.float().mean().item()
print(f"dense {dense:.3f} ms | radial-sparse {sparse:.3f} ms ({dense / sparse:.2f}x) at block density {dens:.3f} "
      f"(ideal {1 / dens:.2f}x); effective {4.0 * h * s * s * hd * dens / sparse / 1e9:.0f} TFLOP/s on the kept blocks")
