import cProfile, pstats, io, os, sys, time, torch
sys.path.insert(0, os.getcwd())
from fastdm_b200 import ops
dev, bf = "cuda", torch.bfloat16
x = torch.randn(512, 3072, device=dev, dtype=bf)
w = torch.randn(3072, 3072, device=dev, dtype=bf) * 0.02
wq, ws = ops.quantize_to_fp8(w)
bias = torch.randn(3072, device=dev, dtype=bf)
wsv = ws.view(-1)
def step():
    xq, xs = ops.quantize_to_fp8(x)
    y = ops.fp8_matmul(xq, wq.t(), xs, wsv, bf, bias)
    return y
for _ in range(50): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per (quant + gemm) pair: {(t1 - t0) / 2000 * 1e6:.1f} us issue, {(t2 - t0) / 2000 * 1e6:.1f} us incl. drain")
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
