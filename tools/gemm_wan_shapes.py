import os, sys, torch
sys.path.insert(0, os.getcwd())
from fastdm_b200 import ops
dev, bf = "cuda", torch.bfloat16
def interleaved(fns, rounds=8):
    """Variants timed in alternation (one call each per round): the board throttles progressively under this load, so
    back-to-back blocks of one variant each are not comparable."""
    for f in fns: f()
    torch.cuda.synchronize()
    tot = [0.0] * len(fns)
    for _ in range(rounds):
        for i, f in enumerate(fns):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            tot[i] += e0.elapsed_time(e1)
    return [x / rounds for x in tot]
for (M, K, N) in ((80640, 5120, 5120), (80640, 13824, 5120), (80640, 5120, 15360), (80640, 5120, 13824)):
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn(M, K, device=dev, generator=g).to(torch.float8_e4m3fn)
    b = (torch.randn(N, K, device=dev, generator=g) * 0.05).to(torch.float8_e4m3fn).t()
    sa = torch.rand(M, 1, device=dev) * 0.1; sb = torch.rand(N, 1, device=dev)
    bias = torch.randn(N, device=dev).to(bf)
    res = torch.randn(M, N, device=dev).to(bf); gate = torch.randn(1, N, device=dev).to(bf).float()
    out = torch.empty(M, N, device=dev, dtype=bf)
    fl = 2.0 * M * N * K
    p, ge, gt, gb = interleaved([
        lambda: ops.fp8_matmul(a, b, sa, sb, bf, bias, out=out),
        lambda: ops.fp8_matmul(a, b, sa, sb, bf, bias, out=out, act="gelu_tanh"),
        lambda: ops.fp8_matmul(a, b, sa, sb, bf, bias, out=out, gate=gate, residual=res, rows_per_batch=M, round_steps=False),
        lambda: ops.fp8_matmul(a, b, sa, sb, bf, bias, out=out, gate=gate, residual=res, rows_per_batch=M, round_steps=True)])
    print(f"M{M} K{K} N{N}: plain {p*1e3:.0f} us {fl/p/1e9:.0f} TF | gelu {ge*1e3:.0f} us {fl/ge/1e9:.0f} | gated fp32 chain (Wan) {gt*1e3:.0f} us {fl/gt/1e9:.0f} | gated bf16 chain {gb*1e3:.0f} us {fl/gb/1e9:.0f}", flush=True)
    del a, b, res, out
