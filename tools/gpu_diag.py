"""GPU-side diagnostics (run under gpurun): prints detailed error patterns for the tcgen05 kernels
so a wrong descriptor / layout can be diagnosed from one trip. Not a test; not a benchmark."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastdm_b200 import ops  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16
out = {}


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_patterns():
    print("== GEMM structured patterns ==")
    for (M, K, N) in ((128, 128, 64), (128, 128, 256), (128, 256, 256), (256, 512, 512), (130, 144, 48)):
        # 1. all ones
        a = torch.ones(M, K, device=DEV).to(torch.float8_e4m3fn)
        b = torch.ones(N, K, device=DEV).to(torch.float8_e4m3fn).t()
        s_m = torch.ones(M, 1, device=DEV)
        s_n = torch.ones(N, 1, device=DEV)
        y = ops.fp8_matmul(a, b, s_m, s_n, BF, None) if N % 16 == 0 and K % 16 == 0 else None
        torch.cuda.synchronize()
        if y is not None:
            bad = (y.float() != K)
            print(f"ones  M{M} K{K} N{N}: wrong={int(bad.sum())}/{y.numel()} sample={y[0, :4].tolist()} uniq={torch.unique(y.float())[:8].tolist()}")
        # 2. random small ints (exact in fp8 and fp32)
        g = torch.Generator(device=DEV).manual_seed(1)
        ai = torch.randint(-3, 4, (M, K), device=DEV, generator=g).float()
        bi = torch.randint(-3, 4, (N, K), device=DEV, generator=g).float()
        a = ai.to(torch.float8_e4m3fn)
        b = bi.to(torch.float8_e4m3fn).t()
        y = ops.fp8_matmul(a, b, s_m, s_n, torch.float16, None)
        ref = ai @ bi.t()
        bad = (y.float() != ref)
        print(f"ints  M{M} K{K} N{N}: wrong={int(bad.sum())}/{y.numel()}")
        if bad.any():
            rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
            cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
            print("   bad rows:", rows, " bad cols:", cols)
            print("   got :", y[:4, :8].float().tolist())
            print("   want:", ref[:4, :8].tolist())
            # does y match ref with K truncated to the first 32 / 64 / 96 elements? (K-advance bug)
            for kk in (32, 64, 96, 128):
                if kk <= K:
                    r2 = ai[:, :kk] @ bi[:, :kk].t()
                    print(f"   match with K[:{kk}] only: {int((y.float() == r2).sum())}/{y.numel()}")
        # int8
        a8 = ai.to(torch.int8)
        b8 = bi.to(torch.int8).t()
        y = ops.int8_matmul(a8, b8, s_m, s_n, torch.float16, None, None, None)
        bad = (y.float() != ref)
        print(f"int8  M{M} K{K} N{N}: wrong={int(bad.sum())}/{y.numel()}")
        if bad.any():
            print("   got :", y[:4, :8].float().tolist())
            print("   want:", ref[:4, :8].tolist())


def gemm_perf():
    print("== GEMM perf (TFLOP/s) ==")
    res = {}
    for (M, K, N) in ((8192, 3072, 9216), (512, 3072, 9216), (8192, 3072, 12288), (8192, 12288, 3072),
                      (8704, 15360, 3072), (4608, 3072, 9216), (8192, 8192, 8192)):
        g = torch.Generator(device=DEV).manual_seed(1)
        a = torch.randn(M, K, device=DEV, generator=g).to(torch.float8_e4m3fn)
        b = torch.randn(N, K, device=DEV, generator=g).to(torch.float8_e4m3fn).t()
        sa = torch.rand(M, 1, device=DEV)
        sb = torch.rand(N, 1, device=DEV)
        bias = torch.randn(N, device=DEV).to(BF)
        ms = timeit(lambda: ops.fp8_matmul(a, b, sa, sb, BF, bias))
        tf = 2.0 * M * N * K / ms / 1e9
        # cuBLASLt rowwise baseline (what the reference's torch backend runs on B200)
        try:
            ms_t = timeit(lambda: torch._scaled_mm(a, b, sa, sb.t(), bias, out_dtype=BF))
            tf_t = 2.0 * M * N * K / ms_t / 1e9
        except Exception as e:  # noqa: BLE001
            ms_t, tf_t = float("nan"), float("nan")
            print("   torch._scaled_mm failed:", str(e)[:200])
        a8 = torch.randint(-128, 128, (M, K), device=DEV, generator=g).to(torch.int8)
        b8 = torch.randint(-128, 128, (N, K), device=DEV, generator=g).to(torch.int8).t()
        adj = torch.randint(-128, 127, (1, N), device=DEV).to(torch.int32)
        azp = torch.randint(-128, 127, (M, 1), device=DEV).to(torch.int32)
        ms8 = timeit(lambda: ops.int8_matmul(a8, b8, sa, sb, BF, adj, azp, bias))
        tf8 = 2.0 * M * N * K / ms8 / 1e9
        print(f"M{M} K{K} N{N}: fp8 {ms:.3f} ms {tf:.0f} TF | cublasLt fp8 {ms_t:.3f} ms {tf_t:.0f} TF | int8 {ms8:.3f} ms {tf8:.0f} TOP")
        res[f"{M}x{K}x{N}"] = dict(fp8_ms=ms, fp8_tflops=tf, cublaslt_ms=ms_t, cublaslt_tflops=tf_t, int8_ms=ms8, int8_tops=tf8)
    out["gemm_perf"] = res


def elementwise_perf():
    print("== elementwise perf (GB/s algorithmic) ==")
    res = {}
    # L2 flush by READING a 512 MB buffer: a memset would leave 126 MB of dirty lines whose write-back
    # then competes with the kernel under test (a ~30 us artefact for kernels that move < 100 MB)
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.float32, device=DEV)

    def timed(fn, iters=10):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.sum()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters

    for (M, K) in ((4608, 3072), (8704, 3072), (8704, 12288), (8704, 15360), (80640, 5120), (80640, 13824)):
        x = torch.randn(M, K, device=DEV, dtype=BF)
        ms = timed(lambda: ops.quantize_to_fp8(x))
        gb = (3 * M * K + 4 * M) / ms / 1e6
        ms8 = timed(lambda: ops.quantize_to_int8(x, False))
        gb8 = (3 * M * K + 8 * M) / ms8 / 1e6
        print(f"quant [{M},{K}]: fp8 {ms*1e3:.1f} us {gb:.0f} GB/s | int8 asym {ms8*1e3:.1f} us {gb8:.0f} GB/s")
        res[f"quant_{M}x{K}"] = dict(fp8_us=ms * 1e3, fp8_gbs=gb, int8_us=ms8 * 1e3, int8_gbs=gb8)
    for shape in ((1, 8704, 24, 128), (1, 80640, 5120), (2, 4685, 24, 64)):
        x = torch.randn(*shape, device=DEV, dtype=BF)
        w = torch.randn(shape[-1], device=DEV, dtype=BF)
        ms = timed(lambda: ops.rms_norm(x, w, 1e-6))
        gb = 4 * x.numel() / ms / 1e6
        print(f"rms_norm {shape}: {ms*1e3:.1f} us {gb:.0f} GB/s")
        res[f"rmsnorm_{'x'.join(map(str, shape))}"] = dict(us=ms * 1e3, gbs=gb)
    for (S, d) in ((8704, 3072), (80640, 5120)):
        q = torch.randn(1, S, d, device=DEV, dtype=BF)
        k = torch.randn(1, S, d, device=DEV, dtype=BF)
        cs = torch.rand(S, 128, device=DEV, dtype=BF)
        ms = timed(lambda: ops.rotary_pos_embedding(q, k, 128, cs, False))
        gb = (8 * S * d + 2 * S * 128) / ms / 1e6
        print(f"rope [{S},{d}]: {ms*1e3:.1f} us {gb:.0f} GB/s")
        res[f"rope_{S}x{d}"] = dict(us=ms * 1e3, gbs=gb)
    for (M, d) in ((8192, 5120), (2048, 10240)):
        x = torch.randn(M, 2 * d, device=DEV, dtype=BF)
        ms = timed(lambda: ops.gelu_and_mul(x))
        gb = 6 * M * d / ms / 1e6
        print(f"gelu_and_mul [{M},{2*d}]: {ms*1e3:.1f} us {gb:.0f} GB/s")
        res[f"gelumul_{M}x{d}"] = dict(us=ms * 1e3, gbs=gb)
    out["elementwise_perf"] = res


def attn_perf():
    print("== attention perf (TFLOP/s) ==")
    res = {}
    for (b, sq, sk, h, hd) in ((1, 4608, 4608, 24, 128), (1, 8704, 8704, 24, 128), (2, 4685, 4685, 24, 64),
                               (1, 80640, 80640, 4, 128), (1, 80640, 512, 40, 128)):
        q = torch.randn(b, sq, h * hd, device=DEV, dtype=BF)
        k = torch.randn(b, sk, h * hd, device=DEV, dtype=BF)
        v = torch.randn(b, sk, h * hd, device=DEV, dtype=BF)
        ms = timeit(lambda: ops.scaled_dot_product_attention(q, k, v, h, h, hd), iters=5, warm=2)
        fl = 4.0 * b * h * sq * sk * hd
        qt, kt, vt = (t.view(b, -1, h, hd).transpose(1, 2) for t in (q, k, v))
        try:
            ms_t = float("nan") if os.environ.get("FDM_DIAG_NO_TORCH") else timeit(
                lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt), iters=5, warm=2)
        except Exception as e:  # noqa: BLE001
            ms_t = float("nan")
            print("   torch sdpa failed:", str(e)[:200])
        print(f"attn b{b} sq{sq} sk{sk} h{h} hd{hd}: {ms:.3f} ms {fl/ms/1e9:.0f} TF | torch sdpa {ms_t:.3f} ms {fl/ms_t/1e9:.0f} TF")
        res[f"{b}x{sq}x{sk}x{h}x{hd}"] = dict(ms=ms, tflops=fl / ms / 1e9, torch_ms=ms_t, torch_tflops=fl / ms_t / 1e9)
    # fp8 (e4m3 q/k/v, P quantised to e4m3, f32 accumulate; hd 128): TFLOP/s and error against the bf16 kernel
    for (b, sq, sk, h, hd) in ((1, 8704, 8704, 24, 128), (1, 80640, 80640, 4, 128)):
        q = torch.randn(b, sq, h * hd, device=DEV, dtype=BF)
        k = torch.randn(b, sk, h * hd, device=DEV, dtype=BF)
        v = torch.randn(b, sk, h * hd, device=DEV, dtype=BF)
        q8, k8, v8 = (t.to(torch.float8_e4m3fn) for t in (q, k, v))
        ms = timeit(lambda: ops.attention(q8, k8, v8, h, hd), iters=5, warm=2)
        fl = 4.0 * b * h * sq * sk * hd
        y8 = ops.attention(q8, k8, v8, h, hd).float()
        y16 = ops.attention(q, k, v, h, hd).float()
        cos = torch.nn.functional.cosine_similarity(y8.flatten(), y16.flatten(), dim=0).item()
        print(f"attn fp8 b{b} sq{sq} sk{sk} h{h} hd{hd}: {ms:.3f} ms {fl/ms/1e9:.0f} TF | cosine vs bf16 {cos:.5f} max abs diff {(y8 - y16).abs().max().item():.4f}")
        res[f"fp8_{b}x{sq}x{sk}x{h}x{hd}"] = dict(ms=ms, tflops=fl / ms / 1e9, cos_vs_bf16=cos)
    out["attn_perf"] = res


def attn_patterns():
    print("== attention structured checks ==")
    import math
    for (b, sq, sk, h, hd) in ((1, 128, 128, 1, 128), (1, 256, 128, 1, 128), (1, 256, 256, 1, 128), (1, 256, 384, 2, 64),
                               (1, 300, 200, 2, 128)):
        torch.manual_seed(0)
        q = torch.randn(b, sq, h * hd, device=DEV).to(BF)
        k = torch.randn(b, sk, h * hd, device=DEV).to(BF)
        v = torch.randn(b, sk, h * hd, device=DEV).to(BF)
        y = ops.scaled_dot_product_attention(q, k, v, h, h, hd)
        torch.cuda.synchronize()
        qf, kf, vf = (t.view(b, -1, h, hd).transpose(1, 2).float() for t in (q, k, v))
        ref = torch.softmax(qf @ kf.transpose(-1, -2) / math.sqrt(hd), -1) @ vf
        ref = ref.transpose(1, 2).reshape(b, sq, h * hd)
        err = (y.float() - ref).abs()
        print(f"attn b{b} sq{sq} sk{sk} h{h} hd{hd}: max err {err.max().item():.4f} mean {err.mean().item():.5f} nan={int(torch.isnan(y.float()).sum())}")
        if err.max().item() > 0.02:
            # which rows / columns are off?  and: uniform-attention hypothesis (P layout wrong -> garbage)
            bad_rows = (err.amax(dim=2) > 0.02)[0].nonzero().flatten()
            print("   bad rows:", bad_rows[:10].tolist(), "... count", len(bad_rows))
            bad_cols = (err.amax(dim=1) > 0.02)[0].nonzero().flatten()
            print("   bad cols:", bad_cols[:10].tolist(), "... count", len(bad_cols))
            print("   got :", y[0, 0, :8].float().tolist())
            print("   want:", ref[0, 0, :8].tolist())
            # V constant test isolates QK/softmax from PV layout
            vc = v[:, :1].expand_as(v).contiguous()
            yc = ops.scaled_dot_product_attention(q, k, vc, h, h, hd)
            print("   const-V err:", (yc.float() - vc[:, :1].float()).abs().max().item())
            # one-hot attention (huge scale on identical q/k rows) isolates P/V ordering
            if sq == sk:
                q1 = torch.zeros_like(q)
                k1 = torch.zeros_like(k)
                eye = torch.eye(hd, device=DEV)
                for hh in range(h):
                    idx = torch.arange(sq, device=DEV) % hd
                    q1[0, :, hh * hd:(hh + 1) * hd] = eye[idx] * 30
                    k1[0, :, hh * hd:(hh + 1) * hd] = eye[idx] * 30
                y1 = ops.scaled_dot_product_attention(q1, k1, v, h, h, hd)
                r1 = torch.softmax(q1.view(b, sq, h, hd).transpose(1, 2).float() @ k1.view(b, sk, h, hd).transpose(1, 2).float().transpose(-1, -2) / math.sqrt(hd), -1) @ vf
                r1 = r1.transpose(1, 2).reshape(b, sq, h * hd)
                print("   periodic-onehot err:", (y1.float() - r1).abs().max().item())


def block_kernels_perf():
    """Fused block kernels + GELU-epilogue GEMMs, timed as back-to-back launches rotating over enough
    buffers to exceed the 126 MB L2 (no host gaps, no flush artefacts)."""
    print("== fused block kernels ==")
    res = {}

    def rot_time(make, run, nbuf, iters=40):
        bufs = [make() for _ in range(nbuf)]
        for b in bufs:
            run(b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            run(bufs[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for (M, K) in ((8704, 3072), (8704, 1536), (80640, 5120)):
        nbuf = max(2, int(400e6 // (M * K * 2)) + 1)
        a = (torch.rand(1, K, device=DEV) + 0.5).to(BF).float()   # bf16-valued, as the AdaLN chain produces them
        c = torch.rand(1, K, device=DEV).to(BF).float()
        ms = rot_time(lambda: torch.randn(M, K, device=DEV, dtype=BF),
                      lambda x: ops.layernorm_modulate_quant(x, a, c, M, torch.float8_e4m3fn), nbuf)
        gb = (3 * M * K + 4 * M) / ms / 1e6
        a16, c16 = a.to(BF), c.to(BF)
        msb = rot_time(lambda: torch.randn(M, K, device=DEV, dtype=BF),
                       lambda x: ops.layernorm_modulate_quant(x, a16, c16, M, torch.float8_e4m3fn), nbuf)
        gbb = (3 * M * K + 4 * M) / msb / 1e6
        a32 = a + 1e-4   # not bf16-valued: the Wan fp32 chain
        msw = rot_time(lambda: torch.randn(M, K, device=DEV, dtype=BF),
                       lambda x: ops.layernorm_modulate_quant(x, a32, c, M, torch.float8_e4m3fn, round_steps=False), nbuf)
        gbw = (3 * M * K + 4 * M) / msw / 1e6
        ms2 = rot_time(lambda: torch.randn(M, K, device=DEV, dtype=BF), lambda x: ops.quantize_to_fp8(x), nbuf)
        gb2 = (3 * M * K + 4 * M) / ms2 / 1e6
        print(f"[{M},{K}] ln_mod_quant fp32 mods {ms*1e3:.1f} us {gb:.0f} GB/s | bf16 mods {msb*1e3:.1f} us {gbb:.0f} GB/s | "
              f"fp32 chain (Wan) {msw*1e3:.1f} us {gbw:.0f} GB/s | quant_fp8 {ms2*1e3:.1f} us {gb2:.0f} GB/s")
        res[f"lnq_{M}x{K}"] = dict(lnq_us=ms * 1e3, lnq_gbs=gb, lnq_bf16mod_gbs=gbb, lnq_wan_gbs=gbw, quant_us=ms2 * 1e3, quant_gbs=gb2)
    for (S, H, across) in ((8704, 24, False), (80640, 40, True)):
        d = H * 128
        nbuf = max(2, int(400e6 // (S * 3 * d * 2)) + 1)
        wq = torch.randn(d if across else 128, device=DEV, dtype=BF)
        cs = torch.rand(S, 128, device=DEV, dtype=BF)
        ms = rot_time(lambda: torch.randn(S, 3 * d, device=DEV, dtype=BF),
                      lambda x: ops.qk_norm_rope_(x, wq, wq, cs, H, H, 128, 0, d, 0, 1e-6, across), nbuf, iters=20)
        gb = 8 * S * d / ms / 1e6
        print(f"qk_norm_rope S{S} H{H} across={across}: {ms*1e3:.1f} us {gb:.0f} GB/s")
        res[f"qknr_{S}x{H}"] = dict(us=ms * 1e3, gbs=gb)
    for (M, K, N) in ((8704, 3072, 12288), (8704, 15360, 3072), (8192, 3072, 3072)):
        g = torch.Generator(device=DEV).manual_seed(1)
        a = torch.randn(M, K, device=DEV, generator=g).to(torch.float8_e4m3fn)
        b = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(torch.float8_e4m3fn).t()
        sa = torch.rand(M, 1, device=DEV) * 0.1
        sb = torch.rand(N, 1, device=DEV)
        bias = torch.randn(N, device=DEV).to(BF)
        resid = torch.randn(M, N, device=DEV).to(BF)
        gate = torch.randn(1, N, device=DEV).to(BF).float()   # bf16-valued, as the AdaLN chain produces it
        obuf = torch.empty(M, N, device=DEV, dtype=BF)
        fl = 2.0 * M * N * K
        t_plain = timeit(lambda: ops.fp8_matmul(a, b, sa, sb, BF, bias, out=obuf))
        t_tanh = timeit(lambda: ops.fp8_matmul(a, b, sa, sb, BF, bias, act="gelu_tanh", out=obuf))
        t_erf = timeit(lambda: ops.fp8_matmul(a, b, sa, sb, BF, bias, act="gelu_erf", out=obuf))
        t_res = timeit(lambda: ops.fp8_matmul(a, b, sa, sb, BF, bias, gate=gate, residual=resid, rows_per_batch=M, out=obuf))
        print(f"gemm M{M} K{K} N{N}: plain {t_plain*1e3:.0f} us {fl/t_plain/1e9:.0f} TF | gelu_tanh {t_tanh*1e3:.0f} us {fl/t_tanh/1e9:.0f} TF"
              f" | gelu_erf {t_erf*1e3:.0f} us {fl/t_erf/1e9:.0f} TF | gate+residual {t_res*1e3:.0f} us {fl/t_res/1e9:.0f} TF")
        res[f"gemm_epi_{M}x{K}x{N}"] = dict(plain_us=t_plain * 1e3, tanh_us=t_tanh * 1e3, erf_us=t_erf * 1e3, resid_us=t_res * 1e3)
    out["block_kernels"] = res


def op_overhead():
    x = torch.randn(8, 64, device=DEV, dtype=BF)
    t0 = time.perf_counter()
    for _ in range(2000):
        ops.quantize_to_fp8(x)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 2000
    print(f"host overhead per custom-op call (tiny quant): {dt*1e6:.1f} us")
    out["custom_op_overhead_us"] = dt * 1e6


if __name__ == "__main__":
    which = sys.argv[1:] or ["patterns", "gemm_perf", "elementwise_perf", "overhead"]
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    if "patterns" in which:
        gemm_patterns()
    if "gemm_perf" in which:
        gemm_perf()
    if "elementwise_perf" in which:
        elementwise_perf()
    if "attn_patterns" in which:
        attn_patterns()
    if "attn_perf" in which:
        attn_perf()
    if "block_kernels" in which:
        block_kernels_perf()
    if "overhead" in which:
        op_overhead()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump(out, f, indent=1)
