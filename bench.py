#!/usr/bin/env python
"""bench.py -- DiT denoise-step latency on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps K --warmup W            # our arm, 1 GPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # Ulysses sequence parallel, N in {2,4,8}
    python bench.py --impl reference ...                     # reference torch backend on the host CPUs

Workload (config.workload): one transformer forward (= one denoise step of one CFG branch) of
  wan  : Wan2.2-T2V-A14B, 768x1280x81 frames -> latent [1,16,21,96,160], 80 640 tokens, 40 blocks,
         d=5120, 40x128 heads, ffn 13824, FP8 per-token x per-channel  (BASELINE configs[4]; default,
         because it is the one configuration that spans 1-8 GPUs -- Ulysses, strong scaling)
  flux : FLUX.1-dev, 1024x2048 -> 8192 image + 512 text tokens, 19 double + 38 single blocks, d=3072,
         24x128 heads, FP8  (BASELINE configs[2]; single GPU; also reported under "flux" at N=1)
Random-init weights of the named architecture, synthetic latents / prompt embeddings.

A "step" is one full forward: embedders, all blocks, output projection. `value` times it with the
inputs resident in HBM (CUDA events, max over ranks); `e2e` times the same call from pinned HOST
buffers including the H2D copies of the step inputs and the D2H read of the predicted latent.
Activations per step (>= 0.8 GB per tensor for wan, 0.14 GB for flux) exceed the 126 MB L2, so no
explicit flush is needed between timed iterations (config.l2).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WAN = dict(frames=21, height=96, width=160, in_ch=16, text_len=512, text_dim=4096, heads=40, head_dim=128,
           ffn=13824, layers=40)
FLUX = dict(img_tokens=8192, txt_tokens=512, heads=24, head_dim=128, double=19, single=38)
# examples/profiling/qwenimg_profiling.py:14-19 (1024x2048 -> img_shapes (1, 64, 128)); prompt length 512
QWEN = dict(grid=(1, 64, 128), txt_tokens=512, txt_dim=3584, heads=24, head_dim=128, layers=60)
# fastdm/model/sd35.py:202-221: SD3.5-medium 1024x1024, batch 2 (CFG), 333 text tokens
SD3 = dict(batch=2, latent=(16, 128, 128), txt_tokens=333, txt_dim=4096, pooled=2048, heads=24, head_dim=64, layers=24)
# algorithmic TFLOP per step (2*M*N*K of every quantised linear + 4*B*H*Sq*Sk*hd of every attention)
STEP_TFLOP = dict(wan=7291.0, flux=165.4, qwen=118.3 + 55.9, sd3=2 * (7.35 + 4.58))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.thread = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle restatement of the reference torch backend on host cores
# ---------------------------------------------------------------------------------------------------
class CpuArm:
    """ONE block of the workload on a token sample with the CPU oracle (oracle/blocks_ref.py, a
    restatement of fastdm/model/*.py + fastdm/kernel/torch/*), scaled to the full step: token-local
    work (GEMMs, norms, quant) linearly in tokens, self-attention quadratically (timed on its own,
    larger, token sample so the x(N/n)^2 scaling is not applied to a noise-level number)."""

    def __init__(self, workload, tokens, threads, attn_tokens=4096):
        from oracle import blocks_ref as B

        torch.set_num_threads(threads)
        self.wl, self.tokens, self.attn_tokens = workload, tokens, attn_tokens
        bf = torch.bfloat16
        g = torch.Generator().manual_seed(0)
        quant = torch.float8_e4m3fn
        if workload == "wan":
            self.d, self.H, self.hd = WAN["heads"] * WAN["head_dim"], WAN["heads"], WAN["head_dim"]
            self.full = WAN["frames"] * (WAN["height"] // 2) * (WAN["width"] // 2)
            sd = B.wan_block_state_dict("blocks.0", self.d, WAN["ffn"], seed=1)
            self.blk = B.WanTransformerBlockRef(sd, "blocks.0", self.H, self.hd, quant)
            self.x = torch.randn(1, tokens, self.d, generator=g).to(bf)
            self.enc = torch.randn(1, WAN["text_len"], self.d, generator=g).to(bf)
            self.temb = torch.randn(1, 6, self.d, generator=g).to(bf)
            self.cos = torch.rand(1, tokens, 1, self.hd, generator=g)
            self.sin = torch.rand(1, tokens, 1, self.hd, generator=g)
        else:
            self.d, self.H, self.hd = FLUX["heads"] * FLUX["head_dim"], FLUX["heads"], FLUX["head_dim"]
            self.full = FLUX["img_tokens"] + FLUX["txt_tokens"]
            txt = max(16, tokens // 17)
            sd = B.flux_double_state_dict("transformer_blocks.0", self.d, self.hd, seed=1)
            sd1 = B.flux_single_state_dict("single_transformer_blocks.0", self.d, self.hd, seed=2)
            self.dbl = B.FluxTransformerBlockRef(sd, "transformer_blocks.0", self.H, self.hd, quant)
            self.sgl = B.FluxSingleTransformerBlockRef(sd1, "single_transformer_blocks.0", self.H, self.hd, quant)
            self.xi = torch.randn(1, tokens - txt, self.d, generator=g).to(bf)
            self.xt = torch.randn(1, txt, self.d, generator=g).to(bf)
            self.temb = torch.randn(1, self.d, generator=g).to(bf)
            self.rope = torch.rand(tokens, self.hd, generator=g).to(bf)
        self.q_small = torch.randn(1, tokens, self.d, generator=g).to(bf)
        self.q_big = torch.randn(1, attn_tokens, self.d, generator=g).to(bf)

    def _attn(self, q):
        from oracle import ops_ref as R

        t0 = time.perf_counter()
        R.scaled_dot_product_attention(q, q, q, self.H, self.H, self.hd, scale=self.hd ** -0.5)
        return time.perf_counter() - t0

    def step(self):
        """-> (estimated full-step milliseconds, description of the sample)"""
        t_small, t_big = self._attn(self.q_small), self._attn(self.q_big)
        r, ra = self.full / self.tokens, self.full / self.attn_tokens
        if self.wl == "wan":
            t0 = time.perf_counter()
            self.blk.forward(self.x, self.enc, self.temb, (self.cos, self.sin))
            t_block = time.perf_counter() - t0
            est = WAN["layers"] * (max(t_block - t_small, 0.0) * r + t_big * ra * ra)
            sample = (f"1 of 40 Wan2.2 blocks (d=5120, ffn 13824, fp8 W8A8) on {self.tokens} of {self.full} tokens + 512 text "
                      f"tokens: {t_block:.2f} s; self-attention timed on {self.attn_tokens} tokens: {t_big:.2f} s; step = 40 x "
                      f"(token-local x{r:.1f} + attention x{ra * ra:.0f})")
        else:
            t0 = time.perf_counter()
            e, h = self.dbl.forward(self.xi, self.xt, self.temb, self.rope)
            t_d = time.perf_counter() - t0
            t0 = time.perf_counter()
            self.sgl.forward(torch.cat([e, h], 1), self.temb, self.rope)
            t_s = time.perf_counter() - t0
            est = (FLUX["double"] * max(t_d - t_small, 0.0) + FLUX["single"] * max(t_s - t_small, 0.0)) * r \
                + (FLUX["double"] + FLUX["single"]) * t_big * ra * ra
            sample = (f"1 double + 1 single FLUX block (d=3072, fp8 W8A8) on {self.tokens} of {self.full} tokens: {t_d:.2f} s + "
                      f"{t_s:.2f} s; attention timed on {self.attn_tokens} tokens: {t_big:.2f} s; step = 19 double + 38 single "
                      f"(token-local x{r:.1f} + attention x{ra * ra:.1f})")
        return est * 1e3, sample


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (its torch backend,
    restated in oracle/ -- the reference itself is not on the GPU box), all host threads."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = args.workload if args.workload != "auto" else "wan"
    if wl not in ("wan", "flux"):
        emit(dict(impl="reference", unavailable=f"the CPU arm restates the Wan and FLUX blocks only (not {wl})"))
        return
    arm = CpuArm(wl, args.cpu_tokens, threads)
    vals, sample = [], ""
    for i in range(args.warmup + args.steps):
        ms, sample = arm.step()
        if i >= args.warmup:
            vals.append(ms)
    v = sum(vals) / len(vals)
    line = dict(metric="dit_denoise_step_ms", value=v, unit="ms", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=v, higher_is_better=False, scaling="strong" if wl == "wan" else "weak", vs_baseline=None,
                dtype="fp8_e4m3 codes, f32 arithmetic on the CPU", data="synthetic", impl="reference",
                config=workload_config(wl, 1),
                cpu_baseline=dict(value=v, unit="ms", cores=threads, kind="port", sample=sample),
                e2e=dict(value=v, unit="ms", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


def workload_config(wl, n):
    if wl == "wan":
        return dict(workload="Wan2.2-T2V-A14B one expert transformer forward, 768x1280x81f (latent 1x16x21x96x160, 80640 tokens, "
                             "40 blocks, d=5120, 40x128 heads, ffn 13824), FP8 per-token x per-channel W8A8, bf16 attention, "
                             "random-init weights",
                    parallelism=f"ulysses-sp{n}" if n > 1 else "single-gpu", l2="per-step activations (>=0.8 GB each) exceed the 126 MB L2; no flush")
    if wl == "qwen":
        return dict(workload="Qwen-Image 20B MMDiT transformer forward, 1024x2048 (8192 image + 512 text tokens, 60 blocks, d=3072, "
                             "24x128 heads), INT8 per-token (asymmetric) x per-channel W8A8, bf16 attention, random-init weights",
                    parallelism=f"ulysses-sp{n}" if n > 1 else "single-gpu",
                    l2="per-step activations (0.16 GB per [8704, 9216] qkv buffer) exceed the 126 MB L2; no flush")
    if wl == "sd3":
        return dict(workload="SD3.5-medium MMDiT transformer forward, 1024x1024, batch 2 (CFG) (4096 image + 333 text tokens, 24 "
                             "blocks, 13 with dual attention, d=1536, 24x64 heads), FP8 W8A8, bf16 attention, random-init weights",
                    parallelism="single-gpu" if n == 1 else f"replicas x{n}",
                    l2="one step touches 2.2 GB of weights + activations, more than the 126 MB L2; no flush",
                    launch="whole step replayed as one CUDA graph (fastdm_b200.graph.GraphedStep); --no-graph launches eagerly")
    return dict(workload="FLUX.1-dev full transformer forward, 1024x2048 (8192 image + 512 text tokens, 19 double + 38 single "
                         "blocks, d=3072, 24x128 heads), FP8 per-token x per-channel W8A8, bf16 attention, random-init weights",
                parallelism="single-gpu" if n == 1 else f"replicas x{n}", l2="per-step activations (0.14 GB each) exceed the 126 MB L2; no flush",
                launch="whole step replayed as one CUDA graph (fastdm_b200.graph.GraphedStep); --no-graph launches eagerly")


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def build_wan(device, layers):
    from fastdm_b200.models import WanTransformer3DModelCore

    model = WanTransformer3DModelCore(num_attention_heads=WAN["heads"], attention_head_dim=WAN["head_dim"],
                                      in_channels=WAN["in_ch"], text_dim=WAN["text_dim"], ffn_dim=WAN["ffn"],
                                      num_layers=layers, device=device, seed=0)
    g = torch.Generator().manual_seed(1)
    host = dict(latent=torch.rand(1, WAN["in_ch"], WAN["frames"], WAN["height"], WAN["width"], generator=g).to(torch.bfloat16).pin_memory(),
                timestep=torch.tensor([999], dtype=torch.int64).pin_memory(),
                prompt=torch.rand(1, WAN["text_len"], WAN["text_dim"], generator=g).to(torch.bfloat16).pin_memory())
    return model, host


def build_flux(device, double, single):
    from fastdm_b200.models import FluxTransformer2DModelCore

    model = FluxTransformer2DModelCore(num_layers=double, num_single_layers=single, device=device, seed=0)
    g = torch.Generator().manual_seed(1)
    bf = torch.bfloat16
    host = dict(latent=torch.rand(1, FLUX["img_tokens"], 64, generator=g).to(bf).pin_memory(),
                prompt=torch.rand(1, FLUX["txt_tokens"], 4096, generator=g).to(bf).pin_memory(),
                pooled=torch.rand(1, 768, generator=g).to(bf).pin_memory(),
                timestep=torch.tensor([1.0]).to(bf).pin_memory(), guidance=torch.tensor([3.5]).to(bf).pin_memory(),
                img_ids=torch.zeros(FLUX["img_tokens"], 3).pin_memory(), txt_ids=torch.zeros(FLUX["txt_tokens"], 3).pin_memory())
    return model, host


def build_qwen(device, layers):
    from fastdm_b200.models import QwenImageTransformer2DModelCore

    model = QwenImageTransformer2DModelCore(num_layers=layers, quant_dtype=torch.int8, device=device, seed=0)
    g = torch.Generator().manual_seed(1)
    f, h, w = QWEN["grid"]
    host = dict(latent=torch.rand(1, f * h * w, 64, generator=g).to(torch.bfloat16).pin_memory(),
                prompt=torch.rand(1, QWEN["txt_tokens"], QWEN["txt_dim"], generator=g).to(torch.bfloat16).pin_memory(),
                timestep=torch.tensor([0.5]).pin_memory())
    return model, host


def build_sd3(device, layers):
    from fastdm_b200.models import SD3TransformerModelCore

    model = SD3TransformerModelCore(num_layers=layers, device=device, seed=0)
    g = torch.Generator().manual_seed(1)
    bf, b = torch.bfloat16, SD3["batch"]
    host = dict(latent=torch.randn(b, *SD3["latent"], generator=g).to(bf).pin_memory(),
                prompt=torch.randn(b, SD3["txt_tokens"], SD3["txt_dim"], generator=g).to(bf).pin_memory(),
                pooled=torch.randn(b, SD3["pooled"], generator=g).to(bf).pin_memory(),
                timestep=torch.tensor([500.0] * b).to(bf).pin_memory())
    return model, host


def to_device(host, device):
    return {k: v.to(device, non_blocking=True) for k, v in host.items()}


def step_fn(wl, model, ulysses):
    if wl == "wan":
        return lambda d: model.forward(d["latent"], d["timestep"], d["prompt"], ulysses=ulysses)[0]
    if wl == "qwen":
        return lambda d: model.forward(d["latent"], d["prompt"], d["timestep"], QWEN["grid"], ulysses=ulysses)[0]
    if wl == "sd3":
        return lambda d: model.forward(d["latent"], d["prompt"], d["pooled"], d["timestep"])[0]
    return lambda d: model.forward(d["latent"], d["prompt"], d["pooled"], d["timestep"], d["img_ids"], d["txt_ids"], d["guidance"])[0]


def timed_steps(fn, dev_inputs, steps, warmup, world, device):
    import torch.distributed as dist

    for _ in range(warmup):
        fn(dev_inputs)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = os.environ.get("FDM_BENCH_PROFILE") == "1"   # ncu --profile-from-start off: capture only the timed steps
    if prof:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(steps):
        fn(dev_inputs)
    e1.record()
    torch.cuda.synchronize()
    if prof:
        torch.cuda.profiler.stop()
        os.environ["FDM_BENCH_PROFILE"] = "0"
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def timed_e2e(fn, host, steps, world, device):
    """Same step from pinned host buffers: H2D of the step inputs + D2H of the predicted latent inside the timed region."""
    import torch.distributed as dist

    out_host = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        d = to_device(host, device)
        y = fn(d)
        out_host = y.to("cpu", non_blocking=False)
    torch.cuda.synchronize()
    ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / steps], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()
    return float(ms.item()), h2d, d2h


def attention_roofline(wl, world, device, pk):
    """The dominant kernel (self-attention: 73% of Wan flops, 32% of FLUX) timed live on its own:
    CUDA events on the launching stream, algorithmic flops 4*B*H*Sq*Sk*hd per launch."""
    from fastdm_b200 import ops

    Bt = 1
    if wl == "wan":
        S = WAN["frames"] * (WAN["height"] // 2) * (WAN["width"] // 2)
        H, hd = WAN["heads"] // world, WAN["head_dim"]
    elif wl == "qwen":
        S, H, hd = math.prod(QWEN["grid"]) + QWEN["txt_tokens"], QWEN["heads"] // world, QWEN["head_dim"]
    elif wl == "sd3":
        Bt, S, H, hd = SD3["batch"], 4096 + SD3["txt_tokens"], SD3["heads"], SD3["head_dim"]
    else:
        S, H, hd = FLUX["img_tokens"] + FLUX["txt_tokens"], FLUX["heads"], FLUX["head_dim"]
    qkv = torch.randn(Bt, S, 3 * H * hd, device=device, dtype=torch.bfloat16)
    d = H * hd
    run = lambda: ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], H, hd)  # noqa: E731
    run()
    torch.cuda.synchronize()
    n = 3 if wl == "wan" else 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 4.0 * Bt * H * S * S * hd
    ach = flops / ms / 1e9
    # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture: profiles/attn_traffic.json
    # (written from the .ncu-rep by tools/ncu_traffic.py; only the N=1 launch shapes were captured)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "attn_traffic.json")) as f:
            traffic = json.load(f).get(f"{wl}_n{world}", {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        pass
    return dict(bound="tensor", kernel=f"attn_fwd_kernel<{hd},bf16>", shape=[Bt, S, S, H, hd], achieved=ach, peak=pk["bf16"], unit="TFLOP/s",
                frac=ach / pk["bf16"], traffic=traffic, traffic_unit="bytes/launch (dram read+write, ncu --set full)",
                algorithmic_bytes=4.0 * Bt * S * H * hd * 2, ms_per_launch=ms, flops_per_launch=flops,
                peak_source=pk["src"] + ", bf16 burst (kernel timed alone)")


def _cuda_ms(fn, iters, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_roofline(wl, device, pk):
    """"FP8 GEMM TFLOPS/peak" (BASELINE.json metric): the workload's largest linear (ff1: Wan 80640x5120x13824,
    FLUX 8704x3072x12288) through ops.fp8_matmul (per-token x per-channel scales, bias, bf16 out), timed alone with
    CUDA events; outputs (>= 0.2 GB) exceed L2. Also the same call through torch._scaled_mm (cuBLASLt rowwise)."""
    from fastdm_b200 import ops

    M, K, N = dict(wan=(80640, 5120, 13824), flux=(8704, 3072, 12288), qwen=(8192, 3072, 12288), sd3=(8192, 1536, 6144))[wl]
    g = torch.Generator(device=device).manual_seed(3)
    sa = torch.rand(M, 1, device=device) * 0.01
    sb = torch.rand(N, 1, device=device)
    bias = torch.randn(N, device=device).to(torch.bfloat16)
    out = torch.empty(M, N, device=device, dtype=torch.bfloat16)
    it = 5 if wl == "wan" else 30
    fl = 2.0 * M * N * K
    if wl == "qwen":   # INT8 W8A8 with the asymmetric-activation zero-point correction (QLinear int8 path)
        a = torch.randint(-128, 128, (M, K), device=device, generator=g).to(torch.int8)
        b = torch.randint(-128, 128, (N, K), device=device, generator=g).to(torch.int8).t()
        adj = b.to(torch.int32).sum(dim=0, keepdim=True, dtype=torch.int32)
        azp = torch.randint(-128, 127, (M, 1), device=device).to(torch.int32)
        ms = _cuda_ms(lambda: ops.int8_matmul(a, b, sa, sb, torch.bfloat16, adj, azp, bias, out=out), it, 2)
        tf = fl / ms / 1e9
        return dict(kernel="gemm_w8a8_kernel<int8>", shape=[M, K, N], ms=ms, tflops=tf, peak=2 * pk["bf16"], frac=tf / (2 * pk["bf16"]),
                    peak_source="2 x measured bf16 burst (" + pk["src"] + "); nominal dense int8 4500", frac_of_nominal=tf / 4500.0,
                    note="the reference torch backend computes int8 GEMMs as fp32 matmuls (kernel/torch/matrixmul.py:67); no library int8 arm timed")
    a = torch.randn(M, K, device=device, generator=g).to(torch.float8_e4m3fn)
    b = torch.randn(N, K, device=device, generator=g).to(torch.float8_e4m3fn).t()
    ms = _cuda_ms(lambda: ops.fp8_matmul(a, b, sa, sb, torch.bfloat16, bias, out=out), it, 2)
    try:
        ms_t = _cuda_ms(lambda: torch._scaled_mm(a, b, sa, sb.t(), bias, out_dtype=torch.bfloat16), it, 2)
    except Exception as e:  # noqa: BLE001
        print("bench: torch._scaled_mm failed:", str(e)[:200], file=sys.stderr)
        ms_t = None
    tf = fl / ms / 1e9
    return dict(kernel="gemm_w8a8_kernel<fp8>", shape=[M, K, N], ms=ms, tflops=tf, peak=2 * pk["bf16"], frac=tf / (2 * pk["bf16"]),
                peak_source="2 x measured bf16 burst (" + pk["src"] + "); nominal dense fp8 4500", frac_of_nominal=tf / 4500.0,
                torch_scaled_mm_ms=ms_t, torch_scaled_mm_tflops=(fl / ms_t / 1e9) if ms_t else None)


def gpu_torch_baseline(wl, device, ours_step_ms, roof, gemm):
    """The reference's B200-runnable backend (KERNEL_BACKEND=torch: fastdm/kernel/torch/matrixmul.py:33
    `torch._scaled_mm` rowwise, attention.py:38-40 `F.scaled_dot_product_attention`, unfused elementwise ops), as
    restated in oracle/blocks_ref.py, moved to this GPU: ONE full-size block of every block type of the workload,
    CUDA-event timed, times the block count (embedders / output projection, < 1 % of a step, are not in the
    estimate). A measured baseline only -- nothing of it is on the product path."""
    from oracle import blocks_ref as B

    bf, quant = torch.bfloat16, torch.float8_e4m3fn
    g = torch.Generator(device=device).manual_seed(0)
    rnd = lambda *sh: torch.randn(*sh, device=device, generator=g).to(bf)  # noqa: E731
    cuda_sd = lambda sd: {k: v.to(device) for k, v in sd.items()}  # noqa: E731
    if wl == "wan":
        d, H, hd = WAN["heads"] * WAN["head_dim"], WAN["heads"], WAN["head_dim"]
        S = WAN["frames"] * (WAN["height"] // 2) * (WAN["width"] // 2)
        blk = B.WanTransformerBlockRef(cuda_sd(B.wan_block_state_dict("blocks.0", d, WAN["ffn"], seed=1)), "blocks.0", H, hd, quant)
        x, enc, temb = rnd(1, S, d), rnd(1, WAN["text_len"], d), rnd(1, 6, d)
        cos = torch.rand(1, S, 1, hd, device=device, generator=g)
        sin = torch.rand(1, S, 1, hd, device=device, generator=g)
        with torch.no_grad():
            t_blk = _cuda_ms(lambda: blk.forward(x, enc, temb, (cos, sin)), 2, 1)
        step = WAN["layers"] * t_blk
        blocks = dict(wan_block_ms=t_blk, count=WAN["layers"])
        del blk, x, cos, sin
    else:
        d, H, hd = FLUX["heads"] * FLUX["head_dim"], FLUX["heads"], FLUX["head_dim"]
        S = FLUX["img_tokens"] + FLUX["txt_tokens"]
        dbl = B.FluxTransformerBlockRef(cuda_sd(B.flux_double_state_dict("transformer_blocks.0", d, hd, seed=1)),
                                        "transformer_blocks.0", H, hd, quant)
        sgl = B.FluxSingleTransformerBlockRef(cuda_sd(B.flux_single_state_dict("single_transformer_blocks.0", d, hd, seed=2)),
                                              "single_transformer_blocks.0", H, hd, quant)
        xi, xt, temb, rope = rnd(1, FLUX["img_tokens"], d), rnd(1, FLUX["txt_tokens"], d), rnd(1, d), \
            torch.rand(S, hd, device=device, generator=g).to(bf)
        xs = rnd(1, S, d)
        with torch.no_grad():
            t_d = _cuda_ms(lambda: dbl.forward(xi, xt, temb, rope), 10, 2)
            t_s = _cuda_ms(lambda: sgl.forward(xs, temb, rope), 10, 2)
        step = FLUX["double"] * t_d + FLUX["single"] * t_s
        blocks = dict(flux_double_block_ms=t_d, flux_single_block_ms=t_s, count=[FLUX["double"], FLUX["single"]])
        del dbl, sgl
    torch.cuda.empty_cache()
    # the reference backend's attention call on the dominant shape (attention.py:33-40: [B,H,S,hd] views of NHD tensors)
    Hh = H
    q, k, v = (rnd(1, S, Hh * hd).view(1, S, Hh, hd).transpose(1, 2) for _ in range(3))
    t_att = _cuda_ms(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v), 2 if wl == "wan" else 20, 1)
    att_tf = 4.0 * Hh * S * S * hd / t_att / 1e9
    del q, k, v
    torch.cuda.empty_cache()
    return dict(backend="reference torch backend on this GPU (torch._scaled_mm rowwise fp8 + F.scaled_dot_product_attention + "
                        "unfused elementwise; oracle/blocks_ref.py on cuda), one full-size block x block count",
                step_ms_est=step, attention_ms=t_att, attention_tflops=att_tf, gemm_ms=gemm.get("torch_scaled_mm_ms"),
                gemm_tflops=gemm.get("torch_scaled_mm_tflops"), **blocks,
                vs_torch_backend=dict(step=step / ours_step_ms, attention=t_att / roof["ms_per_launch"],
                                      gemm=(gemm["torch_scaled_mm_ms"] / gemm["ms"]) if gemm.get("torch_scaled_mm_ms") else None,
                                      note="torch-backend time / our time on the same box: > 1 means we are faster"))


def ulysses_parity(wl, model, dev_inputs, ulysses, device):
    """gather(P-rank output) vs the 1-rank output of the SAME model on the first 2 blocks ("identical up to
    attention-kernel tolerance" -- the per-head math is unchanged, only the order of the token shards' GEMM tiles is)."""
    import torch.distributed as dist

    attr = "blocks" if wl == "wan" else "transformer_blocks"
    all_blocks = getattr(model, attr)
    setattr(model, attr, all_blocks[:2])
    try:
        with torch.no_grad():
            y1 = step_fn(wl, model, None)(dev_inputs).float()
            yp = step_fn(wl, model, ulysses)(dev_inputs).float()
    finally:
        setattr(model, attr, all_blocks)
    cos = torch.nn.functional.cosine_similarity(y1.flatten().double(), yp.flatten().double(), dim=0)
    stats = torch.stack([cos.float(), -(y1 - yp).abs().max(), -y1.abs().max()])
    dist.all_reduce(stats, op=dist.ReduceOp.MIN)   # worst rank
    del y1, yp
    torch.cuda.empty_cache()
    return dict(cos=float(stats[0]), max_abs=float(-stats[1]), out_scale=float(-stats[2]), blocks=2, ranks=ulysses.P,
                what="full forward (embed, 2 blocks, output projection, all-gather) sharded vs unsharded on every rank; worst rank")


def secondary_workload(name, args, world, device, pk, rank):
    """One more BASELINE configuration with the headline run's timing rules (W >= 3, CUDA events, max over ranks)."""
    from fastdm_b200.ulysses import UlyssesAttention

    build = dict(flux=lambda: build_flux(device, FLUX["double"], FLUX["single"]), sd3=lambda: build_sd3(device, SD3["layers"]),
                 qwen=lambda: build_qwen(device, QWEN["layers"]))[name]
    model, host = build()
    uly = UlyssesAttention(QWEN["heads"], QWEN["head_dim"]) if (name == "qwen" and world > 1) else None
    out = dict(config=workload_config(name, world))
    fn = step_fn(name, model, uly)
    dev_in = to_device(host, device)
    if uly is not None:
        out["ulysses_parity"] = ulysses_parity(name, model, dev_in, uly, device)
    eager = fn
    if (name in ("flux", "sd3") or (name == "qwen" and world == 1)) and not args.no_graph:
        from fastdm_b200.graph import GraphedStep
        fn = GraphedStep(fn, dev_in)      # (the Ulysses steps stay eager: NCCL + symmetric-memory barriers)
        out["launch"] = "whole step replayed as one CUDA graph (fastdm_b200.graph.GraphedStep)"
    out["ms_per_step"] = ms = timed_steps(fn, dev_in, 10, 3, world, device)
    e2e, h2d, d2h = timed_e2e(fn, host, 5, world, device)
    out.update(e2e_ms=e2e, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, step_tflops_per_s=STEP_TFLOP[name] * 1e3 / ms)
    if uly is not None:
        uly.stub_comm = True
        stub = timed_steps(fn, dev_in, 3, 1, world, device)
        uly.stub_comm = False
        out.update(a2a_exposed_ms=ms - stub, ms_per_step_comm_stubbed=stub)
    if world == 1:
        from fastdm_b200 import _lib
        c0 = _lib.launch_count
        eager(dev_in)
        torch.cuda.synchronize()
        out["gpu_launches"] = _lib.launch_count - c0
        roof = attention_roofline(name, 1, device, pk)
        out["attention_tflops"] = roof["achieved"]
        out["attention_frac_of_bf16_burst"] = roof["frac"]
        out["gemm"] = gemm_roofline(name, device, pk)
        if name == "flux":
            if not args.no_torch_baseline:
                out["gpu_torch_baseline"] = gpu_torch_baseline("flux", device, ms, roof, out["gemm"])
            # the same model at 1024x1024 (4096 image + 512 text tokens), the "FLUX 1024^2" of BASELINE.json's metric line
            sq = {k: v for k, v in host.items()}
            sq["latent"] = host["latent"][:, :4096].contiguous().pin_memory()
            sq["img_ids"] = host["img_ids"][:4096].contiguous().pin_memory()
            sfn = step_fn("flux", model, None)
            if not args.no_graph:
                from fastdm_b200.graph import GraphedStep
                sfn = GraphedStep(sfn, to_device(sq, device))
            out["ms_per_step_1024x1024"] = timed_steps(sfn, to_device(sq, device), 10, 3, 1, device)
    del model
    torch.cuda.empty_cache()
    return out


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries print there too (NCCL's version banner when NCCL_DEBUG is
    set, torchrun warnings), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a
    duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "wan", "flux", "qwen", "sd3"],
                    help="auto = wan as the headline plus the other BASELINE configurations as secondary entries")
    ap.add_argument("--layers", type=int, default=0, help="debug: fewer blocks (the JSON line then says so and is not a valid result)")
    ap.add_argument("--cpu-tokens", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flux", action="store_true", help="skip the secondary FLUX numbers at N=1")
    ap.add_argument("--no-sd3", action="store_true", help="skip the secondary SD3.5 numbers at N=1")
    ap.add_argument("--no-qwen", action="store_true", help="skip the secondary Qwen-Image numbers (timed at every N)")
    ap.add_argument("--no-sparse", action="store_true", help="skip the radial-sparse Wan variant at N=1")
    ap.add_argument("--no-fp8-attention", action="store_true", help="skip the fp8-attention Wan variant at N=1")
    ap.add_argument("--no-graph", action="store_true", help="launch the FLUX step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the reference-torch-backend-on-this-GPU leg at N=1")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3 and not args.layers:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    import torch.distributed as dist

    from fastdm_b200 import _lib
    from fastdm_b200.ulysses import UlyssesAttention

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's own banner / debug lines (NCCL_DEBUG set on the box) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    wl = "wan" if args.workload == "auto" else args.workload
    pk = peaks()
    _lib.load()
    sequence_parallel = wl in ("wan", "qwen")   # FLUX / SD3.5 stay single-GPU: N ranks = N independent replicas

    ulysses = None
    if wl == "wan":
        model, host = build_wan(device, args.layers or WAN["layers"])
        if world > 1:
            ulysses = UlyssesAttention(WAN["heads"], WAN["head_dim"])
    elif wl == "qwen":
        model, host = build_qwen(device, args.layers or QWEN["layers"])
        if world > 1:
            ulysses = UlyssesAttention(QWEN["heads"], QWEN["head_dim"])
    elif wl == "sd3":
        model, host = build_sd3(device, args.layers or SD3["layers"])
    else:
        model, host = build_flux(device, args.layers or FLUX["double"], args.layers or FLUX["single"])
    fn0 = step_fn(wl, model, ulysses)
    if wl == "wan" and ulysses is not None and args.no_overlap:
        fn = lambda d: model.forward(d["latent"], d["timestep"], d["prompt"], ulysses=ulysses, overlap=False)[0]  # noqa: E731
    else:
        fn = fn0
    dev_inputs = to_device(host, device)
    torch.cuda.synchronize()
    extra = {}
    if ulysses is not None:
        # sharded == unsharded, checked on the device before anything is timed (2 blocks of the same model)
        extra["ulysses_parity"] = ulysses_parity(wl, model, dev_inputs, ulysses, device)
    if (wl in ("flux", "sd3") or (wl == "qwen" and world == 1)) and not args.no_graph:
        from fastdm_b200.graph import GraphedStep
        fn = GraphedStep(fn, dev_inputs)   # 700-1300 short launches per step: replayed as one CUDA graph

    with ClockSampler(local_rank) as clocks:
        c0 = _lib.launch_count
        ms = timed_steps(fn, dev_inputs, args.steps, args.warmup, world, device)
        launches = (_lib.launch_count - c0) // (args.steps + args.warmup)
    if launches == 0:   # graph replay: the library is not entered; count the launches of one eager step
        c0 = _lib.launch_count
        fn0(dev_inputs)
        torch.cuda.synchronize()
        launches = _lib.launch_count - c0
    e2e_ms, h2d, d2h = timed_e2e(fn, host, max(1, min(args.steps, 3)), world, device)

    if ulysses is not None:
        # exposed all-to-all time = step time - step time with the exchange replaced by a local copy
        ulysses.stub_comm = True
        ms_stub = timed_steps(fn, dev_inputs, max(1, min(args.steps, 2)), 1, world, device)
        ulysses.stub_comm = False
        extra["a2a_exposed_ms"] = ms - ms_stub
        extra["ms_per_step_comm_stubbed"] = ms_stub
        extra["ulysses_exchange"] = ("Q | K | V: NCCL all_to_all_single per projection group on a side stream under the next "
                                     "group's GEMM; O: " + ("attention epilogue stores into the owners' symmetric-memory buffers "
                                     "over NVLink (fdm_attn_fwd_scatter) + device-side barrier" if ulysses.scatter else
                                     "NCCL all_to_all_single + unpack kernel"))
    roof = attention_roofline(wl, world if sequence_parallel else 1, device, pk)
    if rank == 0 and world == 1:
        extra["gemm"] = gemm_roofline(wl, device, pk)
        if not args.no_torch_baseline and not args.layers and wl in ("wan", "flux"):
            extra["gpu_torch_baseline"] = gpu_torch_baseline(wl, device, ms, roof, extra["gemm"])

    secondary = wl == "wan" and args.workload == "auto" and not args.layers
    if rank == 0 and secondary and world == 1 and not args.no_sparse:
        # BASELINE configs[4] "dense vs Sparge sparse attention": the same step with the reference's radial block
        # mask (examples/sparse/radial_attn_wan.json: block 64, decay 0.3, first layer dense) on self-attention
        from fastdm_b200.sparse import radial_block_mask, sparge_mask_convert
        tpf = (WAN["height"] // 2) * (WAN["width"] // 2)
        m64 = radial_block_mask(WAN["frames"], tpf, 64, 0.3, "wan", device=device)
        conv = sparge_mask_convert(m64, 64)                                   # [S/128, S/64]
        smask = conv.to(torch.int8)[None, None].expand(1, WAN["heads"], -1, -1).contiguous()
        sfn = lambda d: model.forward(d["latent"], d["timestep"], d["prompt"], sparse_mask=smask, dense_layers=1)[0]  # noqa: E731
        sms = timed_steps(sfn, dev_inputs, max(1, min(args.steps, 3)), 3, 1, device)
        extra["wan_sparse"] = dict(ms_per_step=sms, mask="radial (fastdm/sparse/xsparse.py), block 64, decay_factor 0.3, "
                                   "dense_layers 1, all steps sparse", block_sparsity=1.0 - conv.float().mean().item(),
                                   speedup_vs_dense=ms / sms)
        del smask, m64, conv
    if rank == 0 and secondary and world == 1 and not args.no_fp8_attention:
        # north_star family 2 "bf16 and FP8 QK^T/PV": the same step with e4m3 q/k/v and e4m3 P in self-attention
        # (fdm_attn_fwd with FDM_E4M3 operands; semantics csrc/attention/interface.cu:262-270)
        ref_out = fn(dev_inputs).float()
        for blk in model.blocks:
            blk.fp8_attention = True
        fp8_out = fn(dev_inputs).float()
        fms = timed_steps(fn, dev_inputs, max(1, min(args.steps, 3)), 3, 1, device)
        for blk in model.blocks:
            blk.fp8_attention = False
        cosv = torch.nn.functional.cosine_similarity(ref_out.flatten().double(), fp8_out.flatten().double(), dim=0)
        extra["wan_fp8_attention"] = dict(ms_per_step=fms, speedup_vs_bf16_attention=ms / fms,
                                          cos_vs_bf16_attention_step_output=float(cosv),
                                          note="q/k/v cast to e4m3 with unit scales (one extra pass over the fused qkv buffer "
                                               "per block, inside the timed step), P quantised to e4m3 unscaled, f32 accumulate")
        del ref_out, fp8_out
    if secondary:
        # the other BASELINE configurations, reported next to the headline one (same timing rules). FLUX and SD3.5 are
        # single-GPU models (N = 1 only); Qwen-Image is sequence-parallel and is timed at every N.
        del model, fn, fn0
        torch.cuda.empty_cache()
        for name in ("flux", "sd3", "qwen"):
            if getattr(args, f"no_{name}") or (world > 1 and name != "qwen"):
                continue
            res = secondary_workload(name, args, world, device, pk, rank)
            if rank == 0:
                extra[name] = res

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline and wl in ("wan", "flux"):
        threads = os.cpu_count() or 1
        v, sample = CpuArm(wl, args.cpu_tokens, threads).step()
        cpu = dict(value=v, unit="ms", cores=threads, kind="port", sample=sample)
    cfg = workload_config(wl, world)
    if args.layers:
        cfg["INVALID_debug_layers"] = args.layers
    step_tflop = STEP_TFLOP[wl]
    dtype = "int8 GEMM (s32 accumulate), bf16 attention (f32 accumulate)" if wl == "qwen" else \
        "fp8_e4m3 GEMM (f32 accumulate), bf16 attention (f32 accumulate)"
    line = dict(metric="dit_denoise_step_ms", value=ms, unit="ms", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=False, scaling="strong" if sequence_parallel else "weak", vs_baseline=None,
                dtype=dtype, data="synthetic", config=cfg,
                clocks=clocks.summary(),
                e2e=dict(value=e2e_ms, unit="ms", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                gpu_launches=launches, roofline=roof, cpu_baseline=cpu,
                step_tflops_per_s=step_tflop * 1e3 / ms, **extra)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
