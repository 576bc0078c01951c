"""Qwen-Image 1024x2048 denoise step (BASELINE configs[3]: 60 MMDiT blocks, d = 3072, 24 x 128 heads, INT8 W8A8,
8192 image + 512 text tokens) on 1 GPU or Ulysses sequence-parallel on N GPUs.

    python tools/qwen_step.py                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/qwen_step.py
Prints one JSON line on rank 0 (CUDA-event timing, max over ranks).
"""
import json, os, sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import _lib  # noqa: E402
from fastdm_b200.models import QwenImageTransformer2DModelCore  # noqa: E402
from fastdm_b200.ulysses import UlyssesAttention  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
quant = torch.float8_e4m3fn if os.environ.get("QWEN_QUANT", "int8") == "fp8" else torch.int8
layers = int(os.environ.get("QWEN_LAYERS", "60"))
steps, warmup = int(os.environ.get("STEPS", "5")), 3
model = QwenImageTransformer2DModelCore(num_layers=layers, quant_dtype=quant, device=dev, seed=0)
g = torch.Generator().manual_seed(1)
f, h, w, T = 1, 64, 128, 512
lat = torch.randn(1, f * h * w, 64, generator=g).to(torch.bfloat16).to(dev)
txt = torch.randn(1, T, 3584, generator=g).to(torch.bfloat16).to(dev)
ts = torch.tensor([0.5], device=dev)
uly = UlyssesAttention(24, 128) if world > 1 else None
fn = lambda: model.forward(lat, txt, ts, (f, h, w), ulysses=uly)[0]  # noqa: E731
for _ in range(warmup):
    fn()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
c0 = _lib.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    y = fn()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
exposed = None
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    uly.stub_comm = True
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0.record()
    for _ in range(2):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1) / 2], device=dev)
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    exposed = float(ms.item() - ms2.item())
if rank == 0:
    flops = (111.5e12 + 4.0 * 24 * (f * h * w + T) ** 2 * 128 * layers) * layers / 60 if layers != 60 else None
    print(json.dumps(dict(workload=f"Qwen-Image 1024x2048 transformer forward, {layers} blocks, {f * h * w}+{T} tokens, "
                                   f"{'FP8' if quant != torch.int8 else 'INT8'} W8A8, bf16 attention, random init",
                          n_gpus=world, ms_per_step=float(ms.item()), a2a_exposed_ms=exposed,
                          gpu_launches=(_lib.launch_count - c0) // steps, finite=bool(torch.isfinite(y.float()).all()))))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
