"""SD3.5-medium 1024x1024 denoise step, batch 2 (CFG) -- BASELINE configs[1]: 24 MMDiT blocks, d = 1536, 24 x 64 heads, FP8."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import _lib
from fastdm_b200.models import SD3TransformerModelCore
dev, bf = "cuda", torch.bfloat16
model = SD3TransformerModelCore(device=dev, seed=0)
g = torch.Generator().manual_seed(1)
lat = torch.randn(2, 16, 128, 128, generator=g).to(bf).to(dev)
txt = torch.randn(2, 333, 4096, generator=g).to(bf).to(dev)
pooled = torch.randn(2, 2048, generator=g).to(bf).to(dev)
ts = torch.tensor([500.0, 500.0]).to(bf).to(dev)
from fastdm_b200.graph import GraphedStep
inputs = dict(lat=lat, txt=txt, pooled=pooled, ts=ts)
step = lambda d: model.forward(d["lat"], d["txt"], d["pooled"], d["ts"])[0]  # noqa: E731
graph = os.environ.get("SD3_GRAPH", "1") != "0"
runner = GraphedStep(step, inputs) if graph else step
fn = lambda: runner(inputs)  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
c0 = _lib.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = fn()
e1.record()
torch.cuda.synchronize()
print(json.dumps(dict(workload="SD3.5-medium 1024x1024, batch 2, 4096 image + 333 text tokens, 24 blocks, FP8 W8A8, bf16 attention (hd 64), random init",
                      launch="CUDA graph replay" if graph else "eager", ms_per_step=e0.elapsed_time(e1) / 10, gpu_launches=(_lib.launch_count - c0) // 10,
                      finite=bool(torch.isfinite(y.float()).all()))))
