"""GPU: Ulysses sequence-parallel Wan block on 2 GPUs (NCCL all-to-all) == the same block on 1 GPU.
Spawned with torchrun inside the test; skipped when fewer than 2 GPUs are visible. On one GPU the
pack/unpack kernels are checked against the torch view-op formulation."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_pack_unpack_kernels_match_view_ops(lib):
    from fastdm_b200 import ops

    S, H, hd, P = 50, 8, 128, 4
    d = H * hd
    x = torch.randn(S, 3 * d + 64, device="cuda").to(torch.bfloat16)[:, : 3 * d]   # strided rows
    packed = ops.ulysses_pack_heads(x, H, hd, P, 3)
    want = x.reshape(S, 3, P, H // P, hd).permute(2, 0, 1, 3, 4).reshape(P, S, 3 * d // P)
    assert torch.equal(packed, want)
    back = ops.ulysses_unpack_heads(packed, H, hd, 3)
    assert torch.equal(back, x)
    o = torch.randn(P, S, d // P, device="cuda").to(torch.bfloat16)
    un = ops.ulysses_unpack_heads(o, H, hd, 1)
    assert torch.equal(un, o.reshape(P, S, H // P, hd).permute(1, 0, 2, 3).reshape(S, d))


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FDM_ROOT"])
from fastdm_b200.blocks import WanTransformerBlock
from fastdm_b200.models import random_wan_block_sd, wan_rope_table
from fastdm_b200.ulysses import UlyssesAttention
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
H, hd, ffn = 8, 128, 2048
d = H * hd
g = torch.Generator(device=dev).manual_seed(0)
sd = random_wan_block_sd("blocks.0", d, ffn, g, dev)
blk = WanTransformerBlock(sd, "blocks.0", H, hd, torch.float8_e4m3fn, dev)
frames, hh, ww = 4, 16, 20
S = frames * hh * ww
x = torch.randn(1, S, d, device=dev, generator=g).to(torch.bfloat16)
enc = torch.randn(1, 64, d, device=dev, generator=g).to(torch.bfloat16)
temb = (torch.randn(1, 6, d, device=dev, generator=g) * 0.5).to(torch.bfloat16)
rope = wan_rope_table(frames, hh, ww, hd, torch.bfloat16, dev)
ref = blk.forward(x, enc, temb, rope)                                    # single-GPU path
ul = UlyssesAttention(H, hd)
xs = ul.shard_tokens(x, 1)
res = {}
for overlap in (True, False):
    y = blk.forward(xs, enc, temb, rope, ulysses=ul, pos0=rank * xs.shape[1], overlap=overlap)
    full = ul.gather_tokens(y, 1)
    err = (full.float() - ref.float()).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(full.flatten().double(), ref.flatten().double(), dim=0).item()
    res[overlap] = (err, cos)
# block-sparse self-attention with a DIFFERENT mask per head: every rank must pick its own heads' masks
gm = torch.Generator().manual_seed(5)
mask = (torch.rand(1, H, (S + 127) // 128, (S + 63) // 64, generator=gm) < 0.6).to(torch.int8)
mask[:, :, :, 0] = 1
mask = mask.to(dev)
ref_s = blk.forward(x, enc, temb, rope, mask)
for overlap in (True, False):
    y = blk.forward(xs, enc, temb, rope, mask, ulysses=ul, pos0=rank * xs.shape[1], overlap=overlap)
    full = ul.gather_tokens(y, 1)
    err = (full.float() - ref_s.float()).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(full.flatten().double(), ref_s.flatten().double(), dim=0).item()
    res["sparse_%s" % overlap] = (err, cos)
if rank == 0:
    print("ULYSSES_RESULT", res, float(ref.float().abs().max()))
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.timeout(600)
def test_wan_block_ulysses_2gpu_equals_single_gpu(lib, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, FDM_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", str(script)],
                       capture_output=True, text=True, env=env, timeout=550)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("ULYSSES_RESULT")][0]
    res = eval(line.split("ULYSSES_RESULT", 1)[1].rsplit("}", 1)[0] + "}")
    for overlap, (err, cos) in res.items():
        # identical arithmetic per head; only the GEMM column-group split changes nothing numerically
        assert cos >= 0.9999 and err <= 0.05, f"overlap={overlap}: err {err}, cos {cos}"


QWEN_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FDM_ROOT"])
from fastdm_b200.models import QwenImageTransformer2DModelCore
from fastdm_b200.ulysses import UlyssesAttention
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
dev, bf = "cuda", torch.bfloat16
heads, hd = 4, 128
res = {}
for quant in (torch.int8, torch.float8_e4m3fn):
    model = QwenImageTransformer2DModelCore(num_layers=2, attention_head_dim=hd, num_attention_heads=heads,
                                            joint_attention_dim=256, quant_dtype=quant, device=dev, seed=3)
    g = torch.Generator().manual_seed(1)
    f, h, w, T = 1, 16, 24, 96            # 384 image tokens + 96 text tokens: both divisible by 2 ranks
    lat = torch.randn(1, f * h * w, 64, generator=g).to(bf).to(dev)
    txt = torch.randn(1, T, 256, generator=g).to(bf).to(dev)
    ts = torch.tensor([0.6]).to(dev)
    ref = model.forward(lat, txt, ts, (f, h, w))[0]
    uly = UlyssesAttention(heads, hd)
    out = model.forward(lat, txt, ts, (f, h, w), ulysses=uly)[0]
    torch.cuda.synchronize()
    err = (out.float() - ref.float()).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(out.flatten().double(), ref.flatten().double(), dim=0).item()
    res[str(quant)] = (err, cos, float(ref.float().abs().max()))
if rank == 0:
    print("QWEN_ULYSSES_RESULT", res)
dist.barrier(); dist.destroy_process_group()
'''


@pytest.mark.timeout(600)
def test_qwen_image_model_ulysses_2gpu_equals_single_gpu(lib, tmp_path):
    """Qwen-Image (joint text + image attention, INT8 and FP8): tokens of both streams sharded over 2 ranks,
    head-sharded joint attention through two all-to-alls == the single-GPU forward."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "qwen_worker.py"
    script.write_text(QWEN_WORKER)
    env = dict(os.environ, FDM_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29513", str(script)],
                       capture_output=True, text=True, env=env, timeout=550)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("QWEN_ULYSSES_RESULT")][0]
    res = eval(line.split("QWEN_ULYSSES_RESULT", 1)[1])
    for quant, (err, cos, mag) in res.items():
        assert cos >= 0.9999 and err <= 0.05 * mag, f"{quant}: err {err}, cos {cos}, |ref| {mag}"
