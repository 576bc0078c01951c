"""The drop-in, proven on hardware: the REFERENCE's own op call sequence through fastdm_b200's public op API.

oracle/blocks_ref.py restates FastDM's layer / block glue exactly as the reference writes it -- `QLinear.forward`
(fastdm/layer/qlinear.py:56-81: flatten, quantize_to_fp8 | quantize_to_int8(asym), fp8_matmul | int8_matmul),
`Attention.forward` (fastdm/layer/transformer.py:232-317: column slices of the fused qkv, `.contiguous()` copies,
rms_norm on [B, S, H, hd] views, three torch.cat, in-place rotary_pos_embedding, scaled_dot_product_attention),
FeedForward, the AdaLN variants, the block forwards -- on top of nine op functions. Here those nine names are bound to
`fastdm_b200.ops` (what `fastdm_b200.integration.install()` registers as FastDM's `cuda` backend) instead of the
CPU oracle, tensors live on the GPU, and the outputs must match the fixtures produced by the real reference classes.
Nothing of the fused block path (fastdm_b200/blocks.py) is involved: this is the reference's unfused call pattern,
strided views and all, on our kernels.
"""
import types

import pytest
import torch

from conftest import golden
from oracle import blocks_ref as B

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture()
def b200_ops(monkeypatch, lib):
    from fastdm_b200 import integration, ops

    shim = types.SimpleNamespace(
        quantize_to_fp8=ops.quantize_to_fp8, quantize_to_int8=ops.quantize_to_int8, fp8_matmul=ops.fp8_matmul,
        int8_matmul=ops.int8_matmul, rms_norm=ops.rms_norm, rotary_pos_embedding=ops.rotary_pos_embedding,
        gelu_and_mul=ops.gelu_and_mul, scaled_dot_product_attention=ops.scaled_dot_product_attention,
        sparse_scaled_dot_product_attention=ops.sparse_scaled_dot_product_attention)
    # the nine functions are exactly what install() hands to FastDM's registry
    assert set(integration.OPS.values()) == {getattr(shim, n) for n in vars(shim)}
    monkeypatch.setattr(B, "R", shim)
    return shim


def to_dev(sd):
    return {k: v.to(DEV) for k, v in sd.items()}


def close(got, want, what):
    a, b = got.flatten().double().cpu(), want.flatten().double()
    cos = float((a @ b) / (a.norm() * b.norm()))
    err = float((got.float().cpu() - want.float()).abs().max())
    assert got.dtype == want.dtype and got.shape == want.shape
    assert cos >= 0.999 and err <= 0.05 * float(want.float().abs().max()), f"{what}: cosine {cos}, max abs err {err}"
    return cos


@pytest.mark.parametrize("tag,quant", [("fp8", torch.float8_e4m3fn), ("int8", torch.int8)])
def test_reference_call_sequence_flux_blocks(b200_ops, tag, quant):
    c = golden(f"block_flux_{tag}.pt")
    blk = B.FluxTransformerBlockRef(to_dev(c["sd_double"]), "transformer_blocks.0", c["heads"], c["hd"], quant)
    assert blk.attn.qkv.weight.is_cuda and blk.attn.qkv.weight.dtype == quant       # weights quantised by OUR quant op
    enc, hid = blk.forward(c["img"].to(DEV), c["txt"].to(DEV), c["temb"].to(DEV), c["rope"].to(DEV))
    close(enc, c["enc_out"], "double block / text stream")
    close(hid, c["hid_out"], "double block / image stream")
    sblk = B.FluxSingleTransformerBlockRef(to_dev(c["sd_single"]), "single_transformer_blocks.0", c["heads"], c["hd"], quant)
    close(sblk.forward(c["single_in"].to(DEV), c["temb"].to(DEV), c["rope"].to(DEV)), c["single_out"], "single block")


def test_reference_call_sequence_wan_block(b200_ops):
    c = golden("block_wan_fp8.pt")
    blk = B.WanTransformerBlockRef(to_dev(c["sd"]), "blocks.0", c["heads"], c["hd"], torch.float8_e4m3fn)
    y = blk.forward(c["x"].to(DEV), c["enc"].to(DEV), c["temb"].to(DEV), (c["cos"].to(DEV), c["sin"].to(DEV)))
    close(y, c["y"], "wan block")


def test_reference_call_sequence_qwen_block(b200_ops):
    c = golden("block_qwen_int8.pt")
    blk = B.QwenImageTransformerBlockRef(to_dev(c["sd"]), "transformer_blocks.0", c["heads"], c["hd"], torch.int8)
    enc, hid = blk.forward(c["img"].to(DEV), c["txt"].to(DEV), c["temb"].to(DEV), c["rope"].to(DEV))
    close(enc, c["enc_out"], "qwen block / text stream")
    close(hid, c["hid_out"], "qwen block / image stream")


def test_reference_call_sequence_at_c1_width(b200_ops):
    """The same unfused call pattern at BASELINE config C1 (4096 + 512 tokens, d = 3072, 24 x 128 heads) must agree with the
    fused block path (fastdm_b200/blocks.py) -- two routes to the same block over the same kernels."""
    from fastdm_b200.blocks import FluxTransformerBlock

    dim, heads, hd, quant = 3072, 24, 128, torch.float8_e4m3fn
    g = torch.Generator().manual_seed(77)
    img = torch.randn(1, 4096, dim, generator=g).to(torch.bfloat16).to(DEV)
    txt = torch.randn(1, 512, dim, generator=g).to(torch.bfloat16).to(DEV)
    temb = torch.randn(1, dim, generator=g).to(torch.bfloat16).to(DEV)
    rope = torch.rand(4608, hd, generator=g).to(torch.bfloat16).to(DEV)
    sd = to_dev(B.flux_double_state_dict("transformer_blocks.0", dim, hd, seed=78))
    enc_u, hid_u = B.FluxTransformerBlockRef(sd, "transformer_blocks.0", heads, hd, quant).forward(img, txt, temb, rope)
    enc_f, hid_f = FluxTransformerBlock(sd, "transformer_blocks.0", heads, hd, quant, DEV).forward(img, txt, temb, rope)
    c1 = close(enc_f, enc_u.cpu(), "fused vs unfused / text")
    c2 = close(hid_f, hid_u.cpu(), "fused vs unfused / image")
    print(f"C1 double block, unfused reference call sequence vs fused path on the same kernels: cosine {c1:.6f} / {c2:.6f}")
