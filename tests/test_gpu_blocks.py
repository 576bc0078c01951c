"""GPU parity of the fused block path (fastdm_b200/blocks.py, every kernel through the C ABI)
against (1) the golden block fixtures produced by the REAL reference classes at reduced width and
(2) the CPU oracle (oracle/blocks_ref.py) at the BASELINE C1 shapes (FLUX block pair, 4096 image +
512 text tokens, d = 3072, 24 x 128 heads).

Bar (BASELINE.json north_star): cosine >= 0.999 on full-block outputs; we also bound the max abs
error relative to the output scale. The fused kernels reproduce the reference's rounding chain, so
the residual differences come from GEMM/attention accumulation order only.
"""
import pytest
import torch

from conftest import golden
from oracle import blocks_ref as B

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def cosine(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return float((a @ b) / (a.norm() * b.norm()))


def check(got, want, what, cos_min=0.999, rel=0.05):
    c = cosine(got, want)
    err = (got.float().cpu() - want.float()).abs().max().item()
    scale = want.float().abs().max().item()
    assert c >= cos_min, f"{what}: cosine {c}"
    assert err <= rel * scale, f"{what}: max abs err {err} vs scale {scale}"
    return c


def to_dev(sd):
    return {k: v.to(DEV) for k, v in sd.items()}


@pytest.mark.parametrize("tag,quant", [("fp8", torch.float8_e4m3fn), ("int8", torch.int8)])
def test_flux_blocks_golden(lib, tag, quant):
    from fastdm_b200.blocks import FluxSingleTransformerBlock, FluxTransformerBlock

    c = golden(f"block_flux_{tag}.pt")
    blk = FluxTransformerBlock(to_dev(c["sd_double"]), "transformer_blocks.0", c["heads"], c["hd"], quant)
    enc, hid = blk.forward(c["img"].to(DEV), c["txt"].to(DEV), c["temb"].to(DEV), c["rope"].to(DEV))
    check(enc, c["enc_out"], "double block / text stream")
    check(hid, c["hid_out"], "double block / image stream")
    sblk = FluxSingleTransformerBlock(to_dev(c["sd_single"]), "single_transformer_blocks.0", c["heads"], c["hd"], quant)
    out = sblk.forward(c["single_in"].to(DEV), c["temb"].to(DEV), c["rope"].to(DEV))
    check(out, c["single_out"], "single block")


@pytest.mark.parametrize("tag,quant", [("fp8", torch.float8_e4m3fn), ("int8", torch.int8)])
def test_wan_block_golden(lib, tag, quant):
    from fastdm_b200.blocks import WanTransformerBlock

    c = golden(f"block_wan_{tag}.pt")
    blk = WanTransformerBlock(to_dev(c["sd"]), "blocks.0", c["heads"], c["hd"], quant)
    y = blk.forward(c["x"].to(DEV), c["enc"].to(DEV), c["temb"].to(DEV), (c["cos"].to(DEV), c["sin"].to(DEV)))
    check(y, c["y"], "wan block")


@pytest.mark.parametrize("tag,quant", [("int8", torch.int8), ("fp8", torch.float8_e4m3fn)])
def test_qwen_block_golden(lib, tag, quant):
    from fastdm_b200.blocks import QwenImageTransformerBlock

    c = golden(f"block_qwen_{tag}.pt")
    blk = QwenImageTransformerBlock(to_dev(c["sd"]), "transformer_blocks.0", c["heads"], c["hd"], quant)
    enc, hid = blk.forward(c["img"].to(DEV), c["txt"].to(DEV), None, c["temb"].to(DEV), c["rope"].to(DEV))
    check(enc, c["enc_out"], "qwen block / text stream")
    check(hid, c["hid_out"], "qwen block / image stream")


@pytest.mark.parametrize("name", ["dual", "plain", "last"])
def test_sd3_blocks_golden(lib, name):
    """SD3.5 JointTransformerBlock: dual attention (layers 0-12 of SD3.5-medium), plain, and the
    context_pre_only last block; batch 2 (CFG), head_dim 64, no RoPE."""
    from fastdm_b200.blocks import JointTransformerBlock

    c = golden("block_sd3_fp8.pt")
    blk = c["blocks"][name]
    sd = B.sd3_block_state_dict("transformer_blocks.0", c["dim"], c["hd"], blk["seed"], blk["context_pre_only"], blk["dual"])
    m = JointTransformerBlock(to_dev(sd), "transformer_blocks.0", c["heads"], c["hd"], torch.float8_e4m3fn,
                              context_pre_only=blk["context_pre_only"], use_dual_attention=blk["dual"])
    enc, hid = m.forward(c["img"].to(DEV), c["txt"].to(DEV), c["temb"].to(DEV))
    check(hid, blk["hid_out"], f"sd3 {name} / image stream")
    if blk["context_pre_only"]:
        assert enc is None
    else:
        check(enc, blk["enc_out"], f"sd3 {name} / text stream")


def test_qlinear_weight_quant_matches_reference_cpu_quant(lib):
    # load-time weight quantisation on the GPU == fastdm/utils/quantization.py on the CPU, bit for bit
    from fastdm_b200.layers import load_linear
    from oracle.blocks_ref import QLinearRef

    g = torch.Generator().manual_seed(3)
    sd = {"a.weight": (torch.randn(192, 256, generator=g) * 0.02).to(BF), "a.bias": torch.randn(192, generator=g).to(BF),
          "b.weight": (torch.randn(64, 256, generator=g) * 0.02).to(BF), "b.bias": torch.randn(64, generator=g).to(BF)}
    for quant in (torch.float8_e4m3fn, torch.int8):
        ref = QLinearRef(sd, ["a", "b"], quant)
        lin = load_linear(to_dev(sd), ["a", "b"], quant)
        assert torch.equal(lin.weight.cpu().view(torch.uint8), ref.weight.view(torch.uint8))
        assert torch.equal(lin.weight_quant_scale.cpu(), ref.scale)
        assert torch.equal(lin.bias.cpu(), ref.bias)
        if quant == torch.int8:
            assert torch.equal(lin.weight_asym_sumcol.cpu(), ref.colsum)
        x = torch.randn(2, 37, 256, generator=g).to(BF)
        y = lin.forward(x.to(DEV))
        want = ref.forward(x)
        assert y.shape == want.shape
        torch.testing.assert_close(y.cpu().float(), want.float(), rtol=1.6e-2, atol=2e-2 * float(want.float().abs().mean()))


def test_fused_norm_rope_equals_separate_ops(lib):
    from fastdm_b200 import ops

    g = torch.Generator().manual_seed(5)
    S, H, hd = 333, 24, 128
    d = H * hd
    fused = torch.randn(S, 3 * d, generator=g).to(BF).to(DEV)
    wq = torch.randn(hd, generator=g).to(BF).to(DEV)
    wk = torch.randn(hd, generator=g).to(BF).to(DEV)
    cs = torch.rand(S + 7, hd, generator=g).to(BF).to(DEV)
    a = fused.clone()
    ops.qk_norm_rope_(a, wq, wk, cs, H, H, hd, 0, d, 7, 1e-6)
    q = ops.rms_norm(fused[:, :d].reshape(S, H, hd).contiguous(), wq, 1e-6).view(1, S, d)
    k = ops.rms_norm(fused[:, d:2 * d].reshape(S, H, hd).contiguous(), wk, 1e-6).view(1, S, d)
    ops.rotary_pos_embedding(q, k, hd, cs[7:], False)
    assert torch.equal(a[:, :d], q[0]) and torch.equal(a[:, d:2 * d], k[0]) and torch.equal(a[:, 2 * d:], fused[:, 2 * d:])
    # Wan: across-heads norm
    H2 = 40
    d2 = H2 * hd
    fused = torch.randn(50, 3 * d2, generator=g).to(BF).to(DEV)
    wq = torch.randn(d2, generator=g).to(BF).to(DEV)
    wk = torch.randn(d2, generator=g).to(BF).to(DEV)
    cs = torch.rand(50, hd, generator=g).to(BF).to(DEV)
    a = fused.clone()
    ops.qk_norm_rope_(a, wq, wk, cs, H2, H2, hd, 0, d2, 0, 1e-6, across_heads=True)
    q = ops.rms_norm(fused[:, :d2].contiguous(), wq, 1e-6).view(1, 50, d2)
    k = ops.rms_norm(fused[:, d2:2 * d2].contiguous(), wk, 1e-6).view(1, 50, d2)
    ops.rotary_pos_embedding(q, k, hd, cs, False)
    assert torch.equal(a[:, :d2], q[0]) and torch.equal(a[:, d2:2 * d2], k[0])


def test_layernorm_modulate_quant_equals_unfused(lib):
    import torch.nn.functional as F

    from fastdm_b200 import ops

    g = torch.Generator().manual_seed(6)
    Bn, S, d = 2, 300, 3072
    x = torch.randn(Bn * S, d, generator=g).to(BF).to(DEV)
    scale = (torch.randn(Bn, d, generator=g) * 0.2).to(BF).to(DEV)
    shift = (torch.randn(Bn, d, generator=g) * 0.2).to(BF).to(DEV)
    # FLUX chain (bf16 tensor ops): normalization.py:196
    want = (F.layer_norm(x.view(Bn, S, d), (d,), None, None, 1e-6) * (1 + scale[:, None]) + shift[:, None]).view(Bn * S, d)
    q, s, zp, y = ops.layernorm_modulate_quant(x, (1 + scale).float(), shift.float(), S, torch.float8_e4m3fn, 1e-6,
                                               round_steps=True, want_y=True)
    # a 1-ulp difference in LN(x) (|LN| up to ~4 -> ulp 2^-6) survives the modulate even where the result
    # cancels to ~0, so the bound is absolute: 2 ulps of the largest intermediate; almost all elements equal
    assert float((y.float() - want.float()).abs().max()) <= 0.07
    assert float((y != want).float().mean()) < 5e-3
    rq, rs = ops.quantize_to_fp8(y)
    assert torch.equal(q.view(torch.uint8), rq.view(torch.uint8)) and torch.equal(s, rs)
    q8, s8, zp8, _ = ops.layernorm_modulate_quant(x, (1 + scale).float(), shift.float(), S, torch.int8, 1e-6)
    rq, rs, rzp = ops.quantize_to_int8(y, False)
    assert torch.equal(q8, rq) and torch.equal(s8, rs) and torch.equal(zp8, rzp)
    # Wan chain (fp32): wan.py:95
    sc32, sh32 = scale.float(), shift.float()
    want = (F.layer_norm(x.view(Bn, S, d).float(), (d,), None, None, 1e-6) * (1 + sc32[:, None]) + sh32[:, None]).to(BF)
    _, _, _, y = ops.layernorm_modulate_quant(x, 1 + sc32, sh32, S, None, 1e-6, round_steps=False)
    want = want.view(Bn * S, d)
    assert float((y.float() - want.float()).abs().max()) <= 0.04   # one rounding: 1 ulp at magnitude < 8
    assert float((y != want).float().mean()) < 5e-3


@pytest.mark.parametrize("d", [1536, 3072, 5120, 6144, 1000])
def test_layernorm_modulate_quant_all_paths(lib, d):
    """Every kernel / argument form of the fused LayerNorm-modulate-quant: the warp-per-row kernel (d <= 5120) with bf16
    modulation vectors (packed bf16 chain), with fp32 vectors (element-wise chain) and the fp32 Wan chain; the
    CTA-per-row kernel for wider rows (d = 6144); a width that is not a multiple of 256 (d = 1000); missing mul / add.
    bf16 and fp32 forms of the same bf16-valued vectors must agree bit for bit, codes must equal quantising y."""
    import torch.nn.functional as F

    from fastdm_b200 import ops

    g = torch.Generator().manual_seed(d)
    Bn, S = 2, 173
    x = (torch.randn(Bn * S, d, generator=g) * 2 + 0.3).to(BF).to(DEV)
    x[5] = 7.0                      # constant row: variance 0, LN = 0, y = shift
    x[6, :] = 1000.0
    x[6, 3] = 1001.0                # the E[x^2] - mean^2 cancellation branch
    scale = (torch.randn(Bn, d, generator=g) * 0.2).to(BF).to(DEV)
    shift = (torch.randn(Bn, d, generator=g) * 0.2).to(BF).to(DEV)
    a16, c16 = (1 + scale), shift
    want = (F.layer_norm(x.view(Bn, S, d), (d,), None, None, 1e-6) * a16[:, None] + c16[:, None]).view(Bn * S, d)
    ok = torch.ones(Bn * S, dtype=torch.bool, device=DEV)
    ok[6] = False                   # torch's own bf16 LayerNorm loses this row to cancellation; checked on its own below
    for quant in (torch.float8_e4m3fn, torch.int8):
        qb, sb, zb, yb = ops.layernorm_modulate_quant(x, a16, c16, S, quant, 1e-6, round_steps=True, want_y=True)
        qf, sf, zf, yf = ops.layernorm_modulate_quant(x, a16.float(), c16.float(), S, quant, 1e-6, round_steps=True, want_y=True)
        assert torch.equal(yb, yf) and torch.equal(qb.view(torch.uint8), qf.view(torch.uint8)) and torch.equal(sb, sf)
        assert float((yb.float() - want.float())[ok].abs().max()) <= 0.07
        assert float((yb != want)[ok].float().mean()) < 5e-3
        if quant == torch.int8:
            rq, rs, rzp = ops.quantize_to_int8(yb, False)
            # (constant rows have max == min: scale 0, NaN codes in the reference too -- compare the others)
            live = (yb.float().amax(1) > yb.float().amin(1))
            assert torch.equal(qb[live], rq[live]) and torch.equal(sb[live], rs[live]) and torch.equal(zb[live], rzp[live])
        else:
            rq, rs = ops.quantize_to_fp8(yb)
            assert torch.equal(qb.view(torch.uint8), rq.view(torch.uint8)) and torch.equal(sb, rs)
    # the cancellation row: mean 1000 + 1/d, one element 1 above the rest
    ln6 = F.layer_norm(x[6:7].float(), (d,), None, None, 1e-6)
    y6 = (ln6.to(BF) * a16[0:1]).to(BF) + c16[0:1]
    assert float((yb[6:7].float() - y6.float()).abs().max()) <= 0.07 * max(1.0, float(ln6.abs().max()) / 4)
    # only mul, only add, neither
    for (a, c) in ((a16, None), (None, c16), (None, None)):
        want1 = F.layer_norm(x.view(Bn, S, d), (d,), None, None, 1e-6)
        if a is not None:
            want1 = want1 * a[:, None]
        if c is not None:
            want1 = want1 + c[:, None]
        _, _, _, y1 = ops.layernorm_modulate_quant(x, a, c, S, None, 1e-6, round_steps=True)
        assert float((y1 != want1.view(Bn * S, d))[ok].float().mean()) < 5e-3
    # Wan chain (fp32 vectors, one rounding): wan.py:95
    sc32 = scale.float() + 1e-3 * torch.randn(Bn, d, generator=g).to(DEV)     # not bf16-valued
    want = (F.layer_norm(x.view(Bn, S, d).float(), (d,), None, None, 1e-6) * (1 + sc32[:, None]) + shift.float()[:, None]).to(BF)
    q, s_, _, y = ops.layernorm_modulate_quant(x, 1 + sc32, shift.float(), S, torch.float8_e4m3fn, 1e-6, round_steps=False, want_y=True)
    want = want.view(Bn * S, d)
    assert float((y.float() - want.float())[ok].abs().max()) <= 0.04
    assert float((y != want)[ok].float().mean()) < 5e-3
    rq, rs = ops.quantize_to_fp8(y)
    assert torch.equal(q.view(torch.uint8), rq.view(torch.uint8)) and torch.equal(s_, rs)


def test_gemm_gate_residual_epilogue_equals_unfused(lib):
    from fastdm_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(8)
    M, K, N = 700, 1024, 3072
    a = torch.randn(M, K, device=DEV, generator=g).to(torch.float8_e4m3fn)
    b = (torch.randn(N, K, device=DEV, generator=g) * 0.05).to(torch.float8_e4m3fn).t()
    sa = torch.rand(M, 1, device=DEV, generator=g) * 0.1
    sb = torch.rand(N, 1, device=DEV, generator=g)
    bias = torch.randn(N, device=DEV, generator=g).to(BF)
    res = torch.randn(M, N, device=DEV, generator=g).to(BF)
    gate = torch.randn(2, N, device=DEV, generator=g).to(BF)
    plain = ops.fp8_matmul(a, b, sa, sb, BF, bias)
    rows_per_batch = 350
    gfull = gate.repeat_interleave(rows_per_batch, dim=0)
    want = res + gfull * plain                      # bf16 tensor ops: flux.py:153-154
    got = ops.fp8_matmul(a, b, sa, sb, BF, bias, gate=gate.float(), residual=res, rows_per_batch=rows_per_batch)
    assert torch.equal(got, want)
    want32 = (res.float() + plain * gfull.float()).to(BF)   # wan.py:97
    got32 = ops.fp8_matmul(a, b, sa, sb, BF, bias, gate=gate.float(), residual=res, rows_per_batch=rows_per_batch,
                           round_steps=False)
    assert torch.equal(got32, want32)
    inplace = res.clone()
    ops.fp8_matmul(a, b, sa, sb, BF, bias, residual=inplace, out=inplace)   # wan.py:105
    assert torch.equal(inplace, res + plain)


def test_flux_block_pair_full_size_vs_oracle(lib):
    """BASELINE config C1: 1 double + 1 single FLUX block, 4096 image + 512 text tokens, d = 3072."""
    from fastdm_b200.blocks import FluxSingleTransformerBlock, FluxTransformerBlock

    dim, heads, hd = 3072, 24, 128
    g = torch.Generator().manual_seed(21)
    img = torch.randn(1, 4096, dim, generator=g).to(BF)
    txt = torch.randn(1, 512, dim, generator=g).to(BF)
    temb = torch.randn(1, dim, generator=g).to(BF)
    rope = torch.rand(4608, hd, generator=g).to(BF)
    sd = B.flux_double_state_dict("transformer_blocks.0", dim, hd, seed=31)
    sd1 = B.flux_single_state_dict("single_transformer_blocks.0", dim, hd, seed=32)
    quant = torch.float8_e4m3fn
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref_d = B.FluxTransformerBlockRef(sd, "transformer_blocks.0", heads, hd, quant)
    enc_r, hid_r = ref_d.forward(img, txt, temb, rope)
    cat_r = torch.cat([enc_r, hid_r], dim=1)
    out_r = B.FluxSingleTransformerBlockRef(sd1, "single_transformer_blocks.0", heads, hd, quant).forward(cat_r, temb, rope)

    blk = FluxTransformerBlock(to_dev(sd), "transformer_blocks.0", heads, hd, quant)
    enc, hid = blk.forward(img.to(DEV), txt.to(DEV), temb.to(DEV), rope.to(DEV))
    c1 = check(enc, enc_r, "C1 double / text")
    c2 = check(hid, hid_r, "C1 double / image")
    sblk = FluxSingleTransformerBlock(to_dev(sd1), "single_transformer_blocks.0", heads, hd, quant)
    out = sblk.forward(torch.cat([enc, hid], dim=1), temb.to(DEV), rope.to(DEV))
    c3 = check(out, out_r, "C1 single")
    print(f"C1 block-pair cosines: text {c1:.6f} image {c2:.6f} single {c3:.6f}")


def test_wan_block_full_width_vs_oracle(lib):
    """The headline workload's block at its real width: Wan2.2-A14B, d = 5120, 40 x 128 heads, ffn 13824, FP8, on 2304
    video tokens + 512 text tokens -- enough for the CTA-pair attention kernel (Sk >= 1024), the BN = 256 GEMM tiles on
    multi-wave persistent grids and the across-heads q/k-norm kernel at 5120 columns. Oracle: oracle/blocks_ref.py
    (fastdm/model/wan.py:67-114) on the CPU."""
    from fastdm_b200.blocks import WanTransformerBlock

    dim, heads, hd, ffn, S, T = 5120, 40, 128, 13824, 2304, 512
    g = torch.Generator().manual_seed(51)
    x = torch.randn(1, S, dim, generator=g).to(BF)
    enc = torch.randn(1, T, dim, generator=g).to(BF)
    temb = torch.randn(1, 6, dim, generator=g).to(BF)
    cos = torch.rand(1, S, 1, hd, generator=g)
    sin = torch.rand(1, S, 1, hd, generator=g)
    sd = B.wan_block_state_dict("blocks.0", dim, ffn, seed=52)
    quant = torch.float8_e4m3fn
    want = B.WanTransformerBlockRef(sd, "blocks.0", heads, hd, quant).forward(x, enc, temb, (cos, sin))
    blk = WanTransformerBlock(to_dev(sd), "blocks.0", heads, hd, quant)
    got = blk.forward(x.to(DEV), enc.to(DEV), temb.to(DEV), (cos.to(DEV), sin.to(DEV)))
    c = check(got, want, "Wan block d=5120")
    print(f"Wan full-width block cosine: {c:.6f}")


def test_qwen_block_full_width_vs_oracle(lib):
    """Qwen-Image block at its real width: d = 3072, 24 x 128 heads, INT8 (asymmetric per-token activations), 2304 image +
    128 text tokens (fastdm/model/qwenimage.py:58-124)."""
    from fastdm_b200.blocks import QwenImageTransformerBlock

    dim, heads, hd, S, T = 3072, 24, 128, 2304, 128
    g = torch.Generator().manual_seed(61)
    img = torch.randn(1, S, dim, generator=g).to(BF)
    txt = torch.randn(1, T, dim, generator=g).to(BF)
    temb = torch.randn(1, dim, generator=g).to(BF)
    rope = torch.rand(S + T, hd, generator=g).to(BF)
    sd = B.qwen_block_state_dict("transformer_blocks.0", dim, hd, seed=62)
    quant = torch.int8
    enc_r, hid_r = B.QwenImageTransformerBlockRef(sd, "transformer_blocks.0", heads, hd, quant).forward(img, txt, temb, rope)
    blk = QwenImageTransformerBlock(to_dev(sd), "transformer_blocks.0", heads, hd, quant)
    enc, hid = blk.forward(img.to(DEV), txt.to(DEV), None, temb.to(DEV), rope.to(DEV))
    c1 = check(enc, enc_r, "Qwen block d=3072 / text")
    c2 = check(hid, hid_r, "Qwen block d=3072 / image")
    print(f"Qwen full-width block cosines: text {c1:.6f} image {c2:.6f}")


@pytest.mark.parametrize("dual", [True, False])
def test_sd3_block_full_width_vs_oracle(lib, dual):
    """SD3.5-medium block at its real width: d = 1536, 24 x 64 heads, FP8, batch 2 (CFG), 2304 image + 333 text tokens,
    with and without the second (image-only) attention (fastdm/model/sd35.py:133-200)."""
    from fastdm_b200.blocks import JointTransformerBlock

    dim, heads, hd, S, T = 1536, 24, 64, 2304, 333
    g = torch.Generator().manual_seed(71)
    img = torch.randn(2, S, dim, generator=g).to(BF)
    txt = torch.randn(2, T, dim, generator=g).to(BF)
    temb = torch.randn(2, dim, generator=g).to(BF)
    sd = B.sd3_block_state_dict("transformer_blocks.0", dim, hd, 72, False, dual)
    quant = torch.float8_e4m3fn
    enc_r, hid_r = B.JointTransformerBlockRef(sd, "transformer_blocks.0", heads, hd, quant, False, dual).forward(img, txt, temb)
    blk = JointTransformerBlock(to_dev(sd), "transformer_blocks.0", heads, hd, quant, context_pre_only=False,
                                use_dual_attention=dual)
    enc, hid = blk.forward(img.to(DEV), txt.to(DEV), temb.to(DEV))
    c1 = check(enc, enc_r, "SD3.5 block d=1536 / text")
    c2 = check(hid, hid_r, "SD3.5 block d=1536 / image")
    print(f"SD3.5 full-width block (dual={dual}) cosines: text {c1:.6f} image {c2:.6f}")


@pytest.mark.parametrize("quant", [torch.float8_e4m3fn, torch.int8])
def test_final_latent_cosine_after_n_steps(quant):
    """north_star: "a final-latent cosine after N steps is also reported". An 8-step flow-matching Euler loop
    x <- x + dt * v(x, t) whose velocity model is a FLUX double + single block pair (random init, the step's
    timestep embedding changes every step) runs once on the product path (CUDA, C ABI) and once on the oracle
    (CPU); quantisation error compounds over the steps, so this bounds its drift. Measured on B200: fp8 0.999999,
    int8 1.000000 (printed with -s)."""
    from fastdm_b200.blocks import FluxSingleTransformerBlock, FluxTransformerBlock
    from oracle import blocks_ref as B

    dev, bf = "cuda", torch.bfloat16
    dim, heads, hd, n_img, n_txt, steps = 256, 2, 128, 320, 64, 8
    g = torch.Generator().manual_seed(42)
    x0 = torch.randn(1, n_img, dim, generator=g).to(bf)
    txt = torch.randn(1, n_txt, dim, generator=g).to(bf)
    tembs = [torch.randn(1, dim, generator=g).to(bf) for _ in range(steps)]
    rope = torch.rand(n_img + n_txt, hd, generator=g).to(bf)
    sd = B.flux_double_state_dict("transformer_blocks.0", dim, hd, seed=1)
    sd1 = B.flux_single_state_dict("single_transformer_blocks.0", dim, hd, seed=2)

    def sample(double, single, move):
        x, e, r = move(x0), move(txt), move(rope)
        dt = 1.0 / steps
        for i in range(steps):
            enc, hid = double.forward(x, e, move(tembs[i]), r)
            out = single.forward(torch.cat([enc, hid], 1), move(tembs[i]), r)[:, n_txt:]
            v = (out.float() - x.float())                      # the pair's residual update acts as the velocity
            x = (x.float() + dt * v).to(bf)
        return x

    ref = sample(B.FluxTransformerBlockRef(sd, "transformer_blocks.0", heads, hd, quant),
                 B.FluxSingleTransformerBlockRef(sd1, "single_transformer_blocks.0", heads, hd, quant), lambda t: t)
    to = lambda d: {k: v.to(dev) for k, v in d.items()}  # noqa: E731
    got = sample(FluxTransformerBlock(to(sd), "transformer_blocks.0", heads, hd, quant, dev),
                 FluxSingleTransformerBlock(to(sd1), "single_transformer_blocks.0", heads, hd, quant, dev),
                 lambda t: t.to(dev))
    a, b = got.flatten().double().cpu(), ref.flatten().double()
    cos = float((a @ b) / (a.norm() * b.norm()))
    print(f"final-latent cosine after {steps} steps ({quant}): {cos:.6f}")
    assert cos >= 0.999, cos


def test_qwen_image_model_forward_and_rope_table():
    """Whole Qwen-Image core (2 layers): runs, is finite, AdaLN table == per-block modulation; the rope table
    follows QwenEmbedRope (scale_rope): text rows sit at max(h//2, w//2) + i on all axes, image rows are centred."""
    from fastdm_b200.models import QwenImageTransformer2DModelCore, qwen_rope_table

    dev, bf = "cuda", torch.bfloat16
    tab = qwen_rope_table(1, 4, 6, 5, dtype=torch.float32, device="cpu")
    assert tab.shape == (5 + 24, 128)
    # text row i: the same position on all three axes -> cos of the first frequency (1.0) of each axis block agrees
    pos = max(4 // 2, 6 // 2) + torch.arange(5).float()
    assert torch.allclose(tab[:5, 0], torch.cos(pos)) and torch.allclose(tab[:5, 8], torch.cos(pos)) and torch.allclose(tab[:5, 36], torch.cos(pos))
    # image row (f=0, h=0, w=0): height position -(4 - 2) = -2, width position -(6 - 3) = -3
    assert torch.allclose(tab[5, 8], torch.cos(torch.tensor(-2.0))) and torch.allclose(tab[5, 64 + 36], torch.sin(torch.tensor(-3.0)))
    model = QwenImageTransformer2DModelCore(num_layers=2, attention_head_dim=128, num_attention_heads=2,
                                            joint_attention_dim=256, quant_dtype=torch.int8, device=dev, seed=5)
    g = torch.Generator().manual_seed(2)
    f, h, w, T = 1, 16, 20, 77
    lat = torch.randn(1, f * h * w, 64, generator=g).to(bf).to(dev)
    txt = torch.randn(1, T, 256, generator=g).to(bf).to(dev)
    ts = torch.tensor([0.3]).to(dev)
    y = model.forward(lat, txt, ts, (f, h, w))[0].float()
    assert y.shape == (1, f * h * w, 64) and bool(torch.isfinite(y).all())
    model.use_adaln_table = False
    y2 = model.forward(lat, txt, ts, (f, h, w))[0].float()
    cos = torch.nn.functional.cosine_similarity(y.flatten(), y2.flatten(), dim=0).item()
    assert cos > 0.9999, cos


def test_sd3_model_core_forward():
    """Whole SD3.5-style core (3 layers: dual-attention, plain, context_pre_only last): runs end to end, finite,
    the right output shape, and deterministic."""
    from fastdm_b200.models import SD3TransformerModelCore

    dev, bf = "cuda", torch.bfloat16
    model = SD3TransformerModelCore(num_layers=3, attention_head_dim=64, num_attention_heads=4, joint_attention_dim=256,
                                    pooled_projection_dim=128, pos_embed_max_size=64, dual_attention_layers=(0,),
                                    device=dev, seed=9)
    g = torch.Generator().manual_seed(4)
    lat = torch.randn(2, 16, 48, 64, generator=g).to(bf).to(dev)
    txt = torch.randn(2, 77, 256, generator=g).to(bf).to(dev)
    pooled = torch.randn(2, 128, generator=g).to(bf).to(dev)
    ts = torch.tensor([500.0, 500.0]).to(bf).to(dev)
    y = model.forward(lat, txt, pooled, ts)[0]
    assert y.shape == lat.shape and bool(torch.isfinite(y.float()).all())
    assert torch.equal(y, model.forward(lat, txt, pooled, ts)[0])


@pytest.mark.parametrize("quant", [torch.float8_e4m3fn, torch.int8, None])
def test_prequantized_weight_cache_roundtrip(quant, tmp_path):
    """QLinear -> export -> file -> reload gives the bit-identical forward without touching the bf16 weights again."""
    from fastdm_b200.layers import QLinear, load_linear, load_quantized, save_quantized

    g = torch.Generator().manual_seed(3)
    sd = {"a.weight": (torch.randn(192, 256, generator=g) * 0.05).to(torch.bfloat16), "a.bias": torch.randn(192, generator=g).to(torch.bfloat16),
          "b.weight": (torch.randn(64, 256, generator=g) * 0.05).to(torch.bfloat16), "b.bias": torch.randn(64, generator=g).to(torch.bfloat16)}
    lin = load_linear(sd, ["a", "b"], quant, "cuda")          # fused along N like a qkv projection
    path = str(tmp_path / "w.pt")
    save_quantized({"ab": lin}, path)
    lin2 = load_quantized(path, "cuda")["ab"]
    assert isinstance(lin2, QLinear) and lin2.quant_type == quant
    x = torch.randn(70, 256, generator=g).to(torch.bfloat16).cuda()
    assert torch.equal(lin.forward(x), lin2.forward(x))
