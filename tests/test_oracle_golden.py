"""CPU: the oracle restatement (oracle/ops_ref.py, oracle/blocks_ref.py) against the golden
fixtures that oracle/gen_golden.py produced by running the REAL reference (tests/golden/*.pt).
Integer / byte results bit-exact; the composed blocks are bit-exact too because the restatement
performs the same torch CPU ops in the same order."""
import torch

from conftest import golden
from oracle import blocks_ref as B
from oracle import ops_ref as R

BF = torch.bfloat16


def eq(a, b):
    if a.dtype == torch.float8_e4m3fn:
        a = a.view(torch.uint8)
    if b.dtype == torch.float8_e4m3fn:
        b = b.view(torch.uint8)
    return torch.equal(a, b)


def test_quant_golden():
    for c in golden("quant.pt"):
        q, s = R.quantize_to_fp8(c["x"])
        assert eq(q, c["fp8_q"]) and eq(s, c["fp8_s"])
        q, s, zp = R.quantize_to_int8(c["x"], True)
        assert eq(q, c["s8_q"]) and eq(s, c["s8_s"]) and zp is None
        q, s, zp = R.quantize_to_int8(c["x_asym"], False)
        assert eq(q, c["a8_q"]) and eq(s, c["a8_s"]) and eq(zp, c["a8_zp"])


def test_fp8_scale_floor_constant():
    # torch evaluates clamp(min=1e-12) in bf16: the all-zero row's scale is bf16(1e-12)/448
    x = torch.zeros(1, 16, dtype=BF)
    _, s = R.quantize_to_fp8(x)
    assert torch.equal(s.view(-1), torch.tensor([1e-12]).to(BF).float() / 448.0)


def test_rmsnorm_golden():
    for c in golden("rmsnorm.pt"):
        assert eq(R.rms_norm(c["x"], c["w"], c["eps"]), c["y"])


def test_rope_golden():
    for c in golden("rope.pt"):
        q, k = c["q"].clone(), c["k"].clone()
        assert R.rotary_pos_embedding(q, k, c["hd"], c["cs"], c["neox"]) is None
        assert eq(q, c["q_out"]) and eq(k, c["k_out"])


def test_gelu_and_mul_golden():
    for c in golden("gelu_and_mul.pt"):
        assert eq(R.gelu_and_mul(c["x"]), c["y"])


def test_matmul_golden():
    for c in golden("matmul.pt"):
        b8 = c["b8_t"].t()
        bf = c["bf_t"].view(torch.float8_e4m3fn).t()
        af = c["af"].view(torch.float8_e4m3fn)
        assert eq(R.int8_matmul(c["a8"], b8, c["sa"], c["sb"], BF, c["adj"], c["azp"], c["bias"]), c["y_int8"])
        assert eq(R.int8_matmul(c["a8"], b8, c["sa"], c["sb"], BF, c["adj"], c["azp"], None), c["y_int8_nobias"])
        assert eq(R.fp8_matmul(af, bf, c["sa"], c["sb"], BF, c["bias"]), c["y_fp8"])
        assert eq(R.fp8_matmul(af, bf, c["sa"], c["sb"], BF, None), c["y_fp8_nobias"])


def test_attention_golden():
    for c in golden("attention.pt"):
        q, k, v, h, hd = c["q"], c["k"], c["v"], c["h"], c["hd"]
        y = R.scaled_dot_product_attention(q, k, v, h, h, hd, scale=c["scale"])
        assert eq(y, c["y"])
        b, sq, sk = q.shape[0], q.shape[1], k.shape[1]
        y32 = R.attention_ref(q.view(b, sq, h, hd), k.view(b, sk, h, hd), v.view(b, sk, h, hd), c["scale"])
        # tolerance of the reference's own test: tests/test_attention.py:94
        assert (y32.reshape(b, sq, -1).float() - c["y"].float()).abs().max() <= 1.8e-2


def test_sparse_attention_all_ones_equals_dense():
    # the only masked case the reference tests: tests/test_sparge_attention.py:81 (mask = ones)
    c = golden("attention.pt")[2]
    q, k, v, h, hd = c["q"], c["k"], c["v"], c["h"], c["hd"]
    b, sq, sk = q.shape[0], q.shape[1], k.shape[1]
    mask = torch.ones(b, h, -(-sq // 128), -(-sk // 64), dtype=torch.int8)
    y = R.sparse_scaled_dot_product_attention(q, k, v, h, h, hd, scale=c["scale"], sparse_mask=mask)
    assert (y.float() - c["y"].float()).abs().max() <= 1.8e-2


def test_flux_blocks_golden():
    for tag, quant in (("fp8", torch.float8_e4m3fn), ("int8", torch.int8)):
        c = golden(f"block_flux_{tag}.pt")
        blk = B.FluxTransformerBlockRef(c["sd_double"], "transformer_blocks.0", c["heads"], c["hd"], quant)
        enc, hid = blk.forward(c["img"], c["txt"], c["temb"], c["rope"])
        assert eq(enc, c["enc_out"]) and eq(hid, c["hid_out"])
        sblk = B.FluxSingleTransformerBlockRef(c["sd_single"], "single_transformer_blocks.0", c["heads"], c["hd"], quant)
        assert eq(sblk.forward(c["single_in"], c["temb"], c["rope"]), c["single_out"])


def test_wan_block_golden():
    for tag, quant in (("fp8", torch.float8_e4m3fn), ("int8", torch.int8)):
        c = golden(f"block_wan_{tag}.pt")
        blk = B.WanTransformerBlockRef(c["sd"], "blocks.0", c["heads"], c["hd"], quant)
        y = blk.forward(c["x"], c["enc"], c["temb"], (c["cos"], c["sin"]))
        assert eq(y, c["y"])


def test_qwen_block_golden():
    for tag, quant in (("int8", torch.int8), ("fp8", torch.float8_e4m3fn)):
        c = golden(f"block_qwen_{tag}.pt")
        blk = B.QwenImageTransformerBlockRef(c["sd"], "transformer_blocks.0", c["heads"], c["hd"], quant)
        enc, hid = blk.forward(c["img"], c["txt"], c["temb"], c["rope"])
        assert eq(enc, c["enc_out"]) and eq(hid, c["hid_out"])


def test_sd3_blocks_golden():
    c = golden("block_sd3_fp8.pt")
    for name, blk in c["blocks"].items():
        sd = B.sd3_block_state_dict("transformer_blocks.0", c["dim"], c["hd"], blk["seed"], blk["context_pre_only"], blk["dual"])
        ref = B.JointTransformerBlockRef(sd, "transformer_blocks.0", c["heads"], c["hd"], torch.float8_e4m3fn,
                                         blk["context_pre_only"], blk["dual"])
        enc, hid = ref.forward(c["img"], c["txt"], c["temb"])
        assert eq(hid, blk["hid_out"]), name
        assert (enc is None and blk["enc_out"] is None) or eq(enc, blk["enc_out"]), name
