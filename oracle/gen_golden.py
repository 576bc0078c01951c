"""ORACLE -- TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.pt by running the REAL reference.

    python -m oracle.gen_golden          (in the build container; needs /root/reference)

The reference (FastDM) is imported on CPU through oracle/reference_shim.py and its *own* torch
backend / layer / block classes are executed on seeded inputs; inputs and outputs are stored as
small fixtures that travel to the GPU box (which has no /root/reference). The script also checks
that oracle/ops_ref.py and oracle/blocks_ref.py (the restatement) reproduce every fixture -- that
check is repeated by tests/test_oracle_golden.py.

Shapes are reduced versions of the reference's test tables (tests/test_quant.py:5-50,
tests/test_matmul.py:5-44, tests/test_attention.py:7-21, tests/test_rmsnorm.py:5-14,
tests/test_rope.py:5-7, tests/test_gelu_and_mul.py:5-8) so every file stays well under 1 MB.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import blocks_ref as B  # noqa: E402
from oracle import ops_ref as R  # noqa: E402
from oracle import reference_shim  # noqa: E402

BF = torch.bfloat16


def gen(seed):
    return torch.Generator().manual_seed(seed)


def save(name, obj):
    path = os.path.join(GOLDEN, name)
    torch.save(obj, path)
    print(f"  wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if a.dtype == torch.float8_e4m3fn:
        return torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    return torch.equal(a, b)


def quant_inputs():
    g = gen(1)
    xs = []
    x = torch.randn(37, 320, generator=g) * 3
    x[5] = 0.0                       # all-zero row: fp8 scale floor (clamp(min=1e-12))
    x[6] = -torch.rand(320, generator=g) - 0.5   # all-negative row (reference CUDA max_val bug, elmwise_ops.cu:355)
    x[7] = torch.rand(320, generator=g) * 1e-3   # tiny row (CUDA kernel's 1/(448*512) floor differs from torch)
    x[8, 3] = 1e4                    # outlier
    xs.append(x.to(BF))
    xs.append((torch.randn(14, 3072, generator=g) * 0.7).to(BF))
    xs.append((torch.randn(3, 15360, generator=g) * 2).to(BF))
    xs.append(torch.randn(2, 1536, generator=g).to(BF))
    return xs


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    reference_shim.load()
    T = reference_shim.torch_backend
    ok = True

    # ---------------- quantisation ----------------
    cases = []
    for x in quant_inputs():
        q8, s8 = T("quantize_to_fp8")(x)
        qs, ss, _ = T("quantize_to_int8")(x, True)
        # asym: constant rows give scale 0 -> NaN in the reference; the all-zero row is excluded
        xa = x.clone()
        if xa.shape == (37, 320):
            xa[5, 0] = 1.0
        qa, sa, za = T("quantize_to_int8")(xa, False)
        cases.append(dict(x=x, fp8_q=q8.view(torch.uint8), fp8_s=s8, s8_q=qs, s8_s=ss,
                          x_asym=xa, a8_q=qa, a8_s=sa, a8_zp=za))
        r8, rs8 = R.quantize_to_fp8(x)
        rq, rs, _ = R.quantize_to_int8(x, True)
        ra, rsa, rza = R.quantize_to_int8(xa, False)
        ok &= same(q8, r8) and same(s8, rs8) and same(qs, rq) and same(ss, rs) and same(qa, ra) \
            and same(sa, rsa) and same(za, rza)
    save("quant.pt", cases)

    # ---------------- rms_norm ----------------
    g = gen(2)
    cases = []
    for shape, wshape in (((2, 19, 24, 128), 128), ((2, 5, 24, 64), 64), ((1, 7, 5120), 5120), ((1, 14, 3584), 3584),
                          ((3, 40), 40)):
        x = torch.randn(*shape, generator=g).to(BF)
        w = torch.randn(wshape, generator=g).to(BF)
        y = T("rmsnorm")(x, w, 1e-6)
        cases.append(dict(x=x, w=w, eps=1e-6, y=y))
        ok &= same(y, R.rms_norm(x, w, 1e-6))
    save("rmsnorm.pt", cases)

    # ---------------- rope ----------------
    g = gen(3)
    cases = []
    for (b, s, hq, hk, hd, neox) in ((2, 33, 3, 3, 128, False), (1, 20, 4, 2, 64, False), (1, 17, 2, 2, 128, True)):
        q = torch.randn(b, s, hq * hd, generator=g).to(BF)
        k = torch.randn(b, s, hk * hd, generator=g).to(BF)
        cs = torch.rand(s + 3, hd, generator=g).to(BF)
        q2, k2 = q.clone(), k.clone()
        T("rotembd")(q2, k2, hd, cs, neox)
        cases.append(dict(q=q, k=k, cs=cs, hd=hd, neox=neox, q_out=q2, k_out=k2))
        q3, k3 = q.clone(), k.clone()
        R.rotary_pos_embedding(q3, k3, hd, cs, neox)
        ok &= same(q2, q3) and same(k2, k3)
    save("rope.pt", cases)

    # ---------------- gelu_and_mul ----------------
    g = gen(4)
    cases = []
    for shape in ((9, 640), (2, 5, 2560)):
        x = (torch.randn(*shape, generator=g) * 2).to(BF)
        y = T("gelu_and_mul")(x)
        cases.append(dict(x=x, y=y))
        ok &= same(y, R.gelu_and_mul(x))
    save("gelu_and_mul.pt", cases)

    # ---------------- matmuls (input distributions of tests/test_matmul.py:49-57,98-104) ----------
    g = gen(5)
    cases = []
    for (M, K, N) in ((70, 256, 144), (2, 320, 1280), (130, 768, 64), (14, 1024, 272)):
        a8 = torch.randint(-128, 128, (M, K), generator=g).to(torch.int8)
        b8 = torch.randint(-128, 128, (K, N), generator=g).to(torch.int8).t().contiguous().t()
        af = torch.randn(M, K, generator=g).to(torch.float8_e4m3fn)
        bf = torch.randn(K, N, generator=g).to(torch.float8_e4m3fn).t().contiguous().t()
        sa = torch.randn(M, 1, generator=g)
        sb = torch.randn(N, 1, generator=g)
        adj = torch.randint(-128, 127, (1, N), generator=g).to(torch.int32)
        azp = torch.randint(-128, 127, (M, 1), generator=g).to(torch.int32)
        bias = torch.randn(N, generator=g).to(BF)
        y8 = T("int8_matmul")(a8, b8, sa, sb, BF, adj, azp, bias)
        y8nb = T("int8_matmul")(a8, b8, sa, sb, BF, adj, azp, None)
        yf = T("fp8_matmul")(af, bf, sa, sb, BF, bias)
        yfnb = T("fp8_matmul")(af, bf, sa, sb, BF, None)
        cases.append(dict(a8=a8, b8_t=b8.t().contiguous(), af=af.view(torch.uint8),
                          bf_t=bf.t().contiguous().view(torch.uint8), sa=sa, sb=sb, adj=adj, azp=azp, bias=bias,
                          y_int8=y8, y_int8_nobias=y8nb, y_fp8=yf, y_fp8_nobias=yfnb))
        ok &= same(y8, R.int8_matmul(a8, b8, sa, sb, BF, adj, azp, bias))
        ok &= same(y8nb, R.int8_matmul(a8, b8, sa, sb, BF, adj, azp, None))
        ok &= same(yf, R.fp8_matmul(af, bf, sa, sb, BF, bias))
        ok &= same(yfnb, R.fp8_matmul(af, bf, sa, sb, BF, None))
    save("matmul.pt", cases)

    # ---------------- attention (torch backend sdpa; tolerance-checked, not bit-exact) -------------
    cases = []
    for (b, sq, sk, h, hd) in ((1, 200, 200, 3, 128), (2, 77, 50, 2, 64), (1, 130, 257, 2, 128), (2, 64, 77, 4, 64)):
        torch.manual_seed(0)  # tests/test_attention.py:69
        q = torch.randn(b, sq, h * hd).to(BF)
        k = torch.randn(b, sk, h * hd).to(BF)
        v = torch.randn(b, sk, h * hd).to(BF)
        scale = 1.0 / hd ** 0.5
        y = T("sdpa")(q, k, v, h, h, hd, scale=scale)
        cases.append(dict(q=q, k=k, v=v, h=h, hd=hd, scale=scale, y=y))
        y2 = R.scaled_dot_product_attention(q, k, v, h, h, hd, scale=scale)
        y3 = R.attention_ref(q.view(b, sq, h, hd), k.view(b, sk, h, hd), v.view(b, sk, h, hd), scale).reshape(b, sq, -1)
        ok &= same(y, y2)
        ok &= bool((y.float() - y3.float()).abs().max() <= 1.8e-2)  # tests/test_attention.py:94
    save("attention.pt", cases)

    # ---------------- FLUX block pair + Wan block at reduced width (reference classes) -------------
    from fastdm.model.basemodel import BaseModelCore
    from fastdm.model.flux import FluxSingleTransformerBlock, FluxTransformerBlock
    from fastdm.model.wan import WanTransformerBlock

    def loader(sd):
        core = BaseModelCore()
        core.origin_tensor_dict = sd
        core.unmatched_tensors = list(sd.keys())
        core.device = "cpu"
        return core

    dim, heads, hd = 128, 2, 64
    g = gen(7)
    img = torch.randn(1, 96, dim, generator=g).to(BF)
    txt = torch.randn(1, 32, dim, generator=g).to(BF)
    temb = torch.randn(1, dim, generator=g).to(BF)
    rope = torch.rand(128, hd, generator=g).to(BF)
    for quant, tag in ((torch.float8_e4m3fn, "fp8"), (torch.int8, "int8")):
        # --- double block: loading calls mirror fastdm/model/flux.py:283-308
        sd = B.flux_double_state_dict("transformer_blocks.0", dim, hd, seed=11)
        blk = FluxTransformerBlock(dim, heads, hd)
        c = loader(dict(sd))
        p = "transformer_blocks.0"
        c.init_weight([f"{p}.norm1.linear"], blk.norm1.linear)
        c.init_weight([f"{p}.norm1_context.linear"], blk.norm1_context.linear)
        blk.attn.norm_q_weight = c.init_weight([f"{p}.attn.norm_q.weight"])
        blk.attn.norm_k_weight = c.init_weight([f"{p}.attn.norm_k.weight"])
        c.init_weight([f"{p}.attn.to_q", f"{p}.attn.to_k", f"{p}.attn.to_v"], blk.attn.qkv, quant)
        c.init_weight([f"{p}.attn.add_q_proj", f"{p}.attn.add_k_proj", f"{p}.attn.add_v_proj"], blk.attn.add_qkv_proj, quant)
        c.init_weight([f"{p}.attn.to_out.0"], blk.attn.to_out, quant)
        c.init_weight([f"{p}.attn.to_add_out"], blk.attn.to_add_out, quant)
        blk.attn.norm_added_q_weight = c.init_weight([f"{p}.attn.norm_added_q.weight"])
        blk.attn.norm_added_k_weight = c.init_weight([f"{p}.attn.norm_added_k.weight"])
        c.init_weight([f"{p}.ff.net.0.proj"], blk.ff.act_fn.proj, quant)
        c.init_weight([f"{p}.ff.net.2"], blk.ff.ff_out_proj, quant)
        c.init_weight([f"{p}.ff_context.net.0.proj"], blk.ff_context.act_fn.proj, quant)
        c.init_weight([f"{p}.ff_context.net.2"], blk.ff_context.ff_out_proj, quant)
        assert not c.unmatched_tensors
        enc_o, hid_o = blk.forward(img, txt, temb, image_rotary_emb=rope)
        rb = B.FluxTransformerBlockRef(sd, p, heads, hd, quant)
        enc_r, hid_r = rb.forward(img, txt, temb, rope)
        ok &= same(enc_o, enc_r) and same(hid_o, hid_r)

        # --- single block: fastdm/model/flux.py:310-322
        sd1 = B.flux_single_state_dict("single_transformer_blocks.0", dim, hd, seed=12)
        sblk = FluxSingleTransformerBlock(dim, heads, hd)
        c = loader(dict(sd1))
        p1 = "single_transformer_blocks.0"
        c.init_weight([f"{p1}.norm.linear"], sblk.norm.linear)
        c.init_weight([f"{p1}.proj_mlp"], sblk.proj_mlp, quant)
        c.init_weight([f"{p1}.proj_out"], sblk.proj_out, quant)
        sblk.attn.norm_q_weight = c.init_weight([f"{p1}.attn.norm_q.weight"])
        sblk.attn.norm_k_weight = c.init_weight([f"{p1}.attn.norm_k.weight"])
        c.init_weight([f"{p1}.attn.to_q", f"{p1}.attn.to_k", f"{p1}.attn.to_v"], sblk.attn.qkv, quant)
        assert not c.unmatched_tensors
        cat = torch.cat([enc_o, hid_o], dim=1)
        single_o = sblk.forward(cat, temb, image_rotary_emb=rope)
        rs = B.FluxSingleTransformerBlockRef(sd1, p1, heads, hd, quant)
        ok &= same(single_o, rs.forward(cat, temb, rope))
        save(f"block_flux_{tag}.pt", dict(dim=dim, heads=heads, hd=hd, img=img, txt=txt, temb=temb, rope=rope,
                                          sd_double=sd, sd_single=sd1, enc_out=enc_o, hid_out=hid_o,
                                          single_in=cat, single_out=single_o))

    # --- SD3.5 blocks (dual attention / plain / context_pre_only): fastdm/model/sd35.py:296-326; batch 2 (CFG)
    from fastdm.model.sd35 import JointTransformerBlock
    dim, heads, hd = 128, 2, 64
    g = gen(10)
    img = torch.randn(2, 80, dim, generator=g).to(BF)
    txt = torch.randn(2, 27, dim, generator=g).to(BF)
    temb = torch.randn(2, dim, generator=g).to(BF)
    sd3 = {}
    for name, cpo, dual in (("dual", False, True), ("plain", False, False), ("last", True, False)):
        quant = torch.float8_e4m3fn
        sd = B.sd3_block_state_dict("transformer_blocks.0", dim, hd, seed=20 + len(sd3), context_pre_only=cpo,
                                    use_dual_attention=dual)
        blk = JointTransformerBlock(dim, heads, hd, context_pre_only=cpo, qk_norm="rms_norm", use_dual_attention=dual)
        c = loader(dict(sd))
        p = "transformer_blocks.0"
        c.init_weight([f"{p}.norm1.linear"], blk.norm1.linear)
        c.init_weight([f"{p}.norm1_context.linear"], blk.norm1_context.linear)
        blk.attn.norm_q_weight = c.init_weight([f"{p}.attn.norm_q.weight"])
        blk.attn.norm_k_weight = c.init_weight([f"{p}.attn.norm_k.weight"])
        c.init_weight([f"{p}.attn.to_q", f"{p}.attn.to_k", f"{p}.attn.to_v"], blk.attn.qkv, quant)
        c.init_weight([f"{p}.attn.add_q_proj", f"{p}.attn.add_k_proj", f"{p}.attn.add_v_proj"], blk.attn.add_qkv_proj, quant)
        c.init_weight([f"{p}.attn.to_out.0"], blk.attn.to_out, quant)
        if not cpo:
            c.init_weight([f"{p}.attn.to_add_out"], blk.attn.to_add_out, quant)
        blk.attn.norm_added_q_weight = c.init_weight([f"{p}.attn.norm_added_q.weight"])
        blk.attn.norm_added_k_weight = c.init_weight([f"{p}.attn.norm_added_k.weight"])
        if dual:
            blk.attn2.norm_q_weight = c.init_weight([f"{p}.attn2.norm_q.weight"])
            blk.attn2.norm_k_weight = c.init_weight([f"{p}.attn2.norm_k.weight"])
            c.init_weight([f"{p}.attn2.to_q", f"{p}.attn2.to_k", f"{p}.attn2.to_v"], blk.attn2.qkv, quant)
            c.init_weight([f"{p}.attn2.to_out.0"], blk.attn2.to_out, quant)
        c.init_weight([f"{p}.ff.net.0.proj"], blk.ff.act_fn.proj, quant)
        c.init_weight([f"{p}.ff.net.2"], blk.ff.ff_out_proj, quant)
        if not cpo:
            c.init_weight([f"{p}.ff_context.net.0.proj"], blk.ff_context.act_fn.proj, quant)
            c.init_weight([f"{p}.ff_context.net.2"], blk.ff_context.ff_out_proj, quant)
        assert not c.unmatched_tensors
        enc_o, hid_o = blk.forward(img, txt, temb)
        rb = B.JointTransformerBlockRef(sd, p, heads, hd, quant, cpo, dual)
        enc_r, hid_r = rb.forward(img, txt, temb)
        ok &= same(hid_o, hid_r) and (cpo or same(enc_o, enc_r))
        # weights are regenerated from the seed by oracle.blocks_ref.sd3_block_state_dict (keeps the fixture small)
        sd3[name] = dict(seed=20 + len(sd3), context_pre_only=cpo, dual=dual, enc_out=enc_o, hid_out=hid_o)
    save("block_sd3_fp8.pt", dict(dim=dim, heads=heads, hd=hd, img=img, txt=txt, temb=temb, blocks=sd3))

    # --- Qwen-Image block: fastdm/model/qwenimage.py:215-239 (INT8 is the reference default for this model)
    from fastdm.model.qwenimage import QwenImageTransformerBlock
    dim, heads, hd = 128, 2, 64
    g = gen(9)
    img = torch.randn(1, 96, dim, generator=g).to(BF)
    txt = torch.randn(1, 32, dim, generator=g).to(BF)
    temb = torch.randn(1, dim, generator=g).to(BF)
    rope = torch.rand(128, hd, generator=g).to(BF)
    for quant, tag in ((torch.int8, "int8"), (torch.float8_e4m3fn, "fp8")):
        sd = B.qwen_block_state_dict("transformer_blocks.0", dim, hd, seed=14)
        blk = QwenImageTransformerBlock(dim, heads, hd)
        c = loader(dict(sd))
        p = "transformer_blocks.0"
        c.init_weight([f"{p}.img_mod.1"], blk.img_mod_proj, None)
        c.init_weight([f"{p}.txt_mod.1"], blk.txt_mod_proj, None)
        blk.attn.norm_q_weight = c.init_weight([f"{p}.attn.norm_q.weight"])
        blk.attn.norm_k_weight = c.init_weight([f"{p}.attn.norm_k.weight"])
        c.init_weight([f"{p}.attn.to_q", f"{p}.attn.to_k", f"{p}.attn.to_v"], blk.attn.qkv, quant)
        c.init_weight([f"{p}.attn.add_q_proj", f"{p}.attn.add_k_proj", f"{p}.attn.add_v_proj"], blk.attn.add_qkv_proj, quant)
        c.init_weight([f"{p}.attn.to_out.0"], blk.attn.to_out, quant)
        c.init_weight([f"{p}.attn.to_add_out"], blk.attn.to_add_out, quant)
        blk.attn.norm_added_q_weight = c.init_weight([f"{p}.attn.norm_added_q.weight"])
        blk.attn.norm_added_k_weight = c.init_weight([f"{p}.attn.norm_added_k.weight"])
        c.init_weight([f"{p}.img_mlp.net.0.proj"], blk.img_mlp.act_fn.proj, quant)
        c.init_weight([f"{p}.img_mlp.net.2"], blk.img_mlp.ff_out_proj, quant)
        c.init_weight([f"{p}.txt_mlp.net.0.proj"], blk.txt_mlp.act_fn.proj, quant)
        c.init_weight([f"{p}.txt_mlp.net.2"], blk.txt_mlp.ff_out_proj, quant)
        assert not c.unmatched_tensors
        enc_o, hid_o = blk.forward(img, txt, None, temb, image_rotary_emb=rope)
        rb = B.QwenImageTransformerBlockRef(sd, p, heads, hd, quant)
        enc_r, hid_r = rb.forward(img, txt, temb, rope)
        ok &= same(enc_o, enc_r) and same(hid_o, hid_r)
        save(f"block_qwen_{tag}.pt", dict(dim=dim, heads=heads, hd=hd, img=img, txt=txt, temb=temb, rope=rope, sd=sd,
                                          enc_out=enc_o, hid_out=hid_o))

    # --- Wan block: fastdm/model/wan.py:249-281
    dim, heads, hd, ffn = 128, 2, 64, 384
    g = gen(8)
    x = torch.randn(1, 120, dim, generator=g).to(BF)
    enc = torch.randn(1, 24, dim, generator=g).to(BF)
    temb6 = (torch.randn(1, 6, dim, generator=g) * 0.5).to(BF)
    cos = torch.rand(1, 120, 1, hd, generator=g)
    sin = torch.rand(1, 120, 1, hd, generator=g)
    for quant, tag in ((torch.float8_e4m3fn, "fp8"), (torch.int8, "int8")):
        sd = B.wan_block_state_dict("blocks.0", dim, ffn, seed=13)
        blk = WanTransformerBlock(dim, ffn, heads, cross_attn_norm=True)
        c = loader(dict(sd))
        p = "blocks.0"
        blk.attn1.norm_q_weight = c.init_weight([f"{p}.attn1.norm_q.weight"])
        blk.attn1.norm_k_weight = c.init_weight([f"{p}.attn1.norm_k.weight"])
        c.init_weight([f"{p}.attn1.to_q", f"{p}.attn1.to_k", f"{p}.attn1.to_v"], blk.attn1.qkv, quant)
        c.init_weight([f"{p}.attn1.to_out.0"], blk.attn1.to_out, quant)
        blk.attn2.norm_q_weight = c.init_weight([f"{p}.attn2.norm_q.weight"])
        blk.attn2.norm_k_weight = c.init_weight([f"{p}.attn2.norm_k.weight"])
        c.init_weight([f"{p}.attn2.to_q"], blk.attn2.to_q, quant)
        c.init_weight([f"{p}.attn2.to_k", f"{p}.attn2.to_v"], blk.attn2.to_kv, quant)
        c.init_weight([f"{p}.attn2.to_out.0"], blk.attn2.to_out, quant)
        blk.norm2.weight = c.init_weight([f"{p}.norm2.weight"]).to(torch.float32)
        blk.norm2.bias = c.init_weight([f"{p}.norm2.bias"]).to(torch.float32)
        c.init_weight([f"{p}.ffn.net.0.proj"], blk.ffn.act_fn.proj, quant)
        c.init_weight([f"{p}.ffn.net.2"], blk.ffn.ff_out_proj, quant)
        blk.scale_shift_table = c.init_weight([f"{p}.scale_shift_table"])
        assert not c.unmatched_tensors
        y = blk.forward(x, enc, temb6, (cos, sin))
        rb = B.WanTransformerBlockRef(sd, p, heads, hd, quant)
        ok &= same(y, rb.forward(x, enc, temb6, (cos, sin)))
        save(f"block_wan_{tag}.pt", dict(dim=dim, heads=heads, hd=hd, ffn=ffn, x=x, enc=enc, temb=temb6, cos=cos,
                                         sin=sin, sd=sd, y=y))

    # ---- radial block masks (fastdm/sparse/xsparse.py:71-185, 238-260): the policy FastDM hands to sdpa_sparse
    from fastdm.sparse.config import RadialAttnConfig
    from fastdm.sparse.xsparse import RadialAttn, sparge_mask_convert as ref_convert
    from fastdm_b200.sparse import radial_block_mask, sparge_mask_convert
    cases = []
    for (frames, tpf, bs, decay) in ((6, 300, 64, 0.3), (5, 384, 128, 0.5), (9, 256, 64, 0.3), (16, 480, 64, 0.3)):
        ra = RadialAttn(RadialAttnConfig(sparse_algorithm="radial", block_size=bs, decay_factor=decay, model_type="wan"))
        ra.post_init(video_token_num=frames * tpf, num_frame=frames)
        s = frames * tpf // bs * bs
        RadialAttn._log_mask = None
        m = ra.gen_log_mask_shrinked(s, "cpu")
        conv = ref_convert(m, bs, "sm100") if m.shape[0] % 2 == 0 else None
        mine = radial_block_mask(frames, tpf, bs, decay, "wan", total_tokens=s)
        ok &= torch.equal(m, mine) and (conv is None or torch.equal(conv, sparge_mask_convert(mine, bs)))
        cases.append(dict(frames=frames, tpf=tpf, block=bs, decay=decay, mask=m, converted=conv))
    save("radial_mask.pt", cases)

    print("restatement reproduces every fixture:", ok)
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
