"""ORACLE -- TEST INFRASTRUCTURE ONLY. Whole-model golden vectors from the REAL reference classes.

    python -m oracle.gen_golden_models          (in the build container; needs /root/reference)

Runs the reference's own `FluxTransformer2DModelCore.forward` (fastdm/model/flux.py:334-494) and
`WanTransformer3DModelCore.forward` (fastdm/model/wan.py:283-380) on CPU -- reduced width and depth, FP8 W8A8, weights
loaded through the reference's own `weight_loading(dict)` (fastdm/model/basemodel.py:88-143) -- on seeded inputs and
stores inputs + outputs under tests/golden/model_*.pt. The weights are not stored: tests rebuild them from the seed with
oracle.blocks_ref.{flux,wan}_model_state_dict. tests/test_gpu_models.py pins fastdm_b200/models.py to these fixtures:
embedders, RoPE tables, AdaLN (table), every block, norm_out / proj_out, unpatchify.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import blocks_ref as B  # noqa: E402
from oracle import reference_shim  # noqa: E402

BF = torch.bfloat16
FLUX_CFG = dict(num_layers=2, num_single_layers=2, attention_head_dim=128, num_attention_heads=2, in_channels=16, out_channels=16,
                joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
WAN_CFG = dict(patch_size=(1, 2, 2), num_attention_heads=2, attention_head_dim=128, in_channels=4, out_channels=4, text_dim=64,
               freq_dim=256, ffn_dim=512, num_layers=2, cross_attn_norm=True)


def save(name, obj):
    path = os.path.join(GOLDEN, name)
    torch.save(obj, path)
    print(f"  wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    reference_shim.load()
    torch.set_num_threads(os.cpu_count() or 1)
    quant = torch.float8_e4m3fn

    # ---- FLUX: 96 image tokens on an 8 x 12 grid + 32 text tokens
    from fastdm.model.flux import FluxTransformer2DModelCore
    g = torch.Generator().manual_seed(101)
    n_img, n_txt = 96, 32
    lat = torch.randn(1, n_img, FLUX_CFG["in_channels"], generator=g).to(BF)
    prompt = torch.randn(1, n_txt, FLUX_CFG["joint_attention_dim"], generator=g).to(BF)
    pooled = torch.randn(1, FLUX_CFG["pooled_projection_dim"], generator=g).to(BF)
    timestep = torch.tensor([0.7]).to(BF)
    guidance = torch.tensor([3.5]).to(BF)
    img_ids = torch.zeros(n_img, 3)
    img_ids[:, 1] = torch.arange(8).repeat_interleave(12).float()
    img_ids[:, 2] = torch.arange(12).repeat(8).float()
    txt_ids = torch.zeros(n_txt, 3)
    sd = B.flux_model_state_dict(FLUX_CFG, seed=7)
    m = FluxTransformer2DModelCore(**FLUX_CFG, data_type=BF, quant_dtype=quant)
    m.weight_loading(dict(sd), data_type=BF, device_type="cpu")
    with torch.no_grad():
        y = m.forward(lat, prompt, pooled, timestep, img_ids, txt_ids, guidance)[0]
    print("  flux out", tuple(y.shape), float(y.float().abs().mean()))
    save("model_flux_fp8.pt", dict(cfg=FLUX_CFG, seed=7, latent=lat, prompt=prompt, pooled=pooled, timestep=timestep,
                                   guidance=guidance, img_ids=img_ids, txt_ids=txt_ids, y=y))

    # ---- Wan (t2v): latent [1, 4, 3, 8, 12] -> 3 x 4 x 6 = 72 tokens, 24 text tokens
    from fastdm.model.wan import WanTransformer3DModelCore
    g = torch.Generator().manual_seed(102)
    lat = torch.randn(1, WAN_CFG["in_channels"], 3, 8, 12, generator=g).to(BF)
    prompt = torch.randn(1, 24, WAN_CFG["text_dim"], generator=g).to(BF)
    timestep = torch.tensor([999], dtype=torch.int64)
    sd = B.wan_model_state_dict(WAN_CFG, seed=8)
    m = WanTransformer3DModelCore(**WAN_CFG, data_type=BF, quant_dtype=quant)
    m.weight_loading(dict(sd), data_type=BF, device_type="cpu")
    with torch.no_grad():
        y = m.forward(lat, timestep, prompt)[0]
    print("  wan out", tuple(y.shape), float(y.float().abs().mean()))
    save("model_wan_fp8.pt", dict(cfg=WAN_CFG, seed=8, latent=lat, prompt=prompt, timestep=timestep, y=y))


if __name__ == "__main__":
    main()
