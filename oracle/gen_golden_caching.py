"""ORACLE -- TEST INFRASTRUCTURE ONLY. Step-cache golden vectors from the REAL reference.

    python -m oracle.gen_golden_caching          (in the build container; needs /root/reference)

The reference's own TeaCache / FBCache / DiCache (fastdm/caching/xcaching.py) drive the reference's own block classes
(taken from its FluxTransformer2DModelCore / WanTransformer3DModelCore after `weight_loading`) through `apply_cache` for
a sequence of denoise steps on a synthetic, slowly drifting trajectory. Stored per case: which steps computed the block
stack and every step's output. Inputs and weights are not stored: `case_inputs` below regenerates them from seeds (the
tests import it), the weights come from oracle.blocks_ref.{flux,wan}_model_state_dict.

tests/test_caching_host.py (CPU) runs fastdm_b200/caching.py over the oracle blocks on the same inputs: identical skip
decisions and bit-identical outputs. tests/test_gpu_caching.py runs it over the CUDA blocks: identical decisions,
outputs within the block tolerance.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import blocks_ref as B  # noqa: E402
from oracle.gen_golden_models import FLUX_CFG, WAN_CFG  # noqa: E402

BF = torch.bfloat16
STEPS = 10
# thresholds are chosen so that, on this trajectory, every decision clears its threshold by a margin (printed below)
CASES = {
    "flux_teacache": dict(model="flux", cfg=dict(cache_algorithm="teacache", enable_caching=True, threshold=0.3,
                                                  coefficients=[4.98651651e+02, -2.83781631e+02, 5.58554382e+01,
                                                                -3.82021401e+00, 2.64230861e-01])),   # examples/xcaching/configs/teacache_flux.json
    "flux_fbcache": dict(model="flux", cfg=dict(cache_algorithm="fbcache", enable_caching=True, threshold=0.06, warmup_steps=2)),
    "flux_dicache": dict(model="flux", cfg=dict(cache_algorithm="dicache", enable_caching=True, threshold=0.05, probe_depth=1,
                                                 ret_ratio=0.2, rel_l1_distance_algo="delta_y")),
    "flux_dicache_minus": dict(model="flux", cfg=dict(cache_algorithm="dicache", enable_caching=True, threshold=0.06, probe_depth=1,
                                                       ret_ratio=0.2, rel_l1_distance_algo="delta_minus")),
    "wan_fbcache": dict(model="wan", cfg=dict(cache_algorithm="fbcache", enable_caching=True, threshold=0.05, warmup_steps=2,
                                               negtive_cache=True)),
}


def case_inputs(model: str, step: int):
    """Synthetic apply_cache inputs of denoise step `step` (deterministic CPU ops): a drifting latent, a drifting
    conditioning vector, fixed text states and RoPE table."""
    g = torch.Generator().manual_seed(4242 if model == "flux" else 4343)
    if model == "flux":
        d, hd, n_img, n_txt = 256, 128, 96, 32
        base, drift = torch.randn(1, n_img, d, generator=g), torch.randn(1, n_img, d, generator=g)
        enc = torch.randn(1, n_txt, d, generator=g).to(BF)
        t0, t1 = torch.randn(1, d, generator=g), torch.randn(1, d, generator=g)
        rope = torch.rand(n_txt + n_img, hd, generator=g).to(BF)
    else:
        d, hd, n_img, n_txt = 256, 128, 72, 24
        base, drift = torch.randn(1, n_img, d, generator=g), torch.randn(1, n_img, d, generator=g)
        enc = torch.randn(1, n_txt, d, generator=g).to(BF)
        t0, t1 = torch.randn(1, 6, d, generator=g) * 0.5, torch.randn(1, 6, d, generator=g) * 0.5
        rope = (torch.rand(1, n_img, 1, hd, generator=g), torch.rand(1, n_img, 1, hd, generator=g))
    # the drift accelerates, so early steps get skipped and later ones do not
    s = 0.004 * step * (1 + 0.5 * step)
    hidden = (base + s * drift).to(BF)
    temb = (t0 + 0.5 * s * t1).to(BF)
    return hidden, enc, temb, rope


def reference_blocks(model: str):
    from oracle import reference_shim
    reference_shim.load()
    quant = torch.float8_e4m3fn
    if model == "flux":
        from fastdm.model.flux import FluxTransformer2DModelCore
        m = FluxTransformer2DModelCore(**FLUX_CFG, data_type=BF, quant_dtype=quant)
        m.weight_loading(dict(B.flux_model_state_dict(FLUX_CFG, seed=7)), data_type=BF, device_type="cpu")
        return m.transformer_blocks, m.single_transformer_blocks
    from fastdm.model.wan import WanTransformer3DModelCore
    m = WanTransformer3DModelCore(**WAN_CFG, data_type=BF, quant_dtype=quant)
    m.weight_loading(dict(B.wan_model_state_dict(WAN_CFG, seed=8)), data_type=BF, device_type="cpu")
    return m.blocks, None


def main():
    import contextlib
    import io
    from oracle import reference_shim
    reference_shim.load()
    from fastdm.caching.xcaching import AutoCache
    torch.set_num_threads(os.cpu_count() or 1)
    out = {}
    blocks_of = {}
    for name, case in CASES.items():
        model = case["model"]
        if model not in blocks_of:
            with contextlib.redirect_stdout(io.StringIO()):
                blocks_of[model] = reference_blocks(model)
        blocks, singles = blocks_of[model]
        step_box = [0]
        cfg = dict(case["cfg"], current_steps_callback=lambda: step_box[0], total_steps_callback=lambda: STEPS)
        cache = AutoCache.from_dict(cfg)
        last = (singles or blocks)[-1]
        ran = [False]
        orig_forward = last.forward

        def spy(*a, _f=orig_forward, **k):
            ran[0] = True
            return _f(*a, **k)

        last.forward = spy
        decisions, outputs, accs = [], [], []
        forwards_per_step = 2 if cfg.get("negtive_cache") else 1     # cond / uncond alternate inside one step
        for step in range(STEPS):
            step_box[0] = step
            for branch in range(forwards_per_step):
                hidden, enc, temb, rope = case_inputs(model, step)
                if branch == 1:
                    hidden = (hidden.float() * 0.9).to(BF)               # the "negative prompt" forward sees other activations
                ran[0] = False
                with contextlib.redirect_stdout(io.StringIO()):
                    y = cache.apply_cache(model_type=model, hidden_states=hidden.clone(), encoder_hidden_states=enc, temb=temb,
                                          image_rotary_emb=rope, transformer_blocks=blocks, single_transformer_blocks=singles,
                                          controlnet_single_block_samples=None)
                decisions.append(bool(ran[0]))
                outputs.append(y.clone())
                accs.append(float(cache.accumulated_rel_l1_distance_dict["positive" if branch == 0 else "negative"]))
        last.forward = orig_forward
        print(f"  {name}: computed {sum(decisions)}/{len(decisions)}  decisions {''.join('C' if d else 's' for d in decisions)}")
        print("     accumulated distance after each forward:", " ".join(f"{a:.4f}" for a in accs), " threshold", cfg["threshold"])
        out[name] = dict(model=model, cfg=case["cfg"], steps=STEPS, forwards_per_step=forwards_per_step, decisions=decisions,
                         outputs=torch.stack(outputs))
    path = os.path.join(GOLDEN, "caching.pt")
    torch.save(out, path)
    print(f"  wrote caching.pt: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
