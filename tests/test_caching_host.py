"""Host logic of fastdm_b200/caching.py (TeaCache / FBCache / DiCache) against fixtures produced by the REFERENCE's own
cache classes driving the reference's own blocks (oracle/gen_golden_caching.py): same skip decisions, bit-identical
outputs (residual replay, DiCache's extrapolation, the cond / uncond alternation of `negtive_cache`).

CPU test: the policies run over the oracle blocks (bit-identical to the reference's blocks, tests/test_oracle_golden.py)
with the relative-L1 sums evaluated by a torch expression instead of the CUDA reduction; tests/test_gpu_caching.py
repeats it over the CUDA blocks with the device reduction."""
import pytest
import torch

from conftest import golden
from oracle import blocks_ref as B
from oracle.gen_golden_caching import case_inputs
from oracle.gen_golden_models import FLUX_CFG, WAN_CFG

BF = torch.bfloat16
QUANT = torch.float8_e4m3fn


def torch_sums(a, b):
    # what the device kernel computes: |T(a - b)| summed in fp32, |b| summed in fp32
    return float((a - b).abs().float().sum()), float(b.abs().float().sum())


class _FluxDouble:
    def __init__(self, sd, i):
        self.ref = B.FluxTransformerBlockRef(sd, f"transformer_blocks.{i}", FLUX_CFG["num_attention_heads"],
                                             FLUX_CFG["attention_head_dim"], QUANT)

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None):
        return self.ref.forward(hidden_states, encoder_hidden_states, temb, image_rotary_emb)

    def cache_indicator(self, hidden_states, encoder_hidden_states, temb):
        return B._ada_ln(hidden_states, temb, self.ref.norm1, 6)[0]


class _FluxSingle:
    def __init__(self, sd, i):
        self.ref = B.FluxSingleTransformerBlockRef(sd, f"single_transformer_blocks.{i}", FLUX_CFG["num_attention_heads"],
                                                   FLUX_CFG["attention_head_dim"], QUANT)

    def forward(self, hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None):
        return self.ref.forward(hidden_states, temb, image_rotary_emb)


class _Wan:
    def __init__(self, sd, i):
        self.ref = B.WanTransformerBlockRef(sd, f"blocks.{i}", WAN_CFG["num_attention_heads"], WAN_CFG["attention_head_dim"], QUANT)

    def forward(self, hidden, encoder, temb, rotary_emb, sparse_mask=None):
        return self.ref.forward(hidden, encoder, temb, rotary_emb, sparse_mask)


def oracle_blocks(model):
    if model == "flux":
        sd = B.flux_model_state_dict(FLUX_CFG, seed=7)
        return ([_FluxDouble(sd, i) for i in range(FLUX_CFG["num_layers"])],
                [_FluxSingle(sd, i) for i in range(FLUX_CFG["num_single_layers"])])
    sd = B.wan_model_state_dict(WAN_CFG, seed=8)
    return [_Wan(sd, i) for i in range(WAN_CFG["num_layers"])], None


def run_case(case, blocks, singles, distance_sums, move=lambda t: t):
    from fastdm_b200.caching import AutoCache

    step_box = [0]
    cfg = dict(case["cfg"], current_steps_callback=lambda: step_box[0], total_steps_callback=lambda: case["steps"])
    cache = AutoCache.from_dict(cfg, distance_sums=distance_sums)
    outs = []
    for step in range(case["steps"]):
        step_box[0] = step
        for branch in range(case["forwards_per_step"]):
            hidden, enc, temb, rope = case_inputs(case["model"], step)
            if branch == 1:
                hidden = (hidden.float() * 0.9).to(BF)
            rope = tuple(move(r) for r in rope) if isinstance(rope, tuple) else move(rope)
            y = cache.apply_cache(model_type=case["model"], hidden_states=move(hidden.clone()), encoder_hidden_states=move(enc),
                                  temb=move(temb), image_rotary_emb=rope, transformer_blocks=blocks,
                                  single_transformer_blocks=singles)
            outs.append(y)
    return cache, outs


@pytest.mark.parametrize("name", ["flux_teacache", "flux_fbcache", "flux_dicache", "flux_dicache_minus", "wan_fbcache"])
def test_cache_policy_matches_reference(name):
    case = golden("caching.pt")[name]
    blocks, singles = oracle_blocks(case["model"])
    cache, outs = run_case(case, blocks, singles, torch_sums)
    assert cache.decisions == case["decisions"], (cache.decisions, case["decisions"])
    for i, y in enumerate(outs):
        assert torch.equal(y, case["outputs"][i]), f"forward {i} differs from the reference's cached output"
    thr = case["cfg"]["threshold"]
    margin = min(abs(d - thr) / thr for d in cache.distances)
    print(f"{name}: decisions {''.join('C' if d else 's' for d in cache.decisions)}, closest decision {margin:.1%} from the threshold")
    assert 0 < sum(cache.decisions) < len(cache.decisions)      # the fixture exercises both branches
    # the GPU run repeats these decisions with CUDA blocks: keep them away from the threshold
    assert margin > 0.05, f"fixture decision only {margin:.1%} from the threshold"


def test_config_and_registry():
    from fastdm_b200.caching import AutoCache, CacheConfig, DiCache, FBCache, TeaCache, _round_like

    assert isinstance(AutoCache.from_dict(dict(cache_algorithm="TeaCache", coefficients=[1.0, 0.0])), TeaCache)
    assert isinstance(AutoCache.from_dict(dict(cache_algorithm="fbcache", warmup_steps=3, unknown_key=1)), FBCache)
    assert isinstance(AutoCache.from_dict(dict(cache_algorithm="dicache", probe_depth=2)), DiCache)
    with pytest.raises(ValueError):
        AutoCache.from_dict(dict(cache_algorithm="nope"))
    with pytest.raises(ValueError):
        CacheConfig.from_dict({})
    # cond / uncond alternation (xcaching.py:64-75)
    c = AutoCache.from_dict(dict(cache_algorithm="fbcache", negtive_cache=True))
    assert [c.get_cache_key() for _ in range(4)] == ["positive", "negative", "positive", "negative"]
    # rounding helper == torch's rounding to the tensor dtype
    for x in (0.2, 0.123456, 1e-3, 3.3e4, 0.0):
        assert _round_like(x, torch.bfloat16) == float(torch.tensor(x, dtype=torch.float32).to(torch.bfloat16))
        assert _round_like(x, torch.float16) == float(torch.tensor(x, dtype=torch.float32).to(torch.float16))
