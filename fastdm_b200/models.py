"""Whole-transformer denoise steps assembled from the fused blocks -- host-side mirrors of

    FluxTransformer2DModelCore.forward   fastdm/model/flux.py:334-494
    WanTransformer3DModelCore.forward    fastdm/model/wan.py:283-380

Only the block stack is the hot path; the once-per-step pieces around it
(timestep / text embedders, RoPE tables, patch embedding, output norm + projection) are small
unquantised torch ops exactly as in the reference (fastdm/layer/embeddings.py), kept here so that a
*complete* denoise step can be measured. Weights come as a diffusers-named state dict; for
benchmarks `random_*_state_dict_block` streams random-init weights block by block so the bf16
master copy of a 12-14 B parameter model never has to exist at once.
"""
import math
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops
from .blocks import (AdaLNTable, FluxSingleTransformerBlock, FluxTransformerBlock, JointTransformerBlock,
                     QwenImageTransformerBlock, WanTransformerBlock)
from .layers import QLinear, load_linear


def _caching(model) -> bool:
    """`model.cache` (a fastdm_b200.caching.AutoCache, set by the caller like the reference's `cache=` constructor
    argument) is active: fastdm/model/flux.py:268-271."""
    cache = getattr(model, "cache", None)
    return cache is not None and cache.config.enable_caching


# ---- small embedders (fastdm/layer/embeddings.py) ------------------------------------------------
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    """embeddings.py:18-69."""
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = scale * timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class _MLP:
    """TimestepEmbedding (embeddings.py:400-410) / PixArtAlphaTextProjection (:118-147)."""

    def __init__(self, sd, prefix, act, device, dtype=torch.bfloat16):
        self.l1 = load_linear(sd, [f"{prefix}.linear_1"], None, device, dtype)
        self.l2 = load_linear(sd, [f"{prefix}.linear_2"], None, device, dtype)
        self.act = act

    def forward(self, x):
        h = self.l1.forward(x)
        h = F.silu(h) if self.act == "silu" else F.gelu(h, approximate="tanh")
        return self.l2.forward(h)


def rope_1d(dim, pos, theta=10000.0):
    """angles [len(pos), dim/2] of get_1d_rotary_pos_embed (embeddings.py:160-224)."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64, device=pos.device) / dim))
    return torch.outer(pos.to(torch.float64), freqs)


def flux_rope_table(ids: torch.Tensor, axes_dim=(16, 56, 56), dtype=torch.bfloat16):
    """FluxPosEmbed.forward (embeddings.py:527-549) followed by the cos||sin merge of flux.py:425-428:
    table[s] = [cos(angle_0..63) || sin(angle_0..63)]."""
    ang = torch.cat([rope_1d(axes_dim[i], ids[:, i].float()) for i in range(len(axes_dim))], dim=-1)
    return torch.cat([ang.cos(), ang.sin()], dim=-1).to(dtype).contiguous()


def wan_rope_table(frames, height, width, head_dim=128, dtype=torch.bfloat16, device="cuda"):
    """WanRotaryPosEmbed.forward (embeddings.py:859-923) + the merge of layer/transformer.py:497-498."""
    h_dim = w_dim = 2 * (head_dim // 6)
    t_dim = head_dim - h_dim - w_dim
    af = rope_1d(t_dim, torch.arange(frames, device=device))
    ah = rope_1d(h_dim, torch.arange(height, device=device))
    aw = rope_1d(w_dim, torch.arange(width, device=device))
    ang = torch.cat([af.view(frames, 1, 1, -1).expand(frames, height, width, -1),
                     ah.view(1, height, 1, -1).expand(frames, height, width, -1),
                     aw.view(1, 1, width, -1).expand(frames, height, width, -1)], dim=-1).reshape(frames * height * width, -1)
    return torch.cat([ang.cos(), ang.sin()], dim=-1).to(dtype).contiguous()


# ---- random-init weights with diffusers names -----------------------------------------------------
def _rand_linear(sd, name, out_f, in_f, g, device, std=0.02, dtype=torch.bfloat16, bias=True):
    sd[f"{name}.weight"] = (torch.randn(out_f, in_f, generator=g, device=device, dtype=torch.float32) * std).to(dtype)
    if bias:
        sd[f"{name}.bias"] = (torch.randn(out_f, generator=g, device=device, dtype=torch.float32) * std).to(dtype)


def random_flux_double_sd(prefix, dim, head_dim, g, device):
    sd = {}
    _rand_linear(sd, f"{prefix}.norm1.linear", 6 * dim, dim, g, device)
    _rand_linear(sd, f"{prefix}.norm1_context.linear", 6 * dim, dim, g, device)
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        _rand_linear(sd, f"{prefix}.attn.{n}", dim, dim, g, device)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        sd[f"{prefix}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g, device=device)).to(torch.bfloat16)
    for ff in ("ff", "ff_context"):
        _rand_linear(sd, f"{prefix}.{ff}.net.0.proj", 4 * dim, dim, g, device)
        _rand_linear(sd, f"{prefix}.{ff}.net.2", dim, 4 * dim, g, device)
    return sd


def random_flux_single_sd(prefix, dim, head_dim, g, device):
    sd = {}
    _rand_linear(sd, f"{prefix}.norm.linear", 3 * dim, dim, g, device)
    _rand_linear(sd, f"{prefix}.proj_mlp", 4 * dim, dim, g, device)
    _rand_linear(sd, f"{prefix}.proj_out", dim, 5 * dim, g, device)
    for n in ("to_q", "to_k", "to_v"):
        _rand_linear(sd, f"{prefix}.attn.{n}", dim, dim, g, device)
    for n in ("norm_q", "norm_k"):
        sd[f"{prefix}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g, device=device)).to(torch.bfloat16)
    return sd


def random_wan_block_sd(prefix, dim, ffn_dim, g, device):
    sd = {}
    for a in ("attn1", "attn2"):
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            _rand_linear(sd, f"{prefix}.{a}.{n}", dim, dim, g, device)
        for n in ("norm_q", "norm_k"):
            sd[f"{prefix}.{a}.{n}.weight"] = (1 + 0.1 * torch.randn(dim, generator=g, device=device)).to(torch.bfloat16)
    sd[f"{prefix}.norm2.weight"] = (1 + 0.1 * torch.randn(dim, generator=g, device=device)).to(torch.bfloat16)
    sd[f"{prefix}.norm2.bias"] = (0.1 * torch.randn(dim, generator=g, device=device)).to(torch.bfloat16)
    _rand_linear(sd, f"{prefix}.ffn.net.0.proj", ffn_dim, dim, g, device)
    _rand_linear(sd, f"{prefix}.ffn.net.2", dim, ffn_dim, g, device)
    sd[f"{prefix}.scale_shift_table"] = (torch.randn(1, 6, dim, generator=g, device=device) / dim ** 0.5).to(torch.bfloat16)
    return sd


# ---- FLUX ------------------------------------------------------------------------------------------
class FluxTransformer2DModelCore:
    """fastdm/model/flux.py:180-494. `block_sd(prefix, kind)` supplies each block's weights."""

    def __init__(self, num_layers=19, num_single_layers=38, attention_head_dim=128, num_attention_heads=24,
                 in_channels=64, out_channels=64, joint_attention_dim=4096, pooled_projection_dim=768,
                 guidance_embeds=True, axes_dims_rope=(16, 56, 56), quant_dtype=torch.float8_e4m3fn, device="cuda",
                 seed=0, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.inner_dim = d = self.heads * self.hd
        self.axes_dims_rope = axes_dims_rope
        self.guidance_embeds = guidance_embeds
        self.device = device
        g = torch.Generator(device=device).manual_seed(seed)
        sd = state_dict

        def part(make):  # random init unless a full state dict was handed over
            return sd if sd is not None else make()

        def pre():
            s = {}
            for n in ("timestep_embedder", "guidance_embedder"):
                _rand_linear(s, f"time_text_embed.{n}.linear_1", d, 256, g, device)
                _rand_linear(s, f"time_text_embed.{n}.linear_2", d, d, g, device)
            _rand_linear(s, "time_text_embed.text_embedder.linear_1", d, pooled_projection_dim, g, device)
            _rand_linear(s, "time_text_embed.text_embedder.linear_2", d, d, g, device)
            _rand_linear(s, "context_embedder", d, joint_attention_dim, g, device)
            _rand_linear(s, "x_embedder", d, in_channels, g, device)
            _rand_linear(s, "norm_out.linear", 2 * d, d, g, device)
            _rand_linear(s, "proj_out", out_channels, d, g, device)
            return s

        s = part(pre)
        self.timestep_embedder = _MLP(s, "time_text_embed.timestep_embedder", "silu", device)
        self.guidance_embedder = _MLP(s, "time_text_embed.guidance_embedder", "silu", device) if guidance_embeds else None
        self.text_embedder = _MLP(s, "time_text_embed.text_embedder", "silu", device)
        self.context_embedder = load_linear(s, ["context_embedder"], None, device)
        self.x_embedder = load_linear(s, ["x_embedder"], None, device)
        self.norm_out_linear = load_linear(s, ["norm_out.linear"], None, device)
        self.proj_out = load_linear(s, ["proj_out"], None, device)
        self.transformer_blocks: List[FluxTransformerBlock] = []
        for i in range(num_layers):
            p = f"transformer_blocks.{i}"
            self.transformer_blocks.append(FluxTransformerBlock(
                part(lambda: random_flux_double_sd(p, d, self.hd, g, device)), p, self.heads, self.hd, quant_dtype, device))
        self.single_transformer_blocks: List[FluxSingleTransformerBlock] = []
        for i in range(num_single_layers):
            p = f"single_transformer_blocks.{i}"
            self.single_transformer_blocks.append(FluxSingleTransformerBlock(
                part(lambda: random_flux_single_sd(p, d, self.hd, g, device)), p, self.heads, self.hd, quant_dtype, device))
        # every block's AdaLN modulation comes from the same silu(temb): one stacked GEMM per step
        self.use_adaln_table = True
        self.adaln = AdaLNTable([l for b in self.transformer_blocks for l in (b.norm1_linear, b.norm1_context_linear)]
                                + [b.norm_linear for b in self.single_transformer_blocks])

    def forward(self, hidden_states, encoder_hidden_states, pooled_projections, timestep, img_ids, txt_ids,
                guidance=None):
        hidden_states = self.x_embedder.forward(hidden_states)                                   # flux.py:397
        dt = hidden_states.dtype
        timestep = timestep.to(dt) * 1000
        temb = self.timestep_embedder.forward(
            get_timestep_embedding(timestep, 256, flip_sin_to_cos=True, downscale_freq_shift=0).to(dt))
        if guidance is not None and self.guidance_embedder is not None:
            guidance = guidance.to(dt) * 1000
            temb = temb + self.guidance_embedder.forward(
                get_timestep_embedding(guidance, 256, flip_sin_to_cos=True, downscale_freq_shift=0).to(dt))
        temb = temb + self.text_embedder.forward(pooled_projections)                             # embeddings.py:578-590
        encoder_hidden_states = self.context_embedder.forward(encoder_hidden_states)             # flux.py:410
        if txt_ids.ndim == 3:
            txt_ids = txt_ids[0]
        if img_ids.ndim == 3:
            img_ids = img_ids[0]
        rope = flux_rope_table(torch.cat((txt_ids, img_ids), dim=0), self.axes_dims_rope, dt)    # flux.py:417-428
        if _caching(self):                                                                       # flux.py:430-443
            hidden_states = self.cache.apply_cache(
                model_type="flux", hidden_states=hidden_states, encoder_hidden_states=encoder_hidden_states, temb=temb,
                image_rotary_emb=rope, transformer_blocks=self.transformer_blocks,
                single_transformer_blocks=self.single_transformer_blocks)
            return (self._project_out(hidden_states, temb, dt),)
        use_table = self.use_adaln_table
        tables = self.adaln.compute(F.silu(temb)) if use_table else None
        for i, block in enumerate(self.transformer_blocks):                                      # flux.py:445-452
            mod = (self.adaln.chunks(tables, 2 * i, 6), self.adaln.chunks(tables, 2 * i + 1, 6)) if use_table else None
            encoder_hidden_states, hidden_states = block.forward(hidden_states, encoder_hidden_states, temb, rope, mod=mod)
        t = encoder_hidden_states.shape[1]
        hidden_states = torch.cat([encoder_hidden_states, hidden_states], dim=1)                 # flux.py:466
        n_double = 2 * len(self.transformer_blocks)
        for i, block in enumerate(self.single_transformer_blocks):                               # flux.py:468-474
            hidden_states = block.forward(hidden_states, temb, rope,
                                          mod=self.adaln.chunks(tables, n_double + i, 3) if use_table else None)
        hidden_states = hidden_states[:, t:, ...]
        return (self._project_out(hidden_states, temb, dt),)

    def _project_out(self, hidden_states, temb, dt):
        # AdaLayerNormContinuous (normalization.py:90-128) + proj_out
        emb = self.norm_out_linear.forward(F.silu(temb).to(dt))
        scale, shift = torch.chunk(emb, 2, dim=1)
        hidden_states = F.layer_norm(hidden_states, (self.inner_dim,), None, None, 1e-6) * (1 + scale)[:, None, :] \
            + shift[:, None, :]
        return self.proj_out.forward(hidden_states)


# ---- Wan -------------------------------------------------------------------------------------------
class WanTransformer3DModelCore:
    """fastdm/model/wan.py:116-380, text-to-video (no image embedder)."""

    def __init__(self, patch_size=(1, 2, 2), num_attention_heads=40, attention_head_dim=128, in_channels=16,
                 out_channels=16, text_dim=4096, freq_dim=256, ffn_dim=13824, num_layers=40, cross_attn_norm=True,
                 eps=1e-6, quant_dtype=torch.float8_e4m3fn, device="cuda", seed=0,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.inner_dim = d = self.heads * self.hd
        self.patch_size = patch_size
        self.out_channels = out_channels
        self.eps = eps
        self.device = device
        g = torch.Generator(device=device).manual_seed(seed)
        sd = state_dict

        def part(make):
            return sd if sd is not None else make()

        def pre():
            s = {}
            s["patch_embedding.weight"] = (torch.randn(d, in_channels, *patch_size, generator=g, device=device) * 0.02).to(torch.bfloat16)
            s["patch_embedding.bias"] = torch.zeros(d, device=device, dtype=torch.bfloat16)
            _rand_linear(s, "condition_embedder.time_embedder.linear_1", d, freq_dim, g, device, dtype=torch.float32)
            _rand_linear(s, "condition_embedder.time_embedder.linear_2", d, d, g, device, dtype=torch.float32)
            _rand_linear(s, "condition_embedder.time_proj", 6 * d, d, g, device)
            _rand_linear(s, "condition_embedder.text_embedder.linear_1", d, text_dim, g, device)
            _rand_linear(s, "condition_embedder.text_embedder.linear_2", d, d, g, device)
            _rand_linear(s, "proj_out", out_channels * math.prod(patch_size), d, g, device)
            s["scale_shift_table"] = (torch.randn(1, 2, d, generator=g, device=device) / d ** 0.5).to(torch.bfloat16)
            return s

        s = part(pre)
        self.patch_w = s["patch_embedding.weight"].to(device)
        self.patch_b = s["patch_embedding.bias"].to(device)
        self.time_embedder = _MLP(s, "condition_embedder.time_embedder", "silu", device, torch.float32)
        self.time_proj = load_linear(s, ["condition_embedder.time_proj"], None, device)
        self.text_embedder = _MLP(s, "condition_embedder.text_embedder", "gelu_tanh", device)
        self.proj_out = load_linear(s, ["proj_out"], None, device)
        self.scale_shift_table = s["scale_shift_table"].to(device)
        self.freq_dim = freq_dim
        self.blocks: List[WanTransformerBlock] = []
        for i in range(num_layers):
            p = f"blocks.{i}"
            self.blocks.append(WanTransformerBlock(part(lambda: random_wan_block_sd(p, d, ffn_dim, g, device)), p,
                                                   self.heads, self.hd, quant_dtype, device, cross_attn_norm, eps))

    def embed(self, hidden_states, timestep, encoder_hidden_states):
        """Everything before the block stack (wan.py:296-343)."""
        b, c, f, h, w = hidden_states.shape
        pt, ph, pw = self.patch_size
        grid = (f // pt, h // ph, w // pw)
        rope = wan_rope_table(*grid, head_dim=self.hd, dtype=hidden_states.dtype, device=hidden_states.device)
        x = F.conv3d(hidden_states, self.patch_w, self.patch_b, self.patch_size).flatten(2).transpose(1, 2).contiguous()
        tproj = get_timestep_embedding(timestep, self.freq_dim, flip_sin_to_cos=True, downscale_freq_shift=0)
        temb = self.time_embedder.forward(tproj.to(torch.float32)).type_as(encoder_hidden_states)
        timestep_proj = self.time_proj.forward(F.silu(temb)).unflatten(1, (6, -1))
        enc = self.text_embedder.forward(encoder_hidden_states)
        return x, temb, timestep_proj, enc, rope, grid

    def project_out(self, hidden_states, temb):
        """Output norm + projection on (a shard of) the tokens (wan.py:355-371)."""
        shift, scale = (self.scale_shift_table + temb.unsqueeze(1)).chunk(2, dim=1)
        x = (F.layer_norm(hidden_states.float(), (self.inner_dim,), None, None, self.eps) * (1 + scale) + shift
             ).type_as(hidden_states)
        return self.proj_out.forward(x)

    def unpatchify(self, x, grid, batch):
        """wan.py:373-378."""
        pt, ph, pw = self.patch_size
        x = x.reshape(batch, grid[0], grid[1], grid[2], pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
        return x.flatten(6, 7).flatten(4, 5).flatten(2, 3)

    def forward(self, hidden_states, timestep, encoder_hidden_states, sparse_mask=None, dense_layers=0,
                ulysses=None, overlap=True):
        """`ulysses`: a fastdm_b200.ulysses.UlyssesAttention -> tokens are sharded across its ranks for the
        whole block stack (every rank receives the same inputs and returns the same full output)."""
        batch = hidden_states.shape[0]
        x, temb, timestep_proj, enc, rope, grid = self.embed(hidden_states, timestep, encoder_hidden_states)
        pos0 = 0
        if ulysses is not None and ulysses.P > 1:
            x = ulysses.shard_tokens(x, dim=1)
            pos0 = ulysses.rank * x.shape[1]
        if _caching(self):                                                                       # wan.py:339-349
            if ulysses is not None and ulysses.P > 1:
                raise NotImplementedError("step caches are not combined with Ulysses sharding (the indicator would need an "
                                          "all-reduce per step)")
            x = self.cache.apply_cache(model_type="wan", hidden_states=x, encoder_hidden_states=enc, temb=timestep_proj,
                                       image_rotary_emb=rope, transformer_blocks=self.blocks,
                                       sparse_attn=(sparse_mask, dense_layers) if sparse_mask is not None else None)
        else:
            for i, block in enumerate(self.blocks):
                mask = sparse_mask if (sparse_mask is not None and i >= dense_layers) else None
                x = block.forward(x, enc, timestep_proj, rope, mask, ulysses=ulysses, pos0=pos0, overlap=overlap)
        y = self.project_out(x, temb)
        if ulysses is not None and ulysses.P > 1:
            y = ulysses.gather_tokens(y, dim=1)
        return (self.unpatchify(y, grid, batch),)


# ---- Qwen-Image ------------------------------------------------------------------------------------
def qwen_rope_table(frame, height, width, txt_len, axes_dim=(16, 56, 56), theta=10000.0, dtype=torch.bfloat16,
                    device="cuda"):
    """QwenEmbedRope.forward with scale_rope=True (fastdm/layer/embeddings.py:762-857) + the merge of
    qwenimage.py:309-313: rows [text | image], each [cos(hd/2) | sin(hd/2)]. Image rows: frame index, then height
    and width positions centred on zero; text rows: one position max(h//2, w//2) + i on all three axes."""
    def angles(pos, dim):
        inv = 1.0 / torch.pow(torch.tensor(theta), torch.arange(0, dim, 2, dtype=torch.float32) / dim)
        return torch.outer(pos.to(torch.float32), inv)

    centred = lambda n: torch.cat([torch.arange(-(n - n // 2), 0), torch.arange(0, n // 2)])  # noqa: E731
    f = angles(torch.arange(frame), axes_dim[0]).view(frame, 1, 1, -1).expand(frame, height, width, -1)
    hh = angles(centred(height), axes_dim[1]).view(1, height, 1, -1).expand(frame, height, width, -1)
    ww = angles(centred(width), axes_dim[2]).view(1, 1, width, -1).expand(frame, height, width, -1)
    img = torch.cat([f, hh, ww], dim=-1).reshape(frame * height * width, -1)
    tpos = max(height // 2, width // 2) + torch.arange(txt_len)
    txt = torch.cat([angles(tpos, a) for a in axes_dim], dim=-1)
    ang = torch.cat([txt, img], dim=0)
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(dtype).to(device).contiguous()


def random_qwen_block_sd(prefix, dim, head_dim, g, device):
    sd, p = {}, prefix
    _rand_linear(sd, f"{p}.img_mod.1", 6 * dim, dim, g, device)
    _rand_linear(sd, f"{p}.txt_mod.1", 6 * dim, dim, g, device)
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        _rand_linear(sd, f"{p}.attn.{n}", dim, dim, g, device)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        sd[f"{p}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g, device=device)).to(torch.bfloat16)
    for ff in ("img_mlp", "txt_mlp"):
        _rand_linear(sd, f"{p}.{ff}.net.0.proj", 4 * dim, dim, g, device)
        _rand_linear(sd, f"{p}.{ff}.net.2", dim, 4 * dim, g, device)
    return sd


class QwenImageTransformer2DModelCore:
    """fastdm/model/qwenimage.py:126-352 (60 MMDiT blocks, d = 3072, 24 x 128 heads, INT8 W8A8 by default in
    the reference). `ulysses`: sequence parallelism -- the image and the text tokens are each sharded over the
    ranks, every block's joint attention is head-sharded (two all-to-alls), everything else is token-local;
    every rank receives the same inputs and returns the same full output."""

    def __init__(self, num_layers=60, attention_head_dim=128, num_attention_heads=24, in_channels=64, out_channels=64,
                 joint_attention_dim=3584, patch_size=2, axes_dims_rope=(16, 56, 56), quant_dtype=torch.int8,
                 device="cuda", seed=0, state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self._rope_cache = {}
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.inner_dim = d = self.heads * self.hd
        self.axes_dims_rope = axes_dims_rope
        self.device = device
        g = torch.Generator(device=device).manual_seed(seed)
        sd = state_dict

        def part(make):
            return sd if sd is not None else make()

        def pre():
            s = {}
            _rand_linear(s, "time_text_embed.timestep_embedder.linear_1", d, 256, g, device)
            _rand_linear(s, "time_text_embed.timestep_embedder.linear_2", d, d, g, device)
            s["txt_norm.weight"] = (1 + 0.1 * torch.randn(joint_attention_dim, generator=g, device=device)).to(torch.bfloat16)
            _rand_linear(s, "img_in", d, in_channels, g, device)
            _rand_linear(s, "txt_in", d, joint_attention_dim, g, device)
            _rand_linear(s, "norm_out.linear", 2 * d, d, g, device)
            _rand_linear(s, "proj_out", out_channels, d, g, device)   # patch_size^2 * (out_channels / patch_size^2)
            return s

        s = part(pre)
        self.timestep_embedder = _MLP(s, "time_text_embed.timestep_embedder", "silu", device)
        self.txt_norm_weight = s["txt_norm.weight"].to(device).contiguous()
        self.img_in = load_linear(s, ["img_in"], None, device)
        self.txt_in = load_linear(s, ["txt_in"], None, device)
        self.norm_out_linear = load_linear(s, ["norm_out.linear"], None, device)
        self.proj_out = load_linear(s, ["proj_out"], None, device)
        self.transformer_blocks: List[QwenImageTransformerBlock] = []
        for i in range(num_layers):
            p = f"transformer_blocks.{i}"
            self.transformer_blocks.append(QwenImageTransformerBlock(
                part(lambda: random_qwen_block_sd(p, d, self.hd, g, device)), p, self.heads, self.hd, quant_dtype, device))
        self.use_adaln_table = True
        self.adaln = AdaLNTable([l for b in self.transformer_blocks for l in (b.img_mod_proj, b.txt_mod_proj)])

    def forward(self, hidden_states, encoder_hidden_states, timestep, img_shape, ulysses=None):
        """hidden_states [1, f*h*w, in_channels] (packed latents), encoder_hidden_states [1, T, joint_dim],
        img_shape = (f, h, w) of the packed latent grid."""
        dt = hidden_states.dtype
        x = self.img_in.forward(hidden_states)                                                    # qwenimage.py:286
        enc = ops.rms_norm(encoder_hidden_states.reshape(-1, encoder_hidden_states.shape[-1]).contiguous(),
                           self.txt_norm_weight, 1e-6).view_as(encoder_hidden_states)             # :289
        enc = self.txt_in.forward(enc)                                                            # :290
        temb = self.timestep_embedder.forward(
            get_timestep_embedding(timestep.to(dt), 256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000).to(dt))
        T = enc.shape[1]
        key = (tuple(img_shape), T, dt, str(x.device))
        rope = self._rope_cache.get(key)      # the table only depends on the shapes: built once (on the host), reused
        if rope is None:
            rope = self._rope_cache[key] = qwen_rope_table(*img_shape, T, self.axes_dims_rope, dtype=dt, device=x.device)
        rope_pos = None
        if ulysses is not None and ulysses.P > 1:
            x = ulysses.shard_tokens(x, dim=1)
            enc = ulysses.shard_tokens(enc, dim=1)
            rope_pos = (ulysses.rank * enc.shape[1], T + ulysses.rank * x.shape[1])
        if _caching(self):                                                                        # qwenimage.py:316-329
            if ulysses is not None and ulysses.P > 1:
                raise NotImplementedError("step caches are not combined with Ulysses sharding")
            x = self.cache.apply_cache(model_type="qwenimage", hidden_states=x, encoder_hidden_states=enc, temb=temb,
                                       image_rotary_emb=rope, transformer_blocks=self.transformer_blocks)
        else:
            tables = self.adaln.compute(F.silu(temb)) if self.use_adaln_table else None
            for i, block in enumerate(self.transformer_blocks):                                   # :331-340
                mod = (self.adaln.chunks(tables, 2 * i, 6), self.adaln.chunks(tables, 2 * i + 1, 6)) if tables is not None else None
                enc, x = block.forward(x, enc, None, temb, rope, mod=mod, ulysses=ulysses, rope_pos=rope_pos)
        emb = self.norm_out_linear.forward(F.silu(temb).to(dt))                                   # AdaLayerNormContinuous
        scale, shift = torch.chunk(emb, 2, dim=1)
        x = F.layer_norm(x, (self.inner_dim,), None, None, 1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
        y = self.proj_out.forward(x)
        if ulysses is not None and ulysses.P > 1:
            y = ulysses.gather_tokens(y, dim=1)
        return (y,)


# ---- SD3 / SD3.5 -----------------------------------------------------------------------------------
def random_sd3_block_sd(prefix, dim, head_dim, g, device, context_pre_only=False, use_dual_attention=False):
    sd, p = {}, prefix
    _rand_linear(sd, f"{p}.norm1.linear", (9 if use_dual_attention else 6) * dim, dim, g, device)
    _rand_linear(sd, f"{p}.norm1_context.linear", (2 if context_pre_only else 6) * dim, dim, g, device)
    names = ["to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0"] + ([] if context_pre_only else ["to_add_out"])
    for n in names:
        _rand_linear(sd, f"{p}.attn.{n}", dim, dim, g, device)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        sd[f"{p}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g, device=device)).to(torch.bfloat16)
    if use_dual_attention:
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            _rand_linear(sd, f"{p}.attn2.{n}", dim, dim, g, device)
        for n in ("norm_q", "norm_k"):
            sd[f"{p}.attn2.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g, device=device)).to(torch.bfloat16)
    for ff in (("ff",) if context_pre_only else ("ff", "ff_context")):
        _rand_linear(sd, f"{p}.{ff}.net.0.proj", 4 * dim, dim, g, device)
        _rand_linear(sd, f"{p}.{ff}.net.2", dim, 4 * dim, g, device)
    return sd


class SD3TransformerModelCore:
    """fastdm/model/sd35.py:202-424 (SD3.5-medium defaults: 24 MMDiT blocks, d = 1536, 24 x 64 heads, image-only
    second attention in layers 0-12, the last block `context_pre_only`). Single GPU (BASELINE configs[1])."""

    def __init__(self, patch_size=2, in_channels=16, num_layers=24, attention_head_dim=64, num_attention_heads=24,
                 joint_attention_dim=4096, pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=384,
                 dual_attention_layers=tuple(range(13)), quant_dtype=torch.float8_e4m3fn, device="cuda", seed=0,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None):
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.inner_dim = d = self.heads * self.hd
        self.patch_size, self.out_channels, self.pos_max = patch_size, out_channels, pos_embed_max_size
        g = torch.Generator(device=device).manual_seed(seed)
        sd = state_dict

        def part(make):
            return sd if sd is not None else make()

        def pre():
            s = {}
            s["pos_embed.proj.weight"] = (torch.randn(d, in_channels, patch_size, patch_size, generator=g, device=device) * 0.05).to(torch.bfloat16)
            s["pos_embed.proj.bias"] = torch.zeros(d, device=device, dtype=torch.bfloat16)
            s["pos_embed.pos_embed"] = (torch.randn(1, pos_embed_max_size ** 2, d, generator=g, device=device) * 0.02).to(torch.bfloat16)
            for n in ("timestep_embedder", "text_embedder"):
                _rand_linear(s, f"time_text_embed.{n}.linear_1", d, 256 if n == "timestep_embedder" else pooled_projection_dim, g, device)
                _rand_linear(s, f"time_text_embed.{n}.linear_2", d, d, g, device)
            _rand_linear(s, "context_embedder", d, joint_attention_dim, g, device)
            _rand_linear(s, "norm_out.linear", 2 * d, d, g, device)
            _rand_linear(s, "proj_out", patch_size * patch_size * out_channels, d, g, device)
            return s

        s = part(pre)
        self.proj_weight = s["pos_embed.proj.weight"].to(device)
        self.proj_bias = s["pos_embed.proj.bias"].to(device)
        self.pos_embed = s["pos_embed.pos_embed"].to(device)
        self.timestep_embedder = _MLP(s, "time_text_embed.timestep_embedder", "silu", device)
        self.text_embedder = _MLP(s, "time_text_embed.text_embedder", "silu", device)
        self.context_embedder = load_linear(s, ["context_embedder"], None, device)
        self.norm_out_linear = load_linear(s, ["norm_out.linear"], quant_dtype, device)     # quantised in the reference (sd35.py:327-328)
        self.proj_out = load_linear(s, ["proj_out"], quant_dtype, device)
        self.transformer_blocks: List[JointTransformerBlock] = []
        for i in range(num_layers):
            p, last, dual = f"transformer_blocks.{i}", i == num_layers - 1, i in dual_attention_layers
            self.transformer_blocks.append(JointTransformerBlock(
                part(lambda: random_sd3_block_sd(p, d, self.hd, g, device, last, dual)), p, self.heads, self.hd, quant_dtype,
                device, context_pre_only=last, use_dual_attention=dual))

    def _patch_embed(self, latent):
        """PatchEmbed.forward with the centre-cropped learned position table (layer/embeddings.py)."""
        h, w = latent.shape[-2] // self.patch_size, latent.shape[-1] // self.patch_size
        x = F.conv2d(latent, self.proj_weight, self.proj_bias, stride=self.patch_size).flatten(2).transpose(1, 2)
        top, left = (self.pos_max - h) // 2, (self.pos_max - w) // 2
        pos = self.pos_embed.view(1, self.pos_max, self.pos_max, -1)[:, top:top + h, left:left + w].reshape(1, h * w, -1)
        return (x + pos).to(latent.dtype).contiguous(), h, w

    def forward(self, hidden_states, encoder_hidden_states, pooled_projections, timestep):
        dt = hidden_states.dtype
        x, h, w = self._patch_embed(hidden_states)                                               # sd35.py:380
        temb = self.timestep_embedder.forward(
            get_timestep_embedding(timestep, 256, flip_sin_to_cos=True, downscale_freq_shift=0).to(dt))
        temb = temb + self.text_embedder.forward(pooled_projections)                             # :381
        enc = self.context_embedder.forward(encoder_hidden_states)                               # :382
        if _caching(self):                                                                       # sd35.py:383-393
            x = self.cache.apply_cache(model_type="sd35", hidden_states=x, encoder_hidden_states=enc, temb=temb,
                                       transformer_blocks=self.transformer_blocks)
        else:
            for block in self.transformer_blocks:                                                # :394-400
                enc, x = block.forward(x, enc, temb)
        emb = self.norm_out_linear.forward(F.silu(temb).to(dt))
        scale, shift = torch.chunk(emb, 2, dim=1)
        x = F.layer_norm(x, (self.inner_dim,), None, None, 1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
        x = self.proj_out.forward(x.to(dt))
        ps, c = self.patch_size, self.out_channels                                               # unpatchify :411-424
        x = x.reshape(x.shape[0], h, w, ps, ps, c)
        x = torch.einsum("nhwpqc->nchpwq", x)
        return (x.reshape(x.shape[0], c, h * ps, w * ps),)
