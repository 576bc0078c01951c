"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by fastdm_b200/).

CPU restatement of FastDM's layer / block composition on top of oracle/ops_ref.py:
QLinear (fastdm/layer/qlinear.py:6-81), Attention.forward (fastdm/layer/transformer.py:232-317),
WanAttention.forward (:445-535), FeedForward (:14-62), the AdaLN variants
(fastdm/layer/normalization.py:130-236), FluxTransformerBlock / FluxSingleTransformerBlock
(fastdm/model/flux.py:52-178) and WanTransformerBlock (fastdm/model/wan.py:67-114).

Weights are handed over as a flat dict with diffusers key names (the names the reference's loaders
consume: fastdm/model/flux.py:274-328, fastdm/model/wan.py:249-281), `weight` tensors being
[out_features, in_features] as in a state dict.

Pinned by tests/test_oracle_golden.py against tests/golden/block_*.pt, which oracle/gen_golden.py
produced by running the reference's own classes.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops_ref as R


class QLinearRef:
    """fastdm/layer/qlinear.py:6-81. names: list of state-dict prefixes fused along N (:29)."""

    def __init__(self, sd: Dict[str, torch.Tensor], names, quant_type=None):
        ws = [sd[f"{n}.weight"].transpose(0, 1) for n in names]  # (in, out): basemodel.py:51
        bs = [sd.get(f"{n}.bias") for n in names]
        if len(ws) > 1:
            w = torch.cat(ws, 1).contiguous().transpose(0, 1).contiguous().transpose(0, 1)
            self.bias = torch.cat(bs, 0).contiguous() if bs[0] is not None else None
        else:
            w = ws[0]
            self.bias = bs[0]
        self.scale = None
        self.colsum = None
        if quant_type == torch.float8_e4m3fn:  # qlinear.py:40-44
            q, s = R.quantize_to_fp8(w.transpose(0, 1).contiguous())
            w, self.scale = q.transpose(0, 1), s
        elif quant_type == torch.int8:  # qlinear.py:45-50
            q, s, _ = R.quantize_to_int8(w.transpose(0, 1).contiguous())
            w, self.scale = q.transpose(0, 1), s
            self.colsum = w.to(torch.int32).sum(dim=0, keepdim=True, dtype=torch.int32)
        self.weight = w

    def forward(self, x):
        shp = x.shape
        if len(shp) > 2:
            x = x.reshape(-1, shp[-1])
        if self.weight.dtype == torch.float8_e4m3fn:
            xq, xs = R.quantize_to_fp8(x)
            out = R.fp8_matmul(xq, self.weight, xs, self.scale, x.dtype, bias=self.bias)
        elif self.weight.dtype == torch.int8:
            xq, xs, xzp = R.quantize_to_int8(x, symmetric=False)
            out = R.int8_matmul(xq, self.weight, xs, self.scale, x.dtype, self.colsum, xzp, bias=self.bias)
        else:
            out = torch.addmm(self.bias, x, self.weight) if self.bias is not None else torch.mm(x, self.weight)
        if len(shp) > 2:
            out = out.view(*shp[:-1], self.weight.shape[-1])
        return out


def _ada_ln(x, emb, linear: QLinearRef, chunks: int):
    """AdaLayerNormZero / ZeroSingle: fastdm/layer/normalization.py:186-199, 223-236."""
    emb = linear.forward(F.silu(emb))
    parts = emb.chunk(chunks, dim=1)
    shift, scale = parts[0], parts[1]
    x = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale[:, None]) + shift[:, None]
    return (x,) + tuple(parts[2:])


class FluxAttentionRef:
    """fastdm/layer/transformer.py:232-317 as configured by fastdm/model/flux.py:39-50,102-114."""

    def __init__(self, sd, prefix, heads, head_dim, quant, joint: bool):
        self.heads, self.head_dim = heads, head_dim
        self.inner = heads * head_dim
        self.eps = 1e-6
        self.scale = head_dim ** -0.5
        p = prefix
        self.qkv = QLinearRef(sd, [f"{p}.to_q", f"{p}.to_k", f"{p}.to_v"], quant)
        self.norm_q = sd[f"{p}.norm_q.weight"]
        self.norm_k = sd[f"{p}.norm_k.weight"]
        self.joint = joint
        if joint:
            self.add_qkv = QLinearRef(sd, [f"{p}.add_q_proj", f"{p}.add_k_proj", f"{p}.add_v_proj"], quant)
            self.to_out = QLinearRef(sd, [f"{p}.to_out.0"], quant)
            self.to_add_out = QLinearRef(sd, [f"{p}.to_add_out"], quant)
            self.norm_added_q = sd[f"{p}.norm_added_q.weight"]
            self.norm_added_k = sd[f"{p}.norm_added_k.weight"]

    def _split_norm(self, fused, wq, wk, b):
        i = self.inner
        q, k, v = fused[:, :, :i], fused[:, :, i:2 * i], fused[:, :, 2 * i:]
        q = R.rms_norm(q.unflatten(-1, (self.heads, -1)).contiguous(), wq, self.eps).view(b, -1, i)
        k = R.rms_norm(k.unflatten(-1, (self.heads, -1)).contiguous(), wk, self.eps).view(b, -1, i)
        return q, k, v

    def forward(self, hidden, encoder=None, rope=None):
        b = hidden.shape[0]
        q, k, v = self._split_norm(self.qkv.forward(hidden), self.norm_q, self.norm_k, b)
        if encoder is not None:
            eq, ek, ev = self._split_norm(self.add_qkv.forward(encoder), self.norm_added_q, self.norm_added_k, b)
            q = torch.cat([eq, q], dim=1)
            k = torch.cat([ek, k], dim=1)
            v = torch.cat([ev, v], dim=1)
        if rope is not None:
            R.rotary_pos_embedding(q, k, self.head_dim, rope, is_neox=False)
        o = R.scaled_dot_product_attention(q, k, v, self.heads, self.heads, self.head_dim, scale=self.scale)
        o = o.to(q.dtype)
        if encoder is not None:
            t = encoder.shape[1]
            eo, o = o[:, :t], o[:, t:]
            return self.to_out.forward(o), self.to_add_out.forward(eo)
        return o


class FeedForwardRef:
    """FeedForward(activation_fn="gelu-approximate"): fastdm/layer/transformer.py:44-61,
    fastdm/layer/activations.py:32-41."""

    def __init__(self, sd, prefix, quant):
        self.proj = QLinearRef(sd, [f"{prefix}.net.0.proj"], quant)
        self.out = QLinearRef(sd, [f"{prefix}.net.2"], quant)

    def forward(self, x):
        return self.out.forward(F.gelu(self.proj.forward(x), approximate="tanh"))


class FluxTransformerBlockRef:
    """fastdm/model/flux.py:78-178 (double-stream block)."""

    def __init__(self, sd, prefix, heads, head_dim, quant):
        p = prefix
        self.norm1 = QLinearRef(sd, [f"{p}.norm1.linear"])            # unquantized: flux.py:288
        self.norm1_context = QLinearRef(sd, [f"{p}.norm1_context.linear"])
        self.attn = FluxAttentionRef(sd, f"{p}.attn", heads, head_dim, quant, joint=True)
        self.ff = FeedForwardRef(sd, f"{p}.ff", quant)
        self.ff_context = FeedForwardRef(sd, f"{p}.ff_context", quant)

    def forward(self, hidden, encoder, temb, rope=None):
        n, gate_msa, shift_mlp, scale_mlp, gate_mlp = _ada_ln(hidden, temb, self.norm1, 6)
        ne, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = _ada_ln(encoder, temb, self.norm1_context, 6)
        attn, c_attn = self.attn.forward(n, ne, rope)
        hidden = hidden + gate_msa.unsqueeze(1) * attn
        n = F.layer_norm(hidden, (hidden.shape[-1],), None, None, 1e-6)
        n = n * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        hidden = hidden + gate_mlp.unsqueeze(1) * self.ff.forward(n)
        encoder = encoder + c_gate_msa.unsqueeze(1) * c_attn
        ne = F.layer_norm(encoder, (encoder.shape[-1],), None, None, 1e-6)
        ne = ne * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        encoder = encoder + c_gate_mlp.unsqueeze(1) * self.ff_context.forward(ne)
        return encoder, hidden


class JointTransformerBlockRef:
    """SD3 / SD3.5 block, fastdm/model/sd35.py:31-200 (+ SD35AdaLayerNormZeroX / AdaLayerNormContinuous,
    fastdm/layer/normalization.py:45-128)."""

    def __init__(self, sd, prefix, heads, head_dim, quant, context_pre_only=False, use_dual_attention=False):
        p = prefix
        self.dim = heads * head_dim
        self.heads, self.head_dim = heads, head_dim
        self.context_pre_only, self.dual = context_pre_only, use_dual_attention
        self.norm1 = QLinearRef(sd, [f"{p}.norm1.linear"])
        self.norm1_context = QLinearRef(sd, [f"{p}.norm1_context.linear"])
        self.attn = FluxAttentionRef.__new__(FluxAttentionRef)
        a = self.attn
        a.heads, a.head_dim, a.inner, a.eps, a.scale, a.joint = heads, head_dim, self.dim, 1e-6, head_dim ** -0.5, True
        a.qkv = QLinearRef(sd, [f"{p}.attn.to_q", f"{p}.attn.to_k", f"{p}.attn.to_v"], quant)
        a.add_qkv = QLinearRef(sd, [f"{p}.attn.add_q_proj", f"{p}.attn.add_k_proj", f"{p}.attn.add_v_proj"], quant)
        a.to_out = QLinearRef(sd, [f"{p}.attn.to_out.0"], quant)
        a.to_add_out = None if context_pre_only else QLinearRef(sd, [f"{p}.attn.to_add_out"], quant)
        a.norm_q, a.norm_k = sd[f"{p}.attn.norm_q.weight"], sd[f"{p}.attn.norm_k.weight"]
        a.norm_added_q, a.norm_added_k = sd[f"{p}.attn.norm_added_q.weight"], sd[f"{p}.attn.norm_added_k.weight"]
        if use_dual_attention:
            self.attn2_qkv = QLinearRef(sd, [f"{p}.attn2.to_q", f"{p}.attn2.to_k", f"{p}.attn2.to_v"], quant)
            self.attn2_out = QLinearRef(sd, [f"{p}.attn2.to_out.0"], quant)
            self.attn2_nq, self.attn2_nk = sd[f"{p}.attn2.norm_q.weight"], sd[f"{p}.attn2.norm_k.weight"]
        self.ff = FeedForwardRef(sd, f"{p}.ff", quant)
        self.ff_context = None if context_pre_only else FeedForwardRef(sd, f"{p}.ff_context", quant)

    def _joint_attention(self, hidden, encoder):
        """Attention.forward (layer/transformer.py:232-317) without RoPE; context_pre_only drops to_add_out."""
        a = self.attn
        b = hidden.shape[0]
        q, k, v = a._split_norm(a.qkv.forward(hidden), a.norm_q, a.norm_k, b)
        eq, ek, ev = a._split_norm(a.add_qkv.forward(encoder), a.norm_added_q, a.norm_added_k, b)
        q, k, v = torch.cat([eq, q], 1), torch.cat([ek, k], 1), torch.cat([ev, v], 1)
        o = R.scaled_dot_product_attention(q, k, v, a.heads, a.heads, a.head_dim, scale=a.scale).to(q.dtype)
        t = encoder.shape[1]
        eo, o = o[:, :t], o[:, t:]
        return a.to_out.forward(o), (None if self.context_pre_only else a.to_add_out.forward(eo))

    def forward(self, hidden, encoder, temb):
        d = self.dim
        if self.dual:   # SD35AdaLayerNormZeroX.forward
            emb = self.norm1.forward(F.silu(temb).to(hidden.dtype))
            (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp, shift_msa2, scale_msa2, gate_msa2) = emb.chunk(9, dim=1)
            ln = F.layer_norm(hidden, (d,), None, None, 1e-5)
            n = ln * (1 + scale_msa[:, None]) + shift_msa[:, None]
            n2 = ln * (1 + scale_msa2[:, None]) + shift_msa2[:, None]
        else:
            n, gate_msa, shift_mlp, scale_mlp, gate_mlp = _ada_ln(hidden, temb, self.norm1, 6)
        if self.context_pre_only:   # AdaLayerNormContinuous.forward
            emb = self.norm1_context.forward(F.silu(temb).to(encoder.dtype))
            scale, shift = torch.chunk(emb, 2, dim=1)
            ne = F.layer_norm(encoder, (d,), None, None, 1e-6) * (1 + scale)[:, None, :] + shift[:, None, :]
        else:
            ne, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = _ada_ln(encoder, temb, self.norm1_context, 6)
        attn, c_attn = self._joint_attention(n, ne)
        hidden = hidden + gate_msa.unsqueeze(1) * attn
        if self.dual:
            b = hidden.shape[0]
            f = self.attn2_qkv.forward(n2)
            q, k, v = f[:, :, :d], f[:, :, d:2 * d], f[:, :, 2 * d:]
            q = R.rms_norm(q.unflatten(-1, (self.heads, -1)).contiguous(), self.attn2_nq, 1e-6).view(b, -1, d)
            k = R.rms_norm(k.unflatten(-1, (self.heads, -1)).contiguous(), self.attn2_nk, 1e-6).view(b, -1, d)
            o = R.scaled_dot_product_attention(q, k, v, self.heads, self.heads, self.head_dim, scale=self.head_dim ** -0.5)
            hidden = hidden + gate_msa2.unsqueeze(1) * self.attn2_out.forward(o.to(q.dtype))
        n = F.layer_norm(hidden, (d,), None, None, 1e-6)
        n = n * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        hidden = hidden + gate_mlp.unsqueeze(1) * self.ff.forward(n)
        if self.context_pre_only:
            return None, hidden
        encoder = encoder + c_gate_msa.unsqueeze(1) * c_attn
        ne = F.layer_norm(encoder, (d,), None, None, 1e-6)
        ne = ne * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        encoder = encoder + c_gate_mlp.unsqueeze(1) * self.ff_context.forward(ne)
        return encoder, hidden


class QwenImageTransformerBlockRef:
    """fastdm/model/qwenimage.py:16-124 with Attention.forward_qwen (layer/transformer.py:319-391)."""

    def __init__(self, sd, prefix, heads, head_dim, quant):
        p = prefix
        self.dim = heads * head_dim
        self.eps = 1e-6
        self.img_mod = QLinearRef(sd, [f"{p}.img_mod.1"])
        self.txt_mod = QLinearRef(sd, [f"{p}.txt_mod.1"])
        self.attn = FluxAttentionRef(sd, f"{p}.attn", heads, head_dim, quant, joint=True)  # same op sequence
        self.img_mlp = FeedForwardRef(sd, f"{p}.img_mlp", quant)
        self.txt_mlp = FeedForwardRef(sd, f"{p}.txt_mlp", quant)

    @staticmethod
    def _modulate(x, mod):
        shift, scale, gate = mod.chunk(3, dim=-1)
        return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1), gate.unsqueeze(1)

    def forward(self, hidden, encoder, temb, rope=None):
        img_mod1, img_mod2 = self.img_mod.forward(F.silu(temb)).chunk(2, dim=-1)
        txt_mod1, txt_mod2 = self.txt_mod.forward(F.silu(temb)).chunk(2, dim=-1)
        img_m, img_gate1 = self._modulate(F.layer_norm(hidden, (self.dim,), eps=self.eps), img_mod1)
        txt_m, txt_gate1 = self._modulate(F.layer_norm(encoder, (self.dim,), eps=self.eps), txt_mod1)
        img_attn, txt_attn = self.attn.forward(img_m, txt_m, rope)
        hidden = hidden + img_gate1 * img_attn
        encoder = encoder + txt_gate1 * txt_attn
        img_m2, img_gate2 = self._modulate(F.layer_norm(hidden, (self.dim,), eps=self.eps), img_mod2)
        hidden = hidden + img_gate2 * self.img_mlp.forward(img_m2)
        txt_m2, txt_gate2 = self._modulate(F.layer_norm(encoder, (self.dim,), eps=self.eps), txt_mod2)
        encoder = encoder + txt_gate2 * self.txt_mlp.forward(txt_m2)
        return encoder, hidden


class FluxSingleTransformerBlockRef:
    """fastdm/model/flux.py:17-76 (single-stream block)."""

    def __init__(self, sd, prefix, heads, head_dim, quant):
        p = prefix
        self.norm = QLinearRef(sd, [f"{p}.norm.linear"])
        self.proj_mlp = QLinearRef(sd, [f"{p}.proj_mlp"], quant)
        self.proj_out = QLinearRef(sd, [f"{p}.proj_out"], quant)
        self.attn = FluxAttentionRef(sd, f"{p}.attn", heads, head_dim, quant, joint=False)

    def forward(self, hidden, temb, rope=None):
        residual = hidden
        n, gate = _ada_ln(hidden, temb, self.norm, 3)
        mlp = F.gelu(self.proj_mlp.forward(n))
        attn = self.attn.forward(n, None, rope)
        h = torch.cat([attn, mlp], dim=2)
        h = gate.unsqueeze(1) * self.proj_out.forward(h)
        return residual + h


def _fp32_layer_norm(x, weight=None, bias=None, eps=1e-6):
    """FP32LayerNorm.forward: fastdm/layer/normalization.py:158-160."""
    return F.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps).to(x.dtype)


class WanAttentionRef:
    """fastdm/layer/transformer.py:393-535 (T2V: no added_kv_proj)."""

    def __init__(self, sd, prefix, heads, head_dim, quant, cross: bool):
        p = prefix
        self.heads, self.head_dim = heads, head_dim
        self.inner = heads * head_dim
        self.eps = 1e-6
        self.scale = head_dim ** -0.5
        self.cross = cross
        if cross:
            self.to_q = QLinearRef(sd, [f"{p}.to_q"], quant)
            self.to_kv = QLinearRef(sd, [f"{p}.to_k", f"{p}.to_v"], quant)
        else:
            self.qkv = QLinearRef(sd, [f"{p}.to_q", f"{p}.to_k", f"{p}.to_v"], quant)
        self.to_out = QLinearRef(sd, [f"{p}.to_out.0"], quant)
        self.norm_q = sd[f"{p}.norm_q.weight"]
        self.norm_k = sd[f"{p}.norm_k.weight"]

    def forward(self, hidden, encoder=None, rotary_emb=None, sparse_mask=None, block_q=128, block_k=64):
        i = self.inner
        if self.cross:
            q = self.to_q.forward(hidden)
            kv = self.to_kv.forward(encoder)
            k, v = kv[:, :, :i], kv[:, :, i:]
            q = R.rms_norm(q, self.norm_q, self.eps)
            k = R.rms_norm(k.contiguous(), self.norm_k, self.eps)
        else:
            f = self.qkv.forward(hidden)
            q, k, v = f[:, :, :i], f[:, :, i:2 * i], f[:, :, 2 * i:]
            q = R.rms_norm(q.contiguous(), self.norm_q, self.eps)
            k = R.rms_norm(k.contiguous(), self.norm_k, self.eps)
        if rotary_emb is not None:
            cos, sin = rotary_emb
            merged = torch.cat((cos.squeeze()[:, 0::2], sin.squeeze()[:, 1::2]), dim=-1).to(hidden.dtype)
            R.rotary_pos_embedding(q, k, self.head_dim, merged, is_neox=False)
        if sparse_mask is not None and not self.cross:
            o = R.sparse_scaled_dot_product_attention(q, k, v, self.heads, self.heads, self.head_dim,
                                                      scale=self.scale, sparse_mask=sparse_mask,
                                                      block_q=block_q, block_k=block_k)
        else:
            o = R.scaled_dot_product_attention(q, k, v, self.heads, self.heads, self.head_dim, scale=self.scale)
        return self.to_out.forward(o)


class WanTransformerBlockRef:
    """fastdm/model/wan.py:19-114 (wan2.1 / wan2.2-A14B: temb is [B, 6, dim])."""

    def __init__(self, sd, prefix, heads, head_dim, quant, cross_attn_norm=True):
        p = prefix
        self.attn1 = WanAttentionRef(sd, f"{p}.attn1", heads, head_dim, quant, cross=False)
        self.attn2 = WanAttentionRef(sd, f"{p}.attn2", heads, head_dim, quant, cross=True)
        self.ffn = FeedForwardRef(sd, f"{p}.ffn", quant)
        self.table = sd[f"{p}.scale_shift_table"]
        self.cross_attn_norm = cross_attn_norm
        if cross_attn_norm:
            self.norm2_w = sd[f"{p}.norm2.weight"].to(torch.float32)
            self.norm2_b = sd[f"{p}.norm2.bias"].to(torch.float32)

    def forward(self, hidden, encoder, temb, rotary_emb, sparse_mask=None):
        shift_msa, scale_msa, gate_msa, c_shift, c_scale, c_gate = (self.table + temb.float()).chunk(6, dim=1)
        n = (_fp32_layer_norm(hidden) * (1 + scale_msa) + shift_msa).type_as(hidden)
        a = self.attn1.forward(n, None, rotary_emb, sparse_mask)
        hidden = (hidden.float() + a * gate_msa).type_as(hidden)
        if self.cross_attn_norm:
            n = _fp32_layer_norm(hidden, self.norm2_w, self.norm2_b).type_as(hidden)
        else:
            n = hidden
        hidden = hidden + self.attn2.forward(n, encoder)
        n = (_fp32_layer_norm(hidden) * (1 + c_scale) + c_shift).type_as(hidden)
        ff = self.ffn.forward(n)
        hidden = (hidden.float() + ff.float() * c_gate).type_as(hidden)
        return hidden


# ---- synthetic random-init state dicts (diffusers key names) ------------------------------------
def _lin(sd, name, out_f, in_f, g, std=0.02, bias=True, dtype=torch.bfloat16):
    sd[f"{name}.weight"] = (torch.randn(out_f, in_f, generator=g) * std).to(dtype)
    if bias:
        sd[f"{name}.bias"] = (torch.randn(out_f, generator=g) * std).to(dtype)


def flux_double_state_dict(prefix, dim, head_dim, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = prefix
    _lin(sd, f"{p}.norm1.linear", 6 * dim, dim, g)
    _lin(sd, f"{p}.norm1_context.linear", 6 * dim, dim, g)
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        _lin(sd, f"{p}.attn.{n}", dim, dim, g)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        sd[f"{p}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g)).to(dtype)
    for ff in ("ff", "ff_context"):
        _lin(sd, f"{p}.{ff}.net.0.proj", 4 * dim, dim, g)
        _lin(sd, f"{p}.{ff}.net.2", dim, 4 * dim, g)
    return sd


def sd3_block_state_dict(prefix, dim, head_dim, seed, context_pre_only=False, use_dual_attention=False,
                         dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = prefix
    _lin(sd, f"{p}.norm1.linear", (9 if use_dual_attention else 6) * dim, dim, g)
    _lin(sd, f"{p}.norm1_context.linear", (2 if context_pre_only else 6) * dim, dim, g)
    names = ["to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0"]
    if not context_pre_only:
        names.append("to_add_out")
    for n in names:
        _lin(sd, f"{p}.attn.{n}", dim, dim, g)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        sd[f"{p}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g)).to(dtype)
    if use_dual_attention:
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            _lin(sd, f"{p}.attn2.{n}", dim, dim, g)
        for n in ("norm_q", "norm_k"):
            sd[f"{p}.attn2.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g)).to(dtype)
    for ff in (("ff",) if context_pre_only else ("ff", "ff_context")):
        _lin(sd, f"{p}.{ff}.net.0.proj", 4 * dim, dim, g)
        _lin(sd, f"{p}.{ff}.net.2", dim, 4 * dim, g)
    return sd


def qwen_block_state_dict(prefix, dim, head_dim, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = prefix
    _lin(sd, f"{p}.img_mod.1", 6 * dim, dim, g)
    _lin(sd, f"{p}.txt_mod.1", 6 * dim, dim, g)
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        _lin(sd, f"{p}.attn.{n}", dim, dim, g)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        sd[f"{p}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g)).to(dtype)
    for ff in ("img_mlp", "txt_mlp"):
        _lin(sd, f"{p}.{ff}.net.0.proj", 4 * dim, dim, g)
        _lin(sd, f"{p}.{ff}.net.2", dim, 4 * dim, g)
    return sd


def flux_single_state_dict(prefix, dim, head_dim, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = prefix
    _lin(sd, f"{p}.norm.linear", 3 * dim, dim, g)
    _lin(sd, f"{p}.proj_mlp", 4 * dim, dim, g)
    _lin(sd, f"{p}.proj_out", dim, 5 * dim, g)
    for n in ("to_q", "to_k", "to_v"):
        _lin(sd, f"{p}.attn.{n}", dim, dim, g)
    for n in ("norm_q", "norm_k"):
        sd[f"{p}.attn.{n}.weight"] = (1 + 0.1 * torch.randn(head_dim, generator=g)).to(dtype)
    return sd


def wan_block_state_dict(prefix, dim, ffn_dim, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    p = prefix
    for a in ("attn1", "attn2"):
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            _lin(sd, f"{p}.{a}.{n}", dim, dim, g)
        for n in ("norm_q", "norm_k"):
            sd[f"{p}.{a}.{n}.weight"] = (1 + 0.1 * torch.randn(dim, generator=g)).to(dtype)
    sd[f"{p}.norm2.weight"] = (1 + 0.1 * torch.randn(dim, generator=g)).to(dtype)
    sd[f"{p}.norm2.bias"] = (0.1 * torch.randn(dim, generator=g)).to(dtype)
    _lin(sd, f"{p}.ffn.net.0.proj", ffn_dim, dim, g)
    _lin(sd, f"{p}.ffn.net.2", dim, ffn_dim, g)
    sd[f"{p}.scale_shift_table"] = (torch.randn(1, 6, dim, generator=g) / dim ** 0.5).to(dtype)
    return sd


# ---- whole-model synthetic state dicts (reduced width; tests/golden/model_*.pt regenerate them from the seed) --------
def flux_model_state_dict(cfg, seed, dtype=torch.bfloat16):
    """Keys of fastdm/model/flux.py:274-328 for a FluxTransformer2DModelCore(**cfg)."""
    g = torch.Generator().manual_seed(seed)
    d, hd = cfg["num_attention_heads"] * cfg["attention_head_dim"], cfg["attention_head_dim"]
    sd = {}
    for n in ("timestep_embedder", "guidance_embedder"):
        _lin(sd, f"time_text_embed.{n}.linear_1", d, 256, g)
        _lin(sd, f"time_text_embed.{n}.linear_2", d, d, g)
    _lin(sd, "time_text_embed.text_embedder.linear_1", d, cfg["pooled_projection_dim"], g)
    _lin(sd, "time_text_embed.text_embedder.linear_2", d, d, g)
    _lin(sd, "context_embedder", d, cfg["joint_attention_dim"], g)
    _lin(sd, "x_embedder", d, cfg["in_channels"], g)
    _lin(sd, "norm_out.linear", 2 * d, d, g)
    _lin(sd, "proj_out", cfg["out_channels"], d, g)
    for i in range(cfg["num_layers"]):
        sd.update(flux_double_state_dict(f"transformer_blocks.{i}", d, hd, seed=seed * 100 + i, dtype=dtype))
    for i in range(cfg["num_single_layers"]):
        sd.update(flux_single_state_dict(f"single_transformer_blocks.{i}", d, hd, seed=seed * 100 + 50 + i, dtype=dtype))
    return sd


def wan_model_state_dict(cfg, seed, dtype=torch.bfloat16):
    """Keys of fastdm/model/wan.py:216-281 for a WanTransformer3DModelCore(**cfg) (text-to-video)."""
    import math
    g = torch.Generator().manual_seed(seed)
    d = cfg["num_attention_heads"] * cfg["attention_head_dim"]
    ps = cfg.get("patch_size", (1, 2, 2))
    sd = {}
    sd["patch_embedding.weight"] = (torch.randn(d, cfg["in_channels"], *ps, generator=g) * 0.05).to(dtype)
    sd["patch_embedding.bias"] = (torch.randn(d, generator=g) * 0.02).to(dtype)
    _lin(sd, "condition_embedder.time_embedder.linear_1", d, cfg["freq_dim"], g)
    _lin(sd, "condition_embedder.time_embedder.linear_2", d, d, g)
    _lin(sd, "condition_embedder.time_proj", 6 * d, d, g)
    _lin(sd, "condition_embedder.text_embedder.linear_1", d, cfg["text_dim"], g)
    _lin(sd, "condition_embedder.text_embedder.linear_2", d, d, g)
    _lin(sd, "proj_out", cfg["out_channels"] * math.prod(ps), d, g)
    sd["scale_shift_table"] = (torch.randn(1, 2, d, generator=g) / d ** 0.5).to(dtype)
    for i in range(cfg["num_layers"]):
        sd.update(wan_block_state_dict(f"blocks.{i}", d, cfg["ffn_dim"], seed=seed * 100 + i, dtype=dtype))
    return sd
