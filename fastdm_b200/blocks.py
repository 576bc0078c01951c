"""DiT transformer blocks on the fused B200 kernels -- host-side mirrors of

    FluxTransformerBlock        fastdm/model/flux.py:78-178
    FluxSingleTransformerBlock  fastdm/model/flux.py:17-76
    QwenImageTransformerBlock   fastdm/model/qwenimage.py:16-124
    JointTransformerBlock       fastdm/model/sd35.py:31-200 (SD3.5, incl. dual attention / context_pre_only)
    WanTransformerBlock         fastdm/model/wan.py:19-114   (+ WanAttention, layer/transformer.py:393-535)

Same inputs, outputs and weights (diffusers state-dict names) as the reference classes. The op
sequence of each block is the reference's, regrouped so that every activation makes one trip
through HBM per GEMM:

    AdaLN LayerNorm + modulate + per-token quant   -> ops.layernorm_modulate_quant   (1 kernel)
    qkv projection                                  -> W8A8 GEMM writing the joint [txt|img] buffer
    slice/.contiguous()/rms_norm x2/cat x3/rope     -> ops.qk_norm_rope_ in place     (1 kernel)
    attention                                       -> reads q/k/v as strided views, no copies
    out-proj / ff2 + gate * out + residual          -> GEMM epilogue
    ff1 / proj_mlp + GELU                           -> GEMM epilogue
    cat([attn, mlp])  (single block)                -> both producers write one buffer
"""
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import ops
from .layers import FeedForward, QLinear, Quantized, load_linear, quantize


class ModChunk:
    """One AdaLN modulation vector [B, dim] whose forms were prepared for all blocks at once by `AdaLNTable` (one GEMM
    + two elementwise kernels per step instead of ~20 tiny launches per block): `one_plus` = (1 + x) and `same` = x,
    both in the model dtype (what the reference computes, normalization.py:196), `f32` = x widened to fp32 (the GEMM
    epilogue takes its gate vectors in fp32)."""

    __slots__ = ("one_plus", "same", "f32")

    def __init__(self, one_plus: torch.Tensor, same: torch.Tensor, f32: torch.Tensor):
        self.one_plus, self.same, self.f32 = one_plus, same, f32


def _f32(x) -> torch.Tensor:
    return x.f32 if isinstance(x, ModChunk) else x.float().contiguous()


def _mod(scale, shift) -> Tuple[torch.Tensor, torch.Tensor]:
    """((1 + scale), shift) for the fused LayerNorm-modulate kernel, evaluated in the tensor dtype as the reference
    does (normalization.py:196). bf16 vectors are handed over as they are -- the kernel then runs the reference's
    bf16 op chain as native packed bf16 instructions; other dtypes are widened to fp32 (exact)."""
    if isinstance(scale, ModChunk):
        return scale.one_plus, shift.same
    a, c = (1 + scale), shift
    if a.dtype != torch.bfloat16:
        a, c = a.float(), c.float()
    return a.contiguous(), c.contiguous()


class AdaLNTable:
    """All AdaLN-Zero modulation linears of a model that share one conditioning vector (FLUX: norm1.linear,
    norm1_context.linear of every double block, norm.linear of every single block -- flux.py:288, 99), stacked
    into one [sum N_i, K] weight so that a step computes every block's modulation with ONE GEMM
    (the weights are read once either way: 6.4 GB for FLUX.1-dev). The blocks' own QLinear objects are
    re-pointed at views of the stacked storage, so nothing is duplicated and `block.forward(temb)` without a
    table keeps working."""

    def __init__(self, linears):
        self.offsets = []
        ws, bs, off = [], [], 0
        for lin in linears:
            n = lin.weight.shape[1]
            self.offsets.append((off, n))
            ws.append(lin.weight.t())
            bs.append(lin.bias)
            off += n
        self.weight_store = torch.cat(ws, dim=0).contiguous()          # [sum N, K]
        self.bias = torch.cat(bs, dim=0).contiguous()
        for lin, (o, n) in zip(linears, self.offsets):
            lin.weight = self.weight_store[o:o + n].t()
            lin.bias = self.bias[o:o + n]

    def compute(self, cond: torch.Tensor):
        """cond = silu(temb) [B, K] -> (one_plus, same, f32): (1 + e) and e in the model dtype, e in fp32, [B, sum N]."""
        e = torch.addmm(self.bias, cond, self.weight_store.t())
        return (1 + e), e, e.float()

    def chunks(self, tables, index: int, n_chunks: int):
        one_plus, same, f32 = tables
        o, n = self.offsets[index]
        d = n // n_chunks
        # (batch 1: column slices of one row are contiguous as they are)
        mk = (lambda t, a: t[:, a:a + d]) if one_plus.shape[0] == 1 else (lambda t, a: t[:, a:a + d].contiguous())
        return tuple(ModChunk(mk(one_plus, o + k * d), mk(same, o + k * d), mk(f32, o + k * d)) for k in range(n_chunks))


class _JointDiTBlock:
    """Shared body of the MMDiT double-stream blocks (FLUX double block, Qwen-Image block): two
    streams (image, text) with their own AdaLN modulation, one joint attention over [text | image],
    per-stream FFN. Sub-classes only differ in where the weights and the six modulation vectors per
    stream come from."""

    heads: int
    hd: int
    quant_type: object
    eps: float

    def _load_common(self, sd, p, q, dv, ff_img, ff_txt, context_pre_only=False):
        self.dim = self.heads * self.hd
        self.qkv = load_linear(sd, [f"{p}.attn.to_q", f"{p}.attn.to_k", f"{p}.attn.to_v"], q, dv)
        self.add_qkv_proj = load_linear(sd, [f"{p}.attn.add_q_proj", f"{p}.attn.add_k_proj", f"{p}.attn.add_v_proj"], q, dv)
        self.to_out = load_linear(sd, [f"{p}.attn.to_out.0"], q, dv)
        self.to_add_out = None if context_pre_only else load_linear(sd, [f"{p}.attn.to_add_out"], q, dv)
        self.norm_q_weight = sd[f"{p}.attn.norm_q.weight"].to(dv).contiguous()
        self.norm_k_weight = sd[f"{p}.attn.norm_k.weight"].to(dv).contiguous()
        self.norm_added_q_weight = sd[f"{p}.attn.norm_added_q.weight"].to(dv).contiguous()
        self.norm_added_k_weight = sd[f"{p}.attn.norm_added_k.weight"].to(dv).contiguous()
        self.ff = FeedForward(load_linear(sd, [f"{p}.{ff_img}.net.0.proj"], q, dv), load_linear(sd, [f"{p}.{ff_img}.net.2"], q, dv))
        self.ff_context = None if context_pre_only else FeedForward(
            load_linear(sd, [f"{p}.{ff_txt}.net.0.proj"], q, dv), load_linear(sd, [f"{p}.{ff_txt}.net.2"], q, dv))
        self.scale = self.hd ** -0.5

    def _modulated(self, x, scale, shift, eps):
        """LayerNorm(x) * (1 + scale) + shift in the model dtype (the fused kernel with the quantisation switched off):
        the tensor TeaCache thresholds on (fastdm/caching/xcaching.py:164-185)."""
        B, S, d = x.shape
        a, c = _mod(scale, shift)
        y = ops.layernorm_modulate_quant(x.reshape(B * S, d), a, c, S, None, eps)[3]
        return y.view(B, S, d)

    def _forward_joint(self, hidden_states, encoder_hidden_states, img_mod, txt_mod, image_rotary_emb,
                       dual_mod=None, context_pre_only=False, eps1=None, ulysses=None, rope_pos=None):
        """img_mod / txt_mod: (shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp), each [B, dim].
        SD3.5 extras: `dual_mod` = (shift_msa2, scale_msa2, gate_msa2) runs the image-only second attention
        (self.attn2_*); `context_pre_only` (last block): txt_mod = (shift, scale), the text stream only feeds
        k/v and is dropped afterwards; `eps1` overrides the first LayerNorm's eps.
        Ulysses (Qwen-Image): hidden_states / encoder_hidden_states are this rank's token shards of the two
        streams, `rope_pos` = (first text row, first image row) of the shards in `image_rotary_emb`; the joint
        attention runs head-sharded over the full [text | image] sequence through `ulysses.attention`."""
        hidden_states, encoder_hidden_states = hidden_states.contiguous(), encoder_hidden_states.contiguous()
        B, S_img, d = hidden_states.shape
        S_txt = encoder_hidden_states.shape[1]
        S = S_txt + S_img
        H, hd, qt = self.heads, self.hd, self.quant_type
        eps1 = self.eps if eps1 is None else eps1
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = img_mod
        if context_pre_only:
            c_shift_msa, c_scale_msa = txt_mod
            c_gate_msa = c_shift_mlp = c_scale_mlp = c_gate_mlp = None
        else:
            c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = txt_mod
        hid2 = hidden_states.reshape(B * S_img, d)
        enc2 = encoder_hidden_states.reshape(B * S_txt, d)
        # norm1 + modulate + quant (normalization.py:191-199 / qwenimage.py:77-82) for both streams
        a, c = _mod(scale_msa, shift_msa)
        xq = Quantized(*ops.layernorm_modulate_quant(hid2, a, c, S_img, qt, eps1)[:3])
        a, c = _mod(c_scale_msa, c_shift_msa)
        cq = Quantized(*ops.layernorm_modulate_quant(enc2, a, c, S_txt, qt, self.eps)[:3])
        if dual_mod is not None:   # SD35AdaLayerNormZeroX: the same LayerNorm output, second modulation
            a, c = _mod(dual_mod[1], dual_mod[0])
            xq2 = Quantized(*ops.layernorm_modulate_quant(hid2, a, c, S_img, qt, eps1)[:3])

        # joint [txt | img] qkv buffer: both projections write their rows, no torch.cat (transformer.py:293-295, 370-372)
        qkv = torch.empty((B, S, 3 * d), device=hid2.device, dtype=hidden_states.dtype)
        for b in range(B):
            self.add_qkv_proj.forward(cq.rows(b * S_txt, (b + 1) * S_txt), out=qkv[b, :S_txt])
            self.qkv.forward(xq.rows(b * S_img, (b + 1) * S_img), out=qkv[b, S_txt:])
            pos_txt, pos_img = (0, S_txt) if rope_pos is None else rope_pos
            ops.qk_norm_rope_(qkv[b, :S_txt], self.norm_added_q_weight, self.norm_added_k_weight, image_rotary_emb,
                              H, H, hd, 0, d, pos_txt, self.eps)
            ops.qk_norm_rope_(qkv[b, S_txt:], self.norm_q_weight, self.norm_k_weight, image_rotary_emb,
                              H, H, hd, 0, d, pos_img, self.eps)
        if ulysses is not None and ulysses.P > 1:
            attn = ulysses.attention(qkv, self.scale)      # keys of every rank's shards, this rank's queries
        else:
            attn = ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], H, hd, self.scale)

        # (empty_like would inherit the strides of a transposed / sliced input)
        new_hidden = torch.empty(hidden_states.shape, device=hidden_states.device, dtype=hidden_states.dtype)
        new_encoder = None if context_pre_only else torch.empty(encoder_hidden_states.shape, device=hidden_states.device,
                                                                dtype=encoder_hidden_states.dtype)
        g_msa, g_mlp = _f32(gate_msa), _f32(gate_mlp)
        if not context_pre_only:
            cg_msa, cg_mlp = _f32(c_gate_msa), _f32(c_gate_mlp)
        for b in range(B):
            # hidden = hidden + gate_msa * to_out(attn)      (flux.py:153-154, qwenimage.py:98-99)
            aq = quantize(attn[b, S_txt:], qt)
            self.to_out.forward(aq, gate=g_msa[b:b + 1], residual=hidden_states[b], rows_per_batch=S_img,
                                out=new_hidden[b])
            if not context_pre_only:
                eq = quantize(attn[b, :S_txt], qt)
                self.to_add_out.forward(eq, gate=cg_msa[b:b + 1], residual=encoder_hidden_states[b],
                                        rows_per_batch=S_txt, out=new_encoder[b])
        if dual_mod is not None:
            # hidden = hidden + gate_msa2 * attn2(norm_hidden_states2)      (sd35.py:165-168): image-only attention
            qkv2 = self.attn2_qkv.forward(xq2).view(B, S_img, 3 * d)
            for b in range(B):
                ops.qk_norm_rope_(qkv2[b], self.attn2_norm_q, self.attn2_norm_k, None, H, H, hd, 0, d, 0, self.eps)
            attn2 = ops.attention(qkv2[:, :, :d], qkv2[:, :, d:2 * d], qkv2[:, :, 2 * d:], H, hd, self.scale)
            self.attn2_to_out.forward(quantize(attn2.view(B * S_img, d), qt), gate=dual_mod[2].float().contiguous(),
                                      residual=new_hidden.view(B * S_img, d), rows_per_batch=S_img,
                                      out=new_hidden.view(B * S_img, d))
        # norm2 + modulate + quant, ff with GELU(tanh) epilogue, gate + residual epilogue (flux.py:156-163)
        a, c = _mod(scale_mlp, shift_mlp)
        nq = Quantized(*ops.layernorm_modulate_quant(new_hidden.view(B * S_img, d), a, c, S_img, qt, self.eps)[:3])
        self.ff.forward(nq, gate=g_mlp, residual=new_hidden.view(B * S_img, d), rows_per_batch=S_img,
                        out=new_hidden.view(B * S_img, d))
        if context_pre_only:
            return None, new_hidden
        a, c = _mod(c_scale_mlp, c_shift_mlp)
        nq = Quantized(*ops.layernorm_modulate_quant(new_encoder.view(B * S_txt, d), a, c, S_txt, qt, self.eps)[:3])
        self.ff_context.forward(nq, gate=cg_mlp, residual=new_encoder.view(B * S_txt, d), rows_per_batch=S_txt,
                                out=new_encoder.view(B * S_txt, d))
        return new_encoder, new_hidden


class FluxTransformerBlock(_JointDiTBlock):
    def __init__(self, sd, prefix, num_attention_heads, attention_head_dim, quant_type=torch.float8_e4m3fn,
                 device="cuda", eps=1e-6):
        p, q, dv = prefix, quant_type, device
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.quant_type = q
        self.eps = eps
        self.norm1_linear = load_linear(sd, [f"{p}.norm1.linear"], None, dv)              # unquantized: flux.py:288
        self.norm1_context_linear = load_linear(sd, [f"{p}.norm1_context.linear"], None, dv)
        self._load_common(sd, p, q, dv, "ff", "ff_context")

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None,
                mod=None):
        """`mod` = (img chunks, txt chunks) prepared by an AdaLNTable; without it the block computes its own."""
        if mod is not None:
            return self._forward_joint(hidden_states, encoder_hidden_states, mod[0], mod[1], image_rotary_emb)
        # AdaLN parameters (M = batch GEMMs, unquantized as in the reference)
        emb = self.norm1_linear.forward(F.silu(temb))
        cemb = self.norm1_context_linear.forward(F.silu(temb))
        return self._forward_joint(hidden_states, encoder_hidden_states, emb.chunk(6, dim=1), cemb.chunk(6, dim=1),
                                   image_rotary_emb)

    def cache_indicator(self, hidden_states, encoder_hidden_states, temb):
        """TeaCache: `transformer_blocks[0].norm1.forward(inp, emb=temb)[0]` (xcaching.py:181-183, AdaLayerNormZero)."""
        shift_msa, scale_msa = self.norm1_linear.forward(F.silu(temb)).chunk(6, dim=1)[:2]
        return self._modulated(hidden_states, scale_msa, shift_msa, self.eps)


class QwenImageTransformerBlock(_JointDiTBlock):
    """fastdm/model/qwenimage.py:16-124 (+ Attention.forward_qwen, layer/transformer.py:319-391):
    img_mod / txt_mod give (shift, scale, gate) x 2 per stream; INT8 W8A8 is the reference's default
    for this model, FP8 works the same."""

    def __init__(self, sd, prefix, num_attention_heads, attention_head_dim, quant_type=torch.int8, device="cuda",
                 eps=1e-6, quant_img_txt_mod=False):
        p, q, dv = prefix, quant_type, device
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.quant_type = q
        self.eps = eps
        mq = q if quant_img_txt_mod else None                                            # qwenimage.py:219-220
        self.img_mod_proj = load_linear(sd, [f"{p}.img_mod.1"], mq, dv)
        self.txt_mod_proj = load_linear(sd, [f"{p}.txt_mod.1"], mq, dv)
        self._load_common(sd, p, q, dv, "img_mlp", "txt_mlp")

    def forward(self, hidden_states, encoder_hidden_states, encoder_hidden_states_mask, temb, image_rotary_emb=None,
                joint_attention_kwargs=None, mod=None, ulysses=None, rope_pos=None):
        if mod is not None:    # (img chunks, txt chunks) prepared by an AdaLNTable
            img, txt = mod
        else:
            img = self.img_mod_proj.forward(F.silu(temb)).chunk(6, dim=-1)   # mod1 = (shift, scale, gate), mod2 likewise
            txt = self.txt_mod_proj.forward(F.silu(temb)).chunk(6, dim=-1)
        return self._forward_joint(hidden_states, encoder_hidden_states, img, txt, image_rotary_emb,
                                   ulysses=ulysses, rope_pos=rope_pos)

    def cache_indicator(self, hidden_states, encoder_hidden_states, temb):
        """TeaCache on Qwen-Image thresholds on the modulated TEXT stream of block 0 (xcaching.py:169-180)."""
        shift, scale = self.txt_mod_proj.forward(F.silu(temb)).chunk(6, dim=-1)[:2]
        return self._modulated(encoder_hidden_states, scale, shift, self.eps)


class JointTransformerBlock(_JointDiTBlock):
    """SD3 / SD3.5 MMDiT block, fastdm/model/sd35.py:31-200: no RoPE, optional image-only second
    attention (`use_dual_attention`, SD3.5-medium layers 0-12), `context_pre_only` for the last block."""

    def __init__(self, sd, prefix, num_attention_heads, attention_head_dim, quant_type=torch.float8_e4m3fn,
                 device="cuda", context_pre_only=False, use_dual_attention=False, eps=1e-6):
        p, q, dv = prefix, quant_type, device
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.quant_type = q
        self.eps = eps
        self.context_pre_only, self.use_dual_attention = context_pre_only, use_dual_attention
        self.norm1_linear = load_linear(sd, [f"{p}.norm1.linear"], None, dv)            # 6*dim, or 9*dim (dual)
        self.norm1_context_linear = load_linear(sd, [f"{p}.norm1_context.linear"], None, dv)  # 6*dim, or 2*dim
        self._load_common(sd, p, q, dv, "ff", "ff_context", context_pre_only)
        if use_dual_attention:
            self.attn2_qkv = load_linear(sd, [f"{p}.attn2.to_q", f"{p}.attn2.to_k", f"{p}.attn2.to_v"], q, dv)
            self.attn2_to_out = load_linear(sd, [f"{p}.attn2.to_out.0"], q, dv)
            self.attn2_norm_q = sd[f"{p}.attn2.norm_q.weight"].to(dv).contiguous()
            self.attn2_norm_k = sd[f"{p}.attn2.norm_k.weight"].to(dv).contiguous()

    def forward(self, hidden_states, encoder_hidden_states, temb, joint_attention_kwargs=None):
        emb = self.norm1_linear.forward(F.silu(temb).to(hidden_states.dtype))
        dual = None
        if self.use_dual_attention:      # SD35AdaLayerNormZeroX (normalization.py:45-87): eps 1e-5
            parts = emb.chunk(9, dim=1)
            img_mod, dual, eps1 = parts[:6], parts[6:], 1e-5
        else:
            img_mod, eps1 = emb.chunk(6, dim=1), 1e-6
        cemb = self.norm1_context_linear.forward(F.silu(temb).to(hidden_states.dtype))
        if self.context_pre_only:        # AdaLayerNormContinuous (normalization.py:124-127): scale first, then shift
            scale, shift = cemb.chunk(2, dim=1)
            txt_mod = (shift, scale)
        else:
            txt_mod = cemb.chunk(6, dim=1)
        return self._forward_joint(hidden_states, encoder_hidden_states, img_mod, txt_mod, None, dual_mod=dual,
                                   context_pre_only=self.context_pre_only, eps1=eps1)

    def cache_indicator(self, hidden_states, encoder_hidden_states, temb):
        """TeaCache: first output of norm1 (AdaLayerNormZero, or SD35AdaLayerNormZeroX on the dual-attention blocks)."""
        emb = self.norm1_linear.forward(F.silu(temb).to(hidden_states.dtype))
        shift_msa, scale_msa = emb.chunk(9 if self.use_dual_attention else 6, dim=1)[:2]
        return self._modulated(hidden_states, scale_msa, shift_msa, 1e-5 if self.use_dual_attention else 1e-6)


class FluxSingleTransformerBlock:
    def __init__(self, sd, prefix, num_attention_heads, attention_head_dim, quant_type=torch.float8_e4m3fn,
                 device="cuda", mlp_ratio=4.0, eps=1e-6):
        p, q, dv = prefix, quant_type, device
        self.heads, self.hd = num_attention_heads, attention_head_dim
        self.dim = self.heads * self.hd
        self.mlp_hidden_dim = int(self.dim * mlp_ratio)
        self.quant_type = q
        self.eps = eps
        self.norm_linear = load_linear(sd, [f"{p}.norm.linear"], None, dv)
        self.proj_mlp = load_linear(sd, [f"{p}.proj_mlp"], q, dv)
        self.proj_out = load_linear(sd, [f"{p}.proj_out"], q, dv)
        self.qkv = load_linear(sd, [f"{p}.attn.to_q", f"{p}.attn.to_k", f"{p}.attn.to_v"], q, dv)
        self.norm_q_weight = sd[f"{p}.attn.norm_q.weight"].to(dv).contiguous()
        self.norm_k_weight = sd[f"{p}.attn.norm_k.weight"].to(dv).contiguous()
        self.scale = self.hd ** -0.5

    def forward(self, hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None, mod=None):
        B, S, d = hidden_states.shape
        H, hd, qt = self.heads, self.hd, self.quant_type
        if mod is not None:   # (shift, scale, gate) prepared by an AdaLNTable
            shift_msa, scale_msa, gate = mod
        else:
            emb = self.norm_linear.forward(F.silu(temb))
            shift_msa, scale_msa, gate = emb.chunk(3, dim=1)
        hid2 = hidden_states.reshape(B * S, d)
        a, c = _mod(scale_msa, shift_msa)
        # one quantised copy of norm_hidden_states feeds both proj_mlp and qkv (flux.py:60-67)
        xq = Quantized(*ops.layernorm_modulate_quant(hid2, a, c, S, qt, self.eps)[:3])
        # cat([attn_output, mlp_hidden_states], dim=2) without the cat (flux.py:69)
        cat = torch.empty((B, S, d + self.mlp_hidden_dim), device=hid2.device, dtype=hidden_states.dtype)
        cat2 = cat.view(B * S, d + self.mlp_hidden_dim)
        self.proj_mlp.forward(xq, act="gelu_erf", out=cat2[:, d:])            # F.gelu (erf): flux.py:61
        qkv = self.qkv.forward(xq).view(B, S, 3 * d)
        for b in range(B):
            ops.qk_norm_rope_(qkv[b], self.norm_q_weight, self.norm_k_weight, image_rotary_emb, H, H, hd, 0, d, 0,
                              self.eps)
        ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], H, hd, self.scale, out=cat[:, :, :d])
        cq = quantize(cat2, qt)
        out = torch.empty_like(hidden_states)
        # hidden = residual + gate * proj_out(cat)      (flux.py:70-72)
        self.proj_out.forward(cq, gate=_f32(gate), residual=hid2, rows_per_batch=S,
                              out=out.view(B * S, d))
        return out


class WanTransformerBlock:
    """wan2.1 / wan2.2-A14B text-to-video block (temb [B, 6, dim]; no added_kv_proj)."""

    def __init__(self, sd, prefix, num_heads, head_dim, quant_type=torch.float8_e4m3fn, device="cuda",
                 cross_attn_norm=True, eps=1e-6):
        p, q, dv = prefix, quant_type, device
        self.heads, self.hd = num_heads, head_dim
        self.dim = num_heads * head_dim
        self.quant_type = q
        self.eps = eps
        self.qkv = load_linear(sd, [f"{p}.attn1.to_q", f"{p}.attn1.to_k", f"{p}.attn1.to_v"], q, dv)
        self.to_out1 = load_linear(sd, [f"{p}.attn1.to_out.0"], q, dv)
        self.norm_q1 = sd[f"{p}.attn1.norm_q.weight"].to(dv).contiguous()
        self.norm_k1 = sd[f"{p}.attn1.norm_k.weight"].to(dv).contiguous()
        self.to_q2 = load_linear(sd, [f"{p}.attn2.to_q"], q, dv)
        self.to_kv2 = load_linear(sd, [f"{p}.attn2.to_k", f"{p}.attn2.to_v"], q, dv)
        self.to_out2 = load_linear(sd, [f"{p}.attn2.to_out.0"], q, dv)
        self.norm_q2 = sd[f"{p}.attn2.norm_q.weight"].to(dv).contiguous()
        self.norm_k2 = sd[f"{p}.attn2.norm_k.weight"].to(dv).contiguous()
        self.cross_attn_norm = cross_attn_norm
        if cross_attn_norm:
            self.norm2_weight = sd[f"{p}.norm2.weight"].to(dv).to(torch.float32).reshape(1, -1).contiguous()
            self.norm2_bias = sd[f"{p}.norm2.bias"].to(dv).to(torch.float32).reshape(1, -1).contiguous()
        self.ffn = FeedForward(load_linear(sd, [f"{p}.ffn.net.0.proj"], q, dv), load_linear(sd, [f"{p}.ffn.net.2"], q, dv))
        self.scale_shift_table = sd[f"{p}.scale_shift_table"].to(dv)
        self.scale = self.hd ** -0.5
        # optional: self-attention with e4m3 q/k/v and e4m3 P (the reference's fp8 attention semantics,
        # csrc/attention/interface.cu:262-270: unit descales, P quantised unscaled); single-GPU path
        self.fp8_attention = False

    @staticmethod
    def merge_rotary(rotary_emb, dtype):
        """cos/sin [1, N, 1, hd] pair -> the [N, hd] cos||sin table (layer/transformer.py:497-498)."""
        cos, sin = rotary_emb
        return torch.cat((cos.squeeze()[:, 0::2], sin.squeeze()[:, 1::2]), dim=-1).to(dtype).contiguous()

    def self_attention_qkv(self, xq, B, S, rope_table, pos0=0):
        """qkv projection + across-heads q/k RMSNorm + RoPE, in place (transformer.py:486-499)."""
        d, H, hd = self.dim, self.heads, self.hd
        qkv = self.qkv.forward(xq).view(B, S, 3 * d)
        for b in range(B):
            ops.qk_norm_rope_(qkv[b], self.norm_q1, self.norm_k1, rope_table, H, H, hd, 0, d, pos0, self.eps,
                              across_heads=True)
        return qkv

    def _qkv_group(self, g: int) -> QLinear:
        """Column group g (0 = q, 1 = k, 2 = v) of the fused qkv projection, sharing its storage."""
        if not hasattr(self, "_groups"):
            self._groups = []
            d = self.dim
            for i in range(3):
                lin = QLinear(self.qkv.in_features, d, bias=self.qkv.bias is not None, data_type=self.qkv.dtype,
                              device_type=self.qkv.device)
                lin.weight = self.qkv.weight[:, i * d:(i + 1) * d]
                lin.weight_quant_scale = self.qkv.weight_quant_scale[i * d:(i + 1) * d]
                if self.qkv.weight_asym_sumcol is not None:
                    lin.weight_asym_sumcol = self.qkv.weight_asym_sumcol[:, i * d:(i + 1) * d].contiguous()
                lin.bias = None if self.qkv.bias is None else self.qkv.bias[i * d:(i + 1) * d]
                self._groups.append(lin)
        return self._groups[g]

    def forward(self, hidden_states, encoder_hidden_states, temb, rotary_emb, sparse_mask=None,
                block_q=128, block_k=64, ulysses=None, pos0=0, overlap=True):
        """`ulysses` (fastdm_b200.ulysses.UlyssesAttention): hidden_states is this rank's token shard,
        `pos0` its first position in the full sequence; `rotary_emb` covers the full sequence."""
        B, S, d = hidden_states.shape
        H, hd, qt = self.heads, self.hd, self.quant_type
        if ulysses is not None and ulysses.P > 1 and B != 1:
            # every rank raises (B is the same on all of them), so no rank is left waiting in a collective
            raise NotImplementedError("WanTransformerBlock: the Ulysses path is written for batch 1 "
                                      "(CFG halves are separate forwards); got batch %d" % B)
        shift_msa, scale_msa, gate_msa, c_shift_msa, c_scale_msa, c_gate_msa = (
            self.scale_shift_table + temb.float()).chunk(6, dim=1)                      # wan.py:88-91 (fp32)
        r2 = lambda t: t.reshape(B, d).contiguous()  # noqa: E731
        hid2 = hidden_states.reshape(B * S, d)
        rope_table = rotary_emb if torch.is_tensor(rotary_emb) else self.merge_rotary(rotary_emb, hidden_states.dtype)

        # 1. self-attention: (norm1(x) * (1 + scale) + shift) in fp32, one rounding (wan.py:95)
        xq = Quantized(*ops.layernorm_modulate_quant(hid2, r2(1 + scale_msa), r2(shift_msa), S, qt, self.eps,
                                                     round_steps=False)[:3])
        if ulysses is None or ulysses.P == 1:
            qkv = self.self_attention_qkv(xq, B, S, rope_table, pos0)
            if self.fp8_attention and sparse_mask is None:
                qkv = qkv.to(torch.float8_e4m3fn)      # one cast pass over the fused buffer (unit scales)
            attn = ops.attention(qkv[:, :, :d], qkv[:, :, d:2 * d], qkv[:, :, 2 * d:], H, hd, self.scale,
                                 sparse_mask, block_q, block_k)
        elif overlap:
            norm_w = (self.norm_q1, self.norm_k1, None)

            def group(g):
                def run():
                    x = self._qkv_group(g).forward(xq)
                    if g < 2:
                        ops.qk_norm_rope_(x, norm_w[g], None, rope_table, H, 0, hd, 0, 0, pos0, self.eps,
                                          across_heads=True)
                    return x
                return run

            attn = ulysses.qkv_projection_overlapped([group(0), group(1), group(2)], self.scale, sparse_mask,
                                                     block_q, block_k)
        else:
            qkv = self.self_attention_qkv(xq, B, S, rope_table, pos0)
            attn = ulysses.attention(qkv, self.scale, sparse_mask, block_q, block_k)
        h1 = torch.empty_like(hid2)
        # hidden = (hidden.float() + attn_out * gate).type_as(hidden)      (wan.py:97)
        self.to_out1.forward(quantize(attn.view(B * S, d), qt), gate=r2(gate_msa), residual=hid2, rows_per_batch=S,
                             round_steps=False, out=h1)

        # 2. cross-attention (wan.py:100-105); K/V come from the (replicated) text tokens
        if self.cross_attn_norm:
            nq = Quantized(*ops.layernorm_modulate_quant(h1, self.norm2_weight, self.norm2_bias, B * S, qt, self.eps,
                                                         round_steps=False)[:3])
        else:
            nq = quantize(h1, qt)
        q2 = self.to_q2.forward(nq)
        ops.qk_norm_rope_(q2, self.norm_q2, None, None, H, 0, hd, 0, 0, 0, self.eps, across_heads=True)
        T = encoder_hidden_states.shape[1]
        kv = self.to_kv2.forward(encoder_hidden_states.reshape(B * T, -1))
        ops.qk_norm_rope_(kv, self.norm_k2, None, None, H, 0, hd, 0, 0, 0, self.eps, across_heads=True)
        kv = kv.view(B, T, 2 * d)
        attn2 = ops.attention(q2.view(B, S, d), kv[:, :, :d], kv[:, :, d:], H, hd, self.scale)
        h2 = torch.empty_like(hid2)
        self.to_out2.forward(quantize(attn2.view(B * S, d), qt), residual=h1, rows_per_batch=S, out=h2)  # wan.py:105

        # 3. feed-forward (wan.py:108-112)
        nq = Quantized(*ops.layernorm_modulate_quant(h2, r2(1 + c_scale_msa), r2(c_shift_msa), S, qt, self.eps,
                                                     round_steps=False)[:3])
        out = torch.empty_like(hid2)
        self.ffn.forward(nq, gate=r2(c_gate_msa), residual=h2, rows_per_batch=S, round_steps=False, out=out)
        return out.view(B, S, d)
