"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by fastdm_b200/).

CPU restatement, in plain PyTorch, of the arithmetic of FastDM's `torch` kernel backend for the
DiT block hot path. Each function cites the reference lines it follows. Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Pinning: oracle/gen_golden.py runs the REAL reference (imported from /root/reference with the
shims of oracle/reference_shim.py) on seeded inputs, stores its outputs under tests/golden/, and
tests/test_oracle_golden.py checks every function below against those fixtures bit for bit
(integer / byte results) or to the stated tolerance (floating point).

Two ops have no in-repo reference arithmetic and are therefore "parity unpinned" beyond the
dense case: `sdpa_sparse` (third-party spas_sage_attn, un-vendored,
un-pinned: Dockerfile:31) and the fp8 attention variant (dead code in the reference,
csrc/attention/interface.cu:169-293). Their restatements below define the semantics we build to.
"""
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


# --- a1: fastdm/kernel/torch/quantize.py:45-67 (== fastdm/utils/quantization.py:42-63) -----------
def quantize_to_fp8(input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    assert input.ndim == 2
    finfo = torch.finfo(torch.float8_e4m3fn)
    row_min = input.min(dim=1).values
    row_max = input.max(dim=1).values
    abs_max = torch.max(torch.abs(row_min), torch.abs(row_max)).clamp(min=1e-12)
    scale = abs_max.float() / finfo.max
    q = (input.float() / scale[:, None]).clamp(min=finfo.min, max=finfo.max).to(torch.float8_e4m3fn)
    return q, scale.unsqueeze(-1)


# --- a2: fastdm/kernel/torch/quantize.py:7-43 (== fastdm/utils/quantization.py:5-41) -------------
def quantize_to_int8(input: torch.Tensor, symmetric: bool = True
                     ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    assert input.ndim == 2
    x = input.float()
    row_min = x.min(dim=1).values
    row_max = x.max(dim=1).values
    lo, hi = -128, 127
    if symmetric:
        abs_max = torch.max(torch.abs(row_min), torch.abs(row_max))
        scales = abs_max / hi
        q = torch.clamp(torch.round(x / scales[:, None]), lo, hi).to(torch.int8)
        zp = None
    else:
        scales = (row_max - row_min) / (hi - lo)
        zp = (lo - torch.round(row_min / scales)).to(torch.int32)
        q = torch.clamp(torch.round(x / scales[:, None] + zp.float()[:, None]), lo, hi).to(torch.int8)
    return q, scales.unsqueeze(-1), (zp.unsqueeze(-1) if zp is not None else None)


# --- a3: fastdm/kernel/torch/matrixmul.py:7-35 ---------------------------------------------------
# The reference calls torch._scaled_mm(a, b, scale_a, scale_b.T, bias, out_dtype): fp32 accumulate,
# (acc * sA * sB + bias) rounded once. CPU torch._scaled_mm rejects per-row scales, so the identical
# arithmetic is written out (SURVEY.md 8(c) shim 3).
def fp8_matmul(a, b, scale_a, scale_b, out_dtype, bias=None):
    assert b.shape[0] % 16 == 0 and b.shape[1] % 16 == 0
    if a.is_cuda:
        # on a GPU the reference's own call (matrixmul.py:33) runs as is: this is the "torch backend on the
        # same B200" arm of bench.py's gpu_torch_baseline leg (cuBLASLt rowwise-scaled fp8 GEMM)
        return torch._scaled_mm(a, b, scale_a, scale_b.T, bias, out_dtype=out_dtype)
    acc = a.float() @ b.float()
    out = acc * scale_a.reshape(-1, 1) * scale_b.reshape(1, -1)
    if bias is not None:
        out = out + bias.float()
    return out.to(out_dtype)


# --- a4: fastdm/kernel/torch/matrixmul.py:37-74 --------------------------------------------------
def int8_matmul(a, b, scale_a, scale_b, out_dtype, azp_adj, azp, bias=None):
    assert b.shape[0] % 16 == 0 and b.shape[1] % 16 == 0
    mm = a.float() @ b.float()
    zp_mm = azp.float() @ azp_adj.float()
    scale_c = scale_a.expand(scale_a.size(0), scale_b.size(0)) * \
        scale_b.transpose(0, 1).expand(scale_a.size(0), scale_b.size(0))
    out = ((mm - zp_mm) * scale_c).to(out_dtype)
    return out + bias if bias is not None else out


# --- a6: fastdm/kernel/torch/norm.py:5-27 --------------------------------------------------------
def rms_norm(input: torch.Tensor, scale: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    input_dtype = input.dtype
    variance = input.to(torch.float32).pow(2).mean(-1, keepdim=True)
    input = input * torch.rsqrt(variance + eps)
    if scale is not None:
        input = input.to(scale.dtype)
        return input * scale
    return input.to(input_dtype)


# --- a7: fastdm/kernel/torch/rotemb.py:5-64 (in place, returns None) -----------------------------
def rotary_pos_embedding(query, key, head_size, cos_sin_cache, is_neox=False):
    def rot(x, cos, sin):
        cos = cos.unsqueeze(-2).to(x.dtype)
        sin = sin.unsqueeze(-2).to(x.dtype)
        if is_neox:
            x1, x2 = torch.chunk(x, 2, dim=-1)
        else:
            x1 = x[..., ::2]
            x2 = x[..., 1::2]
        o1 = x1 * cos - x2 * sin
        o2 = x2 * cos + x1 * sin
        if is_neox:
            return torch.cat((o1, o2), dim=-1)
        return torch.stack((o1, o2), dim=-1).flatten(-2)

    pos = torch.arange(query.shape[1], device=query.device)
    cos, sin = cos_sin_cache.index_select(0, pos).chunk(2, dim=-1)
    qs, ks = query.shape, key.shape
    q_rot = rot(query.view(qs[0], qs[1], -1, head_size), cos, sin)
    k_rot = rot(key.view(ks[0], ks[1], -1, head_size), cos, sin)
    query.copy_(q_rot.reshape(qs))
    key.copy_(k_rot.reshape(ks))
    return


# --- a8: fastdm/kernel/torch/gelumul.py:4-16 -----------------------------------------------------
def gelu_and_mul(x: torch.Tensor) -> torch.Tensor:
    x1, x2 = x.chunk(2, dim=-1)
    return x1 * F.gelu(x2)


# --- a9: fp32 reference of tests/test_attention.py:23-63 (what the reference tests compare to) ---
def attention_ref(q, k, v, scale: Optional[float] = None, block_mask=None, mask_bq=128, mask_bk=64,
                  out_dtype=None):
    """q [B,Sq,H,hd], k/v [B,Sk,H,hd] -> [B,Sq,H,hd]; fp32 math, output cast to the input dtype.
    block_mask [B,H,ceil(Sq/bq),ceil(Sk/bk)]: 0 = block excluded from the softmax (sdpa_sparse
    semantics, fastdm/kernel/operators_set.py:181-208; rows whose blocks are all 0 produce 0)."""
    dt = out_dtype or (q.dtype if q.dtype in (torch.bfloat16, torch.float16, torch.float32) else torch.bfloat16)
    b, sq, h, d = q.shape
    sk = k.shape[1]
    if scale is None:
        scale = d ** -0.5
    qf = q.float().transpose(1, 2)
    kf = k.float().transpose(1, 2)
    vf = v.float().transpose(1, 2)
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale  # [B,H,Sq,Sk]
    if block_mask is not None:
        m = block_mask.bool()
        m = m.repeat_interleave(mask_bq, dim=2)[:, :, :sq].repeat_interleave(mask_bk, dim=3)[:, :, :, :sk]
        s = s.masked_fill(~m, float("-inf"))
        p = torch.softmax(s, dim=-1)
        p = torch.nan_to_num(p, nan=0.0)  # fully masked rows
    else:
        p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vf).transpose(1, 2)
    return o.to(dt)


# --- a9 as the torch backend computes it: fastdm/kernel/torch/attention.py:7-43 ------------------
def scaled_dot_product_attention(query, key, value, num_q_heads, num_kv_heads, head_dim, is_causal=False,
                                 scale=None):
    b, t, c = query.size()
    q = query.view(b, t, num_q_heads, head_dim).transpose(1, 2)
    k = key.view(key.size(0), key.size(1), num_kv_heads, head_dim).transpose(1, 2)
    v = value.view(value.size(0), value.size(1), num_kv_heads, head_dim).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=is_causal, scale=scale)
    return o.transpose(1, 2).contiguous().view(b, t, c)


# --- a11: sdpa_sparse -- no in-repo arithmetic (kernel/torch/attention.py:72 raises); semantics of
# the op docstring (operators_set.py:181-208) + padding rule of kernel/cuda/attention.py:97-103 ----
def sparse_scaled_dot_product_attention(query, key, value, num_q_heads, num_kv_heads, head_dim, is_causal=False,
                                        scale=None, sparse_mask=None, block_q=128, block_k=64):
    b, t, c = query.shape
    q = query.view(b, t, num_q_heads, head_dim)
    k = key.view(b, key.shape[1], num_kv_heads, head_dim)
    v = value.view(b, value.shape[1], num_kv_heads, head_dim)
    o = attention_ref(q, k, v, scale, sparse_mask, block_q, block_k)
    return o.reshape(b, t, c)


# --- a10: fp8 attention (csrc/attention/interface.cu:262-270: per-tensor descale 1.0; P -> e4m3
# unscaled: mainloop_fwd_sm90_tma_gmma_ws.hpp:954-955). PARITY UNPINNED: no caller, no test. -------
def attention_fp8_ref(q, k, v, scale: Optional[float] = None):
    """q,k,v float8_e4m3fn [B,S,H,hd] -> bf16. exp(s - rowmax) is quantised to e4m3 before P.V,
    the row sum uses the unquantised probabilities (as the FA3-derived kernel does)."""
    b, sq, h, d = q.shape
    if scale is None:
        scale = d ** -0.5
    qf, kf, vf = (t.float().transpose(1, 2) for t in (q, k, v))
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    p = torch.exp(s - s.amax(dim=-1, keepdim=True))
    denom = p.sum(dim=-1, keepdim=True)
    pq = p.to(torch.float8_e4m3fn).float()
    o = torch.matmul(pq, vf) / denom
    return o.transpose(1, 2).to(torch.bfloat16)
