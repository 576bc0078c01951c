"""Drop-in for the reference's pybind module `fastdm.cuda_ops` (csrc/torch_bindings.cpp:191-201):
the same seven function names, argument order and in-place / allocating conventions, implemented
on libfastdm_b200.so. `fastdm_b200.integration.install()` puts this module at
sys.modules["fastdm.cuda_ops"], after which FastDM's own fastdm/kernel/cuda/*.py wrappers run
unmodified on a B200 (the reference extension has no sm_100 build: setup.py:19-91).
"""
from typing import Optional

import torch

from . import _lib, ops
from ._lib import FDM_BF16, FDM_E4M3, FDM_F16, FDM_F32

_DT = {torch.bfloat16: FDM_BF16, torch.float16: FDM_F16, torch.float32: FDM_F32}


_stream, _on = ops._stream, ops._on   # raw current-stream query, device guard that is a no-op on the current device


def fp8_quant_(out: torch.Tensor, input: torch.Tensor, scale: torch.Tensor, scale_ub: Optional[torch.Tensor] = None) -> None:
    """ops.h:12-14 / elmwise_ops.cu:524-547: caller allocates out (e4m3) and scale [M,1] fp32."""
    if scale_ub is not None:
        raise NotImplementedError("fp8_quant_: scale_ub is never passed by FastDM (kernel/cuda/quantize.py:53)")
    if not (input.is_contiguous() and out.is_contiguous()):
        raise RuntimeError("fp8_quant_: input and out must be contiguous (elmwise_ops.cu:528-529)")
    cols = input.shape[-1]
    rows = input.numel() // cols
    with _on(input):
        rc = _lib.load().fdm_quant_fp8(input.data_ptr(), out.data_ptr(), scale.data_ptr(), rows, cols, cols,
                                       _DT[input.dtype], _stream(input))
    _lib.check(rc, "fp8_quant_")


def int8_quant_(out: torch.Tensor, input: torch.Tensor, scales: torch.Tensor, azp: Optional[torch.Tensor]) -> None:
    """ops.h:9-11 / elmwise_ops.cu:402-431: azp None => symmetric."""
    if not (input.is_contiguous() and out.is_contiguous()):
        raise RuntimeError("int8_quant_: input and out must be contiguous")
    cols = input.shape[-1]
    rows = input.numel() // cols
    with _on(input):
        rc = _lib.load().fdm_quant_int8(input.data_ptr(), out.data_ptr(), scales.data_ptr(),
                                        None if azp is None else azp.data_ptr(), rows, cols, cols,
                                        _DT[input.dtype], _stream(input))
    _lib.check(rc, "int8_quant_")


def rms_norm_(out: torch.Tensor, input: torch.Tensor, weight: torch.Tensor, epsilon: float) -> None:
    """ops.h:15-18 / elmwise_ops.cu:433-449: normalises over input.size(-1) into the caller's out."""
    cols = input.shape[-1]
    rows = input.numel() // cols
    with _on(input):
        rc = _lib.load().fdm_rms_norm(input.data_ptr(), out.data_ptr(), None if weight is None else weight.data_ptr(),
                                      rows, cols, cols, cols, float(epsilon), _DT[input.dtype], _stream(input))
    _lib.check(rc, "rms_norm_")


def rotary_emb_(positions: torch.Tensor, query: torch.Tensor, key: torch.Tensor, head_size: int,
                cos_sin_cache: torch.Tensor, is_neox: bool) -> None:
    """ops.h:20-32 / elmwise_ops.cu:451-522, in place. FastDM always passes positions = arange(seq) per
    batch row (kernel/cuda/rotemb.py:36); anything else is rejected rather than silently ignored."""
    seq = query.shape[1]
    if positions.shape[-1] != seq:
        raise RuntimeError("rotary_emb_: positions must be [batch, seq]")
    ops.rotary_pos_embedding(query, key, head_size, cos_sin_cache, is_neox)


def fp8_scaled_mm_(a, b, scales_a, scales_b, out_dtype, bias=None) -> torch.Tensor:
    """torch_bindings.cpp:24-84."""
    return ops.fp8_matmul(a, b, scales_a, scales_b, out_dtype, bias)


def int8_scaled_mm_(a, b, scales_a, scales_b, out_dtype, azp_adj, azp, bias=None) -> torch.Tensor:
    """torch_bindings.cpp:86-160."""
    return ops.int8_matmul(a, b, scales_a, scales_b, out_dtype, azp_adj, azp, bias)


def flash_attention_fp8_fwd_(q, k, v, softmax_scale: float, is_causal: bool) -> torch.Tensor:
    """torch_bindings.cpp:162-189: q/k/v [B,S,H,hd] e4m3 -> bf16 [B,S,H,hd]. (Never called from FastDM's
    Python: SURVEY.md 0.4.)"""
    if is_causal:
        raise NotImplementedError("flash_attention_fp8_fwd_: causal masking is not on the DiT hot path")
    b, s, h, hd = q.shape
    out = ops.attention(q.reshape(b, s, h * hd), k.reshape(b, k.shape[1], h * hd), v.reshape(b, v.shape[1], h * hd),
                        h, hd, softmax_scale)
    return out.view(b, s, h, hd)
