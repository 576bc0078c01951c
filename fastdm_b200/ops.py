"""Torch-facing wrappers of the C ABI: each op is a `torch.library.custom_op` in the
`fastdm_b200::` namespace that checks/allocates exactly as the reference's pybind layer and cuda
backend did (csrc/torch_bindings.cpp:24-189, fastdm/kernel/cuda/*.py), then calls
libfastdm_b200.so on the current CUDA stream.

Names, argument meaning and error behaviour of the public functions at the bottom mirror
fastdm/kernel/operators_set.py (reference), so they can be registered under `(op, "cuda")` in
FastDM's kernel registry (see fastdm_b200/integration.py).
"""
from typing import Optional, Tuple

import contextlib
import os

import torch

from . import _lib
from ._lib import ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, FDM_BF16, FDM_E4M3, FDM_F16, FDM_F32, FDM_S8

_DT = {torch.bfloat16: FDM_BF16, torch.float16: FDM_F16, torch.float32: FDM_F32,
       torch.float8_e4m3fn: FDM_E4M3, torch.int8: FDM_S8}
_ACT = {None: ACT_NONE, "none": ACT_NONE, "gelu_tanh": ACT_GELU_TANH, "gelu-approximate": ACT_GELU_TANH,
        "gelu_erf": ACT_GELU_ERF, "gelu": ACT_GELU_ERF}


class _Ops:
    """`_ops.<name>` for traced / compiled code, the op's Python body directly otherwise: the
    dispatcher round trip of a `torch.library.custom_op` costs ~30 us per call, more than most of these kernels run
    for on the text streams (a Qwen-Image step issues 1100 of them). FDM_OPS_DISPATCH=1 forces the dispatcher."""

    def __init__(self):
        self._direct = {}
        self._always_dispatch = os.environ.get("FDM_OPS_DISPATCH", "0") == "1"

    def register(self, name, custom_op_def):
        self._direct[name] = custom_op_def._init_fn

    def __getattr__(self, name):
        if self._always_dispatch or torch.compiler.is_compiling():
            return getattr(torch.ops.fastdm_b200, name)
        try:
            return self._direct[name]
        except KeyError:
            raise AttributeError(name) from None


_ops = _Ops()


def _dt(t: torch.Tensor, what: str) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"fastdm_b200.{what}: unsupported dtype {t.dtype}") from None


def _cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"fastdm_b200.{what}: expected a CUDA tensor, got {t.device} "
                           "(there is no CPU implementation)")


# Host cost per call matters on the text streams and at 8-way sequence parallelism (a Qwen-Image step issues ~1100
# launches of 5-40 us): `torch.cuda.current_stream(device).cuda_stream` and a `torch.cuda.device(...)` context were 13 of
# the ~19 us a call took on the host (tools/host_overhead_profile.py); the raw-stream query and a guard that only
# switches when the tensor lives on another device than the current one cost well under 1 us.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)
_NO_GUARD = contextlib.nullcontext()


def _stream(t: torch.Tensor) -> int:
    if _raw_stream is not None:
        return _raw_stream(t.device.index)
    return torch.cuda.current_stream(t.device).cuda_stream


def _on(t: torch.Tensor):
    """Device guard for the launch: a no-op when `t` lives on the current device."""
    if _cur_device is not None and t.device.index == _cur_device():
        return _NO_GUARD
    return torch.cuda.device(t.device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# quantisation
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("fastdm_b200::quant_fp8", mutates_args=())
def _quant_fp8(x: torch.Tensor, act: int) -> Tuple[torch.Tensor, torch.Tensor]:
    _cuda(x, "quantize_to_fp8")
    if x.ndim != 2:
        raise RuntimeError("fastdm_b200.quantize_to_fp8: input must be 2-D [tokens, channels]")
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    out = torch.empty((rows, cols), device=x.device, dtype=torch.float8_e4m3fn)
    scale = torch.empty((rows, 1), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    with _on(x):
        if act == ACT_NONE:
            rc = lib.fdm_quant_fp8(x.data_ptr(), out.data_ptr(), scale.data_ptr(), rows, cols,
                                   x.stride(0) if rows > 1 else cols, _dt(x, "quantize_to_fp8"), _stream(x))
        else:
            rc = lib.fdm_gelu_quant(x.data_ptr(), out.data_ptr(), scale.data_ptr(), None, rows, cols,
                                    x.stride(0) if rows > 1 else cols, act, _dt(x, "gelu_quant"), FDM_E4M3, _stream(x))
    _lib.check(rc, "quantize_to_fp8")
    return out, scale


@_quant_fp8.register_fake
def _(x, act):
    return (x.new_empty(x.shape, dtype=torch.float8_e4m3fn), x.new_empty((x.shape[0], 1), dtype=torch.float32))


@torch.library.custom_op("fastdm_b200::quant_int8", mutates_args=())
def _quant_int8(x: torch.Tensor, symmetric: bool, act: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _cuda(x, "quantize_to_int8")
    if x.ndim != 2:
        raise RuntimeError("fastdm_b200.quantize_to_int8: input must be 2-D [tokens, channels]")
    if x.stride(1) != 1:
        x = x.contiguous()
    rows, cols = x.shape
    out = torch.empty((rows, cols), device=x.device, dtype=torch.int8)
    scale = torch.empty((rows, 1), device=x.device, dtype=torch.float32)
    # symmetric: an empty azp tensor stands for the reference's `None`
    azp = torch.empty((0 if symmetric else rows, 1), device=x.device, dtype=torch.int32)
    lib = _lib.load()
    stride = x.stride(0) if rows > 1 else cols
    with _on(x):
        if act == ACT_NONE:
            rc = lib.fdm_quant_int8(x.data_ptr(), out.data_ptr(), scale.data_ptr(),
                                    None if symmetric else azp.data_ptr(), rows, cols, stride,
                                    _dt(x, "quantize_to_int8"), _stream(x))
        else:
            if symmetric:
                raise RuntimeError("fastdm_b200.gelu_quant: int8 output is asymmetric only")
            rc = lib.fdm_gelu_quant(x.data_ptr(), out.data_ptr(), scale.data_ptr(), azp.data_ptr(), rows, cols,
                                    stride, act, _dt(x, "gelu_quant"), FDM_S8, _stream(x))
    _lib.check(rc, "quantize_to_int8")
    return out, scale, azp


@_quant_int8.register_fake
def _(x, symmetric, act):
    return (x.new_empty(x.shape, dtype=torch.int8), x.new_empty((x.shape[0], 1), dtype=torch.float32),
            x.new_empty((0 if symmetric else x.shape[0], 1), dtype=torch.int32))


# ------------------------------------------------------------------------------------------------
# norms / rope / activation
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("fastdm_b200::rms_norm", mutates_args=())
def _rms_norm(x: torch.Tensor, weight: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    _cuda(x, "rms_norm")
    cols = x.shape[-1]
    if weight is not None:
        if weight.numel() != cols:
            raise RuntimeError(f"fastdm_b200.rms_norm: weight has {weight.numel()} elements, last dim is {cols}")
        weight = weight.contiguous()
    # rows of the last dimension; accept any view whose leading dims collapse to one stride
    xc = x if x.is_contiguous() else x.contiguous()  # callers always pass contiguous (layer/transformer.py:275-290)
    out = torch.empty_like(xc)
    rows = xc.numel() // cols if cols else 0
    with _on(x):
        rc = _lib.load().fdm_rms_norm(xc.data_ptr(), out.data_ptr(), _ptr(weight), rows, cols, cols, cols,
                                      float(eps), _dt(x, "rms_norm"), _stream(x))
    _lib.check(rc, "rms_norm")
    return out


@_rms_norm.register_fake
def _(x, weight, eps):
    return torch.empty_like(x, memory_format=torch.contiguous_format)


@torch.library.custom_op("fastdm_b200::rope_", mutates_args=("query", "key"))
def _rope(query: torch.Tensor, key: torch.Tensor, head_size: int, cos_sin_cache: torch.Tensor,
          is_neox: bool) -> None:
    _cuda(query, "rotary_pos_embedding")
    if query.ndim != 3 or key.ndim != 3:
        raise RuntimeError("fastdm_b200.rotary_pos_embedding: query/key must be [batch, seq, heads*head_size]")
    if query.stride(2) != 1 or key.stride(2) != 1:
        raise RuntimeError("fastdm_b200.rotary_pos_embedding: last dimension must be contiguous (in-place op)")
    b, s, qd = query.shape
    kb, ks, kd = key.shape
    if (kb, ks) != (b, s):
        raise RuntimeError("fastdm_b200.rotary_pos_embedding: query and key must share batch and seq")
    if qd % head_size or kd % head_size:
        raise RuntimeError("fastdm_b200.rotary_pos_embedding: hidden size not a multiple of head_size")
    if cos_sin_cache.ndim != 2 or cos_sin_cache.shape[0] < s or cos_sin_cache.shape[1] != head_size:
        raise RuntimeError("fastdm_b200.rotary_pos_embedding: cos_sin_cache must be [>=seq, head_size]")
    cs = cos_sin_cache
    if cs.dtype != query.dtype or cs.stride(1) != 1:
        cs = cs.to(query.dtype).contiguous()  # reference: cos/sin `.to(x.dtype)` (kernel/torch/rotemb.py:40-41)
    with _on(query):
        rc = _lib.load().fdm_rope(query.data_ptr(), key.data_ptr(), cs.data_ptr(), b, s, qd // head_size,
                                  kd // head_size, head_size, query.stride(0), query.stride(1),
                                  key.stride(0), key.stride(1), cs.stride(0), 1 if is_neox else 0,
                                  _dt(query, "rotary_pos_embedding"), _stream(query))
    _lib.check(rc, "rotary_pos_embedding")


@torch.library.custom_op("fastdm_b200::gelu_and_mul", mutates_args=())
def _gelu_and_mul(x: torch.Tensor) -> torch.Tensor:
    _cuda(x, "gelu_and_mul")
    if x.shape[-1] % 2:
        raise RuntimeError("fastdm_b200.gelu_and_mul: last dimension must be even")
    d = x.shape[-1] // 2
    xc = x if x.is_contiguous() else x.contiguous()
    out = torch.empty(x.shape[:-1] + (d,), device=x.device, dtype=x.dtype)
    rows = xc.numel() // (2 * d) if d else 0
    with _on(x):
        rc = _lib.load().fdm_gelu_and_mul(xc.data_ptr(), out.data_ptr(), rows, d, 2 * d, d,
                                          _dt(x, "gelu_and_mul"), _stream(x))
    _lib.check(rc, "gelu_and_mul")
    return out


@_gelu_and_mul.register_fake
def _(x):
    return x.new_empty(x.shape[:-1] + (x.shape[-1] // 2,))


# ------------------------------------------------------------------------------------------------
# W8A8 GEMMs
# ------------------------------------------------------------------------------------------------
def _check_mm(a, b, scale_a, scale_b, out_dtype, bias, what, qdtype):
    # the checks of csrc/torch_bindings.cpp:31-58 / :93-123
    _cuda(a, what)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[0]:
        raise RuntimeError(f"fastdm_b200.{what}: shapes {tuple(a.shape)} x {tuple(b.shape)} do not multiply")
    if a.dtype != qdtype or b.dtype != qdtype:
        raise RuntimeError(f"fastdm_b200.{what}: a and b must be {qdtype}")
    if a.stride(1) != 1:
        raise RuntimeError(f"fastdm_b200.{what}: a must be row-major")
    if b.stride(0) != 1:
        raise RuntimeError(f"fastdm_b200.{what}: b must be column-major (b.stride(0) == 1)")
    m, k = a.shape
    n = b.shape[1]
    if scale_a.numel() != m or scale_b.numel() != n:
        raise RuntimeError(f"fastdm_b200.{what}: scale_a must have M and scale_b N elements")
    if scale_a.dtype != torch.float32 or scale_b.dtype != torch.float32:
        raise RuntimeError(f"fastdm_b200.{what}: scales must be float32")
    if not (scale_a.is_contiguous() and scale_b.is_contiguous()):
        raise RuntimeError(f"fastdm_b200.{what}: scales must be contiguous")
    if out_dtype not in (torch.bfloat16, torch.float16):
        raise RuntimeError(f"fastdm_b200.{what}: out_dtype must be bfloat16 or float16")
    if bias is not None:
        if bias.numel() != n or not bias.is_contiguous() or bias.dtype != out_dtype:
            raise RuntimeError(f"fastdm_b200.{what}: bias must be contiguous [N] of out_dtype")
    return m, n, k


def _check_fused(out, m, n, out_dtype, gate, residual, rows_per_batch, what):
    if out.shape != (m, n) or out.dtype != out_dtype or out.stride(1) != 1 or not out.is_cuda:
        raise RuntimeError(f"fastdm_b200.{what}: out must be a [M, N] {out_dtype} CUDA tensor with unit column stride")
    if gate is not None:
        if gate.dtype != torch.float32 or gate.ndim != 2 or gate.shape[1] != n or not gate.is_contiguous():
            raise RuntimeError(f"fastdm_b200.{what}: gate must be contiguous float32 [batches, N]")
        if rows_per_batch <= 0 or gate.shape[0] * rows_per_batch < m:
            raise RuntimeError(f"fastdm_b200.{what}: gate rows * rows_per_batch must cover M")
    if residual is not None:
        if residual.shape != (m, n) or residual.dtype != out_dtype or residual.stride(1) != 1:
            raise RuntimeError(f"fastdm_b200.{what}: residual must be [M, N] of out_dtype with unit column stride")


@torch.library.custom_op("fastdm_b200::gemm_fp8_", mutates_args=("out",))
def _gemm_fp8(out: torch.Tensor, a: torch.Tensor, b: torch.Tensor, scale_a: torch.Tensor, scale_b: torch.Tensor,
              bias: Optional[torch.Tensor], act: int, gate: Optional[torch.Tensor],
              residual: Optional[torch.Tensor], rows_per_batch: int, round_steps: bool) -> None:
    m, n, k = _check_mm(a, b, scale_a, scale_b, out.dtype, bias, "fp8_matmul", torch.float8_e4m3fn)
    _check_fused(out, m, n, out.dtype, gate, residual, rows_per_batch, "fp8_matmul")
    with _on(a):
        rc = _lib.load().fdm_gemm_fp8_residual(
            a.data_ptr(), b.data_ptr(), scale_a.data_ptr(), scale_b.data_ptr(), _ptr(bias), out.data_ptr(), m, n, k,
            a.stride(0) if m > 1 else k, b.stride(1) if n > 1 else k, out.stride(0) if m > 1 else n,
            _DT[out.dtype], act, _ptr(gate), _ptr(residual),
            (residual.stride(0) if m > 1 else n) if residual is not None else 0, max(rows_per_batch, 1),
            1 if round_steps else 0, _stream(a))
    _lib.check(rc, "fp8_matmul")


@torch.library.custom_op("fastdm_b200::gemm_int8_", mutates_args=("out",))
def _gemm_int8(out: torch.Tensor, a: torch.Tensor, b: torch.Tensor, scale_a: torch.Tensor, scale_b: torch.Tensor,
               azp_adj: Optional[torch.Tensor], azp: Optional[torch.Tensor], bias: Optional[torch.Tensor], act: int,
               gate: Optional[torch.Tensor], residual: Optional[torch.Tensor], rows_per_batch: int,
               round_steps: bool) -> None:
    m, n, k = _check_mm(a, b, scale_a, scale_b, out.dtype, bias, "int8_matmul", torch.int8)
    _check_fused(out, m, n, out.dtype, gate, residual, rows_per_batch, "int8_matmul")
    if (azp is None) != (azp_adj is None):
        raise RuntimeError("fastdm_b200.int8_matmul: azp and azp_adj must be given together")
    if azp is not None:
        if azp.numel() != m or azp_adj.numel() != n:
            raise RuntimeError("fastdm_b200.int8_matmul: azp must have M and azp_adj N elements")
        if azp.dtype != torch.int32 or azp_adj.dtype != torch.int32:
            raise RuntimeError("fastdm_b200.int8_matmul: azp / azp_adj must be int32")
        if not (azp.is_contiguous() and azp_adj.is_contiguous()):
            raise RuntimeError("fastdm_b200.int8_matmul: azp / azp_adj must be contiguous")
    with _on(a):
        rc = _lib.load().fdm_gemm_int8_residual(
            a.data_ptr(), b.data_ptr(), scale_a.data_ptr(), scale_b.data_ptr(), _ptr(azp_adj), _ptr(azp), _ptr(bias),
            out.data_ptr(), m, n, k, a.stride(0) if m > 1 else k, b.stride(1) if n > 1 else k,
            out.stride(0) if m > 1 else n, _DT[out.dtype], act, _ptr(gate), _ptr(residual),
            (residual.stride(0) if m > 1 else n) if residual is not None else 0, max(rows_per_batch, 1),
            1 if round_steps else 0, _stream(a))
    _lib.check(rc, "int8_matmul")


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("fastdm_b200::attn_fwd_", mutates_args=("out",))
def _attn_fwd(out: torch.Tensor, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, num_heads: int,
              head_dim: int, scale: float, block_mask: Optional[torch.Tensor], mask_bq: int, mask_bk: int) -> None:
    what = "scaled_dot_product_attention"
    _cuda(query, what)
    if query.ndim != 3 or key.ndim != 3 or value.ndim != 3:
        raise RuntimeError(f"fastdm_b200.{what}: q/k/v must be [batch, seq, heads*head_dim]")
    b, sq, c = query.shape
    sk = key.shape[1]
    if c != num_heads * head_dim or key.shape[2] != c or value.shape[2] != c:
        raise RuntimeError(f"fastdm_b200.{what}: hidden size must be num_heads*head_dim for q, k and v (MHA)")
    if key.shape[0] != b or value.shape[0] != b or value.shape[1] != sk:
        raise RuntimeError(f"fastdm_b200.{what}: batch / kv length mismatch")
    if not (query.dtype == key.dtype == value.dtype):
        raise RuntimeError(f"fastdm_b200.{what}: q/k/v dtypes differ")
    if out.shape != query.shape or out.stride(2) != 1 or out.dtype != _attn_out_dtype(query.dtype):
        raise RuntimeError(f"fastdm_b200.{what}: out must be [batch, seq, heads*head_dim] with unit last stride")
    if out.data_ptr() % 16 or out.stride(1) % 8 or (b > 1 and out.stride(0) % 8):
        raise RuntimeError(f"fastdm_b200.{what}: out must be 16-byte aligned with token / batch strides that are "
                           f"multiples of 8 elements (got data_ptr % 16 = {out.data_ptr() % 16}, strides {out.stride()})")
    ts = []
    for t in (query, key, value):
        # last-dim slices of a fused qkv projection are legal (layer/transformer.py:269,300)
        if t.stride(2) != 1 or (t.stride(1) * t.element_size()) % 16 or (t.stride(0) * t.element_size()) % 16 \
                or t.data_ptr() % 16:
            t = t.contiguous()
        ts.append(t)
    q, k, v = ts
    if block_mask is not None:
        block_mask = block_mask.to(torch.int8)
        nbq, nbk = -(-sq // mask_bq), -(-sk // mask_bk)
        mq, mk = block_mask.shape[-2:] if block_mask.ndim == 4 else (-1, -1)
        if block_mask.ndim != 4 or tuple(block_mask.shape[:2]) != (b, num_heads) or mq not in (sq // mask_bq, nbq) \
                or mk not in (sk // mask_bk, nbk):
            raise RuntimeError(f"fastdm_b200.{what}: sparse_mask must be {(b, num_heads, nbq, nbk)} (or floor-sized: "
                               f"{(b, num_heads, sq // mask_bq, sk // mask_bk)}), got {tuple(block_mask.shape)}")
        if (mq, mk) != (nbq, nbk):
            # the reference's mask builders emit floor-sized masks (sparse/xsparse.py gen_log_mask_shrinked: S // block)
            # and its wrapper pads q/k/v and pads the mask with ones (kernel/cuda/attention.py:118-133): the ragged
            # trailing query / key blocks are computed
            block_mask = torch.nn.functional.pad(block_mask, (0, nbk - mk, 0, nbq - mq), value=1)
        block_mask = block_mask.contiguous()
    with _on(query):
        rc = _lib.load().fdm_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), _ptr(block_mask),
                                      b, sq, sk, num_heads, head_dim,
                                      q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1),
                                      out.stride(0), out.stride(1),
                                      mask_bq, mask_bk, float(scale), _dt(q, what), _stream(q))
    _lib.check(rc, what)


def _attn_out_dtype(dt):
    return torch.float16 if dt == torch.float16 else torch.bfloat16


def attention(query, key, value, num_heads, head_dim, scale=None, block_mask=None, mask_bq=128, mask_bk=64,
              out=None):
    """softmax(scale * Q K^T [+ block mask]) V; `out` may be a column slice of a wider buffer."""
    if scale is None:
        scale = head_dim ** -0.5
    if out is None:
        out = torch.empty(query.shape, device=query.device, dtype=_attn_out_dtype(query.dtype))
    _ops.attn_fwd_(out, query, key, value, num_heads, head_dim, scale, block_mask, mask_bq, mask_bk)
    return out


def attention_scatter(query, key, value, num_heads, head_dim, peer_ptrs, rows_per_peer, out_token_stride, scale=None,
                      block_mask=None, mask_bq=128, mask_bk=64):
    """Attention whose epilogue writes query row r into the buffer of the rank that owns it (Ulysses):
    peer_ptrs[r // rows_per_peer] + (r % rows_per_peer) * out_token_stride elements. `peer_ptrs` are device addresses
    (ints) of the peers' output buffers mapped into this process, already offset to this rank's head columns.
    Nothing is returned: the caller owns the buffers and the cross-rank ordering (fdm_attn_fwd_scatter)."""
    what = "attention_scatter"
    _cuda(query, what)
    if query.ndim != 3 or query.shape[0] != 1 or key.shape[0] != 1 or value.shape[0] != 1:
        raise RuntimeError(f"fastdm_b200.{what}: q/k/v must be [1, seq, heads*head_dim]")
    if not (query.dtype == key.dtype == value.dtype == torch.bfloat16) or head_dim != 128:
        raise RuntimeError(f"fastdm_b200.{what}: built for bf16, head_dim 128")
    _, sq, c = query.shape
    sk = key.shape[1]
    if c != num_heads * head_dim or key.shape[2] != c or value.shape[2] != c or value.shape[1] != sk:
        raise RuntimeError(f"fastdm_b200.{what}: shape mismatch")
    n = len(peer_ptrs)
    if not 1 <= n <= 8 or rows_per_peer * n < sq:
        raise RuntimeError(f"fastdm_b200.{what}: need 1..8 peers covering all {sq} query rows")
    ts = []
    for t in (query, key, value):
        if t.stride(2) != 1 or (t.stride(1) * t.element_size()) % 16 or t.data_ptr() % 16:
            t = t.contiguous()
        ts.append(t)
    q, k, v = ts
    if block_mask is not None:
        block_mask = block_mask.to(torch.int8)
        nbq, nbk = -(-sq // mask_bq), -(-sk // mask_bk)
        if tuple(block_mask.shape) != (1, num_heads, nbq, nbk):
            raise RuntimeError(f"fastdm_b200.{what}: sparse_mask must be {(1, num_heads, nbq, nbk)}")
        block_mask = block_mask.contiguous()
    import ctypes
    arr = (ctypes.c_void_p * n)(*[int(p) for p in peer_ptrs])
    if scale is None:
        scale = head_dim ** -0.5
    with _on(query):
        rc = _lib.load().fdm_attn_fwd_scatter(q.data_ptr(), k.data_ptr(), v.data_ptr(), arr, n, int(rows_per_peer),
                                              _ptr(block_mask), sq, sk, num_heads, head_dim, q.stride(1), k.stride(1),
                                              v.stride(1), int(out_token_stride), mask_bq, mask_bk, float(scale),
                                              _dt(q, what), _stream(q))
    _lib.check(rc, what)


# ------------------------------------------------------------------------------------------------
# fused block ops
# ------------------------------------------------------------------------------------------------
@torch.library.custom_op("fastdm_b200::qk_norm_rope_", mutates_args=("buf",))
def _qk_norm_rope(buf: torch.Tensor, wq: Optional[torch.Tensor], wk: Optional[torch.Tensor],
                  cos_sin: Optional[torch.Tensor], q_heads: int, k_heads: int, head_size: int, q_offset: int,
                  k_offset: int, pos0: int, eps: float, across_heads: bool) -> None:
    what = "qk_norm_rope"
    _cuda(buf, what)
    if buf.ndim != 2 or buf.stride(1) != 1:
        raise RuntimeError(f"fastdm_b200.{what}: buf must be [tokens, features] with unit column stride")
    tokens, width = buf.shape
    if q_offset + q_heads * head_size > width or k_offset + k_heads * head_size > width:
        raise RuntimeError(f"fastdm_b200.{what}: q/k column ranges exceed the buffer width")
    for w, hn in ((wq, q_heads), (wk, k_heads)):
        if w is not None:
            need = head_size * hn if across_heads else head_size
            if w.numel() != need or w.dtype != buf.dtype or not w.is_contiguous():
                raise RuntimeError(f"fastdm_b200.{what}: norm weight must be contiguous [{need}] of the buffer dtype")
    cs = cos_sin
    if cs is not None:
        if cs.ndim != 2 or cs.shape[1] != head_size or cs.shape[0] < pos0 + tokens:
            raise RuntimeError(f"fastdm_b200.{what}: cos_sin must be [>= pos0+tokens, head_size]")
        if cs.dtype != buf.dtype or cs.stride(1) != 1:
            cs = cs.to(buf.dtype).contiguous()
    with _on(buf):
        rc = _lib.load().fdm_qk_norm_rope(buf.data_ptr(), _ptr(wq), _ptr(wk), _ptr(cs), tokens, q_heads, k_heads,
                                          head_size, buf.stride(0) if tokens > 1 else width, q_offset, k_offset, pos0,
                                          cs.stride(0) if cs is not None else 0, float(eps),
                                          1 if across_heads else 0, _dt(buf, what), _stream(buf))
    _lib.check(rc, what)


def qk_norm_rope_(buf, wq, wk, cos_sin, q_heads, k_heads, head_size, q_offset=0, k_offset=None, pos0=0, eps=1e-6,
                  across_heads=False):
    """In place on a fused qkv projection [tokens, >= (q_heads+k_heads)*head_size]: rms_norm(q), rms_norm(k)
    and the interleaved rotary embedding, bit-identical to rms_norm + rotary_pos_embedding."""
    if k_offset is None:
        k_offset = q_offset + q_heads * head_size
    _ops.qk_norm_rope_(buf, wq, wk, cos_sin, q_heads, k_heads, head_size, q_offset, k_offset, pos0,
                                        eps, across_heads)


@torch.library.custom_op("fastdm_b200::layernorm_modulate_quant", mutates_args=())
def _ln_mod_quant(x: torch.Tensor, mul: Optional[torch.Tensor], add: Optional[torch.Tensor], rows_per_batch: int,
                  eps: float, round_steps: bool, out_code: int, want_y: bool
                  ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    what = "layernorm_modulate_quant"
    _cuda(x, what)
    if x.ndim != 2 or x.stride(1) != 1 or x.dtype != torch.bfloat16:
        raise RuntimeError(f"fastdm_b200.{what}: x must be bf16 [rows, cols] with unit column stride")
    rows, cols = x.shape
    mod_dt = None
    for t in (mul, add):
        if t is None:
            continue
        if t.dtype not in (torch.float32, torch.bfloat16) or t.ndim != 2 or t.shape[1] != cols or not t.is_contiguous() \
                or t.shape[0] * rows_per_batch < rows or (mod_dt is not None and t.dtype != mod_dt):
            raise RuntimeError(f"fastdm_b200.{what}: mul/add must be contiguous [batches, cols] covering all rows, both "
                               f"float32 or both bfloat16")
        mod_dt = t.dtype
    if mod_dt == torch.bfloat16 and (not round_steps or cols > 5120):
        # bf16 vectors are the bf16 op chain (round_steps) on the warp-per-row kernel; anything else takes them as fp32
        mul, add = (None if t is None else t.float() for t in (mul, add))
        mod_dt = torch.float32
    dev = x.device
    qdt = {FDM_E4M3: torch.float8_e4m3fn, FDM_S8: torch.int8}.get(out_code)
    q = torch.empty((rows, cols) if qdt is not None else (0, cols), device=dev, dtype=qdt or torch.int8)
    scale = torch.empty((rows if qdt is not None else 0, 1), device=dev, dtype=torch.float32)
    azp = torch.empty((rows if out_code == FDM_S8 else 0, 1), device=dev, dtype=torch.int32)
    y = torch.empty((rows, cols) if want_y else (0, cols), device=dev, dtype=x.dtype)
    with _on(x):
        rc = _lib.load().fdm_layernorm_modulate_quant(
            x.data_ptr(), _ptr(mul), _ptr(add), q.data_ptr() if qdt is not None else None,
            scale.data_ptr() if qdt is not None else None, azp.data_ptr() if out_code == FDM_S8 else None,
            y.data_ptr() if want_y else None, rows, cols, x.stride(0) if rows > 1 else cols, cols,
            max(rows_per_batch, 1), float(eps), 1 if round_steps else 0, FDM_BF16,
            out_code if qdt is not None else FDM_BF16, FDM_BF16 if mod_dt == torch.bfloat16 else FDM_F32, _stream(x))
    _lib.check(rc, what)
    return q, scale, azp, y


@_ln_mod_quant.register_fake
def _(x, mul, add, rows_per_batch, eps, round_steps, out_code, want_y):
    rows, cols = x.shape
    has_q = out_code in (FDM_E4M3, FDM_S8)
    return (x.new_empty((rows if has_q else 0, cols), dtype=torch.float8_e4m3fn if out_code == FDM_E4M3 else torch.int8),
            x.new_empty((rows if has_q else 0, 1), dtype=torch.float32),
            x.new_empty((rows if out_code == FDM_S8 else 0, 1), dtype=torch.int32),
            x.new_empty((rows if want_y else 0, cols)))


def layernorm_modulate_quant(x, mul, add, rows_per_batch, quant_dtype, eps=1e-6, round_steps=True, want_y=False):
    """LayerNorm(x) * mul + add (no affine LN) followed by the per-token quantisation the next
    QLinear would do. Returns (codes, scales, azp-or-None, y-or-None)."""
    code = {torch.float8_e4m3fn: FDM_E4M3, torch.int8: FDM_S8, None: FDM_BF16}[quant_dtype]
    q, s, zp, y = _ops.layernorm_modulate_quant(x, mul, add, rows_per_batch, eps, round_steps, code,
                                                                 want_y or quant_dtype is None)
    return (q if quant_dtype is not None else None, s if quant_dtype is not None else None,
            zp if quant_dtype == torch.int8 else None, y if (want_y or quant_dtype is None) else None)


# ------------------------------------------------------------------------------------------------
# Ulysses layout helpers
# ------------------------------------------------------------------------------------------------
def ulysses_pack_heads(x: torch.Tensor, num_heads: int, head_dim: int, world: int, n_seg: int = 1) -> torch.Tensor:
    """x [S, >= n_seg*H*hd] (n_seg head-major segments side by side, e.g. a fused qkv projection)
    -> send buffer [P, S, n_seg * (H/P) * hd]: chunk p = head group p of every segment."""
    _cuda(x, "ulysses_pack_heads")
    s = x.shape[0]
    c = num_heads * head_dim
    if x.stride(1) != 1:
        x = x.contiguous()
    out = torch.empty((world, s, n_seg * c // world), device=x.device, dtype=x.dtype)
    with _on(x):
        rc = _lib.load().fdm_ulysses_pack_heads(x.data_ptr(), out.data_ptr(), s, num_heads, head_dim, world, n_seg,
                                                x.stride(0) if s > 1 else x.shape[1], c, x.element_size(), _stream(x))
    _lib.check(rc, "ulysses_pack_heads")
    return out


def ulysses_unpack_heads(x: torch.Tensor, num_heads: int, head_dim: int, n_seg: int = 1,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [P, S, n_seg * (H/P) * hd] (chunk p = head group p) -> [S, n_seg * H * hd]."""
    _cuda(x, "ulysses_unpack_heads")
    world, s, _ = x.shape
    c = num_heads * head_dim
    x = x.contiguous()
    if out is None:
        out = torch.empty((s, n_seg * c), device=x.device, dtype=x.dtype)
    with _on(x):
        rc = _lib.load().fdm_ulysses_unpack_heads(x.data_ptr(), out.data_ptr(), s, num_heads, head_dim, world, n_seg,
                                                  out.stride(0) if s > 1 else out.shape[1], c, x.element_size(),
                                                  _stream(x))
    _lib.check(rc, "ulysses_unpack_heads")
    return out


# ------------------------------------------------------------------------------------------------
# step-cache indicator
# ------------------------------------------------------------------------------------------------
def rel_l1_sums(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """One pass over a and b -> fp32 device tensor [sum |T(a - b)|, sum |b|] (T = rounding to the tensor dtype):
    the two reductions of `(a - b).abs().mean() / b.abs().mean()` (fastdm/caching/xcaching.py:214-215)."""
    _cuda(a, "rel_l1_distance")
    if a.shape != b.shape or a.dtype != b.dtype or a.device != b.device:
        raise RuntimeError("fastdm_b200.rel_l1_distance: a and b must have the same shape, dtype and device")
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty(2, device=a.device, dtype=torch.float32)
    with _on(a):
        rc = _lib.load().fdm_rel_l1_distance(a.data_ptr(), b.data_ptr(), a.numel(), _dt(a, "rel_l1_distance"),
                                             out.data_ptr(), _stream(a))
    _lib.check(rc, "rel_l1_distance")
    return out


_ops.register("quant_fp8", _quant_fp8)
_ops.register("quant_int8", _quant_int8)
_ops.register("rms_norm", _rms_norm)
_ops.register("rope_", _rope)
_ops.register("gelu_and_mul", _gelu_and_mul)
_ops.register("gemm_fp8_", _gemm_fp8)
_ops.register("gemm_int8_", _gemm_int8)
_ops.register("attn_fwd_", _attn_fwd)
_ops.register("qk_norm_rope_", _qk_norm_rope)
_ops.register("layernorm_modulate_quant", _ln_mod_quant)


# ================================================================================================
# Public op API -- same names and signatures as fastdm/kernel/operators_set.py (reference)
# ================================================================================================
def rms_norm(input: torch.Tensor, scale: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """operators_set.py:9-21; numerics of kernel/torch/norm.py:5-27.

    A weight of another dtype than the input (not used on the DiT hot path): the reference promotes
    (`input.to(scale.dtype) * scale`, norm.py:21-23) and returns the weight's dtype. Here the kernel runs in the
    input dtype with the weight cast to it and the result is returned in the weight's dtype -- same dtype contract,
    values within one rounding of the input dtype."""
    if scale is not None and scale.dtype != input.dtype:
        return _ops.rms_norm(input, scale.to(input.dtype), eps).to(scale.dtype)
    return _ops.rms_norm(input, scale, eps)


def rotary_pos_embedding(query: torch.Tensor, key: torch.Tensor, head_size: int, cos_sin_cache: torch.Tensor,
                         is_neox: bool = False):
    """operators_set.py:23-52; in place on query and key, returns None (kernel/torch/rotemb.py:62-64)."""
    _ops.rope_(query, key, head_size, cos_sin_cache, is_neox)
    return


def gelu_and_mul(input: torch.Tensor) -> torch.Tensor:
    """operators_set.py:54-67: x[..., :d] * gelu(x[..., d:])."""
    return _ops.gelu_and_mul(input)


def quantize_to_int8(input: torch.Tensor, symmetric: bool = True
                     ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """operators_set.py:69-84; returns (int8 codes, scales [M,1], azp [M,1] int32 or None)."""
    q, s, zp = _ops.quant_int8(input, symmetric, ACT_NONE)
    return q, s, (None if symmetric else zp)


def quantize_to_fp8(input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """operators_set.py:86-100; returns (e4m3 codes, scales [M,1])."""
    return _ops.quant_fp8(input, ACT_NONE)


def fp8_matmul(a: torch.Tensor, b: torch.Tensor, scale_a: torch.Tensor, scale_b: torch.Tensor,
               out_dtype: torch.dtype, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
               out: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
               residual: Optional[torch.Tensor] = None, rows_per_batch: int = 1, round_steps: bool = True
               ) -> torch.Tensor:
    """operators_set.py:102-124. Extras beyond the reference signature (all optional): `act` fuses
    F.gelu into the epilogue, `out` writes into a preallocated (possibly column-sliced) tensor,
    `gate`/`residual` fold `residual + gate * linear(x)` into the epilogue."""
    if b.shape[0] % 16 or b.shape[1] % 16:  # kernel/cuda/matrixmul.py:31
        raise AssertionError("fp8_matmul: K and N must be multiples of 16")
    if out is None:
        out = torch.empty((a.shape[0], b.shape[1]), device=a.device, dtype=out_dtype)
    _ops.gemm_fp8_(out, a, b, scale_a, scale_b, bias, _ACT[act], gate, residual, rows_per_batch,
                                    round_steps)
    return out


def int8_matmul(a: torch.Tensor, b: torch.Tensor, scale_a: torch.Tensor, scale_b: torch.Tensor,
                out_dtype: torch.dtype, azp_adj: Optional[torch.Tensor], azp: Optional[torch.Tensor],
                bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
                out: Optional[torch.Tensor] = None, gate: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None, rows_per_batch: int = 1, round_steps: bool = True
                ) -> torch.Tensor:
    """operators_set.py:126-152 (same optional extras as fp8_matmul)."""
    if b.shape[0] % 16 or b.shape[1] % 16:  # kernel/cuda/matrixmul.py:65
        raise AssertionError("int8_matmul: K and N must be multiples of 16")
    if out is None:
        out = torch.empty((a.shape[0], b.shape[1]), device=a.device, dtype=out_dtype)
    _ops.gemm_int8_(out, a, b, scale_a, scale_b, azp_adj, azp, bias, _ACT[act], gate, residual,
                                     rows_per_batch, round_steps)
    return out


def scaled_dot_product_attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, num_q_heads: int,
                                 num_kv_heads: int, head_dim: int, is_causal: bool = False,
                                 scale: Optional[float] = None) -> torch.Tensor:
    """operators_set.py:154-179. Non-causal MHA only (every hot-path call: layer/transformer.py:149,300)."""
    if is_causal:
        raise NotImplementedError("fastdm_b200.scaled_dot_product_attention: is_causal=True is not on the DiT hot path")
    if num_q_heads != num_kv_heads:
        raise NotImplementedError("fastdm_b200.scaled_dot_product_attention: GQA/MQA is not on the DiT hot path")
    if scale is None:
        scale = head_dim ** -0.5
    return attention(query, key, value, num_q_heads, head_dim, scale)


def sparse_scaled_dot_product_attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                                        num_q_heads: int, num_kv_heads: int, head_dim: int, is_causal: bool = False,
                                        scale: Optional[float] = None, sparse_mask: Optional[torch.Tensor] = None,
                                        block_q: int = 128, block_k: int = 64) -> torch.Tensor:
    """operators_set.py:181-208. sparse_mask [B, H, ceil(Sq/128), ceil(Sk/64)], 1 = compute, 0 = skip
    (non-sm90 block geometry of kernel/cuda/attention.py:92-94)."""
    if is_causal:
        raise NotImplementedError("fastdm_b200.sparse_scaled_dot_product_attention: is_causal=True unsupported")
    if num_q_heads != num_kv_heads:
        raise NotImplementedError("fastdm_b200.sparse_scaled_dot_product_attention: GQA/MQA unsupported")
    if scale is None:
        scale = head_dim ** -0.5
    return attention(query, key, value, num_q_heads, head_dim, scale, sparse_mask, block_q, block_k)


# fused extras used by the host-side layers (fastdm_b200/layers.py)
def gelu_quantize_to_fp8(input: torch.Tensor, approximate: str = "tanh"):
    """quantize_to_fp8(F.gelu(x, approximate=...)) in one pass over HBM."""
    return _ops.quant_fp8(input, ACT_GELU_TANH if approximate == "tanh" else ACT_GELU_ERF)


def gelu_quantize_to_int8(input: torch.Tensor, approximate: str = "tanh"):
    """quantize_to_int8(F.gelu(x, approximate=...), symmetric=False) in one pass over HBM."""
    return _ops.quant_int8(input, False, ACT_GELU_TANH if approximate == "tanh" else ACT_GELU_ERF)
