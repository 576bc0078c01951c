"""Host-side mirror of the reference's layer classes for the DiT block hot path, built on the
fused B200 kernels (fastdm_b200/ops.py). Same class names, constructor arguments and weight
loading entry points as the reference (plain Python classes with `.forward`, not nn.Modules):

    QLinear      fastdm/layer/qlinear.py:6-81
    FeedForward  fastdm/layer/transformer.py:14-62   (GELU / GELU-tanh variants used by DiT blocks)

What is different from the reference is *where the work happens*, not what is computed:
  * weights are quantised per output channel on the GPU at load time with the same kernel that
    quantises activations (bit-identical to fastdm/utils/quantization.py on CPU);
  * `forward` can take already-quantised activations (the LayerNorm-modulate-quant kernel produces
    them), fuse GELU into the GEMM epilogue, write into a column slice of a wider buffer and fold
    `residual + gate * out` into the epilogue -- each of those is the reference's op sequence with
    the same roundings, minus the HBM round trips.
"""
from typing import List, Optional

import torch

from . import ops


class Quantized:
    """A per-token quantised activation: codes [M, K], scales [M, 1] and (int8) zero points [M, 1]."""

    __slots__ = ("q", "scale", "zp")

    def __init__(self, q, scale, zp=None):
        self.q, self.scale, self.zp = q, scale, zp

    def rows(self, lo, hi):
        return Quantized(self.q[lo:hi], self.scale[lo:hi], None if self.zp is None else self.zp[lo:hi])


def quantize(x2d: torch.Tensor, quant_type, gelu: Optional[str] = None) -> Quantized:
    """The activation side of QLinear.forward (qlinear.py:69-74), optionally fused with the GELU
    that precedes it (fastdm/layer/activations.py:38-41)."""
    if quant_type == torch.float8_e4m3fn:
        q, s = ops.quantize_to_fp8(x2d) if gelu is None else ops.gelu_quantize_to_fp8(x2d, gelu)
        return Quantized(q, s)
    if quant_type == torch.int8:
        if gelu is None:
            q, s, zp = ops.quantize_to_int8(x2d, symmetric=False)
        else:
            q, s, zp = ops.gelu_quantize_to_int8(x2d, gelu)
        return Quantized(q, s, zp)
    raise ValueError(f"Unsupported quantization type: {quant_type}")


class QLinear:
    def __init__(self, in_features, out_features, bias=True, data_type=torch.bfloat16, device_type="cuda"):
        self.in_features, self.out_features = in_features, out_features
        self.weight = None          # (K, N) view with stride (1, K), as the reference keeps it
        self.bias = None
        self.has_bias = bias
        self.dtype = data_type
        self.device = device_type
        self.weight_quant_scale = None
        self.weight_asym_sumcol = None

    @property
    def quant_type(self):
        return self.weight.dtype if self.weight.dtype in (torch.float8_e4m3fn, torch.int8) else None

    def weight_loading_and_quant(self, src_weight: List[torch.Tensor], src_bias: List[Optional[torch.Tensor]],
                                 quant_type=None):
        """src_weight: list of (in_features, out_features_i) tensors fused along N (qlinear.py:18-54)."""
        w_nk = torch.cat([w.to(self.device).transpose(0, 1) for w in src_weight], 0).to(self.dtype).contiguous()
        if self.has_bias and src_bias[0] is not None:
            self.bias = torch.cat([b.to(self.device) for b in src_bias], 0).to(self.dtype).contiguous()
        else:
            self.bias = None
        if quant_type is None:
            self.weight = w_nk.transpose(0, 1)
        elif quant_type == torch.float8_e4m3fn:
            q, s = ops.quantize_to_fp8(w_nk)                       # per output channel: rows of W^T
            self.weight, self.weight_quant_scale = q.transpose(0, 1), s
        elif quant_type == torch.int8:
            q, s, _ = ops.quantize_to_int8(w_nk, symmetric=True)   # utils/quantization.py:5-41 default
            self.weight, self.weight_quant_scale = q.transpose(0, 1), s
            self.weight_asym_sumcol = q.to(torch.int32).sum(dim=1, dtype=torch.int32).reshape(1, -1).contiguous()
        else:
            raise ValueError(f"Unsupported quantization type: {quant_type}")

    # ---- pre-quantised weights: quantise once, cache, reload without the bf16 weights ----
    def export_quantized(self) -> dict:
        """Everything `forward` needs, on the CPU: 8-bit codes in the [N, K] storage order (as raw bytes, so the
        file does not depend on float8 serialisation), per-channel scales, INT8 column sums, bias."""
        w_nk = self.weight.transpose(0, 1).contiguous()
        qt = self.quant_type
        d = dict(quant_type={None: "none", torch.float8_e4m3fn: "fp8_e4m3", torch.int8: "int8"}[qt],
                 in_features=self.in_features, out_features=self.out_features, dtype=str(self.dtype).split(".")[-1],
                 weight=(w_nk.view(torch.uint8) if qt is not None else w_nk).cpu(),
                 bias=None if self.bias is None else self.bias.cpu())
        if qt is not None:
            d["scale"] = self.weight_quant_scale.cpu()
        if self.weight_asym_sumcol is not None:
            d["sumcol"] = self.weight_asym_sumcol.cpu()
        return d

    @classmethod
    def from_quantized(cls, d: dict, device="cuda") -> "QLinear":
        dt = getattr(torch, d["dtype"])
        lin = cls(d["in_features"], d["out_features"], bias=d["bias"] is not None, data_type=dt, device_type=device)
        w = d["weight"].to(device)
        if d["quant_type"] == "fp8_e4m3":
            w = w.view(torch.float8_e4m3fn)
        elif d["quant_type"] == "int8":
            w = w.view(torch.int8)
        lin.weight = w.transpose(0, 1)
        lin.bias = None if d["bias"] is None else d["bias"].to(device)
        if "scale" in d:
            lin.weight_quant_scale = d["scale"].to(device)
        if "sumcol" in d:
            lin.weight_asym_sumcol = d["sumcol"].to(device)
        return lin

    def forward(self, input_tensor, act: Optional[str] = None, out: Optional[torch.Tensor] = None,
                gate: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                rows_per_batch: int = 1, round_steps: bool = True):
        """input_tensor: a bf16 tensor [..., K] or a `Quantized` [M, K] produced upstream."""
        if isinstance(input_tensor, Quantized):
            xq, lead = input_tensor, None
            out_dtype = self.dtype
        else:
            lead = input_tensor.shape[:-1]
            x2d = input_tensor.reshape(-1, input_tensor.shape[-1])
            out_dtype = input_tensor.dtype
            if self.quant_type is None:
                if act is not None or gate is not None or residual is not None or out is not None:
                    raise NotImplementedError("fused epilogues exist for the quantised linears only")
                y = torch.addmm(self.bias, x2d, self.weight) if self.bias is not None else torch.mm(x2d, self.weight)
                return y.view(*lead, self.weight.shape[-1])
            xq = quantize(x2d, self.quant_type)
        if self.quant_type == torch.float8_e4m3fn:
            y = ops.fp8_matmul(xq.q, self.weight, xq.scale, self.weight_quant_scale, out_dtype, self.bias, act=act,
                               out=out, gate=gate, residual=residual, rows_per_batch=rows_per_batch,
                               round_steps=round_steps)
        else:
            y = ops.int8_matmul(xq.q, self.weight, xq.scale, self.weight_quant_scale, out_dtype,
                                self.weight_asym_sumcol, xq.zp, self.bias, act=act, out=out, gate=gate,
                                residual=residual, rows_per_batch=rows_per_batch, round_steps=round_steps)
        return y if lead is None else y.view(*lead, y.shape[-1])


def load_linear(sd, names, quant_type=None, device="cuda", dtype=torch.bfloat16) -> QLinear:
    """Build a QLinear from state-dict entries `<name>.weight` ([out, in]) / `<name>.bias`, fusing
    several projections along N -- what BaseModelCore.init_weight does (fastdm/model/basemodel.py:33-66)."""
    ws = [sd[f"{n}.weight"].transpose(0, 1) for n in names]
    bs = [sd.get(f"{n}.bias") for n in names]
    lin = QLinear(ws[0].shape[0], sum(w.shape[1] for w in ws), bias=bs[0] is not None, data_type=dtype,
                  device_type=device)
    lin.weight_loading_and_quant(ws, bs, quant_type)
    return lin


class FeedForward:
    """fastdm/layer/transformer.py:14-62 with activation_fn in {"gelu", "gelu-approximate"}: the
    GELU runs in the first GEMM's epilogue, the second linear can fold gate + residual."""

    def __init__(self, proj: QLinear, ff_out_proj: QLinear, activation_fn="gelu-approximate"):
        self.proj, self.ff_out_proj = proj, ff_out_proj
        self.act = {"gelu": "gelu_erf", "gelu-approximate": "gelu_tanh"}[activation_fn]

    def forward(self, hidden_states, gate=None, residual=None, rows_per_batch=1, round_steps=True, out=None):
        h = self.proj.forward(hidden_states, act=self.act)
        hq = quantize(h.reshape(-1, h.shape[-1]), self.ff_out_proj.quant_type)
        return self.ff_out_proj.forward(hq, gate=gate, residual=residual, rows_per_batch=rows_per_batch,
                                        round_steps=round_steps, out=out)


def save_quantized(linears: dict, path: str):
    """{name: QLinear} -> one file of pre-quantised weights (e4m3 / int8 codes + fp32 scales + int32 column sums)."""
    torch.save({name: lin.export_quantized() for name, lin in linears.items()}, path)


def load_quantized(path: str, device="cuda") -> dict:
    return {name: QLinear.from_quantized(d, device) for name, d in torch.load(path, map_location="cpu").items()}
