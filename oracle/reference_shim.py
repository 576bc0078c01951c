"""ORACLE -- TEST INFRASTRUCTURE ONLY. Imports the REAL reference (FastDM, /root/reference) on CPU.

Only usable in the build container (the GPU box has no /root/reference); used by
oracle/gen_golden.py to produce tests/golden/ and by tests that are skipped when the reference is
absent. Shims, all applied outside the reference tree:
  1. fastdm.cuda_ops / fastdm.kernel.cuda / fastdm.kernel.triton are stubbed -- the reference's
     fastdm/kernel/__init__.py:1-3 imports all three and they need a GPU / compiled extension;
  2. QLinear's hard-coded device_type="cuda" default (fastdm/layer/qlinear.py:7) is flipped to "cpu";
  3. the `fp8_matmul` torch backend is replaced by its arithmetic written out, because CPU
     torch._scaled_mm rejects the per-row scales fastdm/kernel/torch/matrixmul.py:33 passes.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FASTDM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fastdm", "kernel"))


_loaded = None


def load():
    """Returns the imported reference package `fastdm` with the torch backend active."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    import torch

    for name in ("fastdm.cuda_ops", "fastdm.kernel.cuda", "fastdm.kernel.triton"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.environ["KERNEL_BACKEND"] = "torch"
    import fastdm  # noqa: F401
    import fastdm.kernel  # noqa: F401
    from fastdm.kernel.registry import kernel_registry
    from fastdm.layer.qlinear import QLinear

    QLinear.__init__.__defaults__ = (True, torch.bfloat16, "cpu")

    def fp8_matmul_cpu(a, b, scale_a, scale_b, out_dtype, bias=None):
        # == torch._scaled_mm(a, b, scale_a, scale_b.T, bias, out_dtype) with row/col scales
        assert b.shape[0] % 16 == 0 and b.shape[1] % 16 == 0
        assert out_dtype is torch.bfloat16
        out = (a.float() @ b.float()) * scale_a * scale_b.transpose(0, 1)
        if bias is not None:
            out = out + bias.float()
        return out.to(out_dtype)

    kernel_registry._registry["fp8_matmul"]["torch"] = fp8_matmul_cpu
    _loaded = fastdm
    return fastdm


def torch_backend(op_name: str):
    """The reference's torch-backend implementation of `op_name` (module-level names are None
    because `register` returns None: fastdm/kernel/registry.py:11-18)."""
    load()
    from fastdm.kernel.registry import kernel_registry

    return kernel_registry._registry[op_name]["torch"]
