"""CPU: fastdm_b200.integration.install() against the real reference package (build container only;
skipped where /root/reference is absent). Checks the drop-in wiring, not arithmetic: the
reference's dispatcher must route all nine op names to our functions when KERNEL_BACKEND=cuda, and
its own fastdm/kernel/cuda/*.py wrappers must import against our `fastdm.cuda_ops` replacement."""
import os
import sys

import pytest
import torch

REF = os.environ.get("FASTDM_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "fastdm")), reason="reference not present")


def test_install_routes_every_op_to_the_b200_kernels(monkeypatch, lib):
    # the reference probes the GPU at import time (fastdm/kernel/cuda/attention.py:8)
    monkeypatch.setattr(torch.cuda, "get_device_capability", lambda *a, **k: (10, 0))
    for m in [m for m in sys.modules if m == "fastdm" or m.startswith("fastdm.")]:
        monkeypatch.delitem(sys.modules, m, raising=False)
    monkeypatch.syspath_prepend(REF)
    monkeypatch.setenv("KERNEL_BACKEND", "cuda")
    from fastdm_b200 import integration, ops

    reg = integration.install()
    import fastdm.cuda_ops as shim
    from fastdm.kernel import operators_set as O

    for name in ("fp8_quant_", "int8_quant_", "rms_norm_", "rotary_emb_", "fp8_scaled_mm_", "int8_scaled_mm_",
                 "flash_attention_fp8_fwd_"):          # csrc/torch_bindings.cpp:191-201
        assert callable(getattr(shim, name))
    for op, fn in integration.OPS.items():
        assert reg._registry[op]["cuda"] is fn
        assert reg.select_backend(op) == "cuda"
    # dispatch really lands in our code: a CPU tensor is refused by OUR wrapper (no CPU fallback)
    x = torch.zeros(4, 16, dtype=torch.bfloat16)
    for call in (lambda: O.quantize_to_fp8(x), lambda: O.quantize_to_int8(x, False), lambda: O.rms_norm(x, None, 1e-6),
                 lambda: O.gelu_and_mul(x),
                 lambda: O.scaled_dot_product_attention(x[None], x[None], x[None], 1, 1, 16)):
        with pytest.raises(RuntimeError, match="fastdm_b200"):
            call()
    # QLinear of the reference is untouched and would now run on our ops
    from fastdm.layer.qlinear import QLinear
    assert QLinear.forward.__module__ == "fastdm.layer.qlinear"
