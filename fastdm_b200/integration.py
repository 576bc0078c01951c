"""Plugging fastdm_b200 into an installed FastDM (the reference) as its `cuda` kernel backend.

    import fastdm_b200.integration as fdi
    fdi.install()                 # before or after `import fastdm`
    # ... FastDMEngine(..., kernel_backend="cuda") / examples/demo/gen.py run unmodified

What install() does
  1. sys.modules["fastdm.cuda_ops"] = fastdm_b200.cuda_ops -- the seven legacy pybind names
     (csrc/torch_bindings.cpp:191-201), so `fastdm/kernel/cuda/*.py` import and work as they are;
  2. registers our implementations under (op_name, "cuda") in fastdm.kernel.registry.kernel_registry
     for all nine ops of fastdm/kernel/operators_set.py -- replacing whatever the reference's cuda
     wrappers registered, so `sdpa` / `sdpa_sparse` hit our attention kernel instead of
     sageattention / spas_sage_attn / cuDNN, and `gelu_and_mul` (hard-wired to the "triton" backend,
     operators_set.py:54) gets our kernel under that name as well.
FastDM itself is not modified and not required: fastdm_b200.ops exposes the same nine functions.
"""
import sys
import types

from . import cuda_ops, ops

OPS = {
    "rmsnorm": ops.rms_norm,
    "rotembd": ops.rotary_pos_embedding,
    "gelu_and_mul": ops.gelu_and_mul,
    "quantize_to_int8": ops.quantize_to_int8,
    "quantize_to_fp8": ops.quantize_to_fp8,
    "fp8_matmul": ops.fp8_matmul,
    "int8_matmul": ops.int8_matmul,
    "sdpa": ops.scaled_dot_product_attention,
    "sdpa_sparse": ops.sparse_scaled_dot_product_attention,
}


def install(stub_triton: bool = True):
    """Returns FastDM's kernel_registry with our ops registered as the `cuda` backend."""
    sys.modules["fastdm.cuda_ops"] = cuda_ops
    if stub_triton and "fastdm.kernel.triton" not in sys.modules:
        # fastdm/kernel/__init__.py:1-3 imports the triton backend unconditionally; it is not needed
        sys.modules["fastdm.kernel.triton"] = types.ModuleType("fastdm.kernel.triton")
    import fastdm.kernel  # noqa: F401  (runs the reference's own registrations)
    from fastdm.kernel.registry import kernel_registry

    for name, fn in OPS.items():
        kernel_registry._registry.setdefault(name, {})["cuda"] = fn
    kernel_registry._registry["gelu_and_mul"]["triton"] = ops.gelu_and_mul  # force_backend="triton" dispatch
    return kernel_registry
