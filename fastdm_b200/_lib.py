"""ctypes binding of libfastdm_b200.so (the C ABI declared in include/fastdm_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
or a call fails, a RuntimeError is raised (reference convention: TORCH_CHECK -> RuntimeError,
csrc/torch_bindings.cpp:31-61).
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfastdm_b200.so")

# every symbol include/fastdm_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "fdm_last_error": (c_char_p, []),
    "fdm_version": (c_char_p, []),
    "fdm_check_device": (c_int, [c_int]),
    "fdm_quant_fp8": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "fdm_quant_int8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "fdm_rms_norm": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_int, c_void_p]),
    "fdm_rope": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                         c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "fdm_gelu_and_mul": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "fdm_gelu_quant": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                               c_int, c_int, c_int, c_void_p]),
    "fdm_gemm_fp8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "fdm_gemm_int8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p]),
    "fdm_gemm_fp8_residual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int,
                                      c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "fdm_gemm_int8_residual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int,
                                       c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    "fdm_attn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int64, c_int64, c_int64, c_int, c_int,
                             c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                             c_int, c_int, c_float, c_int, c_void_p]),
    "fdm_attn_fwd_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p,
                                     c_int64, c_int64, c_int, c_int, c_int64, c_int64, c_int64, c_int64,
                                     c_int, c_int, c_float, c_int, c_void_p]),
    "fdm_debug_set_attn_trace": (c_int, [c_void_p]),
    "fdm_qk_norm_rope": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                 c_int64, c_int64, c_int64, c_int64, c_int64, c_float, c_int, c_int, c_void_p]),
    "fdm_layernorm_modulate_quant": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int64, c_int64, c_int64, c_int64, c_int64, c_float,
                                             c_int, c_int, c_int, c_int, c_void_p]),
    "fdm_ulysses_pack_heads": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p]),
    "fdm_rel_l1_distance": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "fdm_ulysses_unpack_heads": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int, c_void_p]),
}

# dtype / activation enums of the header
FDM_BF16, FDM_F16, FDM_F32, FDM_E4M3, FDM_S8 = 0, 1, 2, 3, 4
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF = 0, 1, 2

_lib = None


def load():
    """Load the shared library (once) and bind every declared symbol. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"fastdm_b200: {LIB_PATH} is missing -- build it with `python -m fastdm_b200.build` "
            "(nvcc, sm_100a). There is no fallback implementation."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().fdm_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


# every C-ABI compute call launches exactly one kernel of ours; bench.py reads this counter
launch_count = 0


def add_launches(n: int):
    """CUDA-graph replays launch the captured kernels without passing through check()."""
    global launch_count
    launch_count += n


def check(rc: int, what: str):
    global launch_count
    launch_count += 1
    if rc != 0:
        kind = {-1: "invalid argument", -2: "unsupported device", -3: "CUDA failure", -4: "unsupported"}.get(rc, "error")
        exc = NotImplementedError if rc in (-2, -4) else RuntimeError
        raise exc(f"fastdm_b200.{what}: {kind} ({rc}): {last_error()}")


def version() -> str:
    return load().fdm_version().decode()
