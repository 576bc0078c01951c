"""fastdm_b200 -- B200-native (sm_100a) kernels for the DiT transformer-block hot path of
KE-AI-ENG/FastDM, behind FastDM's own operator API.

    import fastdm_b200
    from fastdm_b200 import ops            # the 9 op names of fastdm/kernel/operators_set.py
    fastdm_b200.integration.install()      # register them as FastDM's "cuda" backend

The compute lives in libfastdm_b200.so (C ABI: include/fastdm_b200.h); there is no fallback.
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"


def library_path() -> str:
    return _lib.LIB_PATH
