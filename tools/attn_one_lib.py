"""One attention launch through a given build of the library (ctypes), for `ncu -k regex:attn_fwd` A/B captures.
    python tools/attn_one_lib.py <lib.so> [flux|wan]"""
import ctypes, os, sys, torch
from ctypes import c_float, c_int, c_int64, c_void_p
lib = ctypes.CDLL(os.path.abspath(sys.argv[1]))
lib.fdm_attn_fwd.restype = c_int
lib.fdm_attn_fwd.argtypes = [c_void_p] * 5 + [c_int64] * 3 + [c_int, c_int] + [c_int64] * 8 + [c_int, c_int, c_float, c_int, c_void_p]
which = sys.argv[2] if len(sys.argv) > 2 else "flux"
b, s, h, hd = (1, 80640, 40, 128) if which == "wan" else (1, 8704, 24, 128)
q, k, v = (torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16) for _ in range(3))
o = torch.empty_like(q)
for _ in range(3):
    rc = lib.fdm_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), None, b, s, s, h, hd, q.stride(0), q.stride(1),
                          k.stride(0), k.stride(1), v.stride(0), v.stride(1), o.stride(0), o.stride(1), 128, 64, hd ** -0.5, 0,
                          torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
print(float(o.float().abs().mean()))
