"""A/B timing of fdm_attn_fwd between two builds of the library in ONE process on ONE box (box-to-box variance
of this power-capped kernel is larger than most single changes).

    python tools/attn_ab.py tools/ab/libold.so fastdm_b200/libfastdm_b200.so
Libraries are called through ctypes directly; runs alternate A, B, A, B ... per shape.
"""
import ctypes, os, sys, torch
from ctypes import c_float, c_int, c_int64, c_void_p
ARGS = [c_void_p] * 5 + [c_int64] * 3 + [c_int, c_int] + [c_int64] * 8 + [c_int, c_int, c_float, c_int, c_void_p]
libs = []
for path in sys.argv[1:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.fdm_attn_fwd.restype = c_int
    lib.fdm_attn_fwd.argtypes = ARGS
    libs.append((os.path.basename(os.path.dirname(os.path.abspath(path))) + "/" + os.path.basename(path), lib))


def call(lib, q, k, v, o, h, hd):
    b, sq, _ = q.shape
    sk = k.shape[1]
    rc = lib.fdm_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), None, b, sq, sk, h, hd,
                          q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), o.stride(0), o.stride(1),
                          128, 64, hd ** -0.5, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc


def timed(fn, iters):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (b, sq, sk, h, hd, iters) in ((1, 4608, 4608, 24, 128, 20), (1, 8704, 8704, 24, 128, 10), (2, 4685, 4685, 24, 64, 10),
                                  (1, 80640, 80640, 4, 128, 3), (1, 80640, 512, 40, 128, 10), (1, 80640, 80640, 40, 128, 1)):
    q = torch.randn(b, sq, h * hd, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(b, sk, h * hd, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(b, sk, h * hd, device="cuda", dtype=torch.bfloat16)
    outs = [torch.empty_like(q) for _ in libs]
    for (name, lib), o in zip(libs, outs):
        call(lib, q, k, v, o, h, hd)
    res = {name: [] for name, _ in libs}
    for rep in range(3):
        for (name, lib), o in zip(libs, outs):
            res[name].append(timed(lambda: call(lib, q, k, v, o, h, hd), iters))
    fl = 4.0 * b * h * sq * sk * hd
    line = f"b{b} sq{sq} sk{sk} h{h} hd{hd}: "
    for name, _ in libs:
        best = min(res[name])
        line += f"{name} {best:.3f} ms {fl / best / 1e9:.0f} TF ({' '.join(f'{fl / t / 1e9:.0f}' for t in res[name])}) | "
    same = all(torch.equal(outs[0], o) for o in outs[1:])
    print(line + f"outputs identical: {same}", flush=True)
