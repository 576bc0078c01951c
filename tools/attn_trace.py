"""Prints the intra-CTA timeline of the attention kernel (clock64 stamps of CTA 0) -- debug aid.

    python tools/attn_trace.py            # hd 128 bf16: CTA pairs, P through shared memory, two MMA issuers
    FDM_ATTN_CG=1 python tools/attn_trace.py   # single CTA, P in TMEM, one MMA issuer (legacy event meaning)
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import _lib, ops
lib = _lib.load()
b, s, h, hd = 1, 8704, 24, 128
q, k, v = (torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(2):
    ops.scaled_dot_product_attention(q, k, v, h, h, hd)
buf = torch.zeros(7 * 8 * 64, dtype=torch.int64, device="cuda")
lib.fdm_debug_set_attn_trace(buf.data_ptr())
ops.scaled_dot_product_attention(q, k, v, h, h, hd)
torch.cuda.synchronize()
lib.fdm_debug_set_attn_trace(None)
tr = buf.cpu().view(7, 8, 64)
pair = os.environ.get("FDM_ATTN_CG", "2") == "2"
t0 = int(tr[0, 0, 0])
for t in range(8, 20):
    a = [int(tr[0, e, t]) - t0 for e in range(8)]
    bb = [int(tr[1, e, t]) - t0 for e in range(8)]
    # late-store kernels stamp event 7 when all exponentials are packed in registers (before the wait for the P buffer)
    ca = f" (packed +{a[7]-a[3]:4d})" if int(tr[0, 7, t]) else ""
    cb = f" (packed +{bb[7]-bb[3]:4d})" if int(tr[1, 7, t]) else ""
    line = (f"tile {t:2d} | softA waitS {a[0]:6d} Srdy {a[1]:6d} ld {a[2]-a[1]:4d} max {a[3]-a[2]:4d} exp {a[4]-a[3]:4d}{ca} st+arr {a[5]-a[4]:4d} -> {a[5]:6d}"
            f" | softB waitS {bb[0]:6d} Srdy {bb[1]:6d} ld {bb[2]-bb[1]:4d} max {bb[3]-bb[2]:4d} exp {bb[4]-bb[3]:4d}{cb} st+arr {bb[5]-bb[4]:4d} -> {bb[5]:6d}")
    if pair and os.environ.get("FDM_ATTN_ISSUE", "1") == "0":
        for r, nm in ((2, "mmaA"), (3, "mmaB")):
            m = [int(tr[r, e, t]) - t0 for e in range(7)]
            if r == 2:
                line += f" | top {m[4]:6d} probed={int(tr[2, 7, t])} Kfull +{m[5]-m[4]:4d} sfree +{m[6]-m[5]:4d}"
            line += f" | {nm} qk_go {m[0]:6d} issue +{m[1]-m[0]:4d} gap +{m[2]-m[1]:4d} pv_issue +{m[3]-m[2]:4d} -> {m[3]:6d}"
    else:
        m = [int(tr[2, e, t]) - t0 for e in range(6)]
        line += f" | MMA Vrdy {m[0]:6d} P_A {m[1]:6d} pv_issued +{m[4]-m[1]:4d} qk_issued +{m[2]-m[1]:4d} P_B {m[3]:6d}"
    print(line)
    # last arrival over the four warps of each warpgroup (event 6), leader CTA and peer CTA of the pair
    print("        | P arrival per warp (leader CTA): A " + " ".join(f"{int(tr[6, w, t]) - t0:6d}" for w in range(4)) + " | B " + " ".join(f"{int(tr[6, w, t]) - t0:6d}" for w in range(4, 8)))
    pa = [int(tr[4, e, t]) - t0 for e in range(8)]
    pb = [int(tr[5, e, t]) - t0 for e in range(8)]
    print(f"        | last P arrival: leader A {int(tr[0, 6, t]) - t0:6d} B {int(tr[1, 6, t]) - t0:6d} | peer CTA: softA Srdy {pa[1]:6d} packed {pa[7]:6d} done {pa[5]:6d} last {pa[6]:6d}"
          f" | softB Srdy {pb[1]:6d} packed {pb[7]:6d} done {pb[5]:6d} last {pb[6]:6d}")
per = (int(tr[0, 1, 40]) - int(tr[0, 1, 8])) / 32
print("cycles per KV iteration (2 Q tiles x 128 keys):", per, " -> MMA-ideal 2048")
life = [int(tr[3, 7, i]) for i in range(5)]
last = max(int(tr[0, 5, t]) for t in range(64))
print(f"CTA life cycle (cycles from kernel entry): set-up done {life[1]-life[0]}, first S ready {int(tr[0, 1, 0])-life[0]}, "
      f"first P written {int(tr[0, 5, 0])-life[0]}, epilogue start {life[2]-life[0]}, epilogue end {life[3]-life[0]}, exit {life[4]-life[0]}"
      f" | tiles in this launch: {s // 128}")
