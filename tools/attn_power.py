"""Sustained attention throughput next to the SM clock and board power sampled DURING the run (nvidia-smi, 50 ms):
is the kernel clock-limited by the power cap, and how many SM cycles does one 128-key step take at that clock?

    python tools/attn_power.py [seconds]        ours, then torch SDPA (cuDNN), Wan shape 80640^2 x 40 heads x 128
"""
import os, subprocess, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdm_b200 import ops
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
b, s, h, hd = 1, 80640, 40, 128
q, k, v = (torch.randn(b, s, h * hd, device="cuda", dtype=torch.bfloat16) for _ in range(3))
q4 = q.view(b, s, h, hd).transpose(1, 2)
k4 = k.view(b, s, h, hd).transpose(1, 2)
v4 = v.view(b, s, h, hd).transpose(1, 2)
fl = 4.0 * b * h * s * s * hd


def sample(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True)
        try:
            f = [x.strip() for x in r.stdout.strip().split(",")]
            out.append((float(f[0]), float(f[1]), float(f[2]), f[3]))
        except Exception:
            pass
        time.sleep(0.05)


runs = [("fastdm_b200", lambda: ops.scaled_dot_product_attention(q, k, v, h, h, hd))]
if not os.environ.get("FDM_POWER_NO_TORCH"):
    runs += [("torch sdpa", lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4)),
             ("fastdm_b200 again", lambda: ops.scaled_dot_product_attention(q, k, v, h, h, hd))]
for name, fn in runs:
    fn(); torch.cuda.synchronize()
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, out)); th.start()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        fn(); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = e0.elapsed_time(e1) / n
    out = out[len(out) // 4:]     # steady part
    mhz = sorted(o[0] for o in out)[len(out) // 2]; w = sorted(o[1] for o in out)[len(out) // 2]; tc = max(o[2] for o in out)
    tf = fl / ms / 1e9
    peak_at_clock = 148 * 8192 * mhz * 1e6 / 1e12
    print(f"{name:18s}: {ms:7.2f} ms/call {tf:6.0f} TFLOP/s | SM clock {mhz:.0f} MHz, {w:.0f} W, {tc:.0f} C, reasons {out[-1][3]} | "
          f"tensor peak at that clock {peak_at_clock:.0f} -> {tf / peak_at_clock:.3f} of it = {2048 / (tf / peak_at_clock):.0f} cycles per 2x128x128 step", flush=True)
    time.sleep(1.0)
