"""Attention kernel sweep over the experiment knobs (one process per setting; the knobs are read once).

    python tools/attn_sweep.py "FDM_ATTN_ISSUE=0" "FDM_ATTN_ISSUE=1" "FDM_ATTN_ISSUE=1 FDM_ATTN_EMU=8" ...

Each setting runs tools/gpu_diag.py attn_perf (torch SDPA timed only in the first one) and, with
--trace, tools/attn_trace.py.
"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
trace = "--trace" in sys.argv
for i, setting in enumerate(args or [""]):
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("=")
        env[k] = v
    if i > 0:
        env["FDM_DIAG_NO_TORCH"] = "1"
    print(f"--- {setting or 'defaults'}", flush=True)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_diag.py"), "attn_perf"], env=env, capture_output=True, text=True)
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("attn")), flush=True)
    if r.returncode:
        print(r.stderr[-2000:])
    if trace:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "attn_trace.py")], env=env, capture_output=True, text=True)
        print("\n".join(r.stdout.splitlines()[-6:]), flush=True)
        if r.returncode:
            print(r.stderr[-1500:])
