"""Attention kernel timing sweep over the FDM_ATTN_EMU knob (one process per setting)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for emu in (0, 4, 8):
    env = dict(os.environ, FDM_ATTN_EMU=str(emu))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_diag.py"), "attn_perf"], env=env, capture_output=True, text=True)
    print(f"--- FDM_ATTN_EMU={emu}")
    print("\n".join(l for l in r.stdout.splitlines() if l.startswith("attn")))
    if r.returncode: print(r.stderr[-2000:])
