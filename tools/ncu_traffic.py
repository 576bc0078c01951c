"""Reads `ncu --set full` reports and writes the per-launch DRAM traffic of the attention kernel to
profiles/attn_traffic.json (the file bench.py's roofline.traffic comes from).

    python tools/ncu_traffic.py wan_n1=gpurun_out/prof_attn_wan.ncu-rep flux_n1=gpurun_out/prof_attn_flux.ncu-rep
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread")
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
path = os.path.join(ROOT, "profiles", "attn_traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for arg in sys.argv[1:]:
    key, rep = arg.split("=")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[-1]   # last captured launch
    m = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT:
            try:
                m[h] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                pass
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    out[key] = dict(dram_bytes_per_launch=m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0),
                    dram_read=m.get("dram__bytes_read.sum"), dram_write=m.get("dram__bytes_write.sum"),
                    duration_s=m.get("gpu__time_duration.sum"),
                    tensor_pipe_active_pct=m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    xu_pipe_pct=m.get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                    kernel=kname[:120], report=os.path.basename(rep))
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
