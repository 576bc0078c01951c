"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/summarize_launches.py gpurun_out/launches_flux.csv "title" > profiles/rNN_x_summary.txt
"""
import csv, re, sys
from collections import defaultdict

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"<.*", "", r[4]).replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    agg[name][0] += 1
    agg[name][1] += float(r[14]) / 1e6
tot = sum(v[1] for v in agg.values())
print(f"{title}: {len(rows)} launches, sum of kernel durations {tot:.2f} ms")
print("share   launches   total_ms   avg_us   kernel")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * ms / tot:6.2f}%  {n:6d}  {ms:9.2f}  {1e3 * ms / n:8.1f}  {k}")
