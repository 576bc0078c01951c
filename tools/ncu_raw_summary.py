"""Summarise `ncu --page raw --csv` exports: one line per captured launch with duration, DRAM bytes and the pipe utilisations.

    python tools/ncu_raw_summary.py gpurun_out/r02_elem_raw.csv [more.csv ...]
"""
import csv, sys
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
COLS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_thr%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr is None:
        print(path, ": no captured launch"); continue
    h, u = rows[hdr], rows[hdr + 1]
    print(f"== {path}")
    for r in rows[hdr + 2:]:
        if len(r) < len(h): continue
        name = r[h.index("Kernel Name")]
        name = name[:name.index("(")] if "(" in name else name
        out = []
        for col, short in COLS:
            if col not in h: continue
            i = h.index(col)
            try:
                val = float(r[i].replace(",", "")) * UNIT.get(u[i], 1.0)
            except ValueError:
                continue
            if short == "dur": out.append(f"{val * 1e6:9.1f} us")
            elif short in ("rd", "wr"): out.append(f"{short} {val / 1e6:8.1f} MB")
            else: out.append(f"{short} {val:.1f}")
        d = {s: None for _, s in COLS}
        try:
            dur = float(r[h.index('gpu__time_duration.sum')].replace(',', '')) * UNIT.get(u[h.index('gpu__time_duration.sum')], 1.0)
            rd = float(r[h.index('dram__bytes_read.sum')].replace(',', '')) * UNIT.get(u[h.index('dram__bytes_read.sum')], 1.0)
            wr = float(r[h.index('dram__bytes_write.sum')].replace(',', '')) * UNIT.get(u[h.index('dram__bytes_write.sum')], 1.0)
            out.append(f"dram {(rd + wr) / dur / 1e9:6.0f} GB/s")
        except Exception:
            pass
        print(f"{name[:70]:70s} " + " | ".join(out))
