"""Condense an `ncu --page raw --csv` export into one line per launch (duration, DRAM bytes / throughput, occupancy,
registers, issue activity).   python tools/summarize_ncu_raw.py raw.csv > profiles/xxx.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "CTA/SM (smem)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %")]
cols = [(h, hdr.index(k)) for k, h in want if k in hdr]
print("| " + " | ".join(h + (" [" + units[i] + "]" if units[i] and h not in ("kernel", "grid") else "") for h, i in cols) + " |")
print("|" + "---|" * len(cols))
for r in data:
    out = []
    for h, i in cols:
        v = r[i]
        if h == "kernel":
            v = v.split("(")[0][-40:]
        else:
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
        out.append(v)
    print("| " + " | ".join(out) + " |")
