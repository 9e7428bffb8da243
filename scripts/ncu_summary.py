"""markdown table of the interesting metrics of an .ncu-rep (one row per captured launch); usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__issue_active.avg.pct", "issue %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 %"), ("launch__grid_size", "grid")]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print("| " + " | ".join(n for _, n in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    cells = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0]
        else:
            try:
                f = float(v.replace(",", ""))
                v = (f"{f:,.0f}" if f > 1000 else f"{f:.2f}") + (" " + units[i] if units[i] and units[i] != "%" else "")
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
