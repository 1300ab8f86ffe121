"""One line per `ncu --page raw --csv` capture: duration, tensor-pipe activity, issue activity, DRAM bytes, grid, registers."""
import csv
import sys

WANT = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread")
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print(path, "empty capture")
        continue
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    name = next((vals[col[k]] for k in ("Kernel Name", "Function Name", "Demangled Name") if k in col), "?")
    print(path)
    print("  kernel:", name[:110])
    for k in WANT:
        if k in col:
            print(f"  {k}: {vals[col[k]]} {units[col[k]]}")
