"""Markdown table of an `ncu --set full` report: one row per captured launch.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/r01_ncu_full_step_summary_vN.md
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(*parts):
    for i, h in enumerate(hdr):
        if all(p in h for p in parts):
            return i
    return None


cols = [("us", col("gpu__time_duration.sum")), ("dram rd MB", col("dram__bytes_read.sum")), ("dram wr MB", col("dram__bytes_write.sum")),
        ("tensor % active", col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")) ,
        ("issue slots %", col("sm__inst_issued.avg.pct_of_peak_sustained_active")),
        ("L2 hit %", col("lts__t_sector_hit_rate.pct")), ("regs", col("launch__registers_per_thread")),
        ("dyn smem KB", col("launch__shared_mem_per_block_dynamic")), ("SM GHz", col("sm__cycles_elapsed.avg.per_second")),
        ("grid", col("Grid Size")), ("block", col("Block Size"))]
cols = [(n, i) for n, i in cols if i is not None]
name_i = hdr.index("Kernel Name")
print("| # | kernel | " + " | ".join(n for n, _ in cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k, r in enumerate(data):
    cells = []
    for n, i in cols:
        v, u = r[i], units[i]
        try:
            f = float(v.replace(",", ""))
            if u in ("ns",):
                f /= 1e3
            if u in ("byte",) and "MB" in n:
                f /= 1e6
            if u == "Kbyte" and "MB" in n:
                f /= 1e3
            if u == "Gbyte" and "MB" in n:
                f *= 1e3
            if n == "SM GHz" and u in ("hz", "Hz"):
                f /= 1e9
            if n == "SM GHz" and u == "Mhz":
                f /= 1e3
            v = f"{f:.2f}"
        except ValueError:
            pass
        cells.append(v)
    print(f"| {k} | {r[name_i][:44]} | " + " | ".join(cells) + " |")
