#!/usr/bin/env python
"""Summarise .ncu-rep files (read here, on the CPU box) into profiles/*.txt.

    python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep
"""
import csv
import io
import os
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes.sum", "sm__cycles_active.avg", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg"]


def summarise(path: str) -> str:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return f"{path}: no kernels\n"
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append(f"kernel: {name}")
        for i, h in enumerate(hdr):
            if h in WANT:
                out.append(f"  {h:<70} {vals[i]:>16} {units[i]}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    os.makedirs("profiles", exist_ok=True)
    for p in sys.argv[1:]:
        text = summarise(p)
        dst = os.path.join("profiles", os.path.basename(p).replace(".ncu-rep", ".ncu.txt"))
        with open(dst, "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on   ({os.path.basename(p)})\n" + text)
        print(text)
