#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a compact per-launch table (CSV on stdout / file).

    python tools/ncu_summary.py gpurun_out/prof_r1.ncu-rep > profiles/r1_ncu_summary.csv
"""
import csv
import io
import subprocess
import sys

COLS = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_bytes.sum": "l2_MB",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "l1tex__data_pipe_lsu_wavefronts.sum": "l1_wavefronts",
    "smsp__cycles_active.avg": "cycles",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(header)}
    out = csv.writer(sys.stdout)
    names = [c for c in COLS if c in idx]
    out.writerow(["id", "kernel"] + [COLS[c] for c in names])
    for r in data:
        vals = []
        for c in names:
            v, u = r[idx[c]].replace(",", ""), units[idx[c]]
            try:
                f = float(v)
                if COLS[c] == "dur_us":
                    f = f / 1000 if u in ("ns", "nsecond") else (f * 1000 if u in ("ms", "msecond") else f)
                if COLS[c].endswith("_MB"):
                    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * f
                vals.append(f"{f:.3f}" if f < 1e6 else f"{f:.0f}")
            except ValueError:
                vals.append(v)
        out.writerow([r[idx["ID"]], r[idx["Kernel Name"]][:48]] + vals)


if __name__ == "__main__":
    main()
