#!/usr/bin/env python
"""Tabulate an ncu --csv metrics log of one bench step (see profiles/README.md)."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    recs = list(csv.DictReader(lines))
    by = collections.OrderedDict()
    for x in recs:
        d = by.setdefault(x["ID"], {"name": x["Kernel Name"].split("(")[0][-30:], "grid": x["Grid Size"]})
        v = float(x["Metric Value"].replace(",", ""))
        u, m = x["Metric Unit"], x["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        if "bytes" in m:
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[u]
        d[m] = v
    print(f"{'id':>4} {'kernel':30} {'grid':>14} {'us':>8} {'rdMB':>8} {'wrMB':>8} {'l2MB':>8} {'GB/s':>6} {'tens%':>6} {'sm%':>5} {'dram%':>6}")
    tot = 0
    for k, d in by.items():
        t = d["gpu__time_duration.sum"]
        tot += t
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        print(f"{k:>4} {d['name']:30} {d['grid']:>14} {t:8.1f} {rd:8.1f} {wr:8.1f} {d['lts__t_bytes.sum']:8.1f} "
              f"{(rd + wr) / t * 1e3 / 1e3:6.0f} {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']:6.1f} "
              f"{d['sm__throughput.avg.pct_of_peak_sustained_elapsed']:5.1f} "
              f"{d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']:6.1f}")
    print("total us", round(tot, 1))


if __name__ == "__main__":
    main(sys.argv[1])
