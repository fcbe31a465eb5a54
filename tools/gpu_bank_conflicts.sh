#!/bin/bash
# shared-memory wavefronts and bank-conflict wavefronts per launch over one bench step
mkdir -p gpurun_out
ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum \
    --clock-control none --launch-skip 140 --launch-count 100 --csv --log-file gpurun_out/bank_conflicts.csv \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_bc.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/bank_conflicts.csv")) if len(r) > 10]
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[ix["ID"]], r[ix["Kernel Name"]][:60])
    per.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
for (i, name), m in per.items():
    w = m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0); c = m.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0)
    print(f"{i:>4} {name:60s} {m.get('gpu__time_duration.sum', 0) / 1000:8.1f} us  wavefronts {w / 1e6:8.2f} M  conflicts {c / 1e6:8.2f} M ({100 * c / w if w else 0:4.1f} %)")
PY
