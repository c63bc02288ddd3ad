#!/bin/bash
# development aid: per-kernel times of one dense Cholesky (n = 8064)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/chol_launches.csv python -c "
import sys; sys.path.insert(0, '.')
from apex_solver_b200.context import GpuContext
g = GpuContext(); print(g.dense_cholesky_bench(8064, 1))
" > gpurun_out/chol_ncu.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/chol_launches.csv")) if len(r) > 5]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    a = agg[r[ki][:60]]; a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%6.2f%% n=%5d avg=%9.1f  %s" % (100 * a[1] / tot, a[0], a[1] / a[0], k))
print("total", tot, rows[1][h.index("Metric Unit")])
PY
