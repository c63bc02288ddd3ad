#!/bin/bash
# dense Cholesky at the C4 sizes for a few outer panel widths (APEX_CHOL_PANEL)
for p in "$@"; do
APEX_CHOL_PANEL=$p python - <<PY
import sys; sys.path.insert(0, ".")
from apex_solver_b200.context import GpuContext
g = GpuContext()
for n in (14016, 24064, 28032):
    ms = g.dense_cholesky_bench(n, 2)
    print("panel $p n", n, "ms %.1f" % ms, "TF %.2f" % (n ** 3 / 3 / ms / 1e9), flush=True)
PY
done
