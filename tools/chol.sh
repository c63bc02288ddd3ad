#!/bin/bash
# development aid: explicit-Schur tests + dense Cholesky vs cuSOLVER potrf / cuBLAS DGEMM
timeout 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider -k "explicit or solve_augmented or ladybug or trafalgar_full or long_tracks or fixed" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_quick.log | cut -c1-300
timeout 600 python tools/chol_compare.py > gpurun_out/chol_compare.json 2> gpurun_out/chol_compare.err; python - <<PY
import json
d = json.loads(open("gpurun_out/chol_compare.json").read().strip().splitlines()[-1])
print("dgemm", round(d["dgemm_8192_tflops"], 1))
for r in d["rows"]:
    print(r["n"], "ours %.1f ms %.1f TF | cusolver %.1f ms %.1f TF | ratio %.2f" % (r["ours_ms"], r["ours_tflops"], r["cusolver_ms"], r["cusolver_tflops"], r["ours_over_cusolver"]))
PY
tail -3 gpurun_out/chol_compare.err
