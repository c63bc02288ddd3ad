#!/bin/bash
# development aid: compute-sanitizer memcheck + synccheck over a few small GPU tests of every kernel group, racecheck (shared-memory
# hazards; slow) over the PCG-tail and explicit-solve tests
SEL='test_linearize_blocks_cost_matvec and (bal_selfcal_huber or radtan_selfcal or bal_ba_huber) or test_deterministic_operator or test_pcg_fused_tail or test_long_tracks or test_solve_augmented or test_shared_intrinsics or test_per_block or test_lm_with_jacobi or test_observer'
RSEL='test_pcg_fused_tail or test_solve_augmented'
for tool in memcheck synccheck racecheck; do
  sel="$SEL"; [ $tool = racecheck ] && sel="$RSEL"
  timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -q -x --tb=line -p no:cacheprovider -k "$sel" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | sort -rn | head -8
done
