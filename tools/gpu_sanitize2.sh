#!/bin/bash
# compute-sanitizer over the small-size GPU parity tests: memcheck on the whole file, racecheck + synccheck on a few cases
O=gpurun_out/${1:-san2}; mkdir -p $O
echo "== memcheck tests/test_gpu_parity.py"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/memcheck_parity.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_parity.log | tail -3
for tool in racecheck synccheck; do
  echo "== $tool (forward/backward of p2d_vf1, partition weighting, loss total)"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partition_weighting or loss_total or (vs_oracle and p2d_vf1)" > $O/${tool}_subset.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/${tool}_subset.log | tail -3
done
