#!/bin/bash
# the GPU test-suite under the non-default code paths
O=gpurun_out/${1:-var}; mkdir -p $O
for v in "HPV_FWD_TC=0" "HPV_BWD_TC=1" "HPV_BWD_TC=0" "HPV_GRAPH=0" "HPV_FWD_BALANCE=0" "HPV_PDL=0"; do
  tag=$(echo $v | tr '=' '_')
  env $v timeout 900 python -m pytest tests -m gpu -q > $O/pytest_$tag.log 2>&1
  echo "$v: exit $? $(tail -n 1 $O/pytest_$tag.log)"
done 2>&1 | tee $O/summary.txt
