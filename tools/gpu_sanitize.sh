#!/bin/bash
# compute-sanitizer over smoke() (forward + backward + two Adam steps on a small case): memcheck, synccheck, racecheck
O=gpurun_out/${1:-san}; mkdir -p $O
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > $O/$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke:" $O/$tool.log | head -6
  grep -E "=========     at |Race reported|hazard" $O/$tool.log | sort | uniq -c | sort -rn | head -8
done
