#!/bin/bash
# compute-sanitizer over the whole GPU suite (memcheck) and the full-size cases (racecheck)
O=gpurun_out/${1:-san3}; mkdir -p $O
echo "== memcheck: all GPU tests"
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -q > $O/memcheck_all.log 2>&1
echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" $O/memcheck_all.log | tail -3
echo "== racecheck: full-size cases"
timeout 2400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity_regimes.py -m gpu -q > $O/racecheck_fullsize.log 2>&1
echo "exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck_fullsize.log | tail -3
