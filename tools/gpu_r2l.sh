#!/bin/bash
# Forward-kernel timing experiment (no MMA / no projection / no activation variants, results wrong by construction),
# then parity + C3/C4 bench of the working tree.
O=gpurun_out/${1:-r2l}
mkdir -p $O
for v in full fwd_nomma fwd_noproj fwd_noact; do
  lib=tools/variants/libhpv_$v.so; [ $v = full ] && lib=hp-vpinns_b200/libhpv.so
  for w in c3 c4; do
    HPV_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline --no-scaling-base > $O/bench_${v}_$w.json 2> $O/bench_${v}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${v}_$w.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
    print("$v $w  ms/step %.4f fwd %.1f adj %.1f bwd %.1f red %.1f" % (d["ms_per_step"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"]))
except Exception as e:
    print("$v $w unreadable", e)
PY
  done
done 2>&1 | tee $O/summary.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -n 4 $O/pytest.log
