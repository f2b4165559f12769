#!/bin/bash
# parity + C3/C4/C2/C5 bench of the working tree
O=gpurun_out/${1:-r2m}
mkdir -p $O
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; tail -n 4 $O/pytest.log
for w in c3 c4 c2 c5; do
  timeout 300 python bench.py --workload $w --steps 200 --no-cpu-baseline --no-scaling-base > $O/bench_$w.json 2> $O/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$w.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
    print("$w  ms/step %.4f e2e %.4f fwd %.1f adj %.1f bwd %.1f red %.1f fwd_only %.4f loss %.9g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["forward_only"]["ms"], d["loss"]))
except Exception as e:
    print("$w unreadable", e); print(open("$O/bench_$w.err").read()[-1500:])
PY
done 2>&1 | tee $O/summary.txt
