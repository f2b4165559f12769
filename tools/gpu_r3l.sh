#!/bin/bash
# reverse-sweep start stagger between warp rows (HPV_BWD_STAGGER_NS), C3 and C4 kernel times
O=gpurun_out/${1:-r3l}; mkdir -p $O
for ns in 0 300 1000 3000 8000; do
  for w in c3 c4; do
    HPV_BWD_STAGGER_NS=$ns timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline --no-scaling-base > $O/b_${ns}_$w.json 2>/dev/null
    python - <<PY
import json
d=json.loads(open("$O/b_${ns}_$w.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
print("stagger $ns ns  $w  ms/step %.4f bwd %.1f" % (d["ms_per_step"], k["mlpbwd"]["us"]))
PY
  done
done 2>&1 | tee $O/summary.txt
