#!/bin/bash
# A/B: programmatic-launch trigger of the forward kernel at the end of a CTA's work (product) vs right after set-up
O=gpurun_out/${1:-r2y}; mkdir -p $O
for rep in 1 2; do
for v in product pdlearly; do
  lib=hp-vpinns_b200/libhpv.so; [ $v = pdlearly ] && lib=tools/variants/libhpv_pdlearly.so
  for w in c3 c4; do
    HPV_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --steps 300 --no-cpu-baseline --no-scaling-base > $O/bench_${v}_$w.json 2> $O/bench_${v}_$w.err
    python - <<PY
import json
d=json.loads(open("$O/bench_${v}_$w.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
print("$v $w  ms/step %.4f e2e %.4f fwd %.1f adj %.1f bwd %.1f red %.1f sum %.1f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], sum(x["us"] for x in k.values())))
PY
  done
done
done 2>&1 | tee $O/summary.txt
