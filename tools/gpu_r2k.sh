#!/bin/bash
# Timing experiment: the reverse sweep without its weight-gradient GEMMs / without its layer products (libraries built
# with -DHPV_EXP_NO_WGRAD / -DHPV_EXP_NO_PROD -DHPV_EXP_NO_MMA; results are wrong by construction, only the times count).
O=gpurun_out/${1:-r2k}
mkdir -p $O
for v in full nowgrad noprod; do
  for tc in 0 1; do
    lib=tools/variants/libhpv_$v.so; [ $v = full ] && lib=hp-vpinns_b200/libhpv.so
    HPV_LIB=$PWD/$lib HPV_BWD_TC=$tc timeout 300 python bench.py --workload c3 --steps 200 --no-cpu-baseline --no-scaling-base > $O/bench_${v}_tc$tc.json 2> $O/bench_${v}_tc$tc.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${v}_tc$tc.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
    print("$v bwd_tc=$tc  ms/step %.4f fwd %.1f adj %.1f bwd %.1f red %.1f" % (d["ms_per_step"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"]))
except Exception as e:
    print("$v $tc unreadable", e)
PY
  done
done 2>&1 | tee $O/summary.txt
