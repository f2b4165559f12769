#!/bin/bash
# forward structure A/B (old barrier structure + N=24 vs dedicated issuing warp), adjproj persistent grid A/B, ncu of the new forward
O=gpurun_out/${1:-r2o}
mkdir -p $O
run() {  # name, env...
  name=$1; shift
  for w in c3 c4; do
    env "$@" timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline --no-scaling-base > $O/bench_${name}_$w.json 2> $O/bench_${name}_$w.err
    python - <<PY
import json
try:
    d=json.loads(open("$O/bench_${name}_$w.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
    print("$name $w  ms/step %.4f fwd %.1f adj %.1f bwd %.1f red %.1f adj_grid %d" % (d["ms_per_step"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["config"]["launch_geometry"]["adj_grid"]))
except Exception as e:
    print("$name $w unreadable", e)
PY
  done
}
{
run current X=1
run adjpersist3 HPV_ADJ_PERSIST=3
run old_n24 HPV_LIB=$PWD/tools/variants/libhpv_old_n24.so
} 2>&1 | tee $O/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_varfwd_tc -s 2 -c 1 -f -o $O/varfwd_tc python tools/profile_step.py --steps 4 > $O/ncu_varfwd_tc.log 2>&1
python tools/ncu_mix.py $O/varfwd_tc.ncu-rep > $O/varfwd_tc_summary.txt 2>&1
python tools/ncu_segments2.py $O/varfwd_tc.ncu-rep > $O/varfwd_tc_segments.txt 2>&1
head -30 $O/varfwd_tc_summary.txt
