#!/bin/bash
# One GPU-box visit: parity tests, bench (default + A/B variants), ncu launch list and full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh TAG
TAG=${1:-rXX}
VARIANTS=${2:-}
WARPS=${3:-}
EXTRA=${4:-}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench default"; timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 3000 $O/bench_default.json
echo "== bench two-tangent sweep"; HPV_BWD_DIR=0 timeout 300 python bench.py --steps 500 --no-cpu-baseline > $O/bench_dir0.json 2> $O/bench_dir0.err
for v in $VARIANTS; do
  echo "== bench variant lib $v"; HPV_LIB=$PWD/hp-vpinns_b200/libhpv_$v.so timeout 300 python bench.py --steps 500 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
done
echo "== bench c4 on one GPU"; timeout 300 python bench.py --workload c4 --steps 100 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err
# EXTRA: "label:ENV=VAL,ENV2=VAL2 label2:..." -> one short bench per entry with that environment
for x in $EXTRA; do
  lab=${x%%:*}; envs=$(echo ${x#*:} | tr ',' ' ')
  echo "== bench $lab ($envs)"; env $envs timeout 300 python bench.py --steps 500 --no-cpu-baseline > $O/bench_$lab.json 2> $O/bench_$lab.err
done
for w in $WARPS; do
  echo "== bench HPV_BWD_WARPS=$w"; HPV_BWD_WARPS=$w timeout 300 python bench.py --steps 500 --no-cpu-baseline > $O/bench_w$w.json 2> $O/bench_w$w.err
  for v in $VARIANTS; do
    HPV_LIB=$PWD/hp-vpinns_b200/libhpv_$v.so HPV_BWD_WARPS=$w timeout 300 python bench.py --steps 500 --no-cpu-baseline > $O/bench_${v}_w$w.json 2> $O/bench_${v}_w$w.err
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "ms/step %.4f value %.4g e2e %.4g fwd %.1f adj %.1f bwd %.1f red %.1f us frac %.3f geom %s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["roofline"]["frac"], {a:b for a,b in d["config"]["launch_geometry"].items() if a.startswith("bwd")}))
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python tools/profile_step.py --steps 6 > $O/launches.log 2>&1
echo "== ncu full: mlpbwd, varfwd"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_mlpbwd -s 2 -c 1 -f -o $O/mlpbwd python tools/profile_step.py --steps 4 > $O/ncu_mlpbwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_varfwd -s 2 -c 1 -f -o $O/varfwd python tools/profile_step.py --steps 4 > $O/ncu_varfwd.log 2>&1
ls -la $O
