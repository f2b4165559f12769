#!/bin/bash
# Round-2 GPU visit F: reverse sweep with the weight gradients on the tensor cores: parity, bench A/B.
O=gpurun_out/${1:-r2f}
mkdir -p $O
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -25 $O/pytest_gpu.log | cut -c1-300
for w in c3 c4 c5 c2; do
  for v in "tcw:HPV_BWD_TCW=1" "tc:HPV_BWD_TCW=0" "ffma:HPV_BWD_TC=0"; do
    lab=${v%%:*}; envs=${v#*:}
    echo "== bench $w $lab"; env $envs timeout 400 python bench.py --workload $w --steps 100 --no-cpu-baseline > $O/bench_${w}_$lab.json 2> $O/bench_${w}_$lab.err; tail -2 $O/bench_${w}_$lab.err
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "ms/step %.4f value %.4g e2e %.4g fwd %.1f adj %.1f bwd %.1f red %.1f us loss %.9g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["loss"]), {a:b for a,b in d["config"]["launch_geometry"].items() if a.startswith("bwd")})
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== ncu full mlpbwd_tcw"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_mlpbwd_tcw -s 2 -c 1 -f -o $O/mlpbwd_tcw python tools/profile_step.py --steps 4 > $O/ncu_mlpbwd_tcw.log 2>&1
python tools/ncu_mix.py $O/mlpbwd_tcw.ncu-rep > $O/mlpbwd_tcw_summary.txt 2>&1; cat $O/mlpbwd_tcw_summary.txt
