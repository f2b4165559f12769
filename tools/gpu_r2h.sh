#!/bin/bash
# Round-2 GPU visit H: full parity suite on the default configuration + A/B of the micro-fixes.
O=gpurun_out/${1:-r2h}
mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -8 $O/pytest_gpu.log | cut -c1-300
echo "== pytest -m gpu, tensor-core reverse sweep forced"; HPV_BWD_TC=1 timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_bwdtc.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu_bwdtc.log; tail -8 $O/pytest_gpu_bwdtc.log | cut -c1-300
for w in c3 c4; do
  for v in "default:HPV_X=0" "bwdtc:HPV_BWD_TC=1" "bwdtcw:HPV_BWD_TC=1,HPV_BWD_TCW=1"; do
    lab=${v%%:*}; envs=$(echo ${v#*:} | tr ',' ' ')
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
