#!/bin/bash
# Round-2 GPU visit D: tensor-core reverse sweep: parity, bench A/B.
O=gpurun_out/${1:-r2d}
mkdir -p $O
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -25 $O/pytest_gpu.log | cut -c1-300
for w in c3 c4 c5 c2; do
  for tc in 1 0; do
    echo "== bench $w bwd_tc=$tc"; HPV_BWD_TC=$tc timeout 400 python bench.py --workload $w --steps 100 --no-cpu-baseline > $O/bench_${w}_bwdtc$tc.json 2> $O/bench_${w}_bwdtc$tc.err; tail -2 $O/bench_${w}_bwdtc$tc.err
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
