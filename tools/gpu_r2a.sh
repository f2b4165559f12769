#!/bin/bash
# Round-2 GPU visit A: parity tests with the tensor-core forward kernel, bench A/B (tensor-core vs FFMA forward), MN-major probe.
O=gpurun_out/${1:-r2a}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
echo "== bench tc"; timeout 400 python bench.py --steps 300 --no-cpu-baseline > $O/bench_tc.json 2> $O/bench_tc.err; tail -c 1500 $O/bench_tc.json
echo "== bench ffma"; HPV_FWD_TC=0 timeout 400 python bench.py --steps 300 --no-cpu-baseline > $O/bench_ffma.json 2> $O/bench_ffma.err
echo "== bench c4 tc"; timeout 400 python bench.py --workload c4 --steps 60 --no-cpu-baseline > $O/bench_c4_tc.json 2> $O/bench_c4_tc.err
echo "== bench c4 ffma"; HPV_FWD_TC=0 timeout 400 python bench.py --workload c4 --steps 60 --no-cpu-baseline > $O/bench_c4_ffma.json 2> $O/bench_c4_ffma.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "ms/step %.4f value %.4g e2e %.4g fwd %.1f adj %.1f bwd %.1f red %.1f us loss %.9g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["loss"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== umma probe 14"; timeout 60 tools/probes/umma_probe.bin 14 > $O/umma14.txt 2>&1; cat $O/umma14.txt | cut -c1-250
