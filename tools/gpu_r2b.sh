#!/bin/bash
# Round-2 GPU visit B: profile of the tensor-core forward kernel.
O=gpurun_out/${1:-r2b}
mkdir -p $O
echo "== kernel vs size (tc)"; timeout 300 python tools/kernel_vs_size.py > $O/kernel_vs_size_tc.txt 2>&1; cat $O/kernel_vs_size_tc.txt
echo "== kernel vs size (ffma)"; HPV_FWD_TC=0 timeout 300 python tools/kernel_vs_size.py > $O/kernel_vs_size_ffma.txt 2>&1; cat $O/kernel_vs_size_ffma.txt
echo "== ncu full varfwd_tc"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_varfwd_tc -s 2 -c 1 -f -o $O/varfwd_tc python tools/profile_step.py --steps 4 > $O/ncu_varfwd_tc.log 2>&1
tail -3 $O/ncu_varfwd_tc.log
python tools/ncu_mix.py $O/varfwd_tc.ncu-rep > $O/varfwd_tc_summary.txt 2>&1; cat $O/varfwd_tc_summary.txt
echo "== graph A/B"; for g in 1 0; do HPV_GRAPH=$g timeout 300 python bench.py --steps 500 --no-cpu-baseline --no-scaling-base > $O/bench_graph$g.json 2> $O/bench_graph$g.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_graph$g.json").read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
    print("graph=$g ms/step %.4f e2e %.4g launches %d fwd %.1f adj %.1f bwd %.1f red %.1f loss %.9g" % (d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["loss"]))
except Exception as e: print("unreadable", e); print(open("$O/bench_graph$g.err").read()[-2000:])
PY
done
echo "== pytest vpinn/driver (graph path)"; timeout 600 python -m pytest tests/test_gpu_vpinn.py tests/test_gpu_driver_e2e.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest_graph.log 2>&1; tail -5 $O/pytest_graph.log
