#!/bin/bash
# Round-2 final single-GPU visit: parity tests, smoke, the bench lines of record, ncu launch list and full captures.
O=gpurun_out/${1:-final}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log | cut -c1-200
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
echo "== bench default"; timeout 900 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err; tail -2 $O/bench_c3.err
for w in c2 c4 c5; do echo "== bench $w"; timeout 600 python bench.py --workload $w --steps 300 > $O/bench_$w.json 2> $O/bench_$w.err; tail -2 $O/bench_$w.err; done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "ms/step %.4f value %.4g e2e %.4g (%.3f ms) fwd %.1f adj %.1f bwd %.1f red %.1f us frac %.3f loss %.9g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["roofline"]["frac"], d["loss"]))
        print("    clocks", d["clocks"], "cpu", (d.get("cpu_baseline") or {}).get("value"), ((d.get("cpu_baseline") or {}).get("ref_factorised") or {}).get("value"), "base", (d.get("strong_scaling_base") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python tools/profile_step.py --steps 8 > $O/launches.log 2>&1
echo "== ncu full: varfwd_tc, mlpbwd (c3); mlpbwd (c4)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_varfwd_tc -s 2 -c 1 -f -o $O/varfwd_tc python tools/profile_step.py --steps 4 > $O/ncu_varfwd_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_mlpbwd -s 2 -c 1 -f -o $O/mlpbwd python tools/profile_step.py --steps 4 > $O/ncu_mlpbwd.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"hpv_mlpbwd|hpv_varfwd_tc" -s 4 -c 2 --csv --log-file $O/c4_dram.csv python tools/profile_step.py --workload c4 --steps 4 > $O/ncu_c4.log 2>&1
python tools/ncu_mix.py $O/varfwd_tc.ncu-rep > $O/varfwd_tc_summary.txt 2>&1
python tools/ncu_mix.py $O/mlpbwd.ncu-rep > $O/mlpbwd_summary.txt 2>&1
python tools/ncu_segments2.py $O/varfwd_tc.ncu-rep > $O/varfwd_tc_segments.txt 2>&1
for r in varfwd_tc mlpbwd; do ncu -i $O/$r.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[-1]))
for k in rows[0]:
    if any(t in k for t in ['dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_subpipe_hmma_cycles_active_realtime','sm__cycles_elapsed.max','sm__inst_executed_pipe_tmem.avg','gpu__time_duration.sum','sm__warps_active.avg.pct','smsp__issue_active.avg.pct']) and 'per_second' not in k and 'pct_of_peak_sustained_elapsed' not in k:
        print('%-85s %s'%(k,d[k]))
" > $O/${r}_raw_selected.txt; done
cat $O/c4_dram.csv | tail -8; cat $O/varfwd_tc_raw_selected.txt
