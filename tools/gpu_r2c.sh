#!/bin/bash
# Round-2 GPU visit C: tensor-core forward at two CTAs per SM, new bench.py end to end, ncu of the forward kernel.
O=gpurun_out/${1:-r2c}
mkdir -p $O
echo "== kernel vs size (tc)"; timeout 300 python tools/kernel_vs_size.py > $O/kernel_vs_size_tc.txt 2>&1; cat $O/kernel_vs_size_tc.txt
echo "== pytest gpu (parity subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_parity_regimes.py tests/test_gpu_zz_baseline_configs.py -m gpu -q -x > $O/pytest.log 2>&1; tail -4 $O/pytest.log
echo "== bench default (c3 + scaling base + cpu baseline)"; timeout 900 python bench.py --steps 500 > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 2500 $O/bench_c3.json; tail -5 $O/bench_c3.err
for w in c2 c5 c4; do echo "== bench $w"; timeout 600 python bench.py --workload $w --steps 100 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; tail -3 $O/bench_$w.err; done
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cat $O/bench_reference.json | cut -c1-600
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "ms/step %.4f value %.4g e2e %.4g fwd %.1f adj %.1f bwd %.1f red %.1f us frac %.3f loss %.9g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["roofline"]["frac"], d["loss"]), d.get("strong_scaling_base"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== ncu full varfwd_tc"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hpv_varfwd_tc -s 2 -c 1 -f -o $O/varfwd_tc python tools/profile_step.py --steps 4 > $O/ncu_varfwd_tc.log 2>&1
python tools/ncu_mix.py $O/varfwd_tc.ncu-rep > $O/varfwd_tc_summary.txt 2>&1; head -14 $O/varfwd_tc_summary.txt
