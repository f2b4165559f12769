#!/bin/bash
O=gpurun_out/${1:-r2i}
mkdir -p $O
echo "== pytest gpu (parity subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_parity_regimes.py tests/test_gpu_zz_baseline_configs.py tests/test_gpu_vpinn.py -m gpu -q -x > $O/pytest.log 2>&1; tail -4 $O/pytest.log
for w in c3 c4 c5 c2; do echo "== bench $w"; timeout 400 python bench.py --workload $w --steps 100 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; tail -2 $O/bench_$w.err; done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_c*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "ms/step %.4f value %.4g e2e %.4g fwd %.1f adj %.1f bwd %.1f red %.1f us loss %.9g fwd_only %.4f ms" % (d["ms_per_step"], d["value"], d["e2e"]["value"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["loss"], d["forward_only"]["ms"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
