#!/bin/bash
# Round-2 multi-GPU visit: peer-exchange tests, sharded classes with boundary loss, bench with the parity block.
N=${2:-2}
O=gpurun_out/${1:-r2g}
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
echo "== pytest (regimes + peer)"; timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_parity_regimes.py tests/test_gpu_fullsize.py -m gpu -q > $O/pytest_peer.log 2>&1; echo "exit $?" >> $O/pytest_peer.log; tail -6 $O/pytest_peer.log | cut -c1-250
for n in $N; do
  echo "== bench --gpus $n"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 300 --warmup 10 --no-cpu-baseline > $O/bench_n$n.json 2> $O/bench_n$n.err
  tail -3 $O/bench_n$n.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bench_n$n.json").read().strip().splitlines() if l.startswith("{")][-1]); k=d["roofline"]["kernels"]
    print("N=$n ms/step %.4f value %.4g e2e %.4g (%.3f ms) fwd %.1f adj %.1f bwd %.1f red %.1f loss %.9g" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], k["varfwd"]["us"], k["adjproj"]["us"], k["mlpbwd"]["us"], k["gradreduce+unpad"]["us"], d["loss"]))
    print("   parity", d.get("parity"))
except Exception as e: print("unreadable", e)
PY
done
