#!/bin/bash
# Lean multi-GPU visit: C4 strong-scaling bench lines for the listed N (peer-memory exchange; "nccl" as 3rd arg adds NCCL).
TAG=${1:-rXX}; NS=${2:-8}; EXTRA=${3:-}
O=gpurun_out/$TAG; mkdir -p $O
P=29611
for N in $NS; do
  for COLL in peer $EXTRA; do
    P=$((P+1))
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --collective $COLL > $O/scale_n${N}_$COLL.json 2> $O/scale_n${N}_$COLL.err
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/scale_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "ms/step %.4f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
