#!/bin/bash
# Multi-GPU visit: peer-exchange test + C4 strong-scaling bench lines (peer-memory exchange vs NCCL).
# Usage: bash tools/gpu_multi.sh TAG "N1 N2 ..."
TAG=${1:-rXX}
NS=${2:-2}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "== peer test"; timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -q > $O/pytest_peer.log 2>&1; echo "exit $?" >> $O/pytest_peer.log; tail -4 $O/pytest_peer.log
P=29511
for N in $NS; do
  for COLL in peer nccl; do
    P=$((P+1))
    echo "== bench N=$N $COLL"
    if [ "$N" = "1" ]; then
      [ "$COLL" = "peer" ] && timeout 600 python bench.py --gpus 1 --workload c4 --steps 200 --warmup 10 --no-cpu-baseline > $O/scale_n1.json 2> $O/scale_n1.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline --collective $COLL > $O/scale_n${N}_$COLL.json 2> $O/scale_n${N}_$COLL.err
    fi
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/scale_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "ms/step %.4f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d["config"]["step"][-60:])
    except Exception as e:
        print(f, "unreadable", e)
PY
