# Development probe: step time of a workload ($1: c3|c4) under launch-shape / PDL variants ($2..: label:ENV=V,ENV=V)
WL=$1; shift
mkdir -p gpurun_out/probe
for v in "$@"; do
  lab=${v%%:*}; envs=$(echo ${v#*:} | tr ',' ' ')
  env $envs timeout 300 python bench.py --workload $WL --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/probe/${WL}_$lab.json 2> gpurun_out/probe/${WL}_$lab.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/probe/${WL}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d["roofline"]["kernels"]
        print(f.split("/")[-1], "ms/step %.4f fwd %.1f adj %.1f bwd %.1f e2e_ms %.3f geom %s"%(d["ms_per_step"],k["varfwd"]["us"],k["adjproj"]["us"],k["mlpbwd"]["us"],d["e2e"]["ms_per_step"],{a:b for a,b in d["config"]["launch_geometry"].items() if a.startswith("bwd")}))
    except Exception as e: print(f,"unreadable",e)
PY
