#!/bin/bash
# A/B of kernel variants selected by environment (one GPU).  Usage: bash tests/tools/gpu_variants.sh <tag> <ENVVAR> v1 v2 ...
TAG=$1; VAR=$2; shift; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v timeout 600 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/${TAG}_${v}_parity.log 2>&1
  echo "$VAR=$v parity: $(tail -1 gpurun_out/${TAG}_${v}_parity.log)"
  env $VAR=$v timeout 300 python bench.py --no-cpu-baseline --no-train-iteration > gpurun_out/${TAG}_${v}_bench.json 2> gpurun_out/${TAG}_${v}_bench.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_${v}_bench.json") if l.startswith("{")][-1])
    print("  $VAR=$v value %.1f M/s e2e %.1f ms/step %.2f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"]), {k: round(x["ms"], 4) for k, x in d["stages"].items()})
except Exception as ex:
    print("  no line:", ex)
PY
  tail -2 gpurun_out/${TAG}_${v}_bench.err
done
