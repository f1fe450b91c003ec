#!/bin/bash
# Quick GPU iteration: parity tests, one bench line (our arm), optional ncu of the blend kernels.
# Usage (under gpurun, repo root):  bash tests/tools/gpu_quick.sh <tag> [ncu]
TAG=${1:-quick}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench_b200.json 2> gpurun_out/${TAG}_bench_b200.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_b200.json"))
print("value %.1f M/s  e2e %.1f M/s  ms/step %.2f" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"]))
print({k: round(v["ms"], 4) for k, v in d["stages"].items()})
PY
tail -3 gpurun_out/${TAG}_bench_b200.err
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_' -s 12 -c 4 -f -o gpurun_out/${TAG}_prof \
     python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
  ls -la gpurun_out/${TAG}_prof.ncu-rep
fi
