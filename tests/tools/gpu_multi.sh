#!/bin/bash
# Multi-GPU session (under gpurun --gpus N): hardware correctness of the gradient transports, then bench lines.
# Usage: bash tests/tools/gpu_multi.sh <tag> <N> [configs...]   (configs default: c2)
TAG=${1:-multi}; N=${2:-2}; shift; shift
CONFIGS=${@:-c2}
mkdir -p gpurun_out
PORT=29511
TRANSPORTS=${TRANSPORTS:-"nccl multimem multimem_red"}
run() { PORT=$((PORT+1)); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
run tests/tools/vp_check.py > gpurun_out/${TAG}_vp_check.json 2> gpurun_out/${TAG}_vp_check.err
echo "vp_check exit $?"; cat gpurun_out/${TAG}_vp_check.json; tail -5 gpurun_out/${TAG}_vp_check.err
for cfg in $CONFIGS; do
  for tr in $TRANSPORTS; do
    run bench.py --gpus $N --config $cfg --transport $tr --no-cpu-baseline > gpurun_out/${TAG}_${cfg}_${N}gpu_${tr}.json 2> gpurun_out/${TAG}_${cfg}_${N}gpu_${tr}.err
    echo "$cfg $tr exit $?"
    python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${TAG}_${cfg}_${N}gpu_${tr}.json") if l.startswith("{")][-1])
    print("  value %.1f M/s e2e %.1f ms/step %.2f transport %s check %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["config"]["transport"], d.get("allreduce_check")))
except Exception as ex:
    print("  no line:", ex)
PY
    tail -3 gpurun_out/${TAG}_${cfg}_${N}gpu_${tr}.err
  done
done
