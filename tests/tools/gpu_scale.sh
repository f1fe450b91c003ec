#!/bin/bash
# Scaling session on N GPUs of one box (under gpurun --gpus N): hardware correctness of the gradient transports, then
# one bench line per (config, transport).  Usage: bash tests/tools/gpu_scale.sh <tag> <N> "<cfg:transport:steps> ..."
TAG=$1; N=$2; RUNS=$3
mkdir -p gpurun_out
PORT=29600
run() { PORT=$((PORT+1)); timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; }
if [ "$N" -gt 1 ]; then
  run tests/tools/vp_check.py > gpurun_out/${TAG}_${N}gpu_vp_check.json 2> gpurun_out/${TAG}_${N}gpu_vp_check.err
  grep '^{' gpurun_out/${TAG}_${N}gpu_vp_check.json | tail -1 | cut -c1-600; tail -2 gpurun_out/${TAG}_${N}gpu_vp_check.err
fi
for spec in $RUNS; do
  cfg=${spec%%:*}; rest=${spec#*:}; tr=${rest%%:*}; steps=${rest#*:}
  out=gpurun_out/${TAG}_${cfg}_${N}gpu_${tr}
  if [ "$N" -gt 1 ]; then
    run bench.py --gpus $N --config $cfg --transport $tr --steps $steps --no-cpu-baseline > $out.json 2> $out.err
  else
    timeout 700 python bench.py --config $cfg --steps $steps --no-cpu-baseline --no-train-iteration > $out.json 2> $out.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$out.json") if l.startswith("{")][-1])
    print("$cfg N=$N $tr: value %.1f M/s e2e %.1f ms/step %.3f transport %s check %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], d["config"].get("transport"), (d.get("allreduce_check") or {}).get("max_rel_err")))
except Exception as ex:
    print("$cfg N=$N $tr: no line:", ex)
PY
  tail -2 $out.err | cut -c1-300
done
