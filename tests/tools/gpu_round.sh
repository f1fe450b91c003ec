#!/bin/bash
# One GPU-box session: tests, golden vectors, bench (both arms), ncu launch list + full capture.
# Usage (from the repo root, under gpurun):  bash tests/tools/gpu_round.sh [tag] [parts...]
#   parts: tests golden bench trainer ncu  (default: tests golden bench ncu)
TAG=${1:-r01}; shift
PARTS=${@:-tests golden bench ncu}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
for part in $PARTS; do
case $part in
tests)
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -30 gpurun_out/${TAG}_pytest_gpu.log ;;
golden)
  timeout 300 python tests/golden/make_golden.py > gpurun_out/${TAG}_golden.log 2>&1; tail -8 gpurun_out/${TAG}_golden.log ;;
trainer)
  timeout 600 python tests/tools/bench_trainer_ops.py > gpurun_out/${TAG}_trainer_ops.json 2> gpurun_out/${TAG}_trainer_ops.err
  timeout 300 python tests/tools/bench_surface.py > gpurun_out/${TAG}_surface.json 2>> gpurun_out/${TAG}_trainer_ops.err
  cat gpurun_out/${TAG}_trainer_ops.json gpurun_out/${TAG}_surface.json ;;
bench)
  timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  timeout 600 python bench.py > gpurun_out/${TAG}_bench_b200.json 2> gpurun_out/${TAG}_bench_b200.err
  cat gpurun_out/${TAG}_bench_reference.json gpurun_out/${TAG}_bench_b200.json; tail -5 gpurun_out/${TAG}_bench_b200.err ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
     python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline --no-train-iteration --no-settle > gpurun_out/${TAG}_ncu_bench.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_|project_|tile_|scatter|acc_clear' -s 81 -c 10 -f -o gpurun_out/${TAG}_prof \
     python bench.py --steps 2 --warmup 3 --views 2 --no-cpu-baseline --no-train-iteration --no-settle > gpurun_out/${TAG}_ncu_full.log 2>&1
  # the caller-side rows: loss, post-processing, mip filter, raw-parameter projection
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'photometric_(fwd|bwd)|surface_|mip_distance|project_.*true' -s 8 -c 8 -f -o gpurun_out/${TAG}_prof_next \
     python tests/tools/bench_trainer_ops.py --iters 2 > gpurun_out/${TAG}_ncu_next.log 2>&1
  ls -la gpurun_out/ | tail -20 ;;
esac
done
