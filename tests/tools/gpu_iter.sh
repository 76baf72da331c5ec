#!/bin/bash
# One build->measure iteration on the GPU box: parity tests, the step bench, the per-call step profile.
# Usage (under gpurun): bash tests/tools/gpu_iter.sh <tag> [pytest -k filter]
TAG=${1:-it}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$2" > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"
else
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"
fi
tail -4 gpurun_out/gpu_tests_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$TAG.log"))
    print("ms_per_step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "dw fwd frac", d["roofline_depthwise"]["frac"], "dw bwd frac", d["roofline_depthwise_bwd"]["frac"], "launches", d["gpu_launches"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
timeout 600 python tests/tools/step_profile.py 60 > gpurun_out/step_profile_$TAG.txt 2>&1; head -16 gpurun_out/step_profile_$TAG.txt
