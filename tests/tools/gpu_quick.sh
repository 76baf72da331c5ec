#!/bin/bash
# Quick GPU-box check: parity tests, conv micro-bench, step bench.  Usage (under gpurun): bash tests/tools/gpu_quick.sh <tag> [conv filter]
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/gpu_tests_$TAG.log
python tests/tools/conv_bench.py "$2" > gpurun_out/conv_bench_$TAG.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.log"))
print("ms_per_step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "val", d.get("val",{}).get("value"))
PY
