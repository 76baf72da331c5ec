#!/bin/bash
# Round-2 GPU box run: parity tests, smoke, the three bench workloads.  Usage (under gpurun): bash tests/tools/gpu_r2.sh <tag> [quick]
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --timeout=900 > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/gpu_tests_$TAG.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.log
if [ -z "$2" ]; then
  python bench.py --workload feature --steps 10 --warmup 3 > gpurun_out/bench_feature_$TAG.log 2>gpurun_out/bench_feature_$TAG.err; echo "bench feature rc=$?"
  cat gpurun_out/bench_feature_$TAG.log
  python bench.py --workload val > gpurun_out/bench_val_$TAG.log 2>gpurun_out/bench_val_$TAG.err; echo "bench val rc=$?"
  cat gpurun_out/bench_val_$TAG.log
fi
