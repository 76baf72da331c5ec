#!/bin/bash
# One GPU-box round: parity tests, bench, per-call step profile, ncu launch list, ncu full capture of the top kernel.
# Usage (under gpurun): bash tests/tools/gpu_round.sh <tag> [skip-ncu]
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/gpu_tests_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.log
python tests/tools/step_profile.py 400 > gpurun_out/step_profile_$TAG.log 2>&1; echo "profile rc=$?"
head -16 gpurun_out/step_profile_$TAG.log
python tests/tools/input_stage_bench.py > gpurun_out/input_stage_$TAG.log 2>&1; echo "input stage rc=$?"
tail -4 gpurun_out/input_stage_$TAG.log
if [ -z "$2" ]; then
  CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph"
  $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
  echo "launch list rc=$?"
  CMD2="python tests/tools/conv_bench.py dec"
  S2R_BENCH_EAGER=1 $CMD2 > gpurun_out/plain2_$TAG.log 2>&1 &&
  S2R_BENCH_EAGER=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 2 -o gpurun_out/prof_conv_tc_$TAG $CMD2 > gpurun_out/ncu2_$TAG.log 2>&1
  echo "full capture rc=$?"
fi
