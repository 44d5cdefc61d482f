#!/bin/bash
# One GPU-box visit: GPU tests, smoke, bench, ncu launch list (+ optional full capture of the top kernel).
# Usage (via gpurun): bash scripts/gpu_round.sh <tag> [tests|bench|ncu|full ...]
tag=${1:-r01}; shift
what=${@:-tests bench ncu}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests) timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$tag.log;;
    bench) timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err;;
    ncu) for p in bf16 fp32; do
           timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${p}_$tag.csv \
             python bench.py --profile --precision $p --steps 1 --warmup 1 > gpurun_out/ncu_${p}_$tag.log 2>&1; echo "ncu $p rc=$?"; done;;
    full) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 20 -c 3 -o gpurun_out/prof_conv_$tag \
             python bench.py --profile --precision bf16 --steps 1 --warmup 1 > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?";;
  esac
done
