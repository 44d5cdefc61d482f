#!/bin/bash
# One GPU-box visit: GPU tests, smoke, bench, ncu launch lists (+ optional full capture / probes / wait profile).
# Usage (via gpurun): bash scripts/gpu_round.sh <tag> [tests|smoke|bench|ncu|full|probes|waits ...]
# Afterwards, here: python scripts/step_table.py gpurun_out/launches_<p>_<tag>.csv <p> profiles/r01_launches_<p>_final.md \
#                   profiles/r01_step_kernels_final.json      (bench.py reads the JSON for roofline.traffic)
tag=${1:-r01}; shift
what=${@:-tests bench ncu}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for w in $what; do
  case $w in
    tests) timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$tag.log;;
    bench) timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err;;
    ncu) for p in bf16 fp32; do
           FCN8_GRAPHS=0 timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_${p}_$tag.csv \
             python bench.py --profile --precision $p --steps 1 --warmup 1 > gpurun_out/ncu_${p}_$tag.log 2>&1; echo "ncu $p rc=$?"; done;;
    full) FCN8_GRAPHS=0 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|wgrad_gemm_kernel" -s 46 -c 16 -o gpurun_out/prof_gemm_$tag \
             python bench.py --profile --precision fp32 --steps 1 --warmup 1 > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?";;
    probes) for b in mma_rate_probe issue_loop_probe tma_feed_probe pair_gemm_probe; do   # build first: see each file's header
              timeout 60 scripts/_build/$b > gpurun_out/$b.txt 2>&1; echo "$b rc=$?"; done;;
    waits) timeout 300 python scripts/wait_profile.py bf16 fp32 > gpurun_out/wait_profile_$tag.txt 2>&1; echo "waits rc=$?";;
  esac
done
