#!/bin/bash
# Round-2 wave 2: new decoder / fused epilogues: kernel bring-up cases, engine tests, RZ calibration, bench.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=120 timeout 1500 python scripts/bringup.py all > $O/w2_bringup.log 2>&1; echo "bringup rc=$?"
grep -E "^====|FAIL" $O/w2_bringup.log | head -80
timeout 1200 python -m pytest tests/test_gpu_engine.py tests/test_gpu_class_surface.py -m gpu -q -s > $O/w2_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/w2_pytest.log
timeout 600 python scripts/rz_calibrate.py > $O/w2_rz.log 2>&1; echo "rz rc=$?"; tail -8 $O/w2_rz.log
timeout 300 python __graft_entry__.py smoke > $O/w2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/w2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w2_bench_c2.json 2> $O/w2_bench_c2.err; echo "c2 rc=$?"
cut -c1-300 $O/w2_bench_c2.json; tail -n 5 $O/w2_bench_c2.err
