#!/bin/bash
# Round-2 wave 3: promoted accumulation (PROMO kernels): bring-up cases, whole GPU suite, chunk-length sweep, bench.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=200 timeout 900 python scripts/bringup.py promoted_accumulation hwio_pair cta_pair_kernels pair_epilogues_and_wgrad > $O/w3_bringup.log 2>&1; echo "bringup rc=$?"
grep -E "hi\*hi MMAs|FAIL|^case" $O/w3_bringup.log | head -60
timeout 1500 python -m pytest tests -m gpu -q > $O/w3_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $O/w3_pytest.log
timeout 900 python scripts/promo_sweep.py > $O/w3_promo.log 2>&1; echo "promo rc=$?"; tail -8 $O/w3_promo.log
timeout 300 python __graft_entry__.py smoke > $O/w3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/w3_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w3_bench_c2.json 2> $O/w3_bench_c2.err; echo "c2 rc=$?"
cut -c1-300 $O/w3_bench_c2.json; tail -n 5 $O/w3_bench_c2.err
