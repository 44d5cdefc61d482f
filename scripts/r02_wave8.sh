#!/bin/bash
# Round-2 wave 8: CTA-pair halo kernel.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=120 timeout 900 python scripts/bringup.py halo_pair_packed_forward halo_conv fused_pool > $O/w8_bringup.log 2>&1; echo "bringup rc=$?"
grep -E "FAIL|^case .* -> |Error|error" $O/w8_bringup.log | head -40
timeout 1500 python -m pytest tests -m gpu -q -x > $O/w8_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/w8_pytest.log
for mode in pair single; do
  if [ $mode = single ]; then export FCN8_DEBUG="11=1"; else unset FCN8_DEBUG; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w8_bench_c2_$mode.json 2> $O/w8_bench_c2_$mode.err; echo "c2 $mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/w8_bench_c2_$mode.json').read().strip().splitlines()[-1])
print("c2 $mode fp32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "| bf16", d["alt"]["value"], d["alt"]["ms_per_step"])
PY
done
unset FCN8_DEBUG
timeout 300 python scripts/wait_profile.py fp32 bf16 > $O/wait_profile_w8.txt 2>&1; grep -E "^ +[0-3] |^[a-z0-9]+:" $O/wait_profile_w8.txt | cut -c1-230
