#!/bin/bash
# Round-2 wave 9: ALT epilogue + packed-bf16 pool: full GPU suite, bench c2.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/w9_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/w9_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w9_bench_c2.json 2> $O/w9_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w9_bench_c2.json').read().strip().splitlines()[-1])
print("c2 fp32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "| bf16", d["alt"]["value"], d["alt"]["ms_per_step"], "e2e", d["alt"]["e2e"]["value"])
PY
timeout 300 python scripts/wait_profile.py bf16 > $O/wait_profile_w12.txt 2>&1; grep -E "^ +[0-3] " $O/wait_profile_w12.txt | cut -c1-250
