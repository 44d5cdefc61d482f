#!/bin/bash
# Wave 14: packed-bf16 max-pool backward: bring-up (exact), GPU tests, c2 bench line.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=120 timeout 300 python scripts/bringup.py elementwise_kernels pair_elementwise fused_pool 2>&1 | tee $O/w14_bringup.log | grep -E "FAIL|^case .* -> "
timeout 1200 python -m pytest tests -m gpu -q -x > $O/w14_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/w14_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/w14_c2.json 2> $O/w14_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ("w14_c2",):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); a=d.get("alt") or {}
        print(f, "%.1f"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], "| alt", a.get("value"), a.get("ms_per_step"), (a.get("e2e") or {}).get("value"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "failed", e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
FCN8_GRAPHS=0 timeout 300 ncu --metrics $M --clock-control none --csv -k regex:maxpool_bwd --log-file $O/w14_pool.csv \
    python bench.py --profile --precision bf16 --steps 1 --warmup 1 > $O/w14_ncu.log 2>&1; echo "ncu rc=$?"
grep maxpool $O/w14_pool.csv | grep time_duration | tail -5 | cut -d, -f5,12-
