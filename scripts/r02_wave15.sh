#!/bin/bash
# Wave 15: final single-GPU set of the round: bring-up of the reworked pool backward, GPU tests, c2 line, launch lists.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=120 timeout 300 python scripts/bringup.py elementwise_kernels pair_elementwise fused_pool 2>&1 | tee $O/w15_bringup.log | grep -E "FAIL|^case .* -> "
timeout 1200 python -m pytest tests -m gpu -q -x > $O/w15_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/w15_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/w15_c2.json 2> $O/w15_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
for f in ("w15_c2",):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); a=d.get("alt") or {}
        print(f, "%.1f"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], "| alt", a.get("value"), a.get("ms_per_step"), (a.get("e2e") or {}).get("value"), "clocks", d.get("clocks"), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(f, "failed", e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for p in fp32 bf16; do
  FCN8_GRAPHS=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/w15_launches_${p}.csv \
    python bench.py --profile --precision $p --steps 1 --warmup 1 > $O/w15_ncu_${p}.log 2>&1; echo "ncu $p rc=$?"
done
