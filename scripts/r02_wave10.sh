#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=120 timeout 400 python scripts/bringup.py deconv_stride8_loss deconv_stride2 decoder_heads hwio_pair 2>&1 | grep -E "FAIL|^case .* -> "
timeout 1500 python -m pytest tests -m gpu -q > $O/w10_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/w10_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w10_bench_c2.json 2> $O/w10_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w10_bench_c2.json').read().strip().splitlines()[-1])
print("c2 fp32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "| bf16", d["alt"]["value"], d["alt"]["ms_per_step"], "e2e", d["alt"]["e2e"]["value"])
PY
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
FCN8_GRAPHS=0 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/w10_launches_fp32.csv python bench.py --profile --precision fp32 --steps 1 --warmup 1 > $O/w10_ncu.log 2>&1; echo "ncu rc=$?"
