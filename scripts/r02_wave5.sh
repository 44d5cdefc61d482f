#!/bin/bash
# Round-2 wave 5: conv1 kernels with 8+8 warps, wgrad_halo on strided views, 128-bit logits / softmax stores.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=200 timeout 900 python scripts/bringup.py conv1_direct deconv_stride8_loss wgrad_bf16_3x3 pair_epilogues_and_wgrad halo_conv > $O/w5_bringup.log 2>&1; echo "bringup rc=$?"
grep -E "FAIL|^case" $O/w5_bringup.log | head -40
timeout 1500 python -m pytest tests -m gpu -q > $O/w5_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/w5_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w5_bench_c2.json 2> $O/w5_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w5_bench_c2.json').read().strip().splitlines()[-1])
print("c2 fp32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "| bf16", d["alt"]["value"], d["alt"]["ms_per_step"], "e2e", d["alt"]["e2e"]["value"])
PY
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > $O/w5_bench_c4.json 2> $O/w5_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w5_bench_c4.json').read().strip().splitlines()[-1])
print("c4 fp32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["softmax_output"])
PY
