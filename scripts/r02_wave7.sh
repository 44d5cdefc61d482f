#!/bin/bash
# Round-2 wave 7: dynamic tile scheduling (cluster launch control) in conv_gemm_kernel.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=120 timeout 900 python scripts/bringup.py conv_bf16_1x1_single_tile conv_bf16_persistent_many_tiles conv_bf16_splitk_and_7x7 conv_epilogues hwio_bf16 hwio_pair cta_pair_kernels pair_epilogues_and_wgrad promoted_accumulation decoder_heads > $O/w7_bringup.log 2>&1; echo "bringup rc=$?"
grep -E "FAIL|^case .* -> |Error|error" $O/w7_bringup.log | head -40
timeout 1500 python -m pytest tests -m gpu -q -x > $O/w7_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/w7_pytest.log
for mode in dyn static; do
  if [ $mode = static ]; then export FCN8_DEBUG="10=1"; else unset FCN8_DEBUG; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w7_bench_c2_$mode.json 2> $O/w7_bench_c2_$mode.err; echo "c2 $mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/w7_bench_c2_$mode.json').read().strip().splitlines()[-1])
print("c2 $mode fp32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "| bf16", d["alt"]["value"], d["alt"]["ms_per_step"])
PY
done
