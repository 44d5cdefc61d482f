#!/bin/bash
# Round-2 wave 6: coalesced (shared-memory staged) epilogue stores.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=200 timeout 1200 python scripts/bringup.py conv_epilogues hwio_bf16 hwio_pair cta_pair_kernels halo_conv pair_epilogues_and_wgrad fused_pool decoder_heads deconv_stride2 deconv_stride8_loss conv1_direct promoted_accumulation conv_bf16_3x3_bn128_bn256 conv_bf16_persistent_many_tiles > $O/w6_bringup.log 2>&1; echo "bringup rc=$?"
grep -E "FAIL|^case .* -> " $O/w6_bringup.log | head -40
timeout 1500 python -m pytest tests -m gpu -q > $O/w6_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/w6_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/w6_bench_c2.json 2> $O/w6_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w6_bench_c2.json').read().strip().splitlines()[-1])
print("c2 fp32", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "| bf16", d["alt"]["value"], d["alt"]["ms_per_step"], "e2e", d["alt"]["e2e"]["value"])
PY
timeout 300 python scripts/wait_profile.py fp32 bf16 > $O/wait_profile_w6.txt 2>&1; grep -E "^ +[0-3] " $O/wait_profile_w6.txt | cut -c1-260
