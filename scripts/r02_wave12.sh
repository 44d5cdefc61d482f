#!/bin/bash
# Wave 12: label packing in the feed + compact class map for predict(): GPU tests, e2e with / without packing (c3, N = 1),
# c4 line, ncu --set full of the max-pool backward and the loss epilogue kernel.
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/w12_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 $O/w12_pytest.log
for pk in 1 0; do
  FCN8_FEED_PACK_LABELS=$pk timeout 600 python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline > $O/w12_c3_pack$pk.json 2> $O/w12_c3_pack$pk.err; echo "c3 pack=$pk rc=$?"
done
timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline > $O/w12_c4.json 2> $O/w12_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ("w12_c3_pack1","w12_c3_pack0","w12_c4"):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, "%.1f"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e", d["e2e"])
    except Exception as e:
        print(f, "failed", e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
FCN8_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"maxpool_bwd_kernel|conv_gemm_kernel<256, false, true, 1" -c 8 -o $O/w12_prof_pool_loss \
    python bench.py --profile --precision fp32 --steps 1 --warmup 1 > $O/w12_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O/w12_*
