#!/bin/bash
# Wave 11: background Adam (fcn8_adam_chunks) bring-up, side-stream overlap experiments, full GPU test tier.
mkdir -p gpurun_out
O=gpurun_out
BRINGUP_TIMEOUT=120 timeout 300 python scripts/bringup.py adam_chunks elementwise_kernels 2>&1 | tee $O/w11_bringup.log | grep -E "FAIL|adam_chunks|^case .* -> "
timeout 600 python scripts/overlap_exp.py 10 > $O/w11_overlap.jsonl 2> $O/w11_overlap.err; echo "overlap rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/w11_overlap.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print("%-5s %d %-14s %.3f ms  loss %.5f  eq %s mism %s biasrel %s sms %s" % (d["precision"], d["per_gpu"], d["variant"], d["ms_per_step"], d["loss_after_5"], d.get("step1_nonbias_bit_equal"), d.get("step1_mismatching_elements"), d.get("step1_bias_m_max_rel"), d.get("bg_chunks_claimed_sms")))
PY
tail -5 $O/w11_overlap.err
timeout 1200 python -m pytest tests -m gpu -q -x > $O/w11_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/w11_pytest.log
