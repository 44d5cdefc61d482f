#!/bin/bash
# Multi-GPU lines: bash scripts/r02_scale.sh N config [config ...]   (run under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
O=gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q > $O/scale_pytest_dp.log 2>&1; echo "dp pytest rc=$?"; tail -3 $O/scale_pytest_dp.log
fi
for c in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/scale_${c}_n$N.json 2> $O/scale_${c}_n$N.err
  echo "$c n$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_${c}_n$N.json').read().strip().splitlines()[-1])
    a=d.get("alt") or {}
    print("${c} N=$N", d["dtype"][:14], "value %.1f img/s  %.3f ms  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]),
          "exposed allreduce", d.get("allreduce_exposed_ms"), "| alt", a.get("dtype"), a.get("value"), a.get("ms_per_step"), (a.get("e2e") or {}).get("value"), a.get("allreduce_exposed_ms"), "dp", d.get("dp_check"))
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/scale_${c}_n$N.err').read()[-1500:])
PY
done
