#!/bin/bash
# NCCL experiment 2 at N GPUs (c3): SMs reserved for the collective's CTAs while it runs under the conv backward.
N=$1
mkdir -p gpurun_out
O=gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --config c3 --steps 20 --warmup 5 --no-cpu-baseline > $O/nccl_${tag}_n$N.json 2> $O/nccl_${tag}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/nccl_${tag}_n$N.json').read().strip().splitlines()[-1])
    a=d.get("allreduce",{})
    print("${tag} N=$N value %.1f img/s  %.3f ms  e2e %.1f  exposed %.3f ms  without %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], a.get("exposed_ms_per_step",-1), a.get("ms_per_step_without_allreduce",-1)))
except Exception as e:
    print("${tag} parse failed", e); print(open('gpurun_out/nccl_${tag}_n$N.err').read()[-1200:])
PY
}
run r16 FCN8_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16
run r16w4 FCN8_DP_RESERVE_SMS=16 NCCL_MAX_CTAS=16 FCN8_DP_RESERVE_LAYERS=4
run r24w4 FCN8_DP_RESERVE_SMS=24 NCCL_MAX_CTAS=24 FCN8_DP_RESERVE_LAYERS=4
run r32w4 FCN8_DP_RESERVE_SMS=32 NCCL_MAX_CTAS=32 FCN8_DP_RESERVE_LAYERS=4
run r32w6free FCN8_DP_RESERVE_SMS=32 FCN8_DP_RESERVE_LAYERS=6
run noov FCN8_DP_OVERLAP=0
