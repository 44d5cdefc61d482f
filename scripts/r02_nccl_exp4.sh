#!/bin/bash
# NCCL experiment 4 at N GPUs: side-stream wire cast + three chunks; dynamic vs static scheduling.
N=$1
mkdir -p gpurun_out
O=gpurun_out
run() {  # tag, config, env...
  tag=$1; cfg=$2; shift; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-alt > $O/nccl_${tag}_n$N.json 2> $O/nccl_${tag}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/nccl_${tag}_n$N.json').read().strip().splitlines()[-1])
    a=d.get("allreduce",{})
    print("${tag} N=$N value %.1f img/s  %.3f ms  e2e %.1f  exposed %.3f ms  without %.3f  dp %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], a.get("exposed_ms_per_step",-1), a.get("ms_per_step_without_allreduce",-1), d.get("dp_check",{}).get("grad_rel_l2_vs_single_gpu")))
except Exception as e:
    print("${tag} parse failed", e); print(open('gpurun_out/nccl_${tag}_n$N.err').read()[-1500:])
PY
}
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q 2>&1 | tail -2; fi
run c3_dyn c3 FCN8_X=1
run c3_static c3 FCN8_DEBUG=10=1
run c2_dyn c2 FCN8_X=1
run c2_static c2 FCN8_DEBUG=10=1
