#!/bin/bash
# NCCL experiment at N GPUs (c3 = bf16, 2 img/GPU): registered buffers / CTA policy / CTA cap.
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
    print("${tag} N=$N value %.1f img/s  %.3f ms  e2e %.1f  exposed %.3f ms  without %.3f  registered %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], a.get("exposed_ms_per_step",-1), a.get("ms_per_step_without_allreduce",-1), a.get("nccl_registered_buffer")))
except Exception as e:
    print("${tag} parse failed", e); print(open('gpurun_out/nccl_${tag}_n$N.err').read()[-1200:])
PY
}
run base FCN8_NCCL_REGISTER=0
run reg FCN8_NCCL_REGISTER=1 FCN8_DEBUG_DP=1
run reg_eff FCN8_NCCL_REGISTER=1 NCCL_CTA_POLICY=1
run reg_cta8 FCN8_NCCL_REGISTER=1 NCCL_MAX_CTAS=8
run base_cta8 FCN8_NCCL_REGISTER=0 NCCL_MAX_CTAS=8
grep -h "registration unavailable" $O/nccl_reg_n$N.err | head -2
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL,TUNING,REG timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --config c3 --steps 3 --warmup 3 --no-cpu-baseline --profile > /dev/null 2> $O/nccl_info_n$N.log
grep -E "NVLS|Algo|algorithm|AllReduce.*(RING|TREE|NVLS)|registered|Reg" $O/nccl_info_n$N.log | sort | uniq -c | sort -rn | head -30 > $O/nccl_info_n${N}_summary.txt; head -30 $O/nccl_info_n${N}_summary.txt | cut -c1-220
rm -f $O/nccl_info_n$N.log
