#!/bin/bash
# Round-2 single-GPU measurement set: bench lines of every config (+ reference arm), ncu launch lists, one full capture.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > $O/final_c2.json 2> $O/final_c2.err; echo "c2 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/final_ref_c2.json 2> $O/final_ref_c2.err; echo "ref rc=$?"
for c in c3 c4 c5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > $O/final_$c.json 2> $O/final_$c.err; echo "$c rc=$?"
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for p in bf16 fp32; do
  FCN8_GRAPHS=0 timeout 900 ncu --metrics $M --clock-control none --csv --log-file $O/final_launches_${p}.csv \
    python bench.py --profile --precision $p --steps 1 --warmup 1 > $O/final_ncu_${p}.log 2>&1; echo "ncu $p rc=$?"
done
FCN8_GRAPHS=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/final_launches_c4_fp32.csv \
    python bench.py --config c4 --profile --precision fp32 --steps 1 --warmup 1 > $O/final_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
FCN8_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|wgrad_gemm_kernel" -s 40 -c 10 -o $O/final_prof_gemm \
    python bench.py --profile --precision fp32 --steps 1 --warmup 1 > $O/final_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O/final_*
python - <<'PY'
import json
for c in ("c2","c3","c4","c5"):
    try:
        d=json.loads(open('gpurun_out/final_%s.json'%c).read().strip().splitlines()[-1])
        a=d.get("alt") or {}
        print(c, d["dtype"][:12], "%.1f"%d["value"], "%.3f ms"%d["ms_per_step"], "e2e %.1f"%d["e2e"]["value"], "roof %.3f"%d["roofline"]["frac"], "| alt", a.get("value"), a.get("ms_per_step"), (a.get("roofline") or {}).get("frac"), "cpu", d.get("cpu_baseline"))
    except Exception as e: print(c, "failed", e)
print(open('gpurun_out/final_ref_c2.json').read()[:600])
PY
