#!/bin/bash
# Round-2 wave 4: c4 (2048x1024 inference) launch lists + full ncu captures of the kernels under investigation.
mkdir -p gpurun_out
O=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for p in fp32 bf16; do
  FCN8_GRAPHS=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/launches_c4_${p}.csv \
    python bench.py --config c4 --profile --precision $p --steps 1 --warmup 1 > $O/ncu_c4_${p}.log 2>&1; echo "ncu c4 $p rc=$?"
done
FCN8_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel<256, false, true, 1" -c 2 -o $O/prof_c4_loss \
  python bench.py --config c4 --profile --precision fp32 --steps 1 --warmup 1 > $O/ncu_full_c4.log 2>&1; echo "full c4 rc=$?"
FCN8_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo_kernel|conv1_fwd_kernel|conv1_wgrad_kernel|maxpool_bwd" -c 12 -o $O/prof_c2_small \
  python bench.py --profile --precision fp32 --steps 1 --warmup 0 > $O/ncu_full_c2_small.log 2>&1; echo "full c2 rc=$?"
ls -la $O/*.ncu-rep
