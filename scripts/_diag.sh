mkdir -p gpurun_out
python scripts/debug_fullsize.py positive_weights 0 > gpurun_out/dbg_positive.log 2>&1; cat gpurun_out/dbg_positive.log
python scripts/debug_fullsize.py he_normal 0 > gpurun_out/dbg_he.log 2>&1; cat gpurun_out/dbg_he.log
