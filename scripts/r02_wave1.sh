#!/bin/bash
# Round-2 wave 1 on one B200: GPU test tier, the four config bench lines, the backward-terms measurement.
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/w1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/w1_pytest.log
tail -5 $O/w1_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/w1_bench_c2.json 2> $O/w1_bench_c2.err; echo "c2 rc=$?"
timeout 400 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > $O/w1_bench_c3.json 2> $O/w1_bench_c3.err; echo "c3 rc=$?"
timeout 400 python bench.py --config c4 --steps 10 --warmup 3 > $O/w1_bench_c4.json 2> $O/w1_bench_c4.err; echo "c4 rc=$?"
timeout 400 python bench.py --config c5 --steps 10 --warmup 3 > $O/w1_bench_c5.json 2> $O/w1_bench_c5.err; echo "c5 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/w1_ref_c2.json 2> $O/w1_ref_c2.err; echo "ref rc=$?"
timeout 600 python scripts/backward_terms.py $O/w1_backward_terms.md > $O/w1_backward_terms.log 2>&1; echo "bt rc=$?"
for f in $O/w1_bench_c*.json; do echo "== $f"; cut -c1-400 $f; done
tail -3 $O/w1_bench_c*.err
cat $O/w1_backward_terms.md
