#!/bin/bash
# round 2, call L: per-SM queues with bounded helping + sweep launch; sizes 1M / 250k / 125k
mkdir -p gpurun_out
Q="--steps 20 --warmup 5 --no-cpu-baseline --md-steps 0"
timeout 300 python bench.py $Q > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
timeout 300 python bench.py $Q --atoms 250000 --steps 40 > gpurun_out/r2l_bench_250k.json 2> gpurun_out/r2l_bench_250k.err
timeout 300 python bench.py $Q --atoms 125000 --steps 40 > gpurun_out/r2l_bench_125k.json 2> gpurun_out/r2l_bench_125k.err
for pf in 32; do
CPH_EVAL_MAXSCAN=1 timeout 300 python bench.py $Q --no-e2e --no-check > gpurun_out/r2l_bench_scan1.json 2> gpurun_out/r2l_bench_scan1.err; CPH_EVAL_MAXSCAN=16 timeout 300 python bench.py $Q --no-e2e --no-check > gpurun_out/r2l_bench_scan16.json 2> gpurun_out/r2l_bench_scan16.err
done
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
P="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 300 python bench.py $P > gpurun_out/r2l_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_eval -s 12 -c 1 -f -o gpurun_out/r2l_eval python bench.py $P > gpurun_out/r2l_ncu.log 2>&1
ls -la gpurun_out | grep r2l
