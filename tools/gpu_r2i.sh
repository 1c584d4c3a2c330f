#!/bin/bash
# round 2, call I: special partners in the tail chunk of the inner rows; parity suite, bench, ncu
mkdir -p gpurun_out
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0"
timeout 300 python bench.py $Q > gpurun_out/r2i_bench_quick.json 2> gpurun_out/r2i_bench_quick.err
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
P="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 300 python bench.py $P > gpurun_out/r2i_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_eval -s 12 -c 1 -f -o gpurun_out/r2i_eval python bench.py $P > gpurun_out/r2i_ncu.log 2>&1
ls -la gpurun_out | grep r2i
