#!/bin/bash
# round 2, call V: list build with the rows' runs computed by 25 lanes at once and a predicate-only candidate test
mkdir -p gpurun_out
P="--steps 20 --warmup 5 --no-cpu-baseline --md-steps 0"
timeout 900 python -m pytest tests -m gpu -q -x -k "neighbor or config4 or config5 or empty or special or golden or excluded or lj_end or inner" > gpurun_out/r2v_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2v_tests.log
tail -3 gpurun_out/r2v_tests.log
timeout 300 python bench.py $P > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
timeout 300 python bench.py $P --steps 100 --no-e2e > gpurun_out/r2v_bench100.json 2> gpurun_out/r2v_bench100.err
timeout 300 python bench.py $P --no-e2e --no-check > gpurun_out/r2v_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:list_build -s 1 -c 1 -f -o gpurun_out/r2v_list python bench.py $P --no-e2e --no-check > gpurun_out/r2v_ncu.log 2>&1
