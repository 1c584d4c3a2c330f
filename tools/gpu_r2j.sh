#!/bin/bash
# round 2, call J: parity suite + bench after the special-partner / prologue changes
mkdir -p gpurun_out
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0"
timeout 300 python bench.py $Q > gpurun_out/r2j_bench_quick.json 2> gpurun_out/r2j_bench_quick.err
for c in 1 4; do :; done
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2j_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --md-steps 0 > gpurun_out/r2j_bench200.json 2> gpurun_out/r2j_bench200.err
ls -la gpurun_out | grep r2j
