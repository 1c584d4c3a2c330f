#!/bin/bash
# round 2, call X: LJ end-state lists written by the prune itself; thread / warp correction kernels
mkdir -p gpurun_out
P="--steps 20 --warmup 5 --md-steps 0 --no-cpu-baseline"
timeout 900 python -m pytest tests -m gpu -q -k "lj_end or ljstates or inner or config4 or golden or excluded" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2x_tests.log
tail -3 gpurun_out/r2x_tests.log
timeout 300 python bench.py $P --lj-states > gpurun_out/r2x_bench_lj.json 2> gpurun_out/r2x_bench_lj.err
timeout 300 python bench.py $P > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
