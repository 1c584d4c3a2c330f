#!/bin/bash
# round 2, call S: bonded-chain config 4 through the CUDA path, parity suite, bench on the cleaned source
mkdir -p gpurun_out
Q="--steps 20 --warmup 5 --no-cpu-baseline --md-steps 0"
timeout 300 python bench.py $Q > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
timeout 600 python bench.py --config 4 $Q > gpurun_out/r2s_cfg4.json 2> gpurun_out/r2s_cfg4.err
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2s_tests.log
tail -3 gpurun_out/r2s_tests.log
