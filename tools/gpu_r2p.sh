#!/bin/bash
# round 2, call P: parity suite + bench after the lane-parallel lambda update
mkdir -p gpurun_out
Q="--steps 20 --warmup 5 --no-cpu-baseline --md-steps 0"
timeout 300 python bench.py $Q > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
timeout 300 python bench.py $Q --atoms 125000 --steps 40 --no-e2e > gpurun_out/r2p_bench_125k.json 2> gpurun_out/r2p_bench_125k.err
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log
timeout 600 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu-baseline --md-steps 0 > gpurun_out/r2p_cfg5.json 2> gpurun_out/r2p_cfg5.err
