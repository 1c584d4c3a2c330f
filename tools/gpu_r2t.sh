#!/bin/bash
# round 2, call T: LJ end states (correction kernel against the oracle), drop-in mode, bench with config-4 golden check
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "lj_end or ljstates" > gpurun_out/r2t_lj.log 2>&1; echo "lj rc=$?" >> gpurun_out/r2t_lj.log
tail -15 gpurun_out/r2t_lj.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --md-steps 0"
timeout 600 python bench.py --config 4 $Q > gpurun_out/r2t_cfg4.json 2> gpurun_out/r2t_cfg4.err
timeout 600 python bench.py --config 5 $Q > gpurun_out/r2t_cfg5.json 2> gpurun_out/r2t_cfg5.err
timeout 300 python bench.py $Q > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
