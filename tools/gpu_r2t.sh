#!/bin/bash
# round 2, call T: LJ end states (correction kernel against the oracle), drop-in mode, bench with config-4 golden check
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "lj_end or ljstates" > gpurun_out/r2t_lj.log 2>&1; echo "lj rc=$?" >> gpurun_out/r2t_lj.log
tail -15 gpurun_out/r2t_lj.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --md-steps 0"
