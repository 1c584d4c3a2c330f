#!/bin/bash
# round 2, call U2 (2 GPUs): 2-rank tests (pair, bonded, LJ end states, the fix), config 3 and config 4 at 2 ranks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "multi_rank or two_ranks" > gpurun_out/r2u_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_tests2.log
tail -4 gpurun_out/r2u_tests2.log
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733"
Q="--gpus 2 --steps 20 --warmup 5 --md-steps 0 --no-cpu-baseline"
timeout 600 $TR2 bench.py $Q > gpurun_out/r2u_cfg3_n2.json 2> gpurun_out/r2u_cfg3_n2.err
timeout 900 $TR2 bench.py --config 4 $Q > gpurun_out/r2u_cfg4_n2.json 2> gpurun_out/r2u_cfg4_n2.err
ls -la gpurun_out | grep r2u
