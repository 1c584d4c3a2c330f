#!/bin/bash
# round 2, call U4 (4 GPUs): config 3 strong scaling, config 4 weak scaling (2M atoms)
mkdir -p gpurun_out
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29732"
Q="--gpus 4 --steps 20 --warmup 5 --md-steps 0 --no-cpu-baseline"
timeout 600 $TR4 bench.py $Q > gpurun_out/r2u_cfg3_n4.json 2> gpurun_out/r2u_cfg3_n4.err
timeout 900 $TR4 bench.py --config 4 $Q > gpurun_out/r2u_cfg4_n4.json 2> gpurun_out/r2u_cfg4_n4.err
ls -la gpurun_out | grep r2u
