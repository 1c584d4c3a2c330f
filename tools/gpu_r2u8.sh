#!/bin/bash
# round 2, call U8 (8 GPUs): config 3 strong scaling, config 4 weak scaling (4M atoms), config 5, all with `check`
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2u_gpus.txt
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731"
Q="--gpus 8 --steps 20 --warmup 5 --md-steps 0 --no-cpu-baseline"
timeout 600 $TR8 bench.py $Q > gpurun_out/r2u_cfg3_n8.json 2> gpurun_out/r2u_cfg3_n8.err
timeout 900 $TR8 bench.py --config 4 $Q > gpurun_out/r2u_cfg4_n8.json 2> gpurun_out/r2u_cfg4_n8.err
timeout 600 $TR8 bench.py --config 5 $Q > gpurun_out/r2u_cfg5_n8.json 2> gpurun_out/r2u_cfg5_n8.err
tail -c 600 gpurun_out/r2u_cfg3_n8.err; ls -la gpurun_out | grep r2u
