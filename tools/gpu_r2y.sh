#!/bin/bash
# round 2, calls Y8 / Y4 / Y2: config 3 strong scaling on the final tree (usage: gpu_r2y.sh N)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2974$N"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --md-steps 0 --no-cpu-baseline > gpurun_out/r2y_cfg3_n$N.json 2> gpurun_out/r2y_cfg3_n$N.err
tail -c 300 gpurun_out/r2y_cfg3_n$N.err
