#!/bin/bash
# round 2, call M (8 GPUs): strong scaling at 8 and 4 ranks, mailboxes vs NCCL for the small all-reduces
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m_gpus.txt
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29722"
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --md-steps 0 > gpurun_out/r2m_bench8.json 2> gpurun_out/r2m_bench8.err
CPH_MAIL=0 timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 --md-steps 0 --no-check > gpurun_out/r2m_bench8_nomail.json 2> gpurun_out/r2m_bench8_nomail.err
timeout 600 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 --md-steps 0 > gpurun_out/r2m_bench4.json 2> gpurun_out/r2m_bench4.err
timeout 600 $TR8 bench.py --gpus 8 --steps 100 --warmup 5 --md-steps 0 --no-check > gpurun_out/r2m_bench8_100.json 2> gpurun_out/r2m_bench8_100.err
ls -la gpurun_out | grep r2m
