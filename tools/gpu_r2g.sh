#!/bin/bash
# round 2, call G (2 GPUs): multi-rank parity (library + drop-in fix), mailboxes vs NCCL for the small all-reduces
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py tests/test_fix_dropin.py -m gpu -q -k "multi_rank or two_ranks" > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --md-steps 0 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err
CPH_MAIL=0 timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --md-steps 0 --no-check > gpurun_out/r2g_bench2_nomail.json 2> gpurun_out/r2g_bench2_nomail.err
CPH_HALO=nccl timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --md-steps 0 --no-check > gpurun_out/r2g_bench2_nccl.json 2> gpurun_out/r2g_bench2_nccl.err
# the per-rank share of an 8-rank run on 2 GPUs: 250k atoms (2 x 125k)
timeout 900 $TR bench.py --gpus 2 --atoms 250000 --steps 40 --warmup 5 --md-steps 0 --no-check > gpurun_out/r2g_bench2_250k.json 2> gpurun_out/r2g_bench2_250k.err
CPH_MAIL=0 timeout 900 $TR bench.py --gpus 2 --atoms 250000 --steps 40 --warmup 5 --md-steps 0 --no-check > gpurun_out/r2g_bench2_250k_nomail.json 2> gpurun_out/r2g_bench2_250k_nomail.err
tail -3 gpurun_out/r2g_tests.log
ls -la gpurun_out | grep r2g
