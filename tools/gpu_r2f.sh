#!/bin/bash
# round 2, call F: what bounds the evaluation kernel -- memory side alone, arithmetic alone
mkdir -p gpurun_out
V=constant_ph_b200/csrc/variants
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
for v in nomath nogather d2r64; do
  CPH_B200_LIB=$PWD/$V/libcph_b200_$v.so timeout 300 python bench.py $Q > gpurun_out/r2f_bench_$v.json 2> gpurun_out/r2f_bench_$v.err
done
timeout 300 python bench.py $Q > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 python bench.py --config 5 --steps 20 --warmup 5 --no-cpu-baseline --md-steps 0 > gpurun_out/r2f_cfg5.json 2> gpurun_out/r2f_cfg5.err
timeout 600 python bench.py --config 2 --sweep --steps 200 --warmup 10 > gpurun_out/r2f_cfg2_sweep.json 2> gpurun_out/r2f_cfg2_sweep.err
ls -la gpurun_out | grep r2f
