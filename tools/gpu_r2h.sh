#!/bin/bash
# round 2, call H: two pairs per lane (ILP 2) at several register budgets
mkdir -p gpurun_out
V=constant_ph_b200/csrc/variants
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0"
CPH_EVAL_CTAS_PER_SM=12 CPH_B200_LIB=$PWD/$V/libcph_b200_ilp2r80.so timeout 300 python bench.py $Q > gpurun_out/r2h_bench_ilp2r80.json 2> gpurun_out/r2h_bench_ilp2r80.err
CPH_EVAL_CTAS_PER_SM=10 CPH_B200_LIB=$PWD/$V/libcph_b200_ilp2r96.so timeout 300 python bench.py $Q > gpurun_out/r2h_bench_ilp2r96.json 2> gpurun_out/r2h_bench_ilp2r96.err
CPH_EVAL_CTAS_PER_SM=9 CPH_B200_LIB=$PWD/$V/libcph_b200_ilp2r112.so timeout 300 python bench.py $Q > gpurun_out/r2h_bench_ilp2r112.json 2> gpurun_out/r2h_bench_ilp2r112.err
CPH_EVAL_CTAS_PER_SM=8 CPH_B200_LIB=$PWD/$V/libcph_b200_ilp2r128.so timeout 300 python bench.py $Q > gpurun_out/r2h_bench_ilp2r128.json 2> gpurun_out/r2h_bench_ilp2r128.err
P="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
export CPH_EVAL_CTAS_PER_SM=10 CPH_B200_LIB=$PWD/$V/libcph_b200_ilp2r96.so
timeout 300 python bench.py $P > gpurun_out/r2h_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_eval -s 12 -c 1 -f -o gpurun_out/r2h_eval_ilp2r96 python bench.py $P > gpurun_out/r2h_ncu.log 2>&1
ls -la gpurun_out | grep r2h
