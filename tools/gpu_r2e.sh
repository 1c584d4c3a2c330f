#!/bin/bash
# round 2, call E: record pipeline depth, L1 carve-out
mkdir -p gpurun_out
V=constant_ph_b200/csrc/variants
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 300 python bench.py $Q > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
for cv in 0 25 50; do
  CPH_EVAL_CARVEOUT=$cv timeout 300 python bench.py $Q > gpurun_out/r2e_bench_carve$cv.json 2> gpurun_out/r2e_bench_carve$cv.err
done
CPH_B200_LIB=$PWD/$V/libcph_b200_d2r64.so timeout 300 python bench.py $Q > gpurun_out/r2e_bench_d2r64.json 2> gpurun_out/r2e_bench_d2r64.err
CPH_EVAL_CTAS_PER_SM=14 CPH_B200_LIB=$PWD/$V/libcph_b200_d2r72.so timeout 300 python bench.py $Q > gpurun_out/r2e_bench_d2r72.json 2> gpurun_out/r2e_bench_d2r72.err
CPH_EVAL_CTAS_PER_SM=12 CPH_B200_LIB=$PWD/$V/libcph_b200_d2r80.so timeout 300 python bench.py $Q > gpurun_out/r2e_bench_d2r80.json 2> gpurun_out/r2e_bench_d2r80.err
CPH_EVAL_CTAS_PER_SM=10 CPH_B200_LIB=$PWD/$V/libcph_b200_d2r96.so timeout 300 python bench.py $Q > gpurun_out/r2e_bench_d2r96.json 2> gpurun_out/r2e_bench_d2r96.err
CPH_EVAL_CARVEOUT=0 CPH_EVAL_CTAS_PER_SM=14 CPH_B200_LIB=$PWD/$V/libcph_b200_d2r72.so timeout 300 python bench.py $Q > gpurun_out/r2e_bench_d2r72_carve0.json 2> gpurun_out/r2e_bench_d2r72_carve0.err
P="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
export CPH_EVAL_CTAS_PER_SM=14 CPH_B200_LIB=$PWD/$V/libcph_b200_d2r72.so
timeout 300 python bench.py $P > gpurun_out/r2e_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_eval -s 12 -c 1 -f -o gpurun_out/r2e_eval_d2r72 python bench.py $P > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out | grep r2e
