#!/bin/bash
# round 2, call O: register cap / resident CTAs of the evaluation kernel
mkdir -p gpurun_out
V=constant_ph_b200/csrc/variants
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
CPH_EVAL_CTAS_PER_SM=18 CPH_B200_LIB=$PWD/$V/libcph_b200_r56.so timeout 300 python bench.py $Q > gpurun_out/r2o_bench_r56c18.json 2> gpurun_out/r2o_bench_r56c18.err
CPH_EVAL_CTAS_PER_SM=21 CPH_B200_LIB=$PWD/$V/libcph_b200_r48.so timeout 300 python bench.py $Q > gpurun_out/r2o_bench_r48c21.json 2> gpurun_out/r2o_bench_r48c21.err
CPH_EVAL_CTAS_PER_SM=20 CPH_B200_LIB=$PWD/$V/libcph_b200_r48.so timeout 300 python bench.py $Q > gpurun_out/r2o_bench_r48c20.json 2> gpurun_out/r2o_bench_r48c20.err
CPH_EVAL_CTAS_PER_SM=18 CPH_B200_LIB=$PWD/$V/libcph_b200_r56.so timeout 300 python bench.py $Q --atoms 125000 --steps 40 > gpurun_out/r2o_bench_r56c18_125k.json 2> gpurun_out/r2o_bench_r56c18_125k.err
CPH_EVAL_CTAS_PER_SM=21 CPH_B200_LIB=$PWD/$V/libcph_b200_r48.so timeout 300 python bench.py $Q --atoms 125000 --steps 40 > gpurun_out/r2o_bench_r48c21_125k.json 2> gpurun_out/r2o_bench_r48c21_125k.err
for s in 0.5 0.6; do CPH_INNER_SKIN=$s timeout 300 python bench.py $Q > gpurun_out/r2o_bench_skin$s.json 2> gpurun_out/r2o_bench_skin$s.err; done
