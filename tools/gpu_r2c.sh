#!/bin/bash
# round 2, call C: full parity suite, bench, prefetch variants, e2e through the fix, ncu of the evaluation kernel
mkdir -p gpurun_out
V=constant_ph_b200/csrc/variants
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
for pf in 0 32 128; do
  CPH_EVAL_PREFETCH=$pf timeout 300 python bench.py $Q > gpurun_out/r2c_bench_pf$pf.json 2> gpurun_out/r2c_bench_pf$pf.err
done
CPH_EVAL_CTAS_PER_SM=14 CPH_B200_LIB=$PWD/$V/libcph_b200_r72.so timeout 300 python bench.py $Q > gpurun_out/r2c_bench_r72c14.json 2> gpurun_out/r2c_bench_r72c14.err
CPH_EVAL_CTAS_PER_SM=12 CPH_B200_LIB=$PWD/$V/libcph_b200_r80.so timeout 300 python bench.py $Q > gpurun_out/r2c_bench_r80c12.json 2> gpurun_out/r2c_bench_r80c12.err
for s in 0.3 0.5; do
  CPH_INNER_SKIN=$s timeout 300 python bench.py $Q > gpurun_out/r2c_bench_skin$s.json 2> gpurun_out/r2c_bench_skin$s.err
done
timeout 600 python tools/harness_e2e.py > gpurun_out/r2c_harness_e2e.json 2> gpurun_out/r2c_harness_e2e.err
P="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 300 python bench.py $P > gpurun_out/r2c_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_eval -s 12 -c 1 -f -o gpurun_out/r2c_eval python bench.py $P > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out | grep r2c
