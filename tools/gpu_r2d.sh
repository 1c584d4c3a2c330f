#!/bin/bash
# round 2, call D: evaluation-kernel variants (claim size, prefetch distance), parity suite, ncu
mkdir -p gpurun_out
V=constant_ph_b200/csrc/variants
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
for v in claim1 claim2 claim8; do
  CPH_B200_LIB=$PWD/$V/libcph_b200_$v.so timeout 300 python bench.py $Q > gpurun_out/r2d_bench_$v.json 2> gpurun_out/r2d_bench_$v.err
done
CPH_EVAL_CTAS_PER_SM=14 CPH_B200_LIB=$PWD/$V/libcph_b200_r72.so timeout 300 python bench.py $Q > gpurun_out/r2d_bench_r72c14.json 2> gpurun_out/r2d_bench_r72c14.err
for pf in 16 32; do
  CPH_EVAL_PREFETCH=$pf timeout 300 python bench.py $Q > gpurun_out/r2d_bench_pf$pf.json 2> gpurun_out/r2d_bench_pf$pf.err
done
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
P="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 300 python bench.py $P > gpurun_out/r2d_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_eval -s 12 -c 1 -f -o gpurun_out/r2d_eval python bench.py $P > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out | grep r2d
