#!/bin/bash
# round 2, call A: parity suite, bench (product build + tuning variants), ncu of the evaluation kernel
mkdir -p gpurun_out
V=constant_ph_b200/csrc/variants
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for v in r72 ref3 r72ref3; do
  CPH_B200_LIB=$PWD/$V/libcph_b200_$v.so timeout 300 python bench.py $Q > gpurun_out/r2a_bench_$v.json 2> gpurun_out/r2a_bench_$v.err
done
for c in 8 12 24; do
  CPH_EVAL_CTAS_PER_SM=$c timeout 300 python bench.py $Q > gpurun_out/r2a_bench_ctas$c.json 2> gpurun_out/r2a_bench_ctas$c.err
done
for a in 4 16; do
  CPH_EAPW=$a timeout 300 python bench.py $Q > gpurun_out/r2a_bench_eapw$a.json 2> gpurun_out/r2a_bench_eapw$a.err
done
for s in 0.2 0.3; do
  CPH_INNER_SKIN=$s timeout 300 python bench.py $Q > gpurun_out/r2a_bench_skin$s.json 2> gpurun_out/r2a_bench_skin$s.err
done
P="--steps 6 --warmup 3 --no-cpu-baseline --no-e2e --md-steps 0 --no-check"
timeout 300 python bench.py $P > gpurun_out/r2a_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_eval -s 12 -c 1 -f -o gpurun_out/r2a_eval python bench.py $P > gpurun_out/r2a_ncu.log 2>&1
timeout 300 python bench.py $P > gpurun_out/r2a_plain2.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py $P > gpurun_out/r2a_ncu2.log 2>&1
ls -la gpurun_out | tail -30
