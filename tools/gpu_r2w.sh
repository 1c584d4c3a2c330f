#!/bin/bash
# round 2, call W: leaner list build + prune kernels: full parity suite, bench, optional-term cost, fix e2e, smoke, ncu of the prune
mkdir -p gpurun_out
P="--steps 20 --warmup 5 --md-steps 0"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2w_tests.log
tail -3 gpurun_out/r2w_tests.log
timeout 600 python bench.py $P > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
timeout 300 python bench.py $P --no-cpu-baseline --lj-states > gpurun_out/r2w_bench_lj.json 2> gpurun_out/r2w_bench_lj.err
timeout 300 python bench.py $P --no-cpu-baseline --atoms 125000 --steps 40 --no-e2e > gpurun_out/r2w_bench_125k.json 2> gpurun_out/r2w_bench_125k.err
timeout 600 python tools/harness_e2e.py > gpurun_out/r2w_harness_e2e.json 2> gpurun_out/r2w_harness_e2e.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2w_smoke.log 2>&1
timeout 300 python bench.py $P --no-cpu-baseline --no-e2e --no-check > gpurun_out/r2w_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune_kernel -s 2 -c 1 -f -o gpurun_out/r2w_prune python bench.py $P --no-cpu-baseline --no-e2e --no-check > gpurun_out/r2w_ncu.log 2>&1
tail -2 gpurun_out/r2w_smoke.log
