#!/bin/bash
# round 2, call Z: full GPU suite on the tree with the k-space work (host feed, device Ewald), reverse_comm with ghosts;
# headline bench; ncu of the two site kernels at config 5 (the "reduction kernels" of north_star)
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --durations=25 --timeout 300 > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.log
tail -5 gpurun_out/r2z_tests.log
P="--steps 20 --warmup 5 --md-steps 0"
timeout 300 python bench.py $P > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2z_bench.json
timeout 200 python bench.py --config 5 $P --no-cpu-baseline --no-e2e --no-check > gpurun_out/r2z_cfg5_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"site_partition_kernel|lambda_update_kernel" -s 4 -c 4 -f -o gpurun_out/r2z_sites python bench.py --config 5 $P --no-cpu-baseline --no-e2e --no-check > gpurun_out/r2z_ncu.log 2>&1
echo "ncu rc=$?"
