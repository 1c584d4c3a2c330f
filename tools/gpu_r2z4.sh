#!/bin/bash
# round 2, call Z4 (last): row-walking per-atom Ewald kernel -- full GPU suite on the final tree, the k-space tests again
# with the table kernels as default, then the k-space timing of all three variants
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --durations=5 --timeout 200 > gpurun_out/r2z4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z4_tests.log
tail -3 gpurun_out/r2z4_tests.log
CPH_EWALD=tables timeout 120 python -m pytest tests/test_kspace.py -m gpu -q --timeout 100 > gpurun_out/r2z4_tests_tables.log 2>&1; echo "tables rc=$?" >> gpurun_out/r2z4_tests_tables.log
tail -2 gpurun_out/r2z4_tests_tables.log
timeout 100 python tools/ewald_timing.py > gpurun_out/r2z4_ewald_timing.json 2> gpurun_out/r2z4_ewald_timing.err; echo "timing rc=$?"
cut -c1-2200 gpurun_out/r2z4_ewald_timing.json
