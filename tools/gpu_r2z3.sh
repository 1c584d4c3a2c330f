#!/bin/bash
# round 2, call Z3: factorised Ewald kernels -- full GPU suite again on the final tree, then the k-space timing
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --durations=8 --timeout 200 > gpurun_out/r2z3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z3_tests.log
tail -4 gpurun_out/r2z3_tests.log
timeout 120 python tools/ewald_timing.py > gpurun_out/r2z3_ewald_timing.json 2> gpurun_out/r2z3_ewald_timing.err; echo "timing rc=$?"
cat gpurun_out/r2z3_ewald_timing.json | cut -c1-1800
