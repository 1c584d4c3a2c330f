#!/bin/bash
# round 2, call Z7: k-space tests and timing with the launch shapes chosen from the sweep
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_kspace.py -m gpu -q --timeout 90 > gpurun_out/r2z7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z7_tests.log
tail -2 gpurun_out/r2z7_tests.log
timeout 60 python tools/ewald_timing.py > gpurun_out/r2z7_ewald_timing.json 2> gpurun_out/r2z7_ewald_timing.err; echo "timing rc=$?"
cut -c1-2400 gpurun_out/r2z7_ewald_timing.json
