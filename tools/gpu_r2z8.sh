#!/bin/bash
# round 2, call Z8 (last): k-space tests with the size switch between the two per-atom kernels
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_kspace.py -m gpu -q --timeout 90 > gpurun_out/r2z8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z8_tests.log
tail -2 gpurun_out/r2z8_tests.log
