#!/bin/bash
# round 2, call Z2 (2 GPUs): the 2-rank parity tests on the final tree, with the device Ewald sum among them
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_multi_gpu.py tests/test_fix_dropin.py -m gpu -q -k "2- or two_ranks" --durations=10 --timeout 300 > gpurun_out/r2z2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2z2_tests.log
tail -15 gpurun_out/r2z2_tests.log
