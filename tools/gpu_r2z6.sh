#!/bin/bash
# round 2, call Z6: launch-shape sweep of the two big Ewald kernels (config 2, 22 236 wave vectors)
mkdir -p gpurun_out
timeout 100 python tools/ewald_timing.py --tune > gpurun_out/r2z6_tune.json 2> gpurun_out/r2z6_tune.err; echo "rc=$?"
cat gpurun_out/r2z6_tune.json; tail -3 gpurun_out/r2z6_tune.err
