#!/bin/bash
# round 2, call Z5: per-kernel split of the device Ewald pass (ncu launch list, config 2, default kernels)
mkdir -p gpurun_out
timeout 60 python tools/ewald_timing.py --profile-run > gpurun_out/r2z5_plain.log 2>&1 && \
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ewald --csv --log-file gpurun_out/r2z5_ewald_launches.csv python tools/ewald_timing.py --profile-run > gpurun_out/r2z5_ncu.log 2>&1
echo "rc=$?"; tail -8 gpurun_out/r2z5_ewald_launches.csv | cut -c1-200
