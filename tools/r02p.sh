#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:plane_x_kernel -s 6 -c 3 -f -o gpurun_out/r02p_fold python tools/fold_probe.py 2072 > gpurun_out/r02p_probe.log 2>&1
tail -5 gpurun_out/r02p_probe.log
ls -la gpurun_out/r02p_fold.ncu-rep
