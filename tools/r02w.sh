#!/bin/bash
set -u
mkdir -p gpurun_out
BENCH="python bench.py --frames 4144 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --profile-one-step gpurun_out/r02w_step_records.json"
FULL="--set full --clock-control none --import-source on --profile-from-start off -f"
ncu $FULL -k regex:plane_xs_kernel -s 0 -c 1 -o gpurun_out/r02w_xs_stem $BENCH > /dev/null 2> gpurun_out/r02w_ncu_stem.log
ncu $FULL -k regex:lpc_analyze -s 0 -c 1 -o gpurun_out/r02w_lpc_analyze $BENCH > /dev/null 2> gpurun_out/r02w_ncu_lpc.log
ls -la gpurun_out/r02w_*
