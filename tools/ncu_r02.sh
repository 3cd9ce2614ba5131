#!/bin/bash
# Round-2 evidence (run under gpurun, ONE GPU):
#   1. launch list of ONE device-resident step (4,144 frames = two engine passes) with time and DRAM bytes of every launch
#   2. `ncu --set full` captures of the dominant kernels (one launch each)
# -> gpurun_out/r02_*; summarise here with tools/ncu_facts.py
set -u
tag=${1:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --frames 4144 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --profile-one-step gpurun_out/${tag}_step_records.json"
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_launches.log
# launch order per codec: encoder = stem(XS) | T F XS | T F XS | down(X) | ... ; decoder = G(X) F XS | T F XS | up(X, CTA pairs) | ...   (F = folded 20->20 on plane_x_kernel)
FULL="--set full --clock-control none --import-source on --profile-from-start off -f"
ncu $FULL -k regex:plane_xs_kernel -s 1 -c 1 -o gpurun_out/${tag}_xs_20to100 $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_xs.log
ncu $FULL -k regex:plane_t_kernel -s 0 -c 2 -o gpurun_out/${tag}_t_conv1_conv2 $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_t.log
# plane_x_kernel launches of the encoder: folded 20->20 (d1 @512), folded 20->20 (d2 @512, two parity tiles per frame), stride-2 conv
ncu $FULL -k regex:plane_x_kernel -s 0 -c 1 -o gpurun_out/${tag}_x_fold $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_xf.log
ncu $FULL -k regex:plane_x_kernel -s 2 -c 1 -o gpurun_out/${tag}_x_down $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_x.log
ncu $FULL -k regex:lpc_analyze -s 0 -c 1 -o gpurun_out/${tag}_lpc_analyze $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_lpc.log
ls -la gpurun_out/${tag}_*
