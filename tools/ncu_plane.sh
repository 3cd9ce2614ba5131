#!/bin/bash
# ncu captures of the three plane-engine kernels that dominate the cq2 step (run under gpurun, one GPU).
#   tools/ncu_plane.sh <tag>    -> gpurun_out/<tag>_{t20,x20to100,xup}.ncu-rep
set -u
tag=${1:-r01}
CMD="python bench.py --frames 2072 --steps 1 --warmup 1 --no-cpu-baseline --precision tc_f16x3"
mkdir -p gpurun_out
# launch order per codec: encoder = stem(G) T T X | T T X | down(X) | T T X | T T X | head(T); decoder = G T X | T T X | up(X) | ...
ncu --set full --clock-control none --import-source on -k regex:plane_t_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_t20 $CMD > /dev/null 2> gpurun_out/${tag}_ncu_t20.log
ncu --set full --clock-control none --import-source on -k regex:plane_x_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_x20to100 $CMD > /dev/null 2> gpurun_out/${tag}_ncu_x20.log
ncu --set full --clock-control none --import-source on -k regex:plane_x_kernel -s 9 -c 1 -f -o gpurun_out/${tag}_xup $CMD > /dev/null 2> gpurun_out/${tag}_ncu_xup.log
ls -la gpurun_out/${tag}_*.ncu-rep
