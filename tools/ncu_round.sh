#!/bin/bash
# Round evidence (run under gpurun, ONE GPU): the launch list of one bench step and full captures of the dominant kernels.
#   tools/ncu_round.sh <tag>  ->  gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{xs,t1,x100}.ncu-rep
set -u
tag=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --frames 4144 --steps 1 --warmup 1 --no-cpu-baseline --precision tc_f16x3"
# every launch of the timed step with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_launches.log
# launch order per codec: encoder = stem(XS) T T XS | T T XS | down(X) | ... ; decoder = G(X) T XS | T T XS | up(X, CTA pairs) | ...
# -s skips, -c 1 captures one launch: xs #1 = 20->100 + residual @512, t #0 = 100->20 @512, x #2 = up-sampling conv
ncu --set full --clock-control none --import-source on -k regex:plane_xs_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_xs $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_xs.log
ncu --set full --clock-control none --import-source on -k regex:plane_t_kernel -s 0 -c 1 -f -o gpurun_out/${tag}_t1 $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_t1.log
ncu --set full --clock-control none --import-source on -k regex:plane_x_kernel -s 2 -c 1 -f -o gpurun_out/${tag}_x100 $BENCH > /dev/null 2> gpurun_out/${tag}_ncu_x100.log
ls -la gpurun_out/${tag}_*
