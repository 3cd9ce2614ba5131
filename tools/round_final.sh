#!/bin/bash
# Round-end evidence on ONE GPU: ncu launch list + full captures (tools/ncu_round.sh), the headline bench line with its CPU baseline,
# the reference arm, and the training-step line.  Outputs under gpurun_out/<tag>_*.
set -u
tag=${1:-r01v}
mkdir -p gpurun_out
tools/ncu_round.sh $tag
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err
timeout 300 python bench.py --precision tc_f16 --no-cpu-baseline > gpurun_out/${tag}_bench_tc_f16.json 2> /dev/null
for f in bench bench_reference bench_train bench_tc_f16; do echo "== $f"; head -c 700 gpurun_out/${tag}_$f.json; echo; done
