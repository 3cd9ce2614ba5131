#!/bin/bash
# Round-2 evidence on ONE GPU: the GPU test suite, ncu launch list + full captures (tools/ncu_r02.sh), the headline bench line with its
# CPU baseline and sub-records, the reference arm, the training-step line, the reduced-precision line and smoke().
# Outputs: gpurun_out/<tag>_*; summarise here with `python tools/ncu_facts.py <tag>` and copy the bench lines to profiles/.
set -u
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_pytest_gpu.log)"
bash tools/ncu_r02.sh $tag > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err
timeout 300 python bench.py --precision tc_f16 --no-cpu-baseline --no-sub-records > gpurun_out/${tag}_bench_tc_f16.json 2> /dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1
tail -1 gpurun_out/${tag}_smoke.log
for f in bench_default bench_reference bench_train bench_tc_f16; do echo "== $f"; head -c 400 gpurun_out/${tag}_$f.json; echo; done
