#!/bin/bash
# last evidence pass of round 2 (no ncu: the captures of the previous pass are of the same kernels)
set -u
tag=r02
mkdir -p gpurun_out
ok=0; bad=0
for i in 1 2 3 4 5 6 7 8; do
  if timeout 120 python -m pytest tests/test_gpu_block.py -x -q -m gpu -k "three_launches" > /tmp/fl.log 2>&1; then ok=$((ok+1)); else bad=$((bad+1)); tail -3 /tmp/fl.log; fi
done
echo "three_launches x8: ok=$ok bad=$bad" | tee gpurun_out/${tag}_three_launches_repeat.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_pytest_gpu.log)"
timeout 600 python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/${tag}_bench_train.json 2> gpurun_out/${tag}_bench_train.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${tag}_smoke.log 2>&1
tail -1 gpurun_out/${tag}_smoke.log
for f in bench_default bench_reference bench_train; do echo "== $f"; head -c 300 gpurun_out/${tag}_$f.json; echo; done
