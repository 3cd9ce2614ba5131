#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_block.py -x -q -m gpu -k "three_launches or folded or deterministic" 2>&1 | tail -2
done
timeout 600 python -m pytest tests/test_gpu_plane.py -x -q -m gpu -k "bit_identical or full_size or pair_layers" 2>&1 | tail -2
