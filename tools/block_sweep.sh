#!/bin/bash
out=gpurun_out/r02_block_stats3.log
: > $out
run() { echo "=== $*" >> $out; env "$@" timeout 120 python tools/block_stats.py 2072 512 100 1 >> $out 2>&1; }
run NSC_BLOCK_SPLIT=46,36 NSC_BLOCK_RING=256
run NSC_BLOCK_SPLIT=50,36 NSC_BLOCK_RING=256
run NSC_BLOCK_SPLIT=50,40 NSC_BLOCK_RING=256
run NSC_BLOCK_SPLIT=54,40 NSC_BLOCK_RING=256
run NSC_BLOCK_SPLIT=56,36 NSC_BLOCK_RING=256
run NSC_BLOCK_SPLIT=50,36 NSC_BLOCK_RING=128
echo "=== 50-channel block" >> $out
env NSC_BLOCK_SPLIT=44,40 NSC_BLOCK_RING=256 timeout 120 python tools/block_stats.py 2072 512 50 2 >> $out 2>&1
env NSC_BLOCK_SPLIT=52,44 NSC_BLOCK_RING=256 timeout 120 python tools/block_stats.py 2072 512 50 2 >> $out 2>&1
env NSC_BLOCK_SPLIT=56,48 NSC_BLOCK_RING=256 timeout 120 python tools/block_stats.py 2072 512 50 2 >> $out 2>&1
echo "=== L256" >> $out
env NSC_BLOCK_SPLIT=50,36 NSC_BLOCK_RING=256 timeout 120 python tools/block_stats.py 2072 256 100 2 >> $out 2>&1
cat $out
