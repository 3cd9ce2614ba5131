#!/bin/bash
out=gpurun_out/r02_chunk_sweep.log
: > $out
run() {
  echo "=== $*" >> $out
  env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
kb=d['kernel_breakdown']
print('ms/step', round(d['ms_per_step'],2), 'xRT', round(d['value']), 'e2e', round(d['e2e']['value']), 'sm_mhz', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])
print('   blocks', round(sum(v['ms'] for k,v in kb.items() if k.startswith('pB')),2), 'unfused convs of blocks', round(sum(v['ms'] for k,v in kb.items() if k.startswith(('pT2_k9','pX2_k9d1s1_c20'))),2))
" >> $out 2>&1
}
run NSC_BLOCK_FUSED=0
run NSC_BLOCK_FUSED=1
run NSC_BLOCK_FUSED=1 NSC_PLANE_CHUNK=4144
run NSC_BLOCK_FUSED=1 NSC_PLANE_CHUNK=8288
run NSC_BLOCK_FUSED=0 NSC_PLANE_CHUNK=4144
cat $out
