#!/bin/bash
# A/B of two builds of the library on ONE box: nsc_b200/libnsc_b200.so vs nsc_b200/libnsc_b200_alt.so (interleaved runs)
set -u
cp nsc_b200/libnsc_b200.so /tmp/lib_main.so
for rep in 1 2 3; do
  for v in main alt; do
    if [ $v = main ]; then cp /tmp/lib_main.so nsc_b200/libnsc_b200.so; else cp nsc_b200/libnsc_b200_alt.so nsc_b200/libnsc_b200.so; fi
    timeout 200 python bench.py --steps 8 --warmup 3 --no-sub-records --no-cpu-baseline > /tmp/ab.json 2>/dev/null
    python -c "
import json
d=json.load(open('/tmp/ab.json'))
print('$v', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['clocks']['sm_mhz'])
"
  done
done
cp /tmp/lib_main.so nsc_b200/libnsc_b200.so
