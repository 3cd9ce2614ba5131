"""SASS evidence of the tcgen05 / TMEM / bulk-copy path in the built library -> profiles/<tag>_sass_tcgen05.txt
   python tools/sass_extract.py [tag]      (runs `cuobjdump -sass nsc_b200/libnsc_b200.so`; no GPU needed)
Per kernel: counts of the Blackwell-native mnemonics (B200_PROFILING.md: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM,
tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS) and one sample line of each."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
so = os.path.join(ROOT, 'nsc_b200', 'libnsc_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip() or n
WANT = ('UTCHMMA', 'LDTM', 'UTCBAR', 'UBLKCP', 'SYNCS', 'UTMALDG', 'HMMA', 'SHFL', 'FADD2', 'REDG', 'ATOM')
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = demangle(m.group(1))
        cur = re.sub(r'\(anonymous namespace\)::', '', cur)
        kernels[cur] = (collections.Counter(), {})
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)\s*(.*?);', line)
    if m and cur:
        op = m.group(1)
        base = op.split('.')[0]
        key = 'UTCHMMA.2CTA' if op.startswith('UTCHMMA.2CTA') else base
        if base in WANT:
            kernels[cur][0][key] += 1
            kernels[cur][1].setdefault(key, (op + ' ' + m.group(2)).strip())
out = [f"# SASS evidence of the tcgen05 / TMEM / bulk-copy path in nsc_b200/libnsc_b200.so (cuobjdump -sass, sm_100a), {tag}.",
       "# Per kernel: instruction counts of the Blackwell-native mnemonics (B200_PROFILING.md: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM,",
       "# tcgen05.commit -> UTCBAR, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS) and one sample line of each.", ""]
tot = collections.Counter()
legacy = 0
for name, (cnt, sample) in kernels.items():
    legacy += cnt.get('HMMA', 0)
    if not any(k in cnt for k in ('UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UBLKCP')):
        continue
    tot.update(cnt)
    out.append(name)
    out.append('  ' + ', '.join(f"{k} {v}" for k, v in cnt.items()))
    for k in ('LDTM', 'UBLKCP', 'UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'REDG'):
        if k in sample:
            out.append(f"    {k:<13} e.g.  {sample[k]}")
    out.append('')
out.insert(3, f"# whole library: " + ', '.join(f"{k} {v}" for k, v in sorted(tot.items())) + f"; legacy HMMA (mma.sync): {legacy}; UTMALDG (tensor-map TMA): "
           f"{tot.get('UTMALDG', 0)} -- HBM holds the swizzled shared-memory image, so 1-D bulk copies (UBLKCP) are the whole transfer")
open(os.path.join(ROOT, 'profiles', f'{tag}_sass_tcgen05.txt'), 'w').write('\n'.join(out) + '\n')
print('\n'.join(out[:6]))
