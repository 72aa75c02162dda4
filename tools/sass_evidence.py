"""SASS evidence: count the tcgen05 / TMA / mbarrier mnemonics per kernel of the built library.
   python tools/sass_evidence.py > profiles/rNN_sass_tcgen05_tma.txt      (needs cuobjdump; no GPU)"""
import collections, os, re, subprocess, sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'patchgan_b200', 'libpatchgan_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
COLS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'UTCBAR.2CTA.MULTICAST', 'LDTM', 'UTMALDG', 'UTMALDG.2CTA', 'UTMASTG', 'UTMAREDG',
        'UTMAPF', 'SYNCS', 'REDG', 'BAR.SYNC', 'UCGABAR', 'NANOSLEEP']
counts, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        fn = re.sub(r'^_ZN2pg\d+', '', m.group(1))
        counts[fn] = collections.Counter()
        continue
    m = re.search(r'/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and fn:
        op = m.group(1)
        c = counts[fn]
        if op.startswith('UTCHMMA'):
            c['UTCHMMA.2CTA' if '.2CTA' in op else 'UTCHMMA'] += 1
        elif op.startswith('UTCBAR'):
            c['UTCBAR.2CTA.MULTICAST' if 'MULTICAST' in op else 'UTCBAR'] += 1
        elif op.startswith('UTMALDG'):
            c['UTMALDG.2CTA' if '.2CTA' in op else 'UTMALDG'] += 1
        else:
            for k in ('LDTM', 'UTMASTG', 'UTMAREDG', 'UTMAPF', 'SYNCS', 'REDG', 'BAR.SYNC', 'UCGABAR', 'NANOSLEEP'):
                if op.startswith(k):
                    c[k] += 1
print('# SASS evidence: tcgen05 / TMA instructions per kernel of patchgan_b200/libpatchgan_b200.so (sm_100a)')
print('# made by: python tools/sass_evidence.py   (cuobjdump -sass, mnemonics counted per function)')
print('# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit (.2CTA.MULTICAST = the pair form),')
print('# UTMALDG/UTMASTG/UTMAREDG = cp.async.bulk.tensor load / store / reduce, UTMAPF = prefetch.tensormap, SYNCS = mbarrier ops,')
print('# UCGABAR = barrier.cluster')
print()
print(f"{'kernel':64s}" + ''.join(f'{c:>11s}' if len(c) <= 10 else f' {c:>21s}' for c in COLS))
for fn, c in counts.items():
    if c['UTCHMMA'] + c['UTCHMMA.2CTA'] + c['UTMALDG'] + c['UTMALDG.2CTA'] + c['UTMASTG'] + c['UTMAREDG'] == 0:
        continue
    print(f'{fn[:64]:64s}' + ''.join(f'{c[k]:11d}' if len(k) <= 10 else f' {c[k]:21d}' for k in COLS))
