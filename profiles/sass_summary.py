"""Counts the Blackwell-native SASS mnemonics per kernel of the built library (cuobjdump -sass): tcgen05.mma -> UTC*MMA,
tcgen05.ld -> LDTM, TMA -> UTMALDG, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, multimem -> *MC*/LDGMC/REDG.MC; `.2CTA` marks the
cta_group::2 (CTA pair) forms.  Usage: python profiles/sass_summary.py [path/to/libgccnmf_b200.so] > profiles/r02_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'gcc-nmf_b200/libgccnmf_b200.so'
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
pat = re.compile(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)')
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = pat.match(line)
    if m and cur is not None:
        cur[m.group(1)] += 1
demangled = subprocess.run(['cu++filt'] + list(kernels), capture_output=True, text=True).stdout.splitlines()
keys = ('UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UTMALDG.2CTA', 'UTCBAR', 'SYNCS', 'MULTIMEM')
print('# SASS evidence per kernel (`cuobjdump -sass %s`)\n' % lib)
print('| kernel | instructions | ' + ' | '.join(keys) + ' |')
print('|---|---|' + '---|' * len(keys))
total = collections.Counter()
rows = []
for (name, c), dm in zip(kernels.items(), demangled + [''] * len(kernels)):
    row = collections.Counter()
    for op, n in c.items():
        if op.startswith('UTCHMMA'):
            row['UTCHMMA'] += n
            if '.2CTA' in op:
                row['UTCHMMA.2CTA'] += n
        elif op.startswith('LDTM'):
            row['LDTM'] += n
        elif op.startswith('UTMALDG'):
            row['UTMALDG'] += n
            if '.2CTA' in op:
                row['UTMALDG.2CTA'] += n
        elif op.startswith('UTCBAR'):
            row['UTCBAR'] += n
        elif op.startswith('SYNCS'):
            row['SYNCS'] += n
        elif 'MC' in op.split('.')[0] or '.MC' in op or op.startswith('LDGMC') or op.startswith('REDGMC'):
            row['MULTIMEM'] += n
    total.update(row)
    if not any(row.values()):
        continue
    short = (dm or name).replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('tgemm::', '').replace('(int)', '').replace('(bool)', '')
    short = re.sub(r'^void ', '', short)
    depth, end = 0, len(short)
    for i, ch in enumerate(short):            # cut the argument list: the first '(' outside the template brackets
        depth += ch == '<'
        depth -= ch == '>'
        if ch == '(' and depth == 0:
            end = i
            break
    rows.append((short[:end], sum(c.values()), tuple(row[k] for k in keys)))
for name, n, vals in sorted(rows):
    print('| `%s` | %d | ' % (name, n) + ' | '.join(str(v) for v in vals) + ' |')
print('\nTotals over the library: ' + ', '.join('%s %d' % (k, total[k]) for k in keys))
print('\nKernels without any of these mnemonics (FFT, PHAT, masks, real-time block path, SIMT fallbacks): %d of %d.' % (
    sum(1 for c in kernels.values() if not any(op.startswith(('UTCHMMA', 'LDTM', 'UTMALDG', 'UTCBAR')) for op in c)), len(kernels)))
