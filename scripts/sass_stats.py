"""Instruction mix of the SASS region spanned by the HMMA instructions of a cuobjdump -sass listing."""
import re, sys
from collections import Counter
lines = open(sys.argv[1]).read().split('\n')
ins = [l for l in lines if re.search(r'/\*[0-9a-f]{4,}\*/', l)]
idx = [i for i, l in enumerate(ins) if 'HMMA' in l or 'UTCHMMA' in l]
lo, hi = idx[0], idx[-1]
print('total instr', len(ins), 'hmma span', lo, hi, hi - lo, 'hmma count', len(idx))
c = Counter()
for l in ins[lo - 60:hi + 40]:
    m = re.search(r'\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m:
        parts = m.group(2).split('.')
        key = parts[0] + ('.' + '.'.join(parts[1:]) if parts[0] in ('LDS', 'STS', 'LDG') else '')
        c[key] += 1
print(c.most_common(40))
