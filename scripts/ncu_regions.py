"""Aggregate an `ncu --page source --csv` dump into SASS regions: samples, executed instructions, top stalls."""
import csv, sys
path = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 250
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def I(r, k):
    try: return int(r[ix[k]])
    except ValueError: return 0
for b in range(0, len(data), B):
    chunk = data[b:b + B]
    s = sum(I(r, '# Samples') for r in chunk); ie = sum(I(r, 'Instructions Executed') for r in chunk)
    if s < int(sys.argv[3]) if len(sys.argv) > 3 else s < 30: continue
    st = {h: sum(I(r, h) for r in chunk) for h in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    ops = {}
    for r in chunk:
        t = r[ix['Source']].split()
        op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '')).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    print(b, 'samples', s, 'instexec', ie, top, sorted(ops.items(), key=lambda kv: -kv[1])[:5])
print('total samples', sum(I(r, '# Samples') for r in data), 'inst', sum(I(r, 'Instructions Executed') for r in data), 'sass lines', len(data))
