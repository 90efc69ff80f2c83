import csv, gzip, sys, collections, re
def mix(path):
    rows = list(csv.reader(gzip.open(path, 'rt')))
    name = rows[0][1][:60]
    hdr = rows[1]
    isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    ops = collections.Counter(); smp = collections.Counter(); static = collections.Counter()
    tot = 0
    for r in rows[2:]:
        if len(r) <= iex: continue
        src = r[isrc].strip()
        src = re.sub(r'^@!?U?P\d+\s+', '', src)
        op = src.split()[0] if src else '?'
        base = '.'.join(op.split('.')[:3])
        try: n = int(r[iex])
        except: continue
        ops[base] += n; tot += n; smp[base] += int(r[ismp] or 0); static[base] += 1
    print(name, 'total warp insts', tot)
    for op, n in ops.most_common(22):
        print(f'  {op:28s} {n/tot*100:6.2f}%  static {static[op]:5d}  samples {smp[op]}')
    return ops, tot
for p in sys.argv[1:]: mix(p)
