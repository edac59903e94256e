"""Histogram of the hottest inner loop of a kernel's SASS (development aid)."""
import re, collections, sys
lines = open(sys.argv[1]).read().split('\n')
ins = [l for l in lines if re.search(r'/\*[0-9a-f]{4}\*/', l)]
addr = lambda l: int(re.search(r'/\*([0-9a-f]{4})\*/', l).group(1), 16)
want = int(sys.argv[2]) if len(sys.argv) > 2 else 160
for l in ins:
    m = re.search(r'BRA (0x[0-9a-f]+)', l)
    if not m: continue
    tgt, a = int(m.group(1), 16), addr(l)
    if tgt >= a: continue
    body = [x for x in ins if tgt <= addr(x) <= a]
    nd = sum('DFMA' in x for x in body)
    if nd != want: continue
    c = collections.Counter()
    for x in body:
        mm = re.search(r'\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', x)
        if mm: c[mm.group(2).split('.')[0]] += 1
    fp = c['DFMA'] + c['DADD'] + c['DMUL']
    print(f"loop {tgt:#x}-{a:#x}: {len(body)} instrs, FP64 {fp}, other {len(body)-fp}:", c.most_common())
