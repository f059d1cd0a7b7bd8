#!/usr/bin/env python
"""Dynamic instruction count of one kernel from a cuobjdump -sass listing:
    python tools/sass_count.py listing.sass LO:HI:WEIGHT [LO:HI:WEIGHT ...]
Each range [LO, HI) of instruction addresses (hex) is weighted by its trip count (nested ranges:
the largest weight wins), e.g. the per-unit main loop once and the rolled zeta loop 4 times."""
import re,sys,collections
f=sys.argv[1]
ranges=[(int(a,16),int(b,16),float(m)) for a,b,m in (r.split(':') for r in sys.argv[2:])]
ins=[]
for line in open(f):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',line)
    if m:
        addr=int(m.group(1),16); t=m.group(2).strip()
        t=re.sub(r'^@!?U?P\w+\s+','',t)
        op=t.split()[0].split('.')[0]
        ins.append((addr,op))
tot=collections.Counter()
for a,op in ins:
    w=0
    for lo,hi,m in ranges:
        if lo<=a<hi: w=max(w,m)
    tot[op]+=w
fp={'DFMA','DMUL','DADD','MUFU'}
s=sum(tot.values()); f64=sum(v for k,v in tot.items() if k in fp)
print('total',s,'fp64',f64,'other',s-f64)
for k,v in tot.most_common(40):
    if v: print(f'{k:10s} {v:8.0f}')
