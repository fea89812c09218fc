#!/bin/bash
# usage: tools/sass_loop.sh NOPP NB  -> opcode histogram of the kernel's innermost trial loop (backward branch span)
K="_ZN3npk${3:-21equity_uniform_kernel}ILi${1}ELi${2}EEEvNS_12EquityParamsE"
cuobjdump -sass -fun "$K" neuron_poker_b200/libnpk.so | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*//' > /tmp/k.txt
python3 - <<'PY'
import re,collections
lines=[l.rstrip('\n') for l in open('/tmp/k.txt')]
addr=[int(l.split()[0],16) for l in lines]
# find backward branch with the largest span that contains LDS+STS
best=None
for i,l in enumerate(lines):
    m=re.search(r'BRA (0x[0-9a-f]+)',l)
    if m:
        t=int(m.group(1),16)
        if t<addr[i]:
            j=addr.index(t) if t in addr else None
            if j is not None:
                body=lines[j:i+1]
                if any('LDS.U16' in b for b in body) and len(body) > 150:
                    if best is None or (i-j)<best[1]-best[0]: best=(j,i)
j,i=best
c=collections.Counter()
for l in lines[j:i+1]:
    t=l.split()[1:]
    if t[0].startswith('@'): t=t[1:]
    c[t[0].split('.')[0]]+=1
print("loop instrs:",i-j+1)
print("  ".join("%s %d"%(k,v) for k,v in c.most_common()))
alu=sum(v for k,v in c.items() if k in ('LOP3','IADD3','SHF','ISETP','VIMNMX','VIADD','PRMT','SEL','FLO','PLOP3','LEA','MOV','IABS'))
print("alu-pipe approx:",alu," imad:",c['IMAD']," lds+sts:",c['LDS']+c['STS'])
PY
