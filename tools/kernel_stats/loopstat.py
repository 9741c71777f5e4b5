import re,sys
# crude: list instructions, find backward branches (loops), print the two biggest inner loops with opcode histogram
f=sys.argv[1]
ins=[]
for l in open(f):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2).strip()))
loops=[]
for a,t in ins:
    m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)',t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a: loops.append((tgt,a))
loops.sort(key=lambda x:x[1]-x[0])
from collections import Counter
for tgt,a in loops:
    n=(a-tgt)//16+1
    if n<150: continue
    body=[t for x,t in ins if tgt<=x<=a]
    c=Counter()
    for t in body:
        t=re.sub(r'^@!?U?P\d+\s+','',t)
        op=t.split()[0].split('.')[0]
        c[op]+=1
    print(hex(tgt),hex(a),n,dict(c.most_common(24)))
