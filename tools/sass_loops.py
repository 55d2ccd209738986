#!/usr/bin/env python
"""Static look at a kernel's SASS: loops (backward branches) with their instruction counts and
opcode mix.  No GPU needed:
    cuobjdump -sass pumi-pic_b200/_obj/pp_search.cu.o | awk '/Function : .*k_walk_scsILi3ELb1ELb0ELi4/{f=1} f{print} /Function : /{if(f&&!/k_walk_scsILi3ELb1ELb0ELi4/)exit}' > /tmp/k.sass
    python tools/sass_loops.py /tmp/k.sass
"""
import re, collections, sys
def load(path):
    ins=[]
    for l in open(path).read().split('\n'):
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);',l)
        if m: ins.append((int(m.group(1),16),m.group(2).strip()))
    return ins
def loops(ins):
    out=[]
    for a,t in ins:
        if 'BRA' in t:
            mm=re.search(r'0x([0-9a-f]+)',t)
            if mm:
                tgt=int(mm.group(1),16)
                if tgt<a: out.append((tgt,a,sum(1 for x,_ in ins if tgt<=x<=a)))
    return out
def hist(ins,lo,hi):
    c=collections.Counter()
    for a,t in ins:
        if lo<=a<=hi:
            t2=re.sub(r'^@!?U?P\d+\s+','',t)
            op=t2.split()[0]
            op='.'.join(op.split('.')[:2]) if op.startswith(('IMAD','LDS','LDGSTS')) else op.split('.')[0]
            c[op]+=1
    return c
ins=load(sys.argv[1])
print(len(ins),'instructions')
for lo,hi,n in loops(ins):
    print(hex(lo),hex(hi),n, dict(hist(ins,lo,hi).most_common(8)))
