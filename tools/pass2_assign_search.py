#!/usr/bin/env python
"""Search the pass-2 work assignment (lane -> pair, row) of the fused kernel for shared-memory bank conflicts:
row loads (LDS.128), power stores (STS.64) and the parking stores of the self-paired units, with the three
self-paired units pinned to lanes 0..2.  Result: kPass2Tab in auditory_b200/csrc/aud_kernels.cuh."""
import sys, random
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bankconf import wavefronts
PS=554; RS=int(sys.argv[1]) if len(sys.argv)>1 else 22; dE=tuple(int(x) for x in sys.argv[2].split(',')) if len(sys.argv)>2 else (0,10,4)
TRIALS=int(sys.argv[3]) if len(sys.argv)>3 else 60
E=[q*PS+dE[q] for q in range(3)]
Pb=[q*PS for q in range(3)]
def partner(u): return 10 if u==0 else (0 if u==10 else 20-u)
def cost(assign,verbose=False):
    # assign[lane] = (q, rowA) or None
    def w(fn,width,skip=None):
        return wavefronts([None if (a is None or (skip and skip(a))) else fn(a) for a in assign],width)
    c={}
    c['ldA']=10*w(lambda a: 8*(E[a[0]]+RS*a[1]),16)
    c['ldB']=10*w(lambda a: 8*(E[a[0]]+RS*partner(a[1])),16)
    selfp=lambda a: a[1] in (0,10)
    c['p1']=10*w(lambda a: 8*(Pb[a[0]]+a[1]),8,selfp)
    c['p2']=10*w(lambda a: 8*(Pb[a[0]]+partner(a[1])),8,selfp)
    c['park']=20*w(lambda a: 8*(Pb[a[0]]+220),16,lambda a: not selfp(a))
    t=sum(c.values())
    if verbose: print(c)
    return t
cur=[(l//10, l%10) if l<30 else None for l in range(32)]
print('current',cost(cur,True))
random.seed(11)
best=None
for trial in range(TRIALS):
    combos=[(q,u) for q in range(3) for u in range(1,10)]
    random.shuffle(combos)
    assign=[(0,0),(1,0),(2,0)]+combos+[None,None]
    # random orientation
    assign=[None if a is None else (a[0], a[1] if (a[1]==0 or random.random()<0.5) else partner(a[1])) for a in assign]
    c0=cost(assign)
    imp=True
    while imp:
        imp=False
        for a in range(3,30):
            for b in range(a+1,30):
                assign[a],assign[b]=assign[b],assign[a]; c=cost(assign)
                if c<c0: c0=c; imp=True
                else: assign[a],assign[b]=assign[b],assign[a]
            old=assign[a]; assign[a]=(old[0],partner(old[1])); c=cost(assign)
            if c<c0: c0=c; imp=True
            else: assign[a]=old
    if best is None or c0<best[0]: best=(c0,assign[:]); print(trial,c0)
    if c0<=140: break
print(best); cost(best[1],True)
print('ideal: ld 40+40, p 20+20, park 20 = 140')
