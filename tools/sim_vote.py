#!/usr/bin/env python3
"""Monte-Carlo model of the voted scheduler of wf_trace_pt (profiles/r02_history.md section 2): lanes own K rays, every trip the warp
executes the ONE operation (WALK / LEAF / NEXT) most lanes can offer a ray for; op bodies are assumed internally uniform.
Prints thread instructions per warp instruction (incl. a 25-instruction vote at 32 lanes), trips and ray swaps per ray."""
import random, sys
random.seed(1)
# per-ray op sequence model (castle bounce rays): instances ~1.6 (1 + geometric), per instance: NEXT, then walk cells
# (mean 8) with leaf visits (prob per cell 0.15), leaf hit prob 0.5 ends instance; final NEXT = refill.
W,L,N=0,1,2
COST={W:100,L:150,N:120}
def ray_seq():
    seq=[]
    ninst=1+(random.random()<0.45)+(random.random()<0.15)
    done=False
    for i in range(ninst):
        seq.append(N)
        ncell=max(1,int(random.expovariate(1/8.0)))
        for c in range(ncell):
            seq.append(W)
            if random.random()<0.15:
                seq.append(L)
                if random.random()<0.5:
                    done=True;break
        if done:break
    seq.append(N)   # finish/refill
    return seq
def sim(K,nrays=20000,policy="plurality"):
    lanes=[[ray_seq() for _ in range(K)] for _ in range(32)]
    pos=[[0]*K for _ in range(32)]
    issued=0;thread=0;remaining=nrays;trips=0;swaps=0
    active=[0]*32
    while True:
        cnt=[0,0,0]
        want=[[None]*K for _ in range(32)]
        any_=False
        for l in range(32):
            ops=set()
            for k in range(K):
                if lanes[l][k] is not None:
                    ops.add(lanes[l][k][pos[l][k]])
            for o in ops: cnt[o]+=1
            if ops: any_=True
        if not any_:break
        # plurality weighted by nothing
        op=max(range(3),key=lambda o:(cnt[o],-o))
        trips+=1
        served=0
        for l in range(32):
            # prefer active ray
            order=[active[l]]+[k for k in range(K) if k!=active[l]]
            for k in order:
                r=lanes[l][k]
                if r is not None and r[pos[l][k]]==op:
                    if k!=active[l]: swaps+=1; active[l]=k
                    pos[l][k]+=1
                    served+=1
                    if pos[l][k]==len(r):
                        if remaining>0:
                            remaining-=1; lanes[l][k]=ray_seq(); pos[l][k]=0
                        else: lanes[l][k]=None
                    break
        issued+=COST[op]+25
        thread+=COST[op]*served+25*32
    return thread/issued, trips, swaps/nrays
for K in (1,2,3,4):
    d,t,s=sim(K)
    print(K,"density %.1f"%d,"trips",t,"swaps/ray %.2f"%s)
