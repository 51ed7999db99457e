import sys, random, os
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
import synth
from helpers import Sim, Oracle
from pingoo_b200 import Action, Rule
seed=int(sys.argv[1]); rounds=int(sys.argv[2])
rng=random.Random(seed)
base, records = synth.make_geoip(300, config_id=9)
rules=[Rule("r","client.asn == 1",[Action.BLOCK])]
ip=np.zeros((64,16),dtype=np.uint8); v6=np.zeros(64,dtype=np.uint8)
for i in range(64):
    v6[i]=i%2
    for k in range(16 if v6[i] else 4): ip[i,k]=rng.randrange(256)
ok=err=diff=0
for r in range(rounds):
    b=bytearray(base)
    k=rng.random()
    if k<0.3:
        for _ in range(rng.randrange(1,8)): b[rng.randrange(len(b))]=rng.randrange(256)
    elif k<0.5: b=b[:rng.randrange(0,len(b))]
    elif k<0.7:
        i=rng.randrange(len(b)); b[i:i+rng.randrange(1,64)]=bytes(rng.randrange(256) for _ in range(rng.randrange(0,64)))
    elif k<0.85:
        # corrupt metadata region (tail)
        for _ in range(rng.randrange(1,6)): b[len(b)-1-rng.randrange(min(len(b),400))]=rng.randrange(256)
    else:
        for _ in range(rng.randrange(1,4)): b[rng.randrange(min(len(b), 4000))]^=1<<rng.randrange(8)
    data=bytes(b)
    res=[]
    for cls in (Sim, Oracle):
        try:
            e=cls(rules, geoip_mmdb=data)
            if cls is Sim:
                a,c=e.geoip_lookup(ip,v6); res.append(("ok",a.tolist(),c.tolist()))
            else:
                out=[e.geoip_lookup(bytes(ip[i]), int(v6[i])) for i in range(64)]
                res.append(("ok",[o[0] for o in out],[ord(o[1][0])|(ord(o[1][1])<<8) for o in out]))
        except Exception as ex:
            res.append(("err",str(ex)[:80]))
    if res[0][0]=="ok": ok+=1
    else: err+=1
    if res[0][0]!=res[1][0] or (res[0][0]=="ok" and res[0]!=res[1]):
        diff+=1
        if diff<=5: print("DIFF round",r,"engine",str(res[0])[:160],"\n   oracle",str(res[1])[:160])
print("seed",seed,"ok",ok,"err",err,"diff",diff)
