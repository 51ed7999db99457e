import sys, random
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
from helpers import Oracle, Sim
from pingoo_b200 import Action, Rule, ListType, pack_requests
seed=int(sys.argv[1]); rounds=int(sys.argv[2])
rng=random.Random(seed)
EXT=[0,1,-1,2,7,80,443,65535,65536,2**31-1,2**31,-2**31,2**32,2**62,2**63-1,-2**63,-2**63+1,2**63-2,4294967295]
def c(): 
    v=rng.choice(EXT) if rng.random()<0.7 else rng.randint(-2**63,2**63-1)
    return str(v) if v>=0 else ("-"+str(-v) if rng.random()<0.7 else "(0 - "+str(-v-1)+" - 1)")
VARS=["client.asn","client.remote_port","http_request.url.length()","http_request.host.length()"]
def term(d):
    r=rng.random()
    if d<=0 or r<0.35: return rng.choice(VARS) if rng.random()<0.6 else c()
    op=rng.choice(["+","-","*","/","%"])
    if r<0.9: return "("+term(d-1)+" "+op+" "+term(d-1)+")"
    return "(-"+term(d-1)+")"
bad=0
for r in range(rounds):
    rules=[]
    ints=sorted(set(rng.choice(EXT) for _ in range(rng.randint(0,8))))
    lists={"li":(ListType.Int, ("\n".join(str(v) for v in ints)+"\n").encode())}
    for i in range(8):
        k=rng.random()
        if k<0.5: ex=term(2)+" "+rng.choice(["==","!=","<","<=",">",">="])+" "+term(2)
        elif k<0.65: ex='lists["li"].contains('+rng.choice(VARS[:2]+[term(1)])+')'
        elif k<0.8: ex="["+", ".join(c() for _ in range(rng.randint(0,4)))+"].contains("+term(1)+")"
        else: ex=rng.choice(VARS)+" "+rng.choice(["==","!=","<","<=",">",">="])+" "+c()
        if rng.random()<0.2: ex="!("+ex+")"
        rules.append(Rule(f"r{i}", ex, [Action.BLOCK if i%2 else Action.CAPTCHA]))
    reqs=[dict(host="h"*rng.choice([0,1,2,7]), url="/"*rng.choice([1,2,80]), path="/", method="GET", user_agent="M", ip="1.2.3.4", remote_port=rng.choice([0,1,7,80,443,65535]), flags=i%2,
               asn=rng.choice(EXT+[64512, 3]), country="US") for i in range(200)]
    batch=pack_requests(reqs)
    try:
        want=Oracle(rules, lists, eval_gates=False).evaluate(batch, threads=4)
    except Exception as e: print("oracle refused", str(e)[:100], [x.expression for x in rules][:2]); continue
    for one in [None]+list(range(8)):
        rs=rules if one is None else [rules[one]]
        w=want if one is None else Oracle(rs, lists, eval_gates=False).evaluate(batch, threads=2)
        try: got=Sim(rs, lists, eval_gates=False).evaluate(batch)
        except Exception as e:
            msg=str(e)
            if "not supported" in msg: continue
            print("REFUSED", msg[-120:], rs[0].expression); bad+=1; continue
        d=np.nonzero(got!=w)[0]
        if len(d):
            bad+=1; i=int(d[0]); print("MISMATCH",seed,r,one,len(d),hex(int(w[i])),hex(int(got[i])),"asn",int(batch.asn[i]),"port",int(batch.remote_port[i]),[x.expression for x in rs][:2])
            break
print("seed",seed,"rounds",rounds,"bad",bad)
