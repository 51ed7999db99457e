import sys, json, random
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
from helpers import Oracle, Sim
from pingoo_b200 import Action, Rule, pack_requests
seed=int(sys.argv[1]); rounds=int(sys.argv[2])
rng=random.Random(seed)
FR=["a","b","ab","ba","/","x","GET","PUT","","aa","-"]
FIELDS=["host","path","method","url"]
def val(): return "".join(rng.choice(FR) for _ in range(rng.randint(0,3)))
bad=0
for r in range(rounds):
    rules=[]
    for i in range(8):
        parts=[]
        for _ in range(rng.randint(2,4)):
            parts.append("http_request."+rng.choice(FIELDS) if rng.random()<0.6 else json.dumps(val()))
        if not any(p.startswith("http_") for p in parts): parts[0]="http_request.method"
        cat="("+" + ".join(parts)+")"
        C=json.dumps("".join(rng.choice(FR) for _ in range(rng.randint(0,5))))
        k=rng.randrange(7)
        ex={0:f"{cat} == {C}",1:f"{cat} != {C}",2:f"{cat}.contains({C})",3:f"{cat}.starts_with({C})",4:f"{cat}.ends_with({C})",5:f"{cat}.length() {rng.choice(['==','<','>='])} {rng.randint(0,8)}",6:f"{C} == {cat}"}[k]
        if rng.random()<0.2: ex="!("+ex+")"
        rules.append(Rule(f"r{i}", ex, [Action.BLOCK if i%2 else Action.CAPTCHA]))
    reqs=[dict(host=val(), url="/"+val(), path=val(), method=rng.choice(["GET","PUT","","a"]), user_agent="Mozilla/5.0", ip="1.2.3.4", remote_port=1, flags=i%2) for i in range(300)]
    batch=pack_requests(reqs)
    want=Oracle(rules, eval_gates=False).evaluate(batch, threads=4)
    for one in [None]+list(range(len(rules))):
        rs=rules if one is None else [rules[one]]
        w=want if one is None else Oracle(rs, eval_gates=False).evaluate(batch, threads=4)
        try: got=Sim(rs, eval_gates=False).evaluate(batch)
        except Exception as e: print("REFUSED", str(e)[-100:], [x.expression for x in rs]); bad+=1; continue
        d=np.nonzero(got!=w)[0]
        if len(d):
            bad+=1; i=int(d[0])
            print("MISMATCH",seed,r,one,len(d),hex(int(w[i])),hex(int(got[i])),{f:batch.field(f,i) for f in FIELDS},[x.expression for x in rs][:3])
print("seed",seed,"rounds",rounds,"bad",bad)
