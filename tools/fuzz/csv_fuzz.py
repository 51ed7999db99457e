import sys, random
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
from helpers import Oracle, Sim
from pingoo_b200 import Action, Rule, ListType, pack_requests
seed=int(sys.argv[1]); rounds=int(sys.argv[2])
rng=random.Random(seed)
ATOMS={ListType.Ip:["1.2.3.4","10.0.0.0/8","192.168.1.7/32","2001:db8::/32","::1","1.2.3","1.2.3.4/33","300.1.1.1","fe80::1%eth0","0.0.0.0/0","::/0","1.2.3.4/", " 8.8.8.8 ","::ffff:1.2.3.4","1.2.3.4/24","01.2.3.4","2001:db8::1/129","1.2.3.0/255.255.255.0","1.2.3.0/255.0.255.0","10.0.0.0/08","10.0.0.0/+8","10.0.0.0/ 8","1.2.3.4/032","1.2.3.4/0032","2001:db8::/ffff::","10.0.0.0/255.255.255.255","10.1.2.3/0.0.0.0","1.2.3.4/-1","1.2.3.4/8/9","::1/128","2001:DB8::1","1.2.3.4/256","2001:db8::/0128"],
       ListType.Int:["1","-3","+5"," 64500 ","1x","0x10","","9223372036854775807","9223372036854775808","-9223372036854775808","1.0","١"],
       ListType.String:["example.com"," spaced ","a,b",'q"x',"","évil","UPPER","x"*300,"tab\there","#c"]}
def row(t):
    a=rng.choice(ATOMS[t])
    k=rng.random()
    if k<0.2: a='"'+a.replace('"','""')+'"'
    r=a
    if rng.random()<0.3: r+=","+rng.choice(["comment","\"quoted, comment\"",""," x "])
    if rng.random()<0.05: r+=",third"
    return r
probes=[dict(host=h, url="/", path="/", method="GET", user_agent="Mozilla/5.0", ip=ip, remote_port=1, flags=0, asn=asn, country="US")
        for h in ["example.com","spaced"," spaced ","a,b",'q"x',"","évil","UPPER","tab\there","#c","x"*300][:10] for ip,asn in [("1.2.3.4",1),("10.9.9.9",-3),("8.8.8.8",5),("2001:db8::5",64500),("::1",16),("192.168.1.7",0),("::ffff:1.2.3.4",9223372036854775807)]]
bad=0
for r_ in range(rounds):
    t=rng.choice([ListType.Ip, ListType.Int, ListType.String])
    n=rng.randrange(0,6)
    eol=rng.choice(["\n","\r\n"])
    text=eol.join(row(t) for _ in range(n))
    if rng.random()<0.7: text+=eol
    if rng.random()<0.2: text=eol+text
    csv=text.encode()
    expr={ListType.Ip:'lists["l"].contains(client.ip)', ListType.Int:'lists["l"].contains(client.asn)', ListType.String:'lists["l"].contains(http_request.host)'}[t]
    rules=[Rule("r", expr, [Action.BLOCK])]
    res=[]
    for cls in (Oracle, Sim):
        try:
            e=cls(rules, {"l": (t, csv)}, eval_gates=False)
            reqs=[dict(p) for p in probes]
            res.append(("ok", e.evaluate(pack_requests(reqs)).tolist()))
        except Exception as ex:
            res.append(("err", str(ex)))
    if res[0]!=res[1]:
        bad+=1
        if bad<=8: print("DIFF type",t,"csv",repr(csv),"\n oracle:",str(res[0])[:200],"\n engine:",str(res[1])[:200])
print("seed",seed,"rounds",rounds,"bad",bad)
