import sys, random, ipaddress
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
from helpers import Oracle, Sim
from pingoo_b200 import Action, Rule, ListType, pack_requests
seed=int(sys.argv[1]); rounds=int(sys.argv[2])
rng=random.Random(seed)
bad=0
for r in range(rounds):
    lists={}
    rules=[]
    probes=[]
    nl=rng.randint(1,4)
    for li in range(nl):
        ents=[]
        for _ in range(rng.randint(0,40)):
            if rng.random()<0.6:
                pl=rng.choice([0,1,7,8,9,15,16,17,23,24,25,30,31,32,32,32])
                a=rng.getrandbits(32) if rng.random()<0.7 else rng.choice([0,0xFFFFFFFF,0x0A000000,0x7F000001,0xC0A80000])
                net=ipaddress.IPv4Network((a>>(32-pl)<<(32-pl) if pl else 0, pl))
                ents.append(str(net) if rng.random()<0.8 or pl!=32 else str(net.network_address))
                for _ in range(2):
                    probes.append(str(ipaddress.IPv4Address(min(0xFFFFFFFF,max(0,int(net.network_address)+rng.choice([-1,0,1,net.num_addresses-1,net.num_addresses]))))))
            else:
                pl=rng.choice([0,1,15,16,17,32,47,48,49,63,64,65,96,112,127,128,128])
                a=rng.getrandbits(128) if rng.random()<0.7 else rng.choice([0,1,(1<<128)-1,0x20010db8<<96,0xfe80<<112, 0xFFFF<<32|0x01020304])
                net=ipaddress.IPv6Network((a>>(128-pl)<<(128-pl) if pl else 0, pl))
                ents.append(str(net) if rng.random()<0.8 or pl!=128 else str(net.network_address))
                for _ in range(2):
                    v=int(net.network_address)+rng.choice([-1,0,1,net.num_addresses-1,net.num_addresses])
                    probes.append(str(ipaddress.IPv6Address(min((1<<128)-1,max(0,v)))))
        lists[f"l{li}"]=(ListType.Ip, ("\n".join(ents)+"\n").encode())
        ex=f'lists["l{li}"].contains(client.ip)'
        if rng.random()<0.3: ex="!"+ex
        rules.append(Rule(f"r{li}", ex, [Action.BLOCK if li%2 else Action.CAPTCHA]))
    probes+= [str(ipaddress.IPv4Address(rng.getrandbits(32))) for _ in range(50)]+[str(ipaddress.IPv6Address(rng.getrandbits(128))) for _ in range(30)]+["0.0.0.0","255.255.255.255","::","::1","ffff:ffff:ffff:ffff:ffff:ffff:ffff:ffff","::ffff:1.2.3.4","127.0.0.1"]
    batch=pack_requests([dict(host="h",url="/",path="/",method="GET",user_agent="M",ip=p,remote_port=1,flags=i%2) for i,p in enumerate(probes)])
    want=Oracle(rules, lists, eval_gates=False).evaluate(batch, threads=4)
    got=Sim(rules, lists, eval_gates=False).evaluate(batch)
    d=np.nonzero(got!=want)[0]
    if len(d):
        bad+=1; i=int(d[0]); print("MISMATCH",seed,r,len(d),probes[i],hex(int(want[i])),hex(int(got[i])),{k:v[1][:200] for k,v in lists.items()})
print("seed",seed,"rounds",rounds,"bad",bad)
