import sys, random, re, json
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
import numpy as np
import regex as _regex
import test_oracle as T
from helpers import oracle_lib, Sim
from pingoo_b200 import Action, Rule, pack_requests
seed=int(sys.argv[1]); n=int(sys.argv[2])
rng=random.Random(seed)
L=oracle_lib()
alphabet="abcABCx1 2/%=.\n_-é"
bad=0; checked=0
pats=[]
for _ in range(n):
    pat=T.random_pattern(rng)
    if pat.startswith('(?i)') and any(k in pat for k in ('Lu','Ll','Uppercase','Lowercase')): continue   # Rust folds cased properties under (?i), the regex module does not
    try: pre=_regex.compile(T.to_python(pat), _regex.ASCII)
    except (re.error,_regex.error): continue
    hays=["","a","ab","abc","aB1","_a-"]+["".join(rng.choice(alphabet) for _ in range(rng.randint(0,30))) for _ in range(30)]
    pb=pat.encode()
    for h in hays:
        if h=="" and "\\B" in pat: continue
        if "é" in h: continue   # A9: the oracle treats a non-ASCII char as one opaque char; python's ASCII mode agrees only for classes; skip
        hb=h.encode()
        got=L.orc_regex_is_match(pb,len(pb),hb,len(hb))
        ref=T.py_search(pre,h)
        if ref is None: continue
        checked+=1
        if got!=(1 if ref else 0):
            bad+=1
            if bad<=10: print("ORACLE DIFF", repr(pat), repr(h), "oracle", got, "python", ref)
    pats.append(pat)
# product DFA vs python
hays=["","a","ab","abc"]+["".join(rng.choice(alphabet[:-1]) for _ in range(rng.randint(0,40))) for _ in range(200)]
batch=pack_requests([T.req(url=h) for h in hays])
pbad=0
for p in [q for q in pats if len(q)<=60][:400]:
    rules=[Rule("r","http_request.url.matches("+json.dumps(p)+")",[Action.BLOCK])]
    try: sim=Sim(rules, eval_gates=False)
    except ValueError as e:
        print("REFUSED", repr(p), str(e)[-80:]); continue
    got=sim.evaluate(batch)&3
    pre=_regex.compile(T.to_python(p), _regex.ASCII)
    for i,h in enumerate(hays):
        if i==0 and "\\B" in p: continue
        r=T.py_search(pre,h)
        if r is None: continue
        if int(got[i])!=(1 if r else 0):
            pbad+=1
            if pbad<=10: print("PRODUCT DIFF", repr(p), repr(h), int(got[i]), r)
print("seed",seed,"checked",checked,"oracle diffs",bad,"product diffs",pbad)
