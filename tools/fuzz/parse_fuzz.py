import sys, random, ctypes as C
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
from helpers import oracle_lib
from pingoo_b200 import _ffi
import test_fuzz_expressions as t
seed=int(sys.argv[1]); rounds=int(sys.argv[2])
rng=random.Random(seed)
L=oracle_lib(); P=_ffi.load()
P.pgw_compile_expression.argtypes=[C.c_char_p, C.c_char_p, C.c_size_t]; P.pgw_validate_expression.argtypes=[C.c_char_p, C.c_char_p, C.c_size_t]
CH=list("()[]{}\"'\\.,:?!&|=<>+-*/% \t\nabcxyz019_") + ["&&","||","==","!=","<=",">=","r\"","0x","1e","1.","in ","true","null","\\n","\\u00e9","\\x41","b\"","'''",'"""',"-9223372036854775808","9223372036854775808","1u","0xG","..","e5"]
bad=0; acc=0; rej=0
for r in range(rounds):
    e=t.expr(rng, rng.randrange(0,4))
    k=rng.random()
    if k<0.75:
        for _ in range(rng.randrange(1,4)):
            i=rng.randrange(len(e)+1)
            op=rng.random()
            if op<0.4: e=e[:i]+rng.choice(CH)+e[i:]
            elif op<0.8 and e: e=e[:i]+e[i+rng.randrange(1,3):]
            else: e=e[:i]+rng.choice(CH)+e[i+1:]
    eb=e.encode()
    if b"\0" in eb: continue
    for fn in ("compile_expression","validate_expression"):
        e1=C.create_string_buffer(512); e2=C.create_string_buffer(512)
        a=getattr(L,"orc_"+fn)(eb,e1,512); b=getattr(P,"pgw_"+fn)(eb,e2,512)
        if (a==0)!=(b==0):
            bad+=1
            if bad<=12: print("DIFF",fn,repr(e),"oracle",a,e1.value[:90],"engine",b,e2.value[:90])
        elif fn=="compile_expression":
            acc+= a==0; rej+= a!=0
print("seed",seed,"accepted",acc,"rejected",rej,"different",bad)
