import sys, random
import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R); sys.path.insert(0,os.path.join(R,'tests'))
from test_queue import shape_request, make_request
from pingoo_b200 import pack_requests
rng=random.Random(int(sys.argv[1])); n=int(sys.argv[2])
POOL=[b" ", b"\t", b"a", b"Z", b"/", b"~", b"\x7f", b"\x1f", b"\x00", b"\x80", b"\xc3\xa9", b"x"*50, b"-", b".", b"\r", b"\n", b"!", b"%20"]
def rnd(maxlen):
    k=rng.random()
    L=rng.choice([0,1,2,5,40,200,255,256,257,300]) if k<0.5 else rng.randint(0,maxlen)
    out=b""
    while len(out)<L: out+=rng.choice(POOL)
    return out[:L]
bad=0
for i in range(n):
    r=dict(host=rnd(300), url=b"/"+rnd(100).replace(b"\x00",b"0"), path=b"/"+rnd(60).replace(b"\x00",b"0")+rng.choice([b"",b"/",b"//"]), method=rng.choice([b"GET",b"POST",b""]), user_agent=rnd(300))
    got=shape_request(make_request(**r))
    # python restatement works on str for ASCII-visible; emulate by bytes: use pack_requests when decodable as latin-1? The packer takes str: feed latin-1 decoded
    try:
        want=pack_requests([dict({k:(v.decode('latin-1')) for k,v in r.items()}, ip="1.2.3.4", remote_port=1)], encoding='latin-1') if False else None
    except TypeError:
        want=None
    # direct restatement of http_listener.rs:159-165,284-296 + http_utils.rs:114-116
    def hdr(v):
        if any(not (32<=b<=126 or b==9) for b in v): return b""
        v=v.strip(b" \t")   # str::trim on visible ASCII / tab only differs on \x0b\x0c which to_str rejects anyway
        return b"" if len(v)>256 else v
    exp=dict(host=hdr(r["host"]), user_agent=hdr(r["user_agent"]), url=r["url"], path=r["path"].rstrip(b"/"), method=r["method"])
    for f in exp:
        if got[f]!=exp[f]:
            bad+=1
            if bad<=8: print("DIFF",f,repr(r[f][:80]),"got",repr(got[f][:80]),"want",repr(exp[f][:80]))
print("n",n,"bad",bad)
