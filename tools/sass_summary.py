"""cuobjdump -sass opcode summary per kernel of pingoo_b200/libpingoo_waf.so (evidence for profiles/: which memory /
synchronisation / integer-pipe instructions each kernel is made of).  usage: python tools/sass_summary.py > profiles/<file>"""
import os
import re
import subprocess
import sys
from collections import Counter

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(R, "pingoo_b200", "libpingoo_waf.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+[^\n]*)", res):
    usage[m.group(1)] = m.group(2)
print(f"# {os.path.relpath(lib, R)}: SASS summary (cuobjdump -sass, {txt.count('Function : ')} kernels, arch sm_100a)")
INTEREST = ["UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "LDL", "STL", "LDC", "RED", "ATOM", "ATOMS", "ATOMG", "SHFL", "VOTE", "MATCH", "BAR",
            "IMAD", "IMAD.HI", "IMAD.WIDE", "SHF", "LOP3", "PRMT", "IADD3", "VIADD", "ISETP", "CCTL", "BRA", "CALL"]
for part in txt.split("Function : ")[1:]:
    name = part.split("\n", 1)[0].strip()
    short = re.search(r"\d+(waf_\w+?_kernel|geoip_lookup_kernel|captcha_client_id_kernel)", name)
    short = short.group(1) if short else name
    ins = re.findall(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", part)
    base = Counter(i.split(".")[0] for i in ins)
    full = Counter(ins)
    hi = sum(v for k, v in full.items() if k.startswith("IMAD.HI"))
    wide = sum(v for k, v in full.items() if k.startswith("IMAD.WIDE"))
    ldg128 = sum(v for k, v in full.items() if k.startswith("LDG") and ".128" in k)
    print(f"\n## {short}: {len(ins)} instructions; {usage.get(name, '')}")
    print("   top opcodes: " + ", ".join(f"{k} {v}" for k, v in base.most_common(12)))
    sel = []
    for k in INTEREST:
        if k == "IMAD.HI":
            v = hi
        elif k == "IMAD.WIDE":
            v = wide
        else:
            v = base.get(k, 0)
        if v:
            sel.append(f"{k} {v}")
    print("   of interest: " + ", ".join(sel) + (f", LDG.128 {ldg128}" if ldg128 else ""))
