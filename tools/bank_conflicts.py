"""CPU emulation of the shared-memory bank conflicts of the scan's row look-ups (analysis for the next round; no GPU).

For each scan unit of a config-2 rule set it reports the average number of wavefronts per warp-wide row look-up under the
current layout and under alternatives (row stride padded to an odd number of words, 128-byte aligned rows).  The
emulation follows one warp of 32 lanes with the field path's schedule (tests/sim/sim.cpp: pgwsim_bank_stats).
usage: python tools/bank_conflicts.py [n_rules]"""
import ctypes as C
import os
import re
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import synth  # noqa: E402
from helpers import Sim, sim_lib  # noqa: E402
from pingoo_b200 import _ffi  # noqa: E402

lib = sim_lib()
lib.pgwsim_bank_stats.argtypes = [C.c_void_p, C.POINTER(_ffi.Batch), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]

n_rules = int(sys.argv[1]) if len(sys.argv) > 1 else 128
rules, payloads, _ = synth.make_ruleset(n_rules)
batch = synth.RequestStream(config_id=2, payloads=payloads).generate(0, 20_000)
sim = Sim(rules)
cb = batch.as_ctypes()
units = re.findall(r"\[(\w+): states=(\d+) classes=(\d+)\]", sim.describe())
for u, (field, states, classes) in enumerate(units):
    C2 = 2 * int(classes)
    out = (C.c_uint64 * 4)()

    def run(stride):
        lib.pgwsim_bank_stats(sim.h, C.byref(cb), u, stride, None, None, out)
        return out[1] / max(1, out[0]), out[2] / max(1, out[0]), out[3] / max(1, out[0])

    base, lanes, distinct = run(C2)
    odd = (C2 + 3) // 4 * 4
    if (odd // 4) % 2 == 0:
        odd += 4
    alt_odd = run(odd)[0]
    alt_128 = run((C2 + 127) // 128 * 128)[0]
    print(f"unit {u} {field:10s} states {states:>5} classes {classes:>3}: active lanes/step {lanes:4.1f}, distinct states/step {distinct:4.1f}; "
          f"wavefronts per look-up: stride {C2} B -> {base:4.2f}; stride {odd} B (odd words) -> {alt_odd:4.2f}; 128-B rows -> {alt_128:4.2f}")
