"""Deeper run of tests/test_fuzz_patterns.py: python tools/fuzz_patterns.py <seed> <rounds>"""
import os
import random
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
from test_fuzz_patterns import one_round

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rng = random.Random(seed)
bad = 0
for k in range(rounds):
    for p in one_round(rng, n_requests=400):
        bad += 1
        print(f"seed {seed} round {k}: {p}")
print(f"seed {seed}: {rounds} rounds, {bad} problems")
