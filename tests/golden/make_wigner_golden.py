"""Generates tests/golden/wigner_golden.json with sympy (independent of both the
oracle and the product): 3j(000), general 3j (the dipole Wigner-Eckart factors) and 6j
values for integer momenta, including structural zeros.  Run in the build container:  python make_wigner_golden.py"""
import json
import os
import random

from sympy import N
from sympy.physics.wigner import wigner_3j, wigner_6j

random.seed(20261017)
three, six = [], []
for _ in range(150):
    a, b, c = (random.randint(0, 18) for _ in range(3))
    three.append([a, b, c, float(N(wigner_3j(a, b, c, 0, 0, 0), 30))])
for a in range(0, 5):
    for b in range(0, 5):
        for c in range(abs(a - b), a + b + 1):
            three.append([a, b, c, float(N(wigner_3j(a, b, c, 0, 0, 0), 30))])
while len(six) < 400:
    j = [random.randint(0, 16) for _ in range(6)]
    try:
        v = float(N(wigner_6j(*j), 30))
    except ValueError:
        v = 0.0
    if v == 0.0 and random.random() < 0.8:
        continue
    six.append(j + [v])
# the shape ang_k_LS uses: {la lb L; ld lc k} with large k
for la, lb, L, ld, lc, k in [(15, 14, 12, 13, 15, 28), (10, 10, 8, 10, 10, 20), (6, 5, 4, 6, 3, 9),
                             (3, 3, 2, 3, 3, 6), (2, 1, 1, 1, 2, 3), (15, 15, 0, 15, 15, 30)]:
    six.append([la, lb, L, ld, lc, k, float(N(wigner_6j(la, lb, L, ld, lc, k), 30))])
# general 3j with a small middle momentum (the dipole operator has rank 1)
import itertools
three_m = []
for ja, jb, jc in itertools.product(range(0, 7), (0, 1, 2), range(0, 7)):
    if ja + jb < jc or abs(ja - jb) > jc:
        continue
    for ma in range(-ja, ja + 1):
        for mb in range(-jb, jb + 1):
            mc = -ma - mb
            if abs(mc) <= jc:
                three_m.append([ja, jb, jc, ma, mb, mc, float(N(wigner_3j(ja, jb, jc, ma, mb, mc), 30))])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wigner_golden.json")
json.dump({"three_j0": three, "six_j": six, "three_j": three_m}, open(out, "w"))
print(len(three), len(six), len(three_m), out)
