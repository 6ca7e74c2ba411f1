"""Generates tests/golden/numeric_golden.json with numpy/scipy (independent of
the oracle and the product): Gauss-Legendre rules and B-spline values and
second derivatives on a generate_grid-style knot vector.
Run in the build container:  python make_numeric_golden.py"""
import json
import os

import numpy as np
from scipy.interpolate import BSpline

out = {}
out["gl"] = {}
for N in (8, 13, 14, 18, 23):
    x, w = np.polynomial.legendre.leggauss(N)
    out["gl"][str(N)] = [x.tolist(), w.tolist()]

# knot vector in the style of grid_tools.f90 (k-fold ends, geometric then linear)
k = 6
inner = [0.0, 0.125, 0.25, 0.375, 0.421875, 0.474609375, 0.6, 0.9, 1.4, 1.9, 2.4, 2.9]
t = np.array([0.0] * (k - 1) + inner + [inner[-1]] * (k - 1))
n = len(t) - k
rng = np.random.default_rng(7)
pts = []
for cell in range(1, len(inner)):
    a, b = inner[cell - 1], inner[cell]
    for x in a + (b - a) * rng.uniform(0.02, 0.98, 3):
        vals, d2 = [], []
        for s in range(k):
            j = cell - 1 + s  # 0-based full index of the s-th spline alive on this cell
            c = np.zeros(n)
            c[j] = 1.0
            spl = BSpline(t, c, k - 1, extrapolate=False)
            vals.append(float(spl(x)))
            d2.append(float(spl.derivative(2)(x)))
        pts.append([cell, float(x), vals, d2])
out["bspline"] = {"k": k, "knots": t.tolist(), "points": pts}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "numeric_golden.json")
json.dump(out, open(path, "w"))
print(len(pts), path)
