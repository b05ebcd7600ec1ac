"""Tiny GPU check of complex values (pairs of real device arrays): a few fused expressions against NumPy."""
import os, sys, time
t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "baseline", "_ref"), ROOT):
    sys.path.insert(0, p)
import numpy as np
import autograd_b200 as b200
from autograd_b200 import complex_array as CA
b200.install()
r = np.random.RandomState(0)
zh = r.randn(64, 33) + 1j * r.randn(64, 33)
wh = r.randn(64, 33) + 0.5j * r.randn(64, 33)
z, w = CA.as_complex(zh), CA.as_complex(wh)
def f(a, b):
    return (np.exp(a * 0.5) * np.sin(b) + np.sqrt(a) / (b + 2.0), np.tanh(a) + np.log(b) * np.arcsinh(a) + a ** 3 + np.arctanh(b * 0.3),
            np.dot(a, np.conj(b).T), np.abs(a) + np.angle(b), np.prod(a[:4], axis=0) + np.sum(b, axis=0))
worst = 0.0
for g, want in zip(f(z, w), f(zh, wh)):
    g = g.get()
    worst = max(worst, float(np.max(np.abs(g - want)) / np.max(np.abs(want))))
print(f"complex smoke on {b200.backend.get().name}: max rel err {worst:.3e} in {time.time() - t0:.1f} s")
assert worst < 1e-11
