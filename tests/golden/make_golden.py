"""Generates tests/golden/ref_math_kat.json from the REFERENCE's own code.

Run in the dev container only (needs /root/reference):
    make -C oracle ref_math && python tests/golden/make_golden.py
oracle/_ref/libref_math.so is math_meso.h:12-24,143-505 compiled host-side (see
oracle/Makefile, oracle/ref_math_harness.cpp); nothing here comes from our own
restatement, so the file pins the oracle to the reference.  fp64 results are
stored as raw bit patterns.
"""
import ctypes as C
import json
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402

R = oracle.ref_math()
assert R is not None, "build oracle/_ref/libref_math.so first"
rng = np.random.Generator(np.random.PCG64(20140901))


def d2h(x):
    return "%016x" % struct.unpack("<Q", struct.pack("<d", x))[0]


def f2h(x):
    return "%08x" % struct.unpack("<I", struct.pack("<f", x))[0]


out = {"source": "UM/math_meso.h:12-24,143-505 compiled with g++ (oracle/ref_math_harness.cpp)"}
u32 = C.c_uint32
tea = []
for rounds in (1, 4, 8, 16, 64):
    for _ in range(16):
        a, b = (int(v) for v in rng.integers(0, 2 ** 32, 2))
        x, y = u32(a), u32(b)
        R.ref_tea(C.byref(x), C.byref(y), rounds)
        tea.append([rounds, a, b, x.value, y.value])
# the survey's hand-picked vectors
for rounds, a, b in ((1, 0, 0), (4, 1, 2), (8, 1, 0x4D6C3671), (16, 0x80000000, 0)):
    x, y = u32(a), u32(b)
    R.ref_tea(C.byref(x), C.byref(y), rounds)
    tea.append([rounds, a, b, x.value, y.value])
out["tea"] = tea
out["premix64_seed_now"] = [[419084618, t, R.ref_premix64(419084618, t)] for t in (0, 1, 2, 5, 100, 1000, 123456789)]
pm = []
for _ in range(32):
    a, b = (int(v) for v in rng.integers(0, 2 ** 32, 2))
    pm.append([a, b, R.ref_premix16(a, b)])
out["premix16"] = pm
il = []
for _ in range(32):
    i, j, k = (int(v) for v in rng.integers(0, 2048, 3))
    il.append([i, j, k, R.ref_interleave3(i, j, k)])
out["interleave3"] = il
mt = []
for _ in range(32):
    u, v, w = (float(np.float32(t)) for t in rng.normal(size=3))
    mt.append([f2h(u), f2h(v), f2h(w), R.ref_mantissa(u, v, w)])
for u, v, w in ((1.0, 0.5, -0.25), (0.1, 0.2, 0.3)):
    u, v, w = (float(np.float32(t)) for t in (u, v, w))
    mt.append([f2h(u), f2h(v), f2h(w), R.ref_mantissa(u, v, w)])
out["mantissa"] = mt
g = []
for _ in range(256):
    a, b = (int(v) for v in rng.integers(0, 2 ** 32, 2))
    g.append([a, b, d2h(R.ref_gaussian_dp(a, b)), f2h(R.ref_gaussian_sp(a, b))])
out["gaussian"] = g
m = {}
xs = rng.uniform(1e-3, 50.0, 64)
m["rsqrt"] = [[d2h(x), d2h(R.ref_rsqrt(x))] for x in xs]
m["sqrtd"] = [[d2h(x), d2h(R.ref_sqrtd(x))] for x in xs]
m["rcp"] = [[d2h(x), d2h(R.ref_rcp(x))] for x in xs]
xs = rng.uniform(1.0, 2.0, 64)
m["log2d_frac"] = [[d2h(x), d2h(R.ref_log2d_frac(x))] for x in xs]
xs = rng.uniform(0.0, 1.0, 64)
m["exp2d_frac"] = [[d2h(x), d2h(R.ref_exp2d_frac(x))] for x in xs]
m["sinpi"] = [[d2h(x), d2h(R.ref_sinpi(x))] for x in xs]
m["cospi"] = [[d2h(x), d2h(R.ref_cospi(x))] for x in xs]
pw = []
for _ in range(64):
    a, b = float(rng.uniform(1e-6, 1.0)), float(rng.choice([0.25, 0.5, 1.0, 2.0, float(rng.uniform(0, 3))]))
    pw.append([d2h(a), d2h(b), d2h(R.ref_powd(a, b))])
m["powd"] = pw
m["log2u"] = [[int(a), d2h(R.ref_log2u(int(a)))] for a in list(rng.integers(1, 2 ** 32, 60)) + [1, 2, 3, 2 ** 32 - 1]]
out["math"] = m
path = os.path.join(os.path.dirname(__file__), "ref_math_kat.json")
with open(path, "w") as f:
    json.dump(out, f, indent=0, separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes")
