import json, sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import aasist_oracle as O
from tests import gpu_util as g
rng = np.random.default_rng(123)
bad = 0
m32 = g.native_model("AASIST", "fp32"); m16 = g.native_model("AASIST", "f16x3")
l32 = g.native_model("AASIST-L", "fp32"); l16 = g.native_model("AASIST-L", "f16x3")
for it in range(40):
    B = int(rng.integers(1, 9)); L = int(rng.integers(2400, 150000))
    x = O.speech_like(B, L, 1000 + it).to(g.DEV)
    for name, a, b in (("A", m32, m16), ("L", l32, l16)):
        _, o32 = a(x); _, o16 = b(x); torch.cuda.synchronize()
        err = (o16 - o32).abs().max().item()
        flag = "" if err <= 2e-4 and torch.isfinite(o16).all() else "  <<<<<< BAD"
        if flag: bad += 1
        print(name, B, L, f"{err:.2e}", flag)
print("bad:", bad)
