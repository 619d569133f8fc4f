"""Two forwards of one model/batch (the second one is what ncu captures): python tools/one_forward.py [model] [B] [L] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aasist_b200

name = sys.argv[1] if len(sys.argv) > 1 else "AASIST"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
L = int(sys.argv[3]) if len(sys.argv) > 3 else 64600
prec = sys.argv[4] if len(sys.argv) > 4 else None
dev = torch.device("cuda:0")
cls = {"RawGAT-ST": aasist_b200.RawGATSTModel, "AASIST-Robust": aasist_b200.RobustModel}.get(name, aasist_b200.Model)
m = cls(aasist_b200.CONFIGS[name], precision=prec)
m.load_state_dict(torch.load(aasist_b200.weights_path(name), map_location="cpu"))
m = m.to(dev).eval()
x = 0.05 * torch.randn(B, L, device=dev)
with torch.no_grad():
    for _ in range(2):
        out = m(x)[1]
torch.cuda.synchronize()
print("ok", out[0].tolist(), "launches", m.launch_count())
