import torch, time, numpy as np
from aasist_b200 import evaluation as V
from oracle import evaluation_oracle as E
bona, spoof, tar, non, spf = E.make_case(*E.CASES[5][1:])
b=torch.from_numpy(bona).cuda().double(); s=torch.from_numpy(spoof).cuda().double()
for _ in range(3): V.compute_eer(b,s)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10): V.compute_eer(b,s)
torch.cuda.synchronize(); t1=time.perf_counter()
print("device eer 71237 scores: %.3f ms" % ((t1-t0)/10*1e3))
