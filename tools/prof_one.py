"""One forward(+backward) of a named config, for ncu captures:  python tools/prof_one.py dcp [reps]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import rrl_b200
from oracle import synth

CONFIGS = {"demo": (1, 1024, 20000), "dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "fmr": (128, 1024, 15000),
           "large": (1, 500000, 100000)}
name = sys.argv[1] if len(sys.argv) > 1 else "dcp"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, nf, nl = CONFIGS[name]
pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(min(B, 4))]
idx = [i % len(pairs) for i in range(B)]
t1 = torch.from_numpy(np.stack([pairs[i]["tri1"] for i in idx])).cuda().requires_grad_(True)
t2 = torch.from_numpy(np.stack([pairs[i]["tri2"] for i in idx])).cuda()
ln = torch.from_numpy(np.stack([pairs[i]["lines"] for i in idx])).cuda()
for _ in range(reps):
    loss = rrl_b200.intersected_line_loss(t1, t2, ln)
    loss.sum().backward()
torch.cuda.synchronize()
print(name, float(loss.sum()))
