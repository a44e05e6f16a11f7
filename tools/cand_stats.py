"""Candidate volumes of the dense stage per (line, cloud): node candidates, exact-test candidates, hits."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import rrl_b200
from tools import synth
CONFIGS = {"demo": (1, 1024, 20000), "dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "large": (1, 500000, 100000)}
for name in (sys.argv[1:] or ["dcp"]):
    B, nf, nl = CONFIGS[name]
    pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(min(B, 4))]
    idx = [i % len(pairs) for i in range(B)]
    t1, t2, ln = (torch.from_numpy(np.stack([pairs[i][k] for i in idx])).cuda() for k in ("tri1", "tri2", "lines"))
    loss, info = rrl_b200.intersected_line_loss(t1, t2, ln, return_info=True)
    st = info.stats.cpu().numpy().astype(np.float64)
    c1, h1 = info.hits(1)
    c2, h2 = info.hits(2)
    print(name, "node candidates per (line, cloud): %.2f / %.2f of %d nodes; hits per line %.3f / %.3f; selected lines per pair %.0f; band %d" %
          (st[:, 3].mean() / nl, st[:, 4].mean() / nl, nf // (8 if nf < 16384 else 16), c1.float().mean().item(), c2.float().mean().item(),
           st[:, 0].mean(), int(st[:, 5].sum())), flush=True)
