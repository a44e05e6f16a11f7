"""A/B of dense-stage variants in one process: python tools/ab_quick.py 'id=val,id=val' ... (first = baseline)"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import rrl_b200
from tools import synth

L = rrl_b200._native.lib()
CONFIGS = {"demo": (1, 1024, 20000), "dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "fmr": (128, 1024, 15000),
           "large": (1, 500000, 100000)}
DEFAULTS = {1: 0, 2: 16, 3: 32, 4: 0}
variants = sys.argv[1:] or [""]
import os
for name, (B, nf, nl) in CONFIGS.items():
    if os.environ.get("ONLY") and name not in os.environ["ONLY"].split(","):
        continue
    pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(min(B, 4))]
    idx = [i % len(pairs) for i in range(B)]
    t1, t2, ln = (torch.from_numpy(np.stack([pairs[i][k] for i in idx])).cuda() for k in ("tri1", "tri2", "lines"))
    wsb = L.rrl_workspace_bytes(B, nf, nf, nl)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    res = {}
    for rep in range(2):
        for v in variants:
            for k, d in DEFAULTS.items():
                L.rrl_debug_set_param(k, d)
            for kv in filter(None, v.split(",")):
                k, val = kv.split("=")
                L.rrl_debug_set_param(int(k), int(val))
            md, mp = C.c_float(), C.c_float()
            assert L.rrl_measure_dense(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), B, nf, nf, nl, ws.data_ptr(), wsb, 10,
                                       C.byref(md), C.byref(mp), None) == 0
            res.setdefault(v or "default", []).append((round(md.value, 4), round(mp.value, 4)))
    print(name, json.dumps(res), flush=True)
