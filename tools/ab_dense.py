"""A/B sweep of dense-stage launch knobs inside ONE process on ONE GPU (boxes differ between gpurun calls)."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import rrl_b200
from tools import synth

L = rrl_b200._native.lib()
WAVES = (16, 32, 64, 128)
CHUNKS = (16, 32)
CONFIGS = {"fmr": (128, 1024, 15000), "demo": (1, 1024, 20000), "dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "large": (1, 500000, 100000)}


def batch(B, nf, nl):
    pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(min(B, 4))]
    idx = [i % len(pairs) for i in range(B)]
    return tuple(torch.from_numpy(np.stack([pairs[i][k] for i in idx])).cuda() for k in ("tri1", "tri2", "lines"))


def dense_ms(t1, t2, ln, B, nf, nl, iters=10):
    wsb = L.rrl_workspace_bytes(B, nf, nf, nl)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    md, mp = C.c_float(), C.c_float()
    assert L.rrl_measure_dense(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), B, nf, nf, nl, ws.data_ptr(), wsb, iters,
                               C.byref(md), C.byref(mp), None) == 0
    return md.value, mp.value


out = {}
for name in (sys.argv[1:] or ["dcp", "large", "rpm", "demo"]):
    B, nf, nl = CONFIGS[name]
    t1, t2, ln = batch(B, nf, nl)
    res = {}
    for G in (8, 16):
        for waves in WAVES:
            for minch in CHUNKS:
                L.rrl_debug_set_param(1, G); L.rrl_debug_set_param(2, waves); L.rrl_debug_set_param(3, minch)
                d, p = dense_ms(t1, t2, ln, B, nf, nl, 5)
                res["G%d_w%d_c%d" % (G, waves, minch)] = round(d, 4)
    L.rrl_debug_set_param(1, 0); L.rrl_debug_set_param(2, 8); L.rrl_debug_set_param(3, 64)
    d, p = dense_ms(t1, t2, ln, B, nf, nl, 10)
    res["default"] = round(d, 4); res["prep"] = round(p, 4)
    out[name] = res
    print(name, json.dumps(res), flush=True)
json.dump(out, open("gpurun_out/ab_dense.json", "w"), indent=1)
