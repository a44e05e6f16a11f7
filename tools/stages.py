"""Hot per-stage device times of forward+backward for the named configs."""
import ctypes as C, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
import rrl_b200
from tools import synth
L = rrl_b200._native.lib()
CONFIGS = {"demo": (1, 1024, 20000), "dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "fmr": (128, 1024, 15000), "large": (1, 500000, 100000),
           "mid": (8, 8192, 10000), "big": (2, 65536, 20000), "large8": (1, 500000, 12500), "large2": (1, 500000, 50000)}
NAMES = ["prep", "sort", "node", "dense", "select", "build", "median", "welsch", "backward", "total"]
args = [a for a in sys.argv[1:] if "=" not in a]
for kv in [a for a in sys.argv[1:] if "=" in a]:          # e.g. 6=4 -> rrl_debug_set_param(6, 4)
    k, v = kv.split("=")
    L.rrl_debug_set_param(int(k), int(v))
for name in (args or ["dcp"]):
    B, nf, nl = CONFIGS[name]
    pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(min(B, 4))]
    idx = [i % len(pairs) for i in range(B)]
    t1, t2, ln = (torch.from_numpy(np.stack([pairs[i][k] for i in idx])).cuda() for k in ("tri1", "tri2", "lines"))
    wsb = L.rrl_workspace_bytes(B, nf, nf, nl)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    out = (C.c_float * 10)()
    rc = L.rrl_measure_stages(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), B, nf, nf, nl, ws.data_ptr(), wsb, 20, out, None)
    assert rc == 0, rc
    print(name, json.dumps({n: round(out[i] * 1e3, 1) for i, n in enumerate(NAMES)}), "(us)", flush=True)
