"""In-kernel phase timestamps (library built with RRL_MARKS=1): python tools/marks.py dcp"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import rrl_b200
from tools import synth
L = rrl_b200._native.lib()
CONFIGS = {"demo": (1, 1024, 20000), "dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "fmr": (128, 1024, 15000), "large": (1, 500000, 100000)}
for mode in (0, 1, 2):
    tf, ms = C.c_double(), C.c_double()
    L.rrl_measure_fp32_peak(mode, C.byref(tf), C.byref(ms))
    print("peak mode", mode, "(0 FFMA, 1 FFMA2, 2 DFMA):", round(tf.value, 2), "TFLOP/s", flush=True)
for name in (sys.argv[1:] or ["dcp"]):
    B, nf, nl = CONFIGS[name]
    pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(min(B, 4))]
    idx = [i % len(pairs) for i in range(B)]
    t1, t2, ln = (torch.from_numpy(np.stack([pairs[i][k] for i in idx])).cuda() for k in ("tri1", "tri2", "lines"))
    t1.requires_grad_(True)
    for _ in range(3):
        loss = rrl_b200.intersected_line_loss(t1, t2, ln)
        loss.sum().backward()
    torch.cuda.synchronize()
    m = (C.c_ulonglong * 32)()
    L.rrl_debug_read_marks.argtypes = [C.POINTER(C.c_ulonglong)]
    assert L.rrl_debug_read_marks(m) == 0
    v = [int(x) for x in m]
    print(name, "tail marks (us since mark 0):", [round((x - v[0]) / 1e3, 2) if x else None for x in v[:8]], flush=True)
    print(name, "build marks (us since mark 8):", [round((x - v[8]) / 1e3, 2) if x else None for x in v[8:14]], flush=True)
    print(name, "backward marks (us since mark 16):", [round((x - v[16]) / 1e3, 2) if x else None for x in v[16:20]], flush=True)
    if hasattr(L, "rrl_debug_read_cta_times"):
        ct = (C.c_ulonglong * (2 * 8192))()
        L.rrl_debug_read_cta_times.argtypes = [C.POINTER(C.c_ulonglong)]
        assert L.rrl_debug_read_cta_times(ct) == 0
        a = np.array(list(ct), dtype=np.int64).reshape(2, 8192)
        n = min(8192, B * ((nl + 511) // 512))
        st, en = a[0, :n], a[1, :n]
        t0 = st.min()
        print(name, "build CTAs %d: start min/median/max %.2f/%.2f/%.2f us, end min/median/max %.2f/%.2f/%.2f us, life median/max %.2f/%.2f us" %
              (n, 0.0, (np.median(st) - t0) / 1e3, (st.max() - t0) / 1e3, (en.min() - t0) / 1e3, (np.median(en) - t0) / 1e3, (en.max() - t0) / 1e3,
               np.median(en - st) / 1e3, (en - st).max() / 1e3), flush=True)
    if hasattr(L, "rrl_debug_read_marks_prep"):
        L.rrl_debug_read_marks_prep.argtypes = [C.POINTER(C.c_ulonglong)]
        assert L.rrl_debug_read_marks_prep(m) == 0
        v = [int(x) for x in m]
        print(name, "prep marks (us since mark 0; 1 zeroing, 2 thresholds, 3 line extent, 4 sort, 5 indices, 6 k-d refine, 7 node records):",
              [round((x - v[0]) / 1e3, 2) if x else None for x in v[:8]], flush=True)
