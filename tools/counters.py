"""Entries consumed by the queue levels of the dense kernel (build with RRL_DEFS=-DRRL_COUNTERS, load with RRL_LIB_PATH):
python tools/counters.py large [dcp ...]"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench, rrl_b200
L = rrl_b200._native.lib()
L.rrl_debug_read_counters.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
for name in sys.argv[1:] or ["large"]:
    t1, t2, ln = (torch.from_numpy(x).cuda() for x in bench.make_inputs(name, 0, 1)[0])
    out = (C.c_ulonglong * 8)()
    L.rrl_debug_read_counters(out, 1)
    loss, info = rrl_b200.intersected_line_loss(t1, t2, ln, return_info=True)
    torch.cuda.synchronize()
    L.rrl_debug_read_counters(out, 1)
    B, nl = ln.shape[0], ln.shape[1]
    per = B * nl * 2.0
    print(name, "per (line, cloud): fired main-loop records %.1f | (line,group) entries %.1f | (line,node) entries %.1f | point-0 passes %.2f | "
          "to exact %.3f | hits %.3f" % ((float(info.stats[:, 3].sum()) + float(info.stats[:, 4].sum())) / per, out[0] / per, out[1] / per,
                                          out[2] / per, out[3] / per, float(info.hits(1)[0].sum() + info.hits(2)[0].sum()) / per), flush=True)
