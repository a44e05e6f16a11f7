"""Scratch measurement script (GPU box): FP32 peaks, dense-kernel variants, per-stage times."""
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import rrl_b200
from tools import synth

L = rrl_b200._native.lib()
out = {}
for mode in (0, 1):
    tf, ms = C.c_double(), C.c_double()
    assert L.rrl_measure_fp32_peak(mode, C.byref(tf), C.byref(ms)) == 0
    out["fp32_peak_mode%d_tflops" % mode] = tf.value
print(out, flush=True)


def make_batch(B, nf, nl, seed=1000, distinct=4):
    pairs = [synth.make_pair(seed + i, nf, nl) for i in range(min(B, distinct))]
    idx = [i % len(pairs) for i in range(B)]
    t1 = torch.from_numpy(np.stack([pairs[i]["tri1"] for i in idx])).cuda()
    t2 = torch.from_numpy(np.stack([pairs[i]["tri2"] for i in idx])).cuda()
    ln = torch.from_numpy(np.stack([pairs[i]["lines"] for i in idx])).cuda()
    return t1, t2, ln


def measure(B, nf, nl, tag):
    t1, t2, ln = make_batch(B, nf, nl)
    wsb = L.rrl_workspace_bytes(B, nf, nf, nl)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    res = {"B": B, "nf": nf, "nl": nl, "ws_MB": wsb / 2**20}
    for variant in (0, 1):
        L.rrl_debug_set_dense_variant(variant)
        md, mp = C.c_float(), C.c_float()
        rc = L.rrl_measure_dense(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), B, nf, nf, nl, ws.data_ptr(), wsb, 10,
                                 C.byref(md), C.byref(mp), None)
        assert rc == 0, rc
        tests = B * nl * 2.0 * nf
        res["dense_ms_v%d" % variant] = md.value
        res["prep_ms"] = mp.value
        res["tests_per_s_v%d" % variant] = tests / (md.value * 1e-3)
        res["alg_tflops_v%d" % variant] = 48 * tests / (md.value * 1e-3) / 1e12
        res["exec_tflops_v%d" % variant] = 14 * tests / 16 / (md.value * 1e-3) / 1e12
    L.rrl_debug_set_dense_variant(1)
    # whole forward+backward through the autograd op
    t1g = t1.clone().requires_grad_(True)
    for _ in range(3):
        loss, info = rrl_b200.intersected_line_loss(t1g, t2, ln, return_info=True)
        loss.sum().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        loss = rrl_b200.intersected_line_loss(t1g, t2, ln)
        loss.sum().backward()
    e1.record()
    torch.cuda.synchronize()
    res["fwd_bwd_ms"] = e0.elapsed_time(e1) / n
    res["pair_lines_per_s"] = B * nl / (res["fwd_bwd_ms"] * 1e-3)
    st = info.stats.cpu().numpy()
    res["selected_per_pair"] = float(st[:, 0].mean())
    res["cand_groups_per_pair"] = float((st[:, 3] + st[:, 4]).mean())
    res["cand_group_rate"] = float((st[:, 3] + st[:, 4]).sum() / (B * nl * 2.0 * nf / 64))
    res["band"] = int(st[:, 5].sum())
    out[tag] = res
    print(tag, json.dumps(res), flush=True)


measure(1, 1024, 20000, "demo")
measure(32, 1024, 15000, "dcp")
measure(64, 2048, 10000, "rpm")
measure(128, 1024, 15000, "fmr")
t = time.time()
measure(1, 500000, 100000, "large")
print("large wall", time.time() - t)
json.dump(out, open("gpurun_out/quick_perf.json", "w"), indent=1)
