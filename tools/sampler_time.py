"""Device time of the on-device line sampler: the C entry point in a tight loop (preallocated buffers) against the torch
wrapper, for the DCP batch (32 pairs x 15000 lines) and one pair x 100000 lines; acceptance per round for reference."""
import ctypes as C
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import rrl_b200
from tools import synth
L = rrl_b200._native.lib()
for name, (B, nf, nl, rs) in {"dcp": (32, 1024, 15000, 0.5), "dcp-wide-sphere": (32, 1024, 15000, 2.0),
                              "large-lines": (1, 20000, 100000, 0.5)}.items():
    pairs = [synth.make_pair(1000 + i, nf, 256, radius_scale=rs) for i in range(min(B, 4))]
    idx = [i % len(pairs) for i in range(B)]
    v1 = torch.from_numpy(np.stack([pairs[i]["tri1"][:, :3] for i in idx])).cuda().contiguous()
    v2 = torch.from_numpy(np.stack([pairs[i]["tri2"][:, :3] for i in idx])).cuda().contiguous()
    lo2, hi2 = v2.min(1)[0], v2.max(1)[0]
    rad = ((hi2 - lo2).norm(dim=1) * rs).contiguous()
    cen = v2.mean(1).contiguous()
    wsb = L.rrl_sampler_workspace_bytes(B, nl, 10)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    lines = torch.empty(B, nl, 6, device="cuda")
    filled = torch.empty(B, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def raw(off):
        rc = L.rrl_sample_lines(rad.data_ptr(), cen.data_ptr(), v1.data_ptr(), v2.data_ptr(), B, nf, nf, nl, 10, 11, off, None,
                                lines.data_ptr(), filled.data_ptr(), ws.data_ptr(), wsb, st)
        assert rc == 0, rc

    for fn, label in ((raw, "C entry"), (lambda o: rrl_b200.sample_lines(rad, cen, nl, v1, v2, seed=11, offset=o), "torch wrapper")):
        for o in range(3):
            fn(o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for o in range(50):
            fn(o + 3)
        e1.record()
        torch.cuda.synchronize()
        print(name, label, "ms/call %.4f" % (e0.elapsed_time(e1) / 50), "filled %.3f" % (filled.float().mean().item() / nl), flush=True)
    one = rrl_b200.sample_lines(rad, cen, nl, v1, v2, seed=11, offset=0, rounds=1)[1]
    print(name, "acceptance of one round %.3f" % (one.float().mean().item() / nl), flush=True)
