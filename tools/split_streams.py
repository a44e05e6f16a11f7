"""Would cutting a batch into sub-batches on their own streams hide the latency-bound stages under the dense stage of
the other sub-batch?  Graph-captured forward+backward of the DCP batch as 1 x 32, 2 x 16 and 4 x 8 pairs."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import rrl_b200
from tools import synth
name = sys.argv[1] if len(sys.argv) > 1 else "dcp"
B, nf, nl = {"dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "fmr": (128, 1024, 15000)}[name]
pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(8)]
idx = [i % 8 for i in range(B)]
t1, t2, ln = (torch.from_numpy(np.stack([pairs[i][k] for i in idx])).cuda() for k in ("tri1", "tri2", "lines"))


def work(parts, streams):
    outs = []
    cur = torch.cuda.current_stream()
    for s, (a, b) in zip(streams, parts):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            x = t1[a:b].detach().requires_grad_(True)
            loss = rrl_b200.intersected_line_loss(x, t2[a:b], ln[a:b])
            loss.sum().backward()
            outs.append((loss.detach(), x.grad))
    for s in streams:
        cur.wait_stream(s)
    return outs


_pr = torch.cuda.Stream.priority_range()
lo_pri, hi_pri = max(_pr), min(_pr)              # numerically lower = scheduled first
print("stream priority range", lo_pri, hi_pri, flush=True)
for S, prio in ((1, False), (2, False), (4, False), (2, True), (4, True), (6, True), (8, True)):
    parts = [(B * k // S, B * (k + 1) // S) for k in range(S)]
    # with priorities: sub-batch 0 highest, so that the small kernels of an earlier sub-batch are scheduled ahead of the dense
    # CTAs of a later one whenever a slot frees
    streams = [torch.cuda.Stream(priority=min(lo_pri, hi_pri + k) if prio else 0) for k in range(S)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            work(parts, streams)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = work(parts, streams)
    for _ in range(5):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(name, "sub-batches", S, "priorities" if prio else "no priorities", "ms/step %.4f" % (e0.elapsed_time(e1) / 100), "loss sum %.6f" % float(sum(o[0].sum() for o in outs)), flush=True)
