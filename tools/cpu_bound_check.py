"""Is the DCP step CPU-launch bound?  wall time of the launch loop (no sync) vs device time; then the same under a CUDA graph."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import rrl_b200
from tools import synth
B, nf, nl = 32, 1024, 15000
pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(4)]
idx = [i % 4 for i in range(B)]
t1 = torch.from_numpy(np.stack([pairs[i]["tri1"] for i in idx])).cuda().requires_grad_(True)
t2 = torch.from_numpy(np.stack([pairs[i]["tri2"] for i in idx])).cuda()
ln = torch.from_numpy(np.stack([pairs[i]["lines"] for i in idx])).cuda()

def step():
    t1.grad = None
    loss = rrl_b200.intersected_line_loss(t1, t2, ln)
    loss.sum().backward()
    return loss

for _ in range(10): step()
torch.cuda.synchronize()
K = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K): step()
e1.record(); t_launch = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print("eager: cpu launch loop %.3f ms/step, wall incl sync %.3f ms/step, device events %.3f ms/step" % (t_launch / K * 1e3, t_all / K * 1e3, e0.elapsed_time(e1) / K))

# CUDA graph of forward + backward
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
t1.grad = None
with torch.cuda.graph(g):
    loss_static = rrl_b200.intersected_line_loss(t1, t2, ln)
    loss_static.sum().backward()
grad_static = t1.grad
for _ in range(5): g.replay()
torch.cuda.synchronize()
ref = step(); torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
print("graph loss equal:", torch.equal(ref.detach(), loss_static.detach()))
t0 = time.perf_counter(); e0.record()
for _ in range(K): g.replay()
e1.record(); t_launch = time.perf_counter() - t0
torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print("graph: cpu launch loop %.3f ms/step, wall incl sync %.3f ms/step, device events %.3f ms/step" % (t_launch / K * 1e3, t_all / K * 1e3, e0.elapsed_time(e1) / K))
