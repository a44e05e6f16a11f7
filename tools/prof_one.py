"""One forward(+backward) of a named workload ON THE BENCH'S OWN INPUT SET 0, for ncu captures (so that the executed
instruction counts of the capture belong to the launch bench.py times):  python tools/prof_one.py dcp [reps]"""
import sys

import torch

sys.path.insert(0, ".")
import bench
import rrl_b200

name = sys.argv[1] if len(sys.argv) > 1 else "dcp"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t1, t2, ln = (torch.from_numpy(x).cuda() for x in bench.make_inputs(name, 0, 1)[0])
t1.requires_grad_(True)
for _ in range(reps):
    loss = rrl_b200.intersected_line_loss(t1, t2, ln)
    loss.sum().backward()
torch.cuda.synchronize()
print(name, float(loss.sum()))
