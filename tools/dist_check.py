"""Multi-GPU parity check (run under torchrun on >= 2 GPUs): batch shard and line shard -- through the NCCL protocol AND
through the peer-memory exchange inside the kernels (PeerComm) -- against the C oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import rrl_b200
from oracle import c_oracle as co
from tools import synth

dev = torch.device("cuda", local)
# ---- line shard: one pair, lines split over the ranks ----
p = synth.make_pair(4242, 3000, 6001)
lo, hi = rrl_b200.dist.shard_range(6001, rank, world)
t1 = torch.from_numpy(p["tri1"]).to(dev).requires_grad_(True)
t2 = torch.from_numpy(p["tri2"]).to(dev)
loss, status, med = rrl_b200.dist.line_sharded_loss(t1, t2, torch.from_numpy(p["lines"][lo:hi]).to(dev))
loss.sum().backward()
orc = co.loss(p["tri1"], p["tri2"], p["lines"])
assert float(med) == orc.median, (float(med), orc.median)
assert abs(float(loss) - orc.loss) <= 1e-5 * orc.loss, (float(loss), orc.loss)
g = t1.grad.cpu().numpy()
err = np.linalg.norm(g - orc.grad1) / np.linalg.norm(orc.grad1)
assert err <= 1e-5, err
# ---- the same through the peer-memory exchange (CUDA IPC over NVLink): run repeatedly, also inside a CUDA graph ----
comm = rrl_b200.dist.PeerComm.create(6001 // world + 1)
peer = "unavailable (fell back to NCCL)"
if comm is not None:
    ln_local = torch.from_numpy(p["lines"][lo:hi]).to(dev)
    for rep in range(3):
        t1p = torch.from_numpy(p["tri1"]).to(dev).requires_grad_(True)
        lp, sp, mp = rrl_b200.dist.line_sharded_loss(t1p, t2, ln_local, comm=comm)
        lp.sum().backward()
        assert float(mp) == orc.median and float(lp) == float(loss) and int(sp) == 0, (rep, float(mp), float(lp), int(sp))
        errp = np.linalg.norm(t1p.grad.cpu().numpy() - orc.grad1) / np.linalg.norm(orc.grad1)
        assert errp <= 1e-5, errp
    twp = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.02, 0.01], device=dev, requires_grad=True)
    raw_p = torch.from_numpy(p["tri1"]).to(dev)
    ltp, _, _ = rrl_b200.dist.line_sharded_twist_loss(twp, raw_p, t2, ln_local, comm=comm)
    ltp.sum().backward()
    # graph capture of the whole step, replayed: the exchange's sequence numbers live on the device
    side = torch.cuda.Stream(dev)
    static_tw = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.02, 0.01], device=dev)
    def stepfn():
        tw_ = static_tw.clone().requires_grad_(True)
        l_, _, _ = rrl_b200.dist.line_sharded_twist_loss(tw_, raw_p, t2, ln_local, comm=comm)
        l_.sum().backward()
        return l_.detach(), tw_.grad
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3):
            stepfn()
    torch.cuda.current_stream(dev).wait_stream(side)
    dist.barrier(); torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        outs = stepfn()
    for _ in range(5):
        gph.replay()
    torch.cuda.synchronize()
    assert float(outs[0]) == float(ltp) and torch.equal(outs[1], twp.grad), (float(outs[0]), float(ltp))
    assert comm.error() == 0
    peer = "ok (loss bit-identical to the NCCL protocol, graph replay bit-identical to eager)"
    peer_grad = twp.grad.clone()
# ---- line shard through the se(3) twist: only 6 gradient floats are exchanged ----
tw = torch.tensor([0.02, -0.01, 0.03, 0.05, -0.02, 0.01], device=dev, requires_grad=True)
raw = torch.from_numpy(p["tri1"]).to(dev)
loss_t, _, _ = rrl_b200.dist.line_sharded_twist_loss(tw, raw, t2, torch.from_numpy(p["lines"][lo:hi]).to(dev))
loss_t.sum().backward()
tri1_moved = rrl_b200.se3_apply(tw.detach().reshape(1, 6), raw.reshape(1, -1, 3)).reshape(-1, 9).cpu().numpy()
orc_t = co.loss(tri1_moved, p["tri2"], p["lines"])
gt = co.se3_backward(tw.detach().cpu().numpy(), p["tri1"].reshape(-1, 3), orc_t.grad1.reshape(-1, 3))
err_t = np.linalg.norm(tw.grad.cpu().numpy() - gt) / np.linalg.norm(gt)
assert abs(float(loss_t) - orc_t.loss) <= 1e-5 * orc_t.loss and err_t <= 1e-5, (float(loss_t), orc_t.loss, err_t)
if comm is not None:
    err_p = np.linalg.norm(peer_grad.cpu().numpy() - gt) / np.linalg.norm(gt)
    assert float(ltp) == float(loss_t) and err_p <= 1e-5, (float(ltp), float(loss_t), err_p)
    comm.close()
# ---- batch shard: each rank its own pairs ----
pairs = [synth.make_pair(5000 + 10 * rank + i, 512, 2000) for i in range(3)]
a = torch.from_numpy(np.stack([q["tri1"] for q in pairs])).to(dev).requires_grad_(True)
b = torch.from_numpy(np.stack([q["tri2"] for q in pairs])).to(dev)
c = torch.from_numpy(np.stack([q["lines"] for q in pairs])).to(dev)
local_losses, total = rrl_b200.dist.batch_sharded_loss(a, b, c)
total.backward()
want_local = [co.loss(q["tri1"], q["tri2"], q["lines"]) for q in pairs]
mine = torch.tensor([sum(w.loss for w in want_local)], dtype=torch.float64, device=dev)
dist.all_reduce(mine)
assert abs(float(total) - float(mine)) <= 1e-5 * float(mine), (float(total), float(mine))
for i, w in enumerate(want_local):
    assert np.linalg.norm(a.grad[i].cpu().numpy() - w.grad1) <= 1e-5 * np.linalg.norm(w.grad1)
dist.barrier()
if rank == 0:
    print("dist_check ok: world %d, line-shard loss %.6f (oracle %.6f), grad err %.2e, twist-grad err %.2e, batch total %.6f; peer exchange: %s" %
          (world, float(loss), orc.loss, err, err_t, float(total), peer))
dist.destroy_process_group()
