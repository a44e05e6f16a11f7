#!/usr/bin/env python
"""The reference's single-pair demo (code/test_demo_optimized_Lie_Algebra.py) on the B200 path.

Same flow, same five names imported from `loss` -- only the module behind them changes ("shadow the
module").  OBJ IO goes through rrl_b200.io instead of igl (not installed here); TensorBoard logging is dropped.

    python examples/demo_lie_algebra.py --data_path /path/to/sample_data/challenge_data --label1 0 --n_epoch 200
    python examples/demo_lie_algebra.py --synthetic 1024 --n_epoch 200          # no files needed
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rrl_b200  # noqa: E402

sys.modules["loss"] = rrl_b200.loss                 # every `from loss import ...` below resolves to the B200 module
from loss import cal_loss_intersection_batch_whole_median_pts_lines  # noqa: E402
from loss import Reconstruction_point  # noqa: E402
from loss import Random_uniform_distribution_lines_batch_efficient_resample  # noqa: E402
from loss import chamfer_dist, Sample_neighs  # noqa: E402


def adjust_learning_rate(optimizer, epoch, lr):
    if epoch % 1000 == 0:                            # test_demo_optimized_Lie_Algebra.py:15-21
        lr *= 0.5
    for group in optimizer.param_groups:
        group["lr"] = lr


def bounding_box(v):
    lo, hi = v.min(0), v.max(0)                      # igl.bounding_box corner order: first = max corner, last = min corner
    return np.array([[(lo if (q >> a) & 1 else hi)[a] for a in range(3)] for q in range(8)], np.float32)


def test_one_case(data, n_epoch=1000, n_sample_line=20000, device="cuda", log=print):
    """test_demo_optimized_Lie_Algebra.py:27-100 without the file output; returns (module, per-epoch (chamfer, loss))"""
    bbox = data["bounding_box"]
    v1_in, v2 = data["vertics1_tensor"], data["vertics2_tensor"]
    f1_in, f2 = data["vertics1_faces_tensor"], data["vertics2_faces_tensor"]
    centers = data["centers"]
    model = Reconstruction_point().to(device)
    optimize = torch.optim.Adam(model.parameters(), lr=2e-2)
    R = (bbox[0, :] - bbox[-1, :]).norm(p=2).to(device)
    v1 = v1_in
    history = []
    for epoch in range(n_epoch):
        lines = Random_uniform_distribution_lines_batch_efficient_resample(
            torch.FloatTensor([R]).reshape(1, 1).to(device), centers.reshape(1, -1).to(device), n_sample_line,
            v1.view(1, -1, 3).to(device), v2.view(1, -1, 3).to(device), device).detach().view(-1, 6)
        adjust_learning_rate(optimize, epoch, optimize.param_groups[0]["lr"])
        v1, f1 = model(v1_in, f1_in)
        loss_di = cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, f1.reshape(1, -1, 9), f2.reshape(1, -1, 9),
                                                                     lines.reshape(1, -1, 6), device)
        if loss_di is not None and not isinstance(loss_di, tuple):
            optimize.zero_grad()
            loss_di.backward()
            optimize.step()
            loss_cf = chamfer_dist(v1.reshape(-1, v1.shape[0], 3), v2.reshape(-1, v2.shape[0], 3))
            history.append((float(loss_cf.detach()), float(loss_di.detach())))
            if log and epoch % 10 == 0:
                log("epoch %4d  chamfer %.6f  loss_intersection %.6f" % (epoch, history[-1][0], history[-1][1]))
    return model, history


def load_case(args, device):
    if args.synthetic:
        rng = np.random.default_rng(args.seed)
        u = rng.standard_normal((args.synthetic, 3))
        base = (u / np.linalg.norm(u, axis=1, keepdims=True) * np.array([1.0, 0.7, 0.5])).astype(np.float32)
        ang = np.deg2rad(args.angle)
        Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
        v1 = base.astype(np.float32)
        v2 = (base @ Rz.T + np.array([0.05, -0.03, 0.02], np.float32)).astype(np.float32)
    else:
        v1 = rrl_b200.io.read_obj_vertices(os.path.join(args.data_path, args.label1 + "_src_sample.obj")).astype(np.float32)
        v2 = rrl_b200.io.read_obj_vertices(os.path.join(args.data_path, args.label1 + "_tar_sample.obj")).astype(np.float32)
    n1, n2 = Sample_neighs(v1, device=device), Sample_neighs(v2, device=device)
    c1, c2 = v1.mean(0)[None], v2.mean(0)[None]
    v1, v2, n1, n2 = v1 - c1, v2 - c2, n1 - c1, n2 - c2
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(device)
    return {"bounding_box": t(bounding_box(v2)), "vertics1_tensor": t(v1), "vertics2_tensor": t(v2),
            "vertics1_faces_tensor": t(n1).reshape(1, -1, 3), "vertics2_faces_tensor": t(n2), "centers": t(v2).mean(0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data_path", default="./sample_data/challenge_data")
    ap.add_argument("--label1", default="0")
    ap.add_argument("--n_epoch", type=int, default=1000)
    ap.add_argument("--n_sample_line", type=int, default=20000)
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--synthetic", type=int, default=0, help="use a synthetic ellipsoid cloud of this many points")
    ap.add_argument("--angle", type=float, default=20.0, help="rotation of the synthetic target (degrees)")
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    rrl_b200.loss.manual_seed(args.seed)
    data = load_case(args, "cuda")
    model, hist = test_one_case(data, args.n_epoch, args.n_sample_line, "cuda")
    R, T = model.Transform()
    print("final chamfer %.6f (start %.6f), loss %.6f" % (hist[-1][0], hist[0][0], hist[-1][1]))
    print("R =\n", R.detach().cpu().numpy()[0], "\nT =", T.detach().cpu().numpy()[0])


if __name__ == "__main__":
    main()
