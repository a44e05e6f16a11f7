"""Drop-in replacement for the hot-path names of the reference's code/loss.py.

Put this module first on sys.path as `loss` or import it as `rrl_b200.loss`; the demo
(test_demo_optimized_Lie_Algebra.py:7-10) and the DCP / RPM-Net / FMR hooks (Train_DCP.py:30, Train_RPM.py:31,
fmr/model.py:20) use exactly these names and signatures:

    cal_loss_intersection_batch_whole_median_pts_lines(s_m, s_n, e_m, e_n, points1, points2, line, device)
    Random_uniform_distribution_lines_batch_efficient_resample(r, centers, N, vertices1, vertices2, device)
    Reconstruction_point(rotation=None, translation=None)
    chamfer_dist(points_x, points_y)
    Sample_neighs(points, num_sample=5000, num_neigh=3, device='cpu')

Documented deviations from the reference:
  * inputs must be CUDA tensors: there is no CPU path (the `device` argument is accepted and ignored);
  * B > 1 is evaluated per pair and summed (the reference's B > 1 behaviour is a bug nobody relies on);
  * bad ranks raise ValueError instead of print + exit(0); NaN never exits, it is reported in `last_info.status`;
  * the sampler draws from a counter-based Philox stream on the device (seed with `manual_seed`), not from
    torch's CPU generator; pass `uniforms=` to ops.sample_lines to replay recorded draws.
"""
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _native as N
from . import ops

# the reference returns the tuple (None, None, None) when no (k, j) combination is populated (loss.py:232).
# Reproducing that needs one device->host read of the status word; set to False to keep the call
# asynchronous (an empty pair then yields a zero loss with zero gradient).
STRICT_EMPTY_RETURN = True

last_info: Optional[ops.LossInfo] = None        # side channel: status / median / stats of the latest loss call

_sampler_state = {"seed": 0x5EED, "offset": 0}


def manual_seed(seed: int) -> None:
    """Seeds the on-device line sampler (counter-based: same seed + same call order => same lines)."""
    _sampler_state["seed"] = int(seed)
    _sampler_state["offset"] = 0


def cal_loss_intersection_batch_whole_median_pts_lines(s_m, s_n, e_m, e_n, points1, points2, line, device='cpu'):
    """loss.py:170-232.  points1 (B,nf1,9), points2 (B,nf2,9), line (B,nl,6) -> Tensor[1]."""
    global last_info
    if points1.dim() != 3 or points2.dim() != 3 or line.dim() != 3:
        raise ValueError("Input is wrong")            # the reference prints this and exit(0)s (loss.py:69-71)
    per_pair, info = ops.intersected_line_loss(points1, points2, line, (s_m, s_n, e_m, e_n), return_info=True)
    last_info = info
    if STRICT_EMPTY_RETURN and bool((info.status & N.STATUS_EMPTY).ne(0).all().item()):
        return None, None, None
    return per_pair.sum().reshape(1)


def Random_uniform_distribution_lines_batch_efficient_resample(r, centers, N_lines, vertices1, vertices2, device='cpu'):
    """loss.py:415-432.  r (B,1), centers (B,3), vertices (B,n,3) -> lines (B,N,6); unfilled rows are zero."""
    lines, _ = ops.sample_lines(r, centers, int(N_lines), vertices1, vertices2, seed=_sampler_state["seed"],
                                offset=_sampler_state["offset"])
    _sampler_state["offset"] += 1
    return lines


class Reconstruction_point(nn.Module):
    """loss.py:437-463: a 6-vector twist [w | v]; forward applies exp(twist) to the points (row vectors)."""

    def __init__(self, rotation=None, translation=None):
        super().__init__()
        if rotation is None or translation is None:
            tp = np.random.randn(3)
            tp = tp / np.linalg.norm(tp)
            tp_translation = np.random.randn(3) * 0.001
            init = torch.from_numpy(np.concatenate([0.001 * tp, tp_translation], 0).astype(np.float32))
        else:
            init = se3_log(rotation.reshape(3, 3), translation.reshape(3)) + torch.rand(6) * 0.6
        self.parameters_ = nn.Parameter(init)

    def Transform(self):
        """(R (1,3,3), T (1,3)) of the current twist (detached; the demo only saves them)."""
        return ops.se3_exp(self.parameters_.reshape(1, 6))

    def forward(self, points, points_neighbors):
        n = points.reshape(-1, 3).shape[0]
        both = torch.cat([points.reshape(1, -1, 3), points_neighbors.reshape(1, -1, 3)], 1)
        out = ops.se3_apply(self.parameters_.reshape(1, 6), both)
        return out[0, :n].reshape(-1, 3), out[0, n:].reshape(-1, 9)


def se3_log(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """Inverse of exp3 for the (rotation, translation) initialisation (LieAlgebra/se3.py:124-134, so3.log,
    so3.inv_vecs_Xg_ig).  Host-side, float64, runs once at construction."""
    R = R.detach().double().cpu()
    t = t.detach().double().cpu()
    cos = ((torch.trace(R) - 1) / 2).clamp(-1, 1)
    theta = torch.acos(cos)
    skew = torch.stack([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    if theta < 1e-6:
        w = skew
    else:
        w = skew * theta / torch.sin(theta)
    th = w.norm()
    W = torch.tensor([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=torch.float64)
    if th < 1e-6:
        Vinv = torch.eye(3, dtype=torch.float64) - 0.5 * W + W @ W / 12
    else:
        Vinv = (torch.eye(3, dtype=torch.float64) - 0.5 * W +
                (1 / th ** 2 - (1 + torch.cos(th)) / (2 * th * torch.sin(th))) * (W @ W))
    return torch.cat([w, Vinv @ t]).float()


def chamfer_dist(points_x, points_y):
    """loss.py:236-252 (monitoring metric; not differentiated by any caller)."""
    return ops.chamfer(points_x, points_y)


def Sample_neighs(points, num_sample=5000, num_neigh=3, device='cuda'):
    """loss.py:473-485: farthest-point sample + the num_neigh nearest neighbours of every sample, returned as the
    (3*nsample, 3) array the demo reshapes to (nf, 9).  Pre-processing, run once per cloud."""
    from . import prep
    return prep.sample_neighs(points, num_sample, num_neigh, device)
