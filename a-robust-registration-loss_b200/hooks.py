"""Batched forms of the three training hooks of the reference.

Each function replaces the per-pair Python loop of one hook by ONE native evaluation per predicted transform, with the
hook's own scaling, and takes the UNTRANSFORMED source triplets together with the predicted transform, so that the
backward delivers dL/dR, dL/dt (12 floats per pair) through the fused rigid transform instead of a dense (B, nf, 9)
point gradient that autograd would then push through a matmul:

    dcp_loss   Train_DCP.py:252-297   sum_j loss_j / 5.0 / batch_size,  points = R p + t  (utils.py:32-37)
    rpm_loss   Train_RPM.py:218-258   sum_i 0.5^(n-i-1) (sum_j loss_j) / n,  points = a R^T + t  (rpm/common/math_torch/se3.py:55-82)
    fmr_loss   fmr/model.py:285-313   sum_{i in last 3} 0.5^(maxiter-i-1) sum_j loss_j / 5.0 / batch_size,
                                      points = R a + p of g_series[i]  (fmr/se_math/se3.py:110-124)

The unchanged hooks also run fast through `rrl_b200.loss` (its slice batching); these helpers additionally fuse the
transform and are what `bench.py --api hooks` times.  `Exp` is FMR's se3.Exp with the reference's ExpMap gradient
(fmr/se_math/se3.py:133-165).
"""
from typing import Optional, Sequence

import torch

from . import ops

Exp = ops.se3_Exp


def _faces(points_rows: torch.Tensor) -> torch.Tensor:
    """(B, 3nf, 3) rows [p0; p1; p2; p0; ...] -> (B, nf, 9)"""
    return points_rows.reshape(points_rows.shape[0], -1, 9)


def _pair_sum(src_rows, R, t, tar_faces, lines, window):
    moved = ops.rigid_apply(R, t, src_rows)                     # R p + t, differentiable w.r.t. R, t (and the points)
    return ops.intersected_line_loss(_faces(moved), tar_faces, lines, window).sum()


def dcp_loss(points_based_neighs_src: torch.Tensor, rotation_ab_pred: torch.Tensor, translation_ab_pred: torch.Tensor,
             points_based_neighs_tar: torch.Tensor, lines: torch.Tensor, window=(1, 1, 5, 5)) -> torch.Tensor:
    """Train_DCP.py:252-297.  points_based_neighs_* (B, 3, 3nf) as the DCP loader stores them (channels first),
    rotation (B,3,3), translation (B,3), lines (B,nl,6) -> `tp_loss_intersection / batch_size`, shape (1,)."""
    B = points_based_neighs_src.shape[0]
    src_rows = points_based_neighs_src.transpose(2, 1).contiguous()            # (B, 3nf, 3)
    tar_faces = points_based_neighs_tar.transpose(2, 1).reshape(B, -1, 9)      # Train_DCP.py:241-242
    total = _pair_sum(src_rows, rotation_ab_pred, translation_ab_pred, tar_faces, lines, window)
    return (total / 5.0 / B).reshape(1)


def rpm_loss(pred_transforms: Sequence[torch.Tensor], points_based_neighs_src: torch.Tensor,
             points_based_neighs_tar: torch.Tensor, lines: torch.Tensor, window=(1, 1, 5, 5),
             discount_factor: float = 0.5) -> torch.Tensor:
    """Train_RPM.py:218-258.  pred_transforms: num_iter tensors (B,3,4) [R | t]; points_based_neighs_* (B, 3nf, 3);
    -> losses['loss_intersection'], shape (1,)."""
    B = points_based_neighs_src.shape[0]
    n = len(pred_transforms)
    tar_faces = points_based_neighs_tar.reshape(B, -1, 9)
    src_rows = points_based_neighs_src.contiguous()
    total = None
    for i, g in enumerate(pred_transforms):
        li = _pair_sum(src_rows, g[:, :3, :3], g[:, :3, 3], tar_faces, lines, window) / n
        li = li * discount_factor ** (n - i - 1)
        total = li if total is None else total + li
    return total.reshape(1)


def fmr_loss(g_series: torch.Tensor, points_based_neighs_src: torch.Tensor, points_based_neighs_tar: torch.Tensor,
             lines: torch.Tensor, maxiter: Optional[int] = None, last: int = 3, window=(1, 1, 5, 5)) -> torch.Tensor:
    """fmr/model.py:285-313.  g_series (maxiter + 1, B, 4, 4) = the IC iterates (`self.g_series_gpu`);
    points_based_neighs_* (B, 3nf, 3) -> `loss_intersection / batch_size`, shape (1,)."""
    if maxiter is None:
        maxiter = g_series.shape[0] - 1
    B = points_based_neighs_src.shape[0]
    tar_faces = points_based_neighs_tar.reshape(B, -1, 9)
    src_rows = points_based_neighs_src.contiguous()
    total = None
    for i in range(maxiter - last, maxiter):
        g = g_series[i]
        li = _pair_sum(src_rows, g[:, :3, :3], g[:, :3, 3], tar_faces, lines, window) / 5.0
        li = li * 0.5 ** (maxiter - i - 1)
        total = li if total is None else total + li
    return (total / B).reshape(1)


def fmr_twist_loss(twist: torch.Tensor, points_based_neighs_src: torch.Tensor, points_based_neighs_tar: torch.Tensor,
                   lines: torch.Tensor, window=(1, 1, 5, 5)) -> torch.Tensor:
    """BASELINE configs[3] as SURVEY 8(d) states it: gradient to the twist (B,6) through FMR's Exp.  One IC iterate
    g = Exp(twist); per-pair losses (B,)."""
    g = Exp(twist)
    moved = ops.rigid_apply(g[:, :3, :3], g[:, :3, 3], points_based_neighs_src.contiguous())
    return ops.intersected_line_loss(_faces(moved), points_based_neighs_tar.reshape(twist.shape[0], -1, 9), lines, window)
