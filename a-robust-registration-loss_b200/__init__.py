"""B200-native intersected-line robust registration loss (hot path of Dengzhi-USTC/A-robust-registration-loss).

The directory name carries a hyphen, so import it through the repo-root shim:  `import rrl_b200`.

    rrl_b200.intersected_line_loss(tri1, tri2, lines)      batched native API, per-pair losses (B,)
    rrl_b200.twist_loss(twist, raw_tri1, tri2, lines)      the registration step in one op (gradient to the twist only)
    rrl_b200.se3_apply / se3_exp / rigid_apply             fused se(3) exponential + transform
    rrl_b200.sample_lines                                  on-device line sampler (Philox4x32-10)
    rrl_b200.chamfer                                       monitoring metric
    rrl_b200.loss                                          drop-in module with the reference's names (code/loss.py)
    rrl_b200.dist                                          batch-/line-sharded multi-GPU evaluation
    rrl_b200.prep                                          farthest-point sampling + kNN triplets (Sample_neighs)
    rrl_b200.io                                            dataset file formats of the reference's DL loaders
    rrl_b200.hooks                                         batched DCP / RPM-Net / FMR hook losses, FMR's se3.Exp
"""
from . import _native
from ._native import NativeError, launch_count
from .ops import (LossInfo, LossSession, chamfer, intersected_line_loss, rigid_apply, sample_lines, se3_apply, se3_exp, se3_Exp, twist_loss)
from . import loss  # noqa: E402  (reference-compatible names)
from . import dist  # noqa: E402
from . import prep  # noqa: E402
from . import io  # noqa: E402
from . import hooks  # noqa: E402

__all__ = ["NativeError", "launch_count", "LossInfo", "chamfer", "intersected_line_loss", "rigid_apply",
           "sample_lines", "se3_apply", "se3_exp", "se3_Exp", "twist_loss", "loss", "dist", "prep", "io", "hooks"]
