"""torch-facing operators: thin autograd.Function wrappers that hand raw device pointers and the current CUDA
stream to the C ABI (include/rrl_b200.h).  torch is plumbing here (memory, streams, autograd graph); every
number is produced by the sm_100a kernels."""
from typing import Optional, Tuple

import torch

from . import _native as N


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _on(t: torch.Tensor):
    """Context that makes t's device current for the native call: the C entry points launch on the current device with
    the stream they are handed, so a tensor on cuda:1 while cuda:0 is current would otherwise meet a foreign stream."""
    return torch.cuda.device(t.device)


def _same_device(*ts):
    dev = ts[0].device
    for t in ts[1:]:
        if t is not None and t.device != dev:
            raise ValueError("all inputs must live on one device, got %s and %s" % (dev, t.device))


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise N.NativeError("%s must live on a CUDA device: the B200 kernels are the only implementation "
                            "(no CPU fallback)" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class LossInfo:
    """Side channel of one forward call (everything stays on the device; reading it synchronises)."""

    def __init__(self, status, median, stats, workspace, geom):
        self.status = status            # (B,) int32 RRL_STATUS_* bits
        self.median = median            # (B,) float32
        self.stats = stats              # (B, 8) int64, see include/rrl_b200.h
        self._ws = workspace
        self._geom = geom

    @property
    def n_selected_lines(self):
        return self.stats[:, 0]

    @property
    def n_entries(self):
        return self.stats[:, 1]

    @property
    def band_tests(self):
        """tests whose distance lies within 1 ulp of the threshold (reported separately, see north star)"""
        return self.stats[:, 5]

    def hits(self, cloud: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(counts (B,nl) int32, hits (B,nl,5) int32 ascending, -1 padded) of cloud 1 or 2."""
        B, nf1, nf2, nl = self._geom
        counts = torch.empty(B, nl, dtype=torch.int32, device=self._ws.device)
        hits = torch.empty(B, nl, N.HIT_CAP, dtype=torch.int32, device=self._ws.device)
        with _on(self._ws):
            N.check(N.lib().rrl_loss_export_hits(self._ws.data_ptr(), self._ws.numel(), B, nf1, nf2, nl, cloud,
                                                 counts.data_ptr(), hits.data_ptr(), _stream(self._ws)), "rrl_loss_export_hits")
        return counts, hits


class LossSession:
    """Keeps the workspace of ONE geometry alive across calls, so that the spatial order of both clouds computed by
    one forward is reused by the next (RRL_REUSE_ORDER, include/rrl_b200.h): the steps of a registration loop, the
    iterations of RPM-Net / FMR on one batch.  Any order gives the same results; a cloud that moved rigidly keeps a good
    one.  The backward of a call must run before the next forward on the same session (they share the workspace).
    `reset()` forces a fresh sort (e.g. when the clouds were replaced by different ones)."""

    def __init__(self, static_target: bool = False):
        """static_target: cloud 2 (the registration target) is the SAME tensor contents in every call of this session --
        its thresholds / records / bounding spheres are then kept too (RRL_REUSE_TARGET; clouds above 4096 triplets)."""
        self._ws, self._key, self._ready = None, None, False
        self._flags = N.REUSE_ORDER | (N.REUSE_TARGET if static_target else 0)

    def reset(self):
        self._ready = False

    def acquire(self, geom, dev, nbytes):
        key = (tuple(geom), str(dev), int(nbytes))
        if self._ws is None or self._key != key:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._key, self._ready = key, False
        flags = self._flags if self._ready else 0
        self._ready = True                      # the forward about to be queued leaves a complete order behind
        return self._ws, flags


class _IntersectedLineLoss(torch.autograd.Function):
    """loss[b] = reference loss of pair b (loss.py:170-232 on the B=1 slice); lines carry no gradient."""

    @staticmethod
    def forward(ctx, tri1, tri2, lines, window, holder, session=None):
        B, nf1, _ = tri1.shape
        nf2, nl = tri2.shape[1], lines.shape[1]
        dev = tri1.device
        L = N.lib()
        ws_bytes = L.rrl_workspace_bytes(B, nf1, nf2, nl)
        if ws_bytes == 0:
            raise ValueError("unsupported geometry B=%d nf1=%d nf2=%d nl=%d" % (B, nf1, nf2, nl))
        if session is not None:
            ws, flags = session.acquire((B, nf1, nf2, nl), dev, ws_bytes)
        else:
            ws, flags = torch.empty(ws_bytes, dtype=torch.uint8, device=dev), 0
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        median = torch.empty(B, dtype=torch.float32, device=dev)
        stats = torch.empty(B, N.NSTAT, dtype=torch.int64, device=dev)
        with _on(tri1):
            N.check(L.rrl_loss_forward_ex(tri1.data_ptr(), tri2.data_ptr(), lines.data_ptr(), B, nf1, nf2, nl,
                                          window[0], window[1], window[2], window[3], ws.data_ptr(), ws_bytes,
                                          loss.data_ptr(), status.data_ptr(), median.data_ptr(), stats.data_ptr(), flags,
                                          _stream(tri1)), "rrl_loss_forward_ex")
        ctx.ws = ws
        ctx.geom = (B, nf1, nf2, nl)
        ctx.mark_non_differentiable(status, median, stats)
        if holder is not None:
            holder.append(LossInfo(status, median, stats, ws, ctx.geom))
        return loss, status, median, stats

    @staticmethod
    def backward(ctx, grad_loss, *_unused):
        B, nf1, nf2, nl = ctx.geom
        ws = ctx.ws
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g = grad_loss.contiguous().float()
        g1 = torch.empty(B, nf1, 9, dtype=torch.float32, device=ws.device) if need1 else None
        g2 = torch.empty(B, nf2, 9, dtype=torch.float32, device=ws.device) if need2 else None
        with _on(ws):
            N.check(N.lib().rrl_loss_backward(ws.data_ptr(), ws.numel(), g.data_ptr(), B, nf1, nf2, nl,
                                              g1.data_ptr() if need1 else None, g2.data_ptr() if need2 else None,
                                              _stream(ws)), "rrl_loss_backward")
        return g1, g2, None, None, None, None


class _TwistLoss(torch.autograd.Function):
    """twist (B,6) -> exp3 -> raw_tri1 @ R + T -> loss (B,), with the backward contracted to pose space inside the kernels
    (rrl_loss_backward_twist): no dense point gradient is ever written.  `shard` = (PeerComm, NativeShardBackend factory)
    is used by dist.line_sharded_twist_loss."""

    @staticmethod
    def forward(ctx, twist, raw_tri1, tri2, lines, window, holder, session, comm):
        B, nf1, _ = raw_tri1.shape
        nf2, nl = tri2.shape[1], lines.shape[1]
        dev = raw_tri1.device
        L = N.lib()
        tri1 = torch.empty_like(raw_tri1)
        ws_bytes = L.rrl_workspace_bytes(B, nf1, nf2, nl)
        if ws_bytes == 0:
            raise ValueError("unsupported geometry B=%d nf1=%d nf2=%d nl=%d" % (B, nf1, nf2, nl))
        if session is not None:
            ws, flags = session.acquire((B, nf1, nf2, nl), dev, ws_bytes)
        else:
            ws, flags = torch.empty(ws_bytes, dtype=torch.uint8, device=dev), 0
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        median = torch.empty(B, dtype=torch.float32, device=dev)
        stats = torch.empty(B, N.NSTAT, dtype=torch.int64, device=dev)
        st = _stream(raw_tri1)
        with _on(raw_tri1):
            N.check(L.rrl_se3_apply(twist.data_ptr(), raw_tri1.data_ptr(), B, nf1 * 3, tri1.data_ptr(), st), "rrl_se3_apply")
            if comm is None:
                N.check(L.rrl_loss_forward_ex(tri1.data_ptr(), tri2.data_ptr(), lines.data_ptr(), B, nf1, nf2, nl,
                                              window[0], window[1], window[2], window[3], ws.data_ptr(), ws_bytes,
                                              loss.data_ptr(), status.data_ptr(), median.data_ptr(), stats.data_ptr(), flags, st),
                        "rrl_loss_forward_ex")
            else:                                           # one pair, this rank's line shard, exchange inside the kernels
                N.check(L.rrl_shard_stage1_ex(tri1.data_ptr(), tri2.data_ptr(), lines.data_ptr(), nf1, nf2, nl,
                                              window[0], window[1], window[2], window[3], ws.data_ptr(), ws_bytes, flags, st),
                        "rrl_shard_stage1_ex")
                N.check(L.rrl_shard_tail(ws.data_ptr(), ws_bytes, nf1, nf2, nl, comm.handle, loss.data_ptr(), status.data_ptr(),
                                         median.data_ptr(), stats.data_ptr(), st), "rrl_shard_tail")
        ctx.save_for_backward(twist, raw_tri1)
        ctx.ws, ctx.geom, ctx.comm = ws, (B, nf1, nf2, nl), comm
        ctx.mark_non_differentiable(status, median, stats)
        if holder is not None:
            holder.append(LossInfo(status, median, stats, ws, ctx.geom))
            holder.append(tri1)
        return loss, status, median, stats

    @staticmethod
    def backward(ctx, grad_loss, *_unused):
        twist, raw = ctx.saved_tensors
        B, nf1, nf2, nl = ctx.geom
        ws = ctx.ws
        g = grad_loss.contiguous().float()
        acc = torch.empty(B, 12, dtype=torch.float64, device=ws.device)
        gt = torch.empty(B, 6, dtype=torch.float32, device=ws.device)
        L = N.lib()
        with _on(ws):
            if ctx.comm is None:
                N.check(L.rrl_loss_backward_twist(ws.data_ptr(), ws.numel(), g.data_ptr(), B, nf1, nf2, nl, twist.data_ptr(),
                                                  raw.data_ptr(), acc.data_ptr(), gt.data_ptr(), _stream(ws)), "rrl_loss_backward_twist")
            else:                                           # the 12 pose-space sums of every rank's line shard, then the chain
                N.check(L.rrl_loss_backward_twist(ws.data_ptr(), ws.numel(), g.data_ptr(), B, nf1, nf2, nl, None,
                                                  raw.data_ptr(), acc.data_ptr(), None, _stream(ws)), "rrl_loss_backward_twist")
                ctx.comm.allreduce_f64_(acc.view(-1))
                N.check(L.rrl_se3_chain(twist.data_ptr(), acc.data_ptr(), B, gt.data_ptr(), _stream(ws)), "rrl_se3_chain")
        return gt, None, None, None, None, None, None, None


def twist_loss(twist: torch.Tensor, raw_tri1: torch.Tensor, tri2: torch.Tensor, lines: torch.Tensor, window=(1, 1, 5, 5),
               return_info: bool = False, session: Optional["LossSession"] = None, comm=None):
    """The registration step in one op (test_demo_optimized_Lie_Algebra.py:57-66): cloud 1 = exp(twist) applied to
    `raw_tri1` (B,nf1,9) (Reconstruction_point.forward, loss.py:455-463), per-pair losses (B,) against `tri2`, differentiable
    w.r.t. the twist only; forward = fused exp + transform + loss, backward = pose-space contraction + closed-form exp3
    derivative.  With return_info also returns (LossInfo, transformed tri1)."""
    for name, t, last in (("points1", raw_tri1, 9), ("points2", tri2, 9), ("line", lines, 6)):
        if t.dim() != 3 or t.shape[-1] != last:
            raise ValueError("%s must have shape (B, n, %d), got %s" % (name, last, tuple(t.shape)))
    B = raw_tri1.shape[0]
    if twist.shape != (B, 6) or tri2.shape[0] != B or lines.shape[0] != B:
        raise ValueError("expected twist (B,6) and matching batch sizes")
    w = tuple(int(v) for v in window)
    if not (1 <= w[0] < w[2] <= 5 and 1 <= w[1] < w[3] <= 5):
        raise ValueError("hit-count window must satisfy 1 <= lo < hi <= 5, got %s" % (w,))
    twist = _cuda_f32(twist, "twist")
    raw_tri1, tri2 = _cuda_f32(raw_tri1.detach(), "points1"), _cuda_f32(tri2.detach(), "points2")
    lines = _cuda_f32(lines.detach(), "line")
    _same_device(twist, raw_tri1, tri2, lines)
    holder = [] if return_info else None
    loss, _, _, _ = _TwistLoss.apply(twist, raw_tri1, tri2, lines, w, holder, session, comm)
    return (loss, holder[0], holder[1]) if return_info else loss


def intersected_line_loss(tri1: torch.Tensor, tri2: torch.Tensor, lines: torch.Tensor,
                          window=(1, 1, 5, 5), return_info: bool = False, session: Optional[LossSession] = None):
    """Batched native API: tri1 (B,nf1,9), tri2 (B,nf2,9), lines (B,nl,6) -> per-pair losses (B,).

    `window` = the reference's (s_m, s_n, e_m, e_n).  Differentiable w.r.t. tri1 and tri2.
    A pair with no populated (k,j) combo yields loss 0 with zero gradient and status bit RRL_STATUS_EMPTY.
    `session`: a LossSession that carries the clouds' spatial order from call to call (see there).
    """
    for name, t, last in (("points1", tri1, 9), ("points2", tri2, 9), ("line", lines, 6)):
        if t.dim() != 3 or t.shape[-1] != last:
            raise ValueError("%s must have shape (B, n, %d), got %s" % (name, last, tuple(t.shape)))
    if not (tri1.shape[0] == tri2.shape[0] == lines.shape[0]):
        raise ValueError("batch sizes differ")
    w = tuple(int(v) for v in window)
    if not (1 <= w[0] < w[2] <= 5 and 1 <= w[1] < w[3] <= 5):
        raise ValueError("hit-count window must satisfy 1 <= lo < hi <= 5, got %s" % (w,))
    tri1, tri2 = _cuda_f32(tri1, "points1"), _cuda_f32(tri2, "points2")
    lines = _cuda_f32(lines.detach(), "line")
    _same_device(tri1, tri2, lines)
    holder = [] if return_info else None
    loss, _, _, _ = _IntersectedLineLoss.apply(tri1, tri2, lines, w, holder, session)
    return (loss, holder[0]) if return_info else loss


# --------------------------------------------------------------------------------------------------------
class _Se3Apply(torch.autograd.Function):
    """out = points @ R(twist) + T(twist)   (Reconstruction_point.forward, loss.py:458-463)"""

    @staticmethod
    def forward(ctx, twist, points):
        B, n, _ = points.shape
        out = torch.empty_like(points)
        with _on(points):
            N.check(N.lib().rrl_se3_apply(twist.data_ptr(), points.data_ptr(), B, n, out.data_ptr(), _stream(points)),
                    "rrl_se3_apply")
        ctx.save_for_backward(twist, points)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        twist, points = ctx.saved_tensors
        B, n, _ = points.shape
        g = grad_out.contiguous().float()
        gt = torch.empty(B, 6, dtype=torch.float32, device=points.device)
        scratch = torch.empty(B * 12, dtype=torch.float64, device=points.device)
        with _on(points):
            N.check(N.lib().rrl_se3_apply_backward(twist.data_ptr(), points.data_ptr(), g.data_ptr(), B, n, gt.data_ptr(),
                                                   scratch.data_ptr(), _stream(points)), "rrl_se3_apply_backward")
        return gt, None


def se3_apply(twist: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """twist (B,6) [w|v], points (B,n,3) -> (B,n,3); differentiable w.r.t. the twist."""
    if twist.dim() != 2 or twist.shape[1] != 6 or points.dim() != 3 or points.shape[2] != 3:
        raise ValueError("expected twist (B,6) and points (B,n,3)")
    twist, points = _cuda_f32(twist, "twist"), _cuda_f32(points.detach(), "points")
    _same_device(twist, points)
    return _Se3Apply.apply(twist, points)


def se3_exp(twist: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """twist (B,6) -> R (B,3,3), T (B,3)   (se3.exp3, LieAlgebra/se3.py:83-106); no autograd."""
    twist = _cuda_f32(twist.detach().reshape(-1, 6), "twist")
    B = twist.shape[0]
    R = torch.empty(B, 3, 3, dtype=torch.float32, device=twist.device)
    T = torch.empty(B, 3, dtype=torch.float32, device=twist.device)
    with _on(twist):
        N.check(N.lib().rrl_se3_exp(twist.data_ptr(), B, R.data_ptr(), T.data_ptr(), _stream(twist)), "rrl_se3_exp")
    return R, T


class _Se3ExpMap(torch.autograd.Function):
    """FMR's se3.Exp (exps_deep_learning/fmr/se_math/se3.py:133-165): x (B,6) -> g (B,4,4); the backward is the
    reference's ExpMap.backward (sum_ij grad_g[i][j] (gen_k g)[i][j]), not the analytic derivative of exp."""

    @staticmethod
    def forward(ctx, x):
        B = x.shape[0]
        g = torch.empty(B, 4, 4, dtype=torch.float32, device=x.device)
        with _on(x):
            N.check(N.lib().rrl_se3_exp4(x.data_ptr(), B, g.data_ptr(), _stream(x)), "rrl_se3_exp4")
        ctx.save_for_backward(x)
        return g

    @staticmethod
    def backward(ctx, grad_g):
        x, = ctx.saved_tensors
        go = grad_g.contiguous().float()
        gx = torch.empty_like(x)
        with _on(x):
            N.check(N.lib().rrl_se3_expmap_backward(x.data_ptr(), go.data_ptr(), x.shape[0], gx.data_ptr(), _stream(x)),
                    "rrl_se3_expmap_backward")
        return gx


def se3_Exp(x: torch.Tensor) -> torch.Tensor:
    """Drop-in for fmr/se_math/se3.Exp: x (..., 6) -> g (..., 4, 4) with the reference's ExpMap gradient convention."""
    lead = x.shape[:-1]
    g = _Se3ExpMap.apply(_cuda_f32(x.reshape(-1, 6), "x"))
    return g.reshape(*lead, 4, 4)


class _RigidApply(torch.autograd.Function):
    """out = R p + t for the DL hooks (utils.py:32-37, rpm se3.transform, fmr se3.transform)."""

    @staticmethod
    def forward(ctx, R, t, points):
        B, n, _ = points.shape
        out = torch.empty_like(points)
        with _on(points):
            N.check(N.lib().rrl_rigid_apply(R.data_ptr(), t.data_ptr(), points.data_ptr(), B, n, out.data_ptr(),
                                            _stream(points)), "rrl_rigid_apply")
        ctx.save_for_backward(R, points)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        R, points = ctx.saved_tensors
        B, n, _ = points.shape
        g = grad_out.contiguous().float()
        gR = torch.empty(B, 3, 3, dtype=torch.float32, device=points.device)
        gt = torch.empty(B, 3, dtype=torch.float32, device=points.device)
        gp = torch.empty_like(points) if ctx.needs_input_grad[2] else None
        scratch = torch.empty(B * 12, dtype=torch.float64, device=points.device)
        with _on(points):
            N.check(N.lib().rrl_rigid_apply_backward(R.data_ptr(), points.data_ptr(), g.data_ptr(), B, n, gR.data_ptr(),
                                                     gt.data_ptr(), gp.data_ptr() if gp is not None else None,
                                                     scratch.data_ptr(), _stream(points)), "rrl_rigid_apply_backward")
        return gR, gt, gp


def rigid_apply(R: torch.Tensor, t: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """R (B,3,3), t (B,3), points (B,n,3) -> R p + t, differentiable w.r.t. all three."""
    R, t, points = _cuda_f32(R, "R"), _cuda_f32(t.reshape(-1, 3), "t"), _cuda_f32(points, "points")
    _same_device(R, t, points)
    return _RigidApply.apply(R, t, points)


# --------------------------------------------------------------------------------------------------------
def sample_lines(radius: torch.Tensor, centers: torch.Tensor, n_lines: int, verts1: torch.Tensor, verts2: torch.Tensor,
                 seed: int = 0, offset: int = 0, rounds: int = 10, uniforms: Optional[torch.Tensor] = None):
    """On-device sampler: radius (B,) or (B,1), centers (B,3), verts (B,n,3) -> (lines (B,N,6), filled (B,) int32)."""
    verts1, verts2 = _cuda_f32(verts1.detach(), "vertices1"), _cuda_f32(verts2.detach(), "vertices2")
    B = verts1.shape[0]
    dev = verts1.device
    radius = _cuda_f32(radius.detach().reshape(-1).to(dev), "r")
    centers = _cuda_f32(centers.detach().reshape(-1, 3).to(dev), "centers")
    if radius.numel() != B or centers.shape[0] != B or verts2.shape[0] != B:
        raise ValueError("batch sizes differ")
    if uniforms is not None:
        uniforms = _cuda_f32(uniforms.to(dev), "uniforms")
        if tuple(uniforms.shape) != (B, rounds, 4, n_lines):
            raise ValueError("uniforms must have shape (B, rounds, 4, N)")
    L = N.lib()
    wsb = L.rrl_sampler_workspace_bytes(B, n_lines, rounds)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    lines = torch.empty(B, n_lines, 6, dtype=torch.float32, device=dev)
    filled = torch.empty(B, dtype=torch.int32, device=dev)
    _same_device(verts1, verts2, radius, centers, uniforms)
    with _on(verts1):
        N.check(L.rrl_sample_lines(radius.data_ptr(), centers.data_ptr(), verts1.data_ptr(), verts2.data_ptr(), B,
                                   verts1.shape[1], verts2.shape[1], n_lines, rounds, seed & (2 ** 64 - 1),
                                   offset & (2 ** 64 - 1), uniforms.data_ptr() if uniforms is not None else None,
                                   lines.data_ptr(), filled.data_ptr(), ws.data_ptr(), wsb, _stream(verts1)),
                "rrl_sample_lines")
    return lines, filled


class _Chamfer(torch.autograd.Function):
    """chamfer_dist (loss.py:236-252) with its autograd: the forward records the argmin of every directed minimum."""

    @staticmethod
    def forward(ctx, x, y):
        B, M, _ = x.shape
        Nn = y.shape[1]
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        scratch = torch.empty(B * (M + Nn), dtype=torch.float32, device=x.device)
        arg = torch.empty(B * (M + Nn), dtype=torch.int32, device=x.device)
        with _on(x):
            N.check(N.lib().rrl_chamfer_forward(x.data_ptr(), y.data_ptr(), B, M, Nn, out.data_ptr(), scratch.data_ptr(),
                                                arg.data_ptr(), _stream(x)), "rrl_chamfer_forward")
        ctx.save_for_backward(x, y, arg)
        return out[0]

    @staticmethod
    def backward(ctx, grad_out):
        x, y, arg = ctx.saved_tensors
        B, M, _ = x.shape
        Nn = y.shape[1]
        go = grad_out.reshape(1).contiguous().float()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gy = torch.empty_like(y) if ctx.needs_input_grad[1] else None
        with _on(x):
            N.check(N.lib().rrl_chamfer_backward(x.data_ptr(), y.data_ptr(), arg.data_ptr(), go.data_ptr(), B, M, Nn,
                                                 gx.data_ptr() if gx is not None else None,
                                                 gy.data_ptr() if gy is not None else None, _stream(x)),
                    "rrl_chamfer_backward")
        return gx, gy


def chamfer(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """chamfer_dist (loss.py:236-252): x (B,M,3), y (B,N,3) -> 0-dim tensor, differentiable w.r.t. both clouds like the
    reference's (Train_DCP.py:248,297 returns it in the loss tuple)."""
    x, y = _cuda_f32(x, "points_x"), _cuda_f32(y, "points_y")
    _same_device(x, y)
    if x.dim() != 3 or y.dim() != 3 or x.shape[2] != 3 or y.shape[2] != 3 or x.shape[0] != y.shape[0]:
        raise ValueError("expected points_x (B,M,3) and points_y (B,N,3)")
    if not (torch.is_grad_enabled() and (x.requires_grad or y.requires_grad)):
        B, M, _ = x.shape
        Nn = y.shape[1]
        out = torch.empty(1, dtype=torch.float32, device=x.device)
        scratch = torch.empty(B * (M + Nn), dtype=torch.float32, device=x.device)
        with _on(x):
            N.check(N.lib().rrl_chamfer(x.data_ptr(), y.data_ptr(), B, M, Nn, out.data_ptr(), scratch.data_ptr(), _stream(x)),
                    "rrl_chamfer")
        return out[0]
    return _Chamfer.apply(x, y)
