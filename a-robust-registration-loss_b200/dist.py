"""Multi-GPU evaluation (one process per GPU, torch.distributed; NCCL on the B200 box, gloo in CPU tests of the
host logic).  SURVEY 8(e):

  batch shard  pairs are independent -> each rank evaluates its own pairs; the only exchange is an all-reduce of the
               scalar loss (and whatever the caller logs).
  line shard   one pair, lines split across ranks; the global lower median and the 1/(n_kj k) normalisers couple
               all lines, so there is one exchange step between the dense and the Welsch stage -- an all-reduce of
               the 18 counts and two all-reduces of 16-bit key histograms from which every rank picks the same
               median (no host round trip, no variable-length gather) -- and one all-reduce of the fixed-point
               partial sums; in backward the point gradient is all-reduced, or, with reduce_grad=False, left
               local so that the caller reduces it in pose space (line_sharded_twist_loss: 6 floats).
"""
import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist_

from . import _native as N


class PeerComm:
    """The ranks' peer-memory exchange buffers (include/rrl_b200.h, rrl_comm_*): one per rank, mapped by every rank of the
    group over CUDA IPC / NVLink.  With it the line shard's exchange step runs INSIDE its kernels (rrl_shard_tail,
    rrl_comm_allreduce_f64): no NCCL collective, no host round trip, graph-capturable.  torch.distributed only carries
    the 64-byte IPC handles here, once.  `PeerComm.create` returns None when peer mapping is not available (the callers
    then fall back to the NCCL protocol)."""

    def __init__(self, handle, rank, world, device, nl_capacity):
        self.handle, self.rank, self.world, self.device, self.nl_capacity = handle, rank, world, device, nl_capacity

    @staticmethod
    def slot_bytes_for(nl_local: int) -> int:
        return 160 + 64 * int(nl_local)

    @classmethod
    def create(cls, nl_capacity: int, device=None, group=None) -> Optional["PeerComm"]:
        """Collective over `group` (every rank must call it).  nl_capacity = the largest line shard any rank will pass."""
        L = N.lib()
        world = dist_.get_world_size(group) if dist_.is_initialized() else 1
        rank = dist_.get_rank(group) if dist_.is_initialized() else 0
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        h = C.c_void_p()
        with torch.cuda.device(device):
            rc = L.rrl_comm_create(rank, world, cls.slot_bytes_for(nl_capacity), C.byref(h))
            ok = rc == 0
            mine = torch.zeros(64, dtype=torch.uint8)
            if ok and world > 1:
                buf = (C.c_ubyte * 64)()
                ok = L.rrl_comm_ipc_handle(h, buf) == 0
                mine = torch.tensor(list(buf), dtype=torch.uint8)
            if world > 1:
                # one all-gather of (ok flag + handle): every rank learns whether EVERY rank can take part
                send = torch.cat([torch.tensor([1 if ok else 0], dtype=torch.uint8), mine]).to(device)
                recv = torch.empty(world * 65, dtype=torch.uint8, device=device)
                dist_.all_gather_into_tensor(recv, send, group=group)
                recv = recv.cpu().reshape(world, 65)
                all_ok = bool(recv[:, 0].all())
                if all_ok:
                    handles = recv[:, 1:].contiguous().numpy().tobytes()
                    all_ok = L.rrl_comm_connect_ipc(h, handles) == 0
                # ... and whether every rank managed to map every peer
                flag = torch.tensor([1 if all_ok else 0], dtype=torch.int32, device=device)
                dist_.all_reduce(flag, op=dist_.ReduceOp.MIN, group=group)
                ok = bool(flag.item())
        if not ok:
            if h:
                L.rrl_comm_destroy(h)
            return None
        return cls(h, rank, world, device, int(nl_capacity))

    def allreduce_f64_(self, buf: torch.Tensor) -> torch.Tensor:
        """in place sum over the ranks of a float64 CUDA tensor of <= 512 elements (rank order: bit-identical everywhere)"""
        assert buf.dtype == torch.float64 and buf.is_contiguous() and buf.numel() <= 512
        with torch.cuda.device(buf.device):
            N.check(N.lib().rrl_comm_allreduce_f64(self.handle, buf.data_ptr(), buf.numel(),
                                                   torch.cuda.current_stream(buf.device).cuda_stream), "rrl_comm_allreduce_f64")
        return buf

    def error(self) -> int:
        return int(N.lib().rrl_comm_error(self.handle))

    def close(self):
        if self.handle:
            N.lib().rrl_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block partition of n items: the first n % world ranks get one extra item"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_sum_with_local_grad(total: torch.Tensor, group=None, average: bool = False) -> torch.Tensor:
    """all-reduce(sum) of a scalar: the value is the sum over ranks, the gradient flows to the local term only"""
    if dist_.is_available() and dist_.is_initialized() and dist_.get_world_size(group) > 1:
        red = total.detach().clone()
        dist_.all_reduce(red, op=dist_.ReduceOp.SUM, group=group)
        if average:
            red = red / dist_.get_world_size(group)
        total = total + (red - total.detach())
    return total


def batch_sharded_loss(tri1, tri2, lines, window=(1, 1, 5, 5), group=None, average: bool = False):
    """Each rank passes ITS OWN pairs (any B, possibly different per rank).  Returns (local per-pair losses (B,),
    global sum of all losses as a 0-dim tensor that carries gradient for the local pairs only)."""
    from . import ops
    local = ops.intersected_line_loss(tri1, tri2, lines, window)
    return local, global_sum_with_local_grad(local.sum(), group, average)


def combine_counts(local18: torch.Tensor, group=None) -> torch.Tensor:
    out = local18.clone()
    dist_.all_reduce(out, op=dist_.ReduceOp.SUM, group=group)
    return out


def gather_entries(local: torch.Tensor, n_local: int, counts: List[int], group=None) -> torch.Tensor:
    """all-gather of variable-length float arrays: pad to the maximum, gather, strip"""
    world = dist_.get_world_size(group)
    cap = max(max(counts), 1)
    buf = torch.zeros(cap, dtype=torch.float32, device=local.device)
    buf[:n_local] = local[:n_local]
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist_.all_gather(outs, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, counts)]) if sum(counts) else buf[:0]


class NativeShardBackend:
    """The per-rank stages of the line-sharded evaluation on the C ABI (include/rrl_b200.h, rrl_shard_*)."""

    def __init__(self, tri1, tri2, lines_local, window, session=None):
        self.L = N.lib()
        self.dev = tri1.device
        self.nf1, self.nf2, self.nl = tri1.shape[0], tri2.shape[0], lines_local.shape[0]
        self.tri1, self.tri2, self.lines, self.window = tri1, tri2, lines_local, window
        self.wsb = self.L.rrl_workspace_bytes(1, self.nf1, self.nf2, self.nl)
        if session is not None:                 # ops.LossSession: the clouds' spatial order survives from call to call
            self.ws, self.flags = session.acquire((1, self.nf1, self.nf2, self.nl), self.dev, self.wsb)
        else:
            self.ws, self.flags = torch.empty(self.wsb, dtype=torch.uint8, device=self.dev), 0

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream

    def fused_forward(self, comm: "PeerComm"):
        """stage 1 + rrl_shard_tail: the whole line-sharded forward of this rank in the kernels' own exchange"""
        if self.nl > comm.nl_capacity:
            raise ValueError("line shard of %d lines exceeds the communicator's capacity (%d)" % (self.nl, comm.nl_capacity))
        w = self.window
        loss = torch.empty(1, dtype=torch.float32, device=self.dev)
        status = torch.empty(1, dtype=torch.int32, device=self.dev)
        med = torch.empty(1, dtype=torch.float32, device=self.dev)
        with torch.cuda.device(self.dev):
            N.check(self.L.rrl_shard_stage1_ex(self.tri1.data_ptr(), self.tri2.data_ptr(), self.lines.data_ptr(), self.nf1,
                                               self.nf2, self.nl, w[0], w[1], w[2], w[3], self.ws.data_ptr(), self.wsb,
                                               self.flags, self._stream()), "rrl_shard_stage1_ex")
            N.check(self.L.rrl_shard_tail(self.ws.data_ptr(), self.wsb, self.nf1, self.nf2, self.nl, comm.handle,
                                          loss.data_ptr(), status.data_ptr(), med.data_ptr(), None, self._stream()),
                    "rrl_shard_tail")
        return loss, status, med

    def _geom(self):
        return self.ws.data_ptr(), self.wsb, self.nf1, self.nf2, self.nl

    def stage1_counts(self):
        w = self.window
        N.check(self.L.rrl_shard_stage1_ex(self.tri1.data_ptr(), self.tri2.data_ptr(), self.lines.data_ptr(), self.nf1,
                                           self.nf2, self.nl, w[0], w[1], w[2], w[3], self.ws.data_ptr(), self.wsb,
                                           self.flags, self._stream()), "rrl_shard_stage1_ex")
        counts = torch.empty(18, dtype=torch.int64, device=self.dev)
        N.check(self.L.rrl_shard_counts(*self._geom(), counts.data_ptr(), self._stream()), "rrl_shard_counts")
        return counts

    def pack_entries(self, n_local):
        out = torch.empty(max(n_local, 1), dtype=torch.float32, device=self.dev)
        N.check(self.L.rrl_shard_pack_entries(*self._geom(), out.data_ptr(), n_local, self._stream()),
                "rrl_shard_pack_entries")
        return out

    def median(self, entries):
        med = torch.empty(1, dtype=torch.float32, device=self.dev)
        N.check(self.L.rrl_select_lower_median(entries.data_ptr() if entries.numel() else None, entries.numel(),
                                               med.data_ptr(), self._stream()), "rrl_select_lower_median")
        return med

    def select_hist(self, rnd, state):
        hist = torch.empty(65536, dtype=torch.int32, device=self.dev)
        N.check(self.L.rrl_shard_select_hist(*self._geom(), int(rnd), state.data_ptr(), hist.data_ptr(), self._stream()),
                "rrl_shard_select_hist")
        return hist

    def select_pick(self, rnd, ghist, gcounts, state, med):
        N.check(self.L.rrl_shard_select_pick(int(rnd), ghist.data_ptr(), gcounts.data_ptr(), state.data_ptr(), med.data_ptr(),
                                             self._stream()), "rrl_shard_select_pick")

    def stage2_sums(self, gcounts, med):
        sums = torch.empty(32, dtype=torch.int64, device=self.dev)
        N.check(self.L.rrl_shard_stage2(*self._geom(), gcounts.data_ptr(), med.data_ptr(), sums.data_ptr(),
                                        self._stream()), "rrl_shard_stage2")
        return sums

    def stage3_loss(self, gsums):
        loss = torch.empty(1, dtype=torch.float32, device=self.dev)
        status = torch.empty(1, dtype=torch.int32, device=self.dev)
        N.check(self.L.rrl_shard_stage3(*self._geom(), gsums.data_ptr(), loss.data_ptr(), status.data_ptr(),
                                        self._stream()), "rrl_shard_stage3")
        return loss, status


def line_shard_forward(backend, group=None):
    """The exchange protocol of SURVEY 8(e), independent of where the stages run (the CPU tests drive it with an
    oracle-backed backend over gloo).  Three small all-reduces, nothing read by the host, nothing of variable length:
    the 18 counts, two 65536-bin key histograms (distributed lower median: high then low 16 bits of the float bit
    pattern), the 32 fixed-point partial sums."""
    lcounts = backend.stage1_counts()
    state = torch.zeros(2, dtype=torch.int64, device=lcounts.device)
    med = torch.zeros(1, dtype=torch.float32, device=lcounts.device)
    # the first histogram does not depend on the global counts (only the pick does): counts and histogram share ONE
    # all-reduce (int64 payload: the 18 counts behind the 65536 bins)
    hist0 = backend.select_hist(0, state)
    buf = torch.empty(65536 + 18, dtype=torch.int64, device=lcounts.device)
    buf[:65536] = hist0
    buf[65536:] = lcounts
    dist_.all_reduce(buf, op=dist_.ReduceOp.SUM, group=group)
    gcounts = buf[65536:].clone()
    backend.select_pick(0, buf[:65536].to(torch.int32), gcounts, state, med)
    hist = backend.select_hist(1, state)
    dist_.all_reduce(hist, op=dist_.ReduceOp.SUM, group=group)
    backend.select_pick(1, hist, gcounts, state, med)
    sums = backend.stage2_sums(gcounts, med)
    dist_.all_reduce(sums, op=dist_.ReduceOp.SUM, group=group)
    loss, status = backend.stage3_loss(sums)
    return loss, status, med


class _LineShardedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tri1, tri2, lines_local, window, group, reduce_grad, session=None, comm=None):
        backend = NativeShardBackend(tri1, tri2, lines_local, window, session)
        if comm is not None:
            loss, status, med = backend.fused_forward(comm)
        else:
            with torch.cuda.device(tri1.device):
                loss, status, med = line_shard_forward(backend, group)
        ctx.ws, ctx.geom, ctx.group, ctx.reduce_grad = backend.ws, (backend.nf1, backend.nf2, backend.nl), group, reduce_grad
        ctx.mark_non_differentiable(status, med)
        return loss, status, med

    @staticmethod
    def backward(ctx, grad_loss, *_):
        nf1, nf2, nl = ctx.geom
        ws = ctx.ws
        dev = ws.device
        g = grad_loss.contiguous().float()
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g1 = torch.empty(nf1, 9, dtype=torch.float32, device=dev) if need1 else None
        g2 = torch.empty(nf2, 9, dtype=torch.float32, device=dev) if need2 else None
        with torch.cuda.device(dev):
            N.check(N.lib().rrl_loss_backward(ws.data_ptr(), ws.numel(), g.data_ptr(), 1, nf1, nf2, nl,
                                              g1.data_ptr() if need1 else None, g2.data_ptr() if need2 else None,
                                              torch.cuda.current_stream(dev).cuda_stream), "rrl_loss_backward")
        # every rank holds the replicated clouds, so the point gradient is summed over the line shards
        if ctx.reduce_grad:
            for t in (g1, g2):
                if t is not None:
                    dist_.all_reduce(t, op=dist_.ReduceOp.SUM, group=ctx.group)
        return g1, g2, None, None, None, None, None, None


class _AllReduceGrad(torch.autograd.Function):
    """identity whose gradient is summed over the ranks (through the peer exchange when a PeerComm is given: a few floats)"""

    @staticmethod
    def forward(ctx, x, group, comm):
        ctx.group, ctx.comm = group, comm
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        if ctx.comm is not None and g.numel() <= 512:
            g64 = g.contiguous().double()
            ctx.comm.allreduce_f64_(g64.view(-1))
            return g64.to(g.dtype), None, None
        g = g.contiguous().clone()
        dist_.all_reduce(g, op=dist_.ReduceOp.SUM, group=ctx.group)
        return g, None, None


class ShardedSampler:
    """Candidate-sharded line sampler for ONE pair (loss.py:415-432): rank r evaluates the candidate
    chunks r (mod world) of every round, the ranks sum their per-chunk accepted counts (ONE all-reduce of
    rrl_sampler_num_chunks ints -- the "scan of accepted counts" of the survey), and every rank keeps the accepted
    candidates of its own chunks that fall among the first N of the reference's ordered compaction, plus its share of the
    unfilled (all-zero) rows.  The union of the ranks' rows is exactly the line set `ops.sample_lines` returns; the loss is
    independent of the order of the lines, so no line is ever exchanged.  The number of local rows is data dependent: one
    3-int device->host read per call (the loss kernels need the row count on the host)."""

    def __init__(self, n_lines: int, rounds: int = 10, device=None, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None):
        self.L = N.lib()
        self.N, self.rounds, self.group = int(n_lines), int(rounds), group
        if rank is None:
            rank = dist_.get_rank(group) if dist_.is_initialized() else 0
            world = dist_.get_world_size(group) if dist_.is_initialized() else 1
        self.rank, self.world = int(rank), int(world)
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nchunks = int(self.L.rrl_sampler_num_chunks(self.N, self.rounds))
        if self.nchunks <= 0:
            raise ValueError("unsupported sampler geometry N=%d rounds=%d" % (self.N, self.rounds))
        self.wsb = self.L.rrl_sampler_workspace_bytes(1, self.N, self.rounds)
        self.ws = torch.empty(self.wsb, dtype=torch.uint8, device=self.dev)

    def _args(self, radius, centers, uniforms):
        from .ops import _cuda_f32
        radius = _cuda_f32(radius.detach().reshape(-1).to(self.dev), "r")
        centers = _cuda_f32(centers.detach().reshape(-1, 3).to(self.dev), "centers")
        if radius.numel() != 1 or centers.shape[0] != 1:
            raise ValueError("the sharded sampler handles one pair (B = 1)")
        if uniforms is not None:
            uniforms = _cuda_f32(uniforms.to(self.dev), "uniforms")
            if tuple(uniforms.shape) != (1, self.rounds, 4, self.N):
                raise ValueError("uniforms must have shape (1, rounds, 4, N)")
        return radius, centers, uniforms

    def local_counts(self, radius, centers, verts1, verts2, seed=0, offset=0, uniforms=None) -> torch.Tensor:
        """step 1: evaluate this rank's candidate chunks; returns its per-chunk accepted counts (nchunks,) int32"""
        from .ops import _cuda_f32
        radius, centers, uniforms = self._args(radius, centers, uniforms)
        verts1, verts2 = _cuda_f32(verts1.detach(), "vertices1"), _cuda_f32(verts2.detach(), "vertices2")
        counts = torch.empty(self.nchunks, dtype=torch.int32, device=self.dev)
        self._call = (radius, centers, uniforms, int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1))
        with torch.cuda.device(self.dev):
            N.check(self.L.rrl_sample_lines_shard_flags(radius.data_ptr(), centers.data_ptr(), verts1.data_ptr(), verts2.data_ptr(), 1,
                                                        verts1.shape[1], verts2.shape[1], self.N, self.rounds, self._call[3], self._call[4],
                                                        uniforms.data_ptr() if uniforms is not None else None, self.rank, self.world,
                                                        counts.data_ptr(), self.ws.data_ptr(), self.wsb,
                                                        torch.cuda.current_stream(self.dev).cuda_stream), "rrl_sample_lines_shard_flags")
        return counts

    def place(self, local_counts: torch.Tensor, global_counts: torch.Tensor):
        """step 3: given the counts summed over the ranks; returns (lines_local (rows, 6), rows filled globally)"""
        radius, centers, uniforms, seed, offset = self._call
        lines = torch.empty(self.N, 6, dtype=torch.float32, device=self.dev)
        out3 = torch.empty(3, dtype=torch.int32, device=self.dev)
        g, l = global_counts.clone(), local_counts.clone()
        with torch.cuda.device(self.dev):
            N.check(self.L.rrl_sample_lines_shard_scatter(radius.data_ptr(), centers.data_ptr(), 1, self.N, self.rounds, seed, offset,
                                                          uniforms.data_ptr() if uniforms is not None else None, self.rank, self.world,
                                                          g.data_ptr(), l.data_ptr(), lines.data_ptr(), out3.data_ptr(),
                                                          self.ws.data_ptr(), self.wsb,
                                                          torch.cuda.current_stream(self.dev).cuda_stream), "rrl_sample_lines_shard_scatter")
        placed, zeros, filled = out3.tolist()                   # the one host read: the loss needs the row count
        return lines[:placed + zeros], filled

    def sample(self, radius, centers, verts1, verts2, seed=0, offset=0, uniforms=None):
        """all three steps with the all-reduce over `group`"""
        local = self.local_counts(radius, centers, verts1, verts2, seed, offset, uniforms)
        total = local.clone()
        if self.world > 1:
            dist_.all_reduce(total, op=dist_.ReduceOp.SUM, group=self.group)
        return self.place(local, total)


def line_sharded_twist_loss(twist, raw_tri1, tri2, lines_local, window=(1, 1, 5, 5), group=None, session=None, comm=None):
    """The demo's / large-scan configuration: cloud 1 = se(3) transform of `raw_tri1` (nf1,9) by `twist` (6,), one pair,
    lines sharded.  The sparse point gradient of every rank's line shard is reduced to pose space locally (closed-form
    se(3) backward) and only the 6 twist-gradient floats cross NVLink.  Returns (loss (1,), status, median)."""
    from . import ops
    if comm is not None:
        # everything inside the kernels: exp + transform + dense + exchange + loss forward; pose-space contraction + 12-double
        # peer all-reduce + exp3 derivative backward
        loss, info, _ = ops.twist_loss(twist.reshape(1, 6), raw_tri1.reshape(1, -1, 9), tri2.reshape(1, -1, 9),
                                       lines_local.reshape(1, -1, 6), window, return_info=True, session=session, comm=comm)
        return loss, info.status, info.median
    tw = _AllReduceGrad.apply(twist.reshape(1, 6), group, comm)
    tri1 = ops.se3_apply(tw, raw_tri1.reshape(1, -1, 3)).reshape(-1, 9)
    return line_sharded_loss(tri1, tri2, lines_local, window, group, reduce_grad=False, session=session, comm=comm)


def line_sharded_loss(tri1, tri2, lines_local, window=(1, 1, 5, 5), group=None, reduce_grad=True, session=None, comm=None):
    """ONE pair: tri1 (nf1,9) and tri2 (nf2,9) replicated on every rank, lines_local (nl_r,6) = this rank's shard.
    Returns (loss (1,), status (1,), median (1,)) -- identical on all ranks; the point gradients are all-reduced unless
    reduce_grad=False (then each rank keeps the gradient of its own line shard).  With `comm` (a PeerComm) the exchange
    step of the forward runs inside the kernels over peer memory instead of through four NCCL collectives."""
    from .ops import _cuda_f32
    if comm is None and not (dist_.is_available() and dist_.is_initialized()):
        raise RuntimeError("line_sharded_loss needs an initialised torch.distributed process group (or a PeerComm)")
    w = tuple(int(v) for v in window)
    return _LineShardedLoss.apply(_cuda_f32(tri1, "points1"), _cuda_f32(tri2, "points2"),
                                  _cuda_f32(lines_local.detach(), "line"), w, group, bool(reduce_grad), session, comm)
