"""Line shard of ONE pair (BASELINE configs[4]) on a GPU: the per-rank stage kernels of the NCCL protocol
(rrl_shard_counts / pack_entries / select_hist / select_pick / stage2 / stage3, rrl_select_lower_median) and the fused
peer-memory tail (rrl_shard_tail, rrl_comm_*), driven on ONE device so that the driver's single-GPU `-m gpu` run covers
them:

  * emulation: R "ranks" = R workspaces on R slices of the lines; what the collectives would sum is added on the device;
  * the peer exchange for real: R communicators in one process connected by pointers, one stream per rank, the ranks'
    kernels spin on each other's flags exactly as they do across GPUs;
  * world-size-1 process group over NCCL through the public rrl_b200.dist entry points.

Bars: global lower median bit-identical to the single-GPU forward and to the oracle (loss.py:221-224 over ALL lines),
loss within 1e-5 of the oracle (bit-identical between the sharded and the unsharded evaluation: fixed-point sums), the
ranks' gradient shares sum to the oracle's gradient within 1e-5."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


@pytest.fixture(scope="module")
def rrl():
    import rrl_b200
    assert torch.cuda.is_available()
    return rrl_b200


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _splits(nl, R):
    from rrl_b200.dist import shard_range
    return [shard_range(nl, r, R) for r in range(R)]


def _single(rrl, p):
    t1 = torch.from_numpy(p["tri1"]).cuda()[None].requires_grad_(True)
    t2 = torch.from_numpy(p["tri2"]).cuda()[None]
    ln = torch.from_numpy(p["lines"]).cuda()[None]
    loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
    loss.sum().backward()
    return loss.item(), float(info.median[0]), t1.grad[0].cpu().numpy()


@pytest.mark.parametrize("nf,nl,R", [(700, 3001, 2), (900, 2500, 3), (20000, 1500, 2)])
def test_nccl_protocol_stages_emulated_on_one_gpu(rrl, nf, nl, R):
    """every rrl_shard_* stage kernel with the collectives replaced by device-side additions"""
    from rrl_b200.dist import NativeShardBackend
    p = synth.make_pair(600 + nf % 97, nf, nl, zero_frac=0.05)
    t1 = torch.from_numpy(p["tri1"]).cuda(); t2 = torch.from_numpy(p["tri2"]).cuda()
    lines = torch.from_numpy(p["lines"]).cuda()
    ranks = [NativeShardBackend(t1, t2, lines[lo:hi].contiguous(), (1, 1, 5, 5)) for lo, hi in _splits(nl, R)]
    counts = [b.stage1_counts() for b in ranks]
    gcounts = torch.stack(counts).sum(0)
    # (a) gather path: pack the entries of every rank, select on the concatenation
    ents = [b.pack_entries(int(c[17])) [:int(c[17])] for b, c in zip(ranks, counts)]
    med_gather = ranks[0].median(torch.cat(ents))
    # (b) histogram path: two rounds, histograms summed
    state = torch.zeros(2, dtype=torch.int64, device="cuda")
    med = torch.zeros(1, dtype=torch.float32, device="cuda")
    for rnd in (0, 1):
        hist = torch.stack([b.select_hist(rnd, state) for b in ranks]).sum(0).to(torch.int32)
        ranks[0].select_pick(rnd, hist, gcounts, state, med)
    sums = torch.stack([b.stage2_sums(gcounts, med) for b in ranks]).sum(0)
    outs = [b.stage3_loss(sums) for b in ranks]
    orc = co.loss(p["tri1"], p["tri2"], p["lines"])
    assert int(gcounts[16]) == orc.n_selected and int(gcounts[17]) == orc.n_entries
    assert float(med) == orc.median and float(med_gather) == orc.median
    loss1, med1, grad1 = _single(rrl, p)
    for loss, status in outs:
        assert loss.item() == loss1 and int(status) == 0           # fixed-point sums: sharding changes no bit
        assert abs(loss.item() - orc.loss) <= REL_TOL * orc.loss
    # backward of every rank's share (global counts and median are in its workspace), summed = the full gradient
    L = rrl._native.lib()
    total = torch.zeros(nf, 9, device="cuda")
    go = torch.ones(1, device="cuda")
    for b in ranks:
        g = torch.empty(nf, 9, device="cuda")
        assert L.rrl_loss_backward(b.ws.data_ptr(), b.wsb, go.data_ptr(), 1, b.nf1, b.nf2, b.nl, g.data_ptr(), None, None) == 0
        total += g
        # stage 2 may be re-run on the same forward (gradient vectors live in their own buffer): same sums, same gradient
        again = b.stage2_sums(gcounts, med)
        g2 = torch.empty(nf, 9, device="cuda")
        assert L.rrl_loss_backward(b.ws.data_ptr(), b.wsb, go.data_ptr(), 1, b.nf1, b.nf2, b.nl, g2.data_ptr(), None, None) == 0
        assert torch.equal(again, b.stage2_sums(gcounts, med)) and _rel(g2.cpu().numpy(), g.cpu().numpy()) <= 1e-6
    assert _rel(total.cpu().numpy(), orc.grad1) <= REL_TOL


def _make_comms(L, R, nl_cap):
    hs = []
    for r in range(R):
        h = C.c_void_p()
        assert L.rrl_comm_create(r, R, 160 + 64 * nl_cap, C.byref(h)) == 0
        hs.append(h)
    bases = (C.c_void_p * R)(*[L.rrl_comm_local_base(h) for h in hs])
    for h in hs:
        assert L.rrl_comm_connect_ptrs(h, bases) == 0
    return hs


@pytest.mark.parametrize("nf,nl,R", [(700, 3001, 2), (1024, 20000, 4), (20000, 1500, 2), (300, 97, 8)])
def test_fused_peer_exchange_tail(rrl, nf, nl, R):
    """rrl_shard_tail: R ranks on one GPU, one stream each, exchanging through their mapped buffers.  Run twice on the
    same communicators (sequence numbers and slot parity carry over) with different lines."""
    L = rrl._native.lib()
    p = synth.make_pair(640 + nf % 89, nf, nl, zero_frac=0.03)
    q = synth.make_pair(641 + nf % 89, nf, nl)
    comms = _make_comms(L, R, (nl + R - 1) // R + 1)
    streams = [torch.cuda.Stream() for _ in range(R)]
    try:
        for case in (p, q, p):
            t1 = torch.from_numpy(case["tri1"]).cuda(); t2 = torch.from_numpy(case["tri2"]).cuda()
            lines = torch.from_numpy(case["lines"]).cuda()
            torch.cuda.synchronize()
            ws, outs = [], []
            for r, (lo, hi) in enumerate(_splits(nl, R)):
                n_r = hi - lo
                wsb = L.rrl_workspace_bytes(1, nf, nf, n_r)
                w = torch.empty(wsb, dtype=torch.uint8, device="cuda")
                ln = lines[lo:hi].contiguous()
                o = (torch.zeros(1, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda"), torch.zeros(1, device="cuda"),
                     torch.empty(nf, 9, device="cuda"))
                ws.append((w, wsb, n_r, ln)); outs.append(o)
            torch.cuda.synchronize()
            for r in range(R):                                  # stage 1 of every rank first, then the exchanging tails
                w, wsb, n_r, ln = ws[r]
                assert L.rrl_shard_stage1(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), nf, nf, n_r, 1, 1, 5, 5, w.data_ptr(), wsb,
                                          streams[r].cuda_stream) == 0
            go = torch.ones(1, device="cuda")
            for r in range(R):
                w, wsb, n_r, ln = ws[r]
                o = outs[r]
                assert L.rrl_shard_tail(w.data_ptr(), wsb, nf, nf, n_r, comms[r], o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                        None, streams[r].cuda_stream) == 0
                assert L.rrl_loss_backward(w.data_ptr(), wsb, go.data_ptr(), 1, nf, nf, n_r, o[3].data_ptr(), None,
                                           streams[r].cuda_stream) == 0
            torch.cuda.synchronize()
            for h in comms:
                assert L.rrl_comm_error(h) == 0
            orc = co.loss(case["tri1"], case["tri2"], case["lines"])
            loss1, med1, grad1 = _single(rrl, case)
            total = torch.zeros(nf, 9, device="cuda")
            for o in outs:
                assert int(o[1]) == 0
                assert float(o[2]) == orc.median == med1                     # the global lower median, bit for bit
                assert o[0].item() == loss1                                  # identical on every rank and to the unsharded call
                assert abs(o[0].item() - orc.loss) <= REL_TOL * orc.loss
                total += o[3]
            assert _rel(total.cpu().numpy(), orc.grad1) <= REL_TOL
    finally:
        torch.cuda.synchronize()
        for h in comms:
            L.rrl_comm_destroy(h)


def test_fused_tail_with_a_rank_that_selects_nothing(rrl):
    """one rank's lines miss both clouds entirely (no record, no entry): it still takes part in both exchanges"""
    L = rrl._native.lib()
    nf, nl = 500, 1200
    p = synth.make_pair(660, nf, nl)
    lines = p["lines"].copy()
    lines[600:, 3:] += 1000.0                                    # second half: far away from everything
    t1 = torch.from_numpy(p["tri1"]).cuda(); t2 = torch.from_numpy(p["tri2"]).cuda(); ln = torch.from_numpy(lines).cuda()
    comms = _make_comms(L, 2, 600)
    streams = [torch.cuda.Stream() for _ in range(2)]
    try:
        res = []
        wss = []
        torch.cuda.synchronize()
        for r in range(2):
            wsb = L.rrl_workspace_bytes(1, nf, nf, 600)
            w = torch.empty(wsb, dtype=torch.uint8, device="cuda")
            part = ln[r * 600:(r + 1) * 600].contiguous()
            out = (torch.zeros(1, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda"), torch.zeros(1, device="cuda"))
            wss.append((w, part)); res.append(out)
            assert L.rrl_shard_stage1(t1.data_ptr(), t2.data_ptr(), part.data_ptr(), nf, nf, 600, 1, 1, 5, 5, w.data_ptr(), wsb,
                                      streams[r].cuda_stream) == 0
        for r in range(2):
            w, part = wss[r]
            assert L.rrl_shard_tail(w.data_ptr(), w.numel(), nf, nf, 600, comms[r], res[r][0].data_ptr(), res[r][1].data_ptr(),
                                    res[r][2].data_ptr(), None, streams[r].cuda_stream) == 0
        torch.cuda.synchronize()
        orc = co.loss(p["tri1"], p["tri2"], lines)
        half = co.loss(p["tri1"], p["tri2"], lines[:600])
        assert orc.n_selected == half.n_selected                  # the premise: rank 1 contributes nothing
        for o in res:
            assert float(o[2]) == orc.median and abs(o[0].item() - orc.loss) <= REL_TOL * orc.loss
    finally:
        torch.cuda.synchronize()
        for h in comms:
            L.rrl_comm_destroy(h)


def test_peer_allreduce_f64(rrl):
    """rrl_comm_allreduce_f64 between 3 same-process ranks: rank-ordered sums, identical on every rank, repeated calls"""
    L = rrl._native.lib()
    R = 3
    comms = _make_comms(L, R, 64)
    streams = [torch.cuda.Stream() for _ in range(R)]
    try:
        for rep, n in enumerate((6, 12, 511)):
            vals = [torch.arange(n, dtype=torch.float64, device="cuda") * (r + 1) + 0.25 * rep for r in range(R)]
            torch.cuda.synchronize()
            for r in range(R):
                assert L.rrl_comm_allreduce_f64(comms[r], vals[r].data_ptr(), n, streams[r].cuda_stream) == 0
            torch.cuda.synchronize()
            want = torch.arange(n, dtype=torch.float64, device="cuda") * 6 + 0.75 * rep
            for v in vals:
                assert torch.equal(v, want)
        assert L.rrl_comm_allreduce_f64(comms[0], vals[0].data_ptr(), 513, None) == -1
    finally:
        torch.cuda.synchronize()
        for h in comms:
            L.rrl_comm_destroy(h)


def test_world_size_one_process_group_through_the_public_entry_points(rrl):
    """rrl_b200.dist.line_sharded_loss / line_sharded_twist_loss over NCCL with one rank, with the NCCL protocol and with
    the peer exchange (PeerComm), against the unsharded call and the oracle"""
    import torch.distributed as dist
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29641")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        p = synth.make_pair(670, 800, 2600)
        orc = co.loss(p["tri1"], p["tri2"], p["lines"])
        loss1, med1, grad1 = _single(rrl, p)
        comm = rrl.dist.PeerComm.create(2600)
        assert comm is not None and comm.world == 1
        for cm in (None, comm):
            t1 = torch.from_numpy(p["tri1"]).cuda().requires_grad_(True)
            t2 = torch.from_numpy(p["tri2"]).cuda(); ln = torch.from_numpy(p["lines"]).cuda()
            loss, status, med = rrl.dist.line_sharded_loss(t1, t2, ln, comm=cm)
            loss.sum().backward()
            assert loss.item() == loss1 and float(med) == orc.median and int(status) == 0
            assert _rel(t1.grad.cpu().numpy(), orc.grad1) <= REL_TOL
            # pose-space reduction: twist -> transform -> sharded loss -> 6 floats
            tw = torch.tensor([0.01, -0.02, 0.015, 0.03, -0.01, 0.02], device="cuda", requires_grad=True)
            l2, _, _ = rrl.dist.line_sharded_twist_loss(tw, torch.from_numpy(p["tri1"]).cuda(), t2, ln, comm=cm)
            l2.sum().backward()
            moved = rrl.se3_apply(tw.detach().reshape(1, 6), torch.from_numpy(p["tri1"]).cuda().reshape(1, -1, 3)).reshape(-1, 9)
            o2 = co.loss(moved.cpu().numpy(), p["tri2"], p["lines"])
            gt = co.se3_backward(tw.detach().cpu().numpy(), p["tri1"].reshape(-1, 3), o2.grad1.reshape(-1, 3))
            assert abs(l2.item() - o2.loss) <= REL_TOL * o2.loss and _rel(tw.grad.cpu().numpy(), gt) <= REL_TOL
        comm.close()
    finally:
        if created:
            dist.destroy_process_group()


def test_stale_workspace_is_inert(rrl):
    """backward / export / shard stages on a workspace that holds no forward of the stated geometry read its header on
    the device and leave zero gradients, zero counts, -1 slots (include/rrl_b200.h, RRL_ERR_STATE note)"""
    L = rrl._native.lib()
    nf, nl = 300, 500
    wsb = L.rrl_workspace_bytes(1, nf, nf, nl)
    ws = torch.randint(0, 255, (wsb,), dtype=torch.uint8, device="cuda")          # garbage
    g1 = torch.full((nf, 9), 7.0, device="cuda"); go = torch.ones(1, device="cuda")
    assert L.rrl_loss_backward(ws.data_ptr(), wsb, go.data_ptr(), 1, nf, nf, nl, g1.data_ptr(), None, None) == 0
    counts = torch.full((nl,), 9, dtype=torch.int32, device="cuda"); hits = torch.zeros(nl, 5, dtype=torch.int32, device="cuda")
    assert L.rrl_loss_export_hits(ws.data_ptr(), wsb, 1, nf, nf, nl, 1, counts.data_ptr(), hits.data_ptr(), None) == 0
    c18 = torch.full((18,), 5, dtype=torch.int64, device="cuda")
    assert L.rrl_shard_counts(ws.data_ptr(), wsb, nf, nf, nl, c18.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert not g1.any() and not counts.any() and bool((hits == -1).all()) and not c18.any()
    # a forward of ANOTHER geometry in the same memory is refused the same way
    p = synth.make_pair(680, nf, nl)
    t1 = torch.from_numpy(p["tri1"]).cuda(); t2 = torch.from_numpy(p["tri2"]).cuda(); ln = torch.from_numpy(p["lines"]).cuda()
    loss = torch.zeros(1, device="cuda")
    assert L.rrl_loss_forward(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), 1, nf, nf, nl, 1, 1, 5, 5, ws.data_ptr(), wsb,
                              loss.data_ptr(), None, None, None, None) == 0
    small = L.rrl_workspace_bytes(1, nf, nf, nl - 100)
    g1.fill_(7.0)
    assert L.rrl_loss_backward(ws.data_ptr(), small, go.data_ptr(), 1, nf, nf, nl - 100, g1.data_ptr(), None, None) == 0
    torch.cuda.synchronize()
    assert not g1.any()
    assert L.rrl_loss_backward(ws.data_ptr(), wsb, go.data_ptr(), 1, nf, nf, nl, g1.data_ptr(), None, None) == 0
    torch.cuda.synchronize()
    assert _rel(g1.cpu().numpy(), co.loss(p["tri1"], p["tri2"], p["lines"]).grad1) <= REL_TOL


@pytest.mark.parametrize("R,n,rscale", [(2, 5000, 0.5), (3, 4097, 1.0), (8, 20000, 2.5)])
def test_candidate_sharded_sampler_reproduces_the_unsharded_line_set(rrl, R, n, rscale):
    """R ranks each evaluate the candidate chunks r (mod R), sum their per-chunk accepted counts (here: a
    device-side addition in place of the all-reduce) and keep their own rows.  The union of the ranks' rows must be EXACTLY
    the N rows of the unsharded sampler (same Philox stream): same accepted lines bit for bit, same number of all-zero
    rows -- for a geometry that fills all rows early, one that fills them late, and one that leaves rows unfilled."""
    p = synth.make_pair(690 + R, 800, 64)
    v1 = torch.from_numpy(p["tri1"][:, :3]).cuda()[None].contiguous()
    v2 = torch.from_numpy(p["tri2"][:, :3]).cuda()[None].contiguous()
    lo2, hi2 = v2[0].min(0)[0], v2[0].max(0)[0]
    radius = ((hi2 - lo2).norm() * rscale).reshape(1)
    center = v2[0].mean(0)[None]
    ref, ref_filled = rrl.sample_lines(radius, center, n, v1, v2, seed=21, offset=5)
    ref = ref[0].cpu().numpy(); ref_filled = int(ref_filled[0])
    ranks = [rrl.dist.ShardedSampler(n, rank=r, world=R) for r in range(R)]
    local = [s.local_counts(radius, center, v1, v2, seed=21, offset=5) for s in ranks]
    total = torch.stack(local).sum(0).to(torch.int32)
    rows, filled = [], set()
    for s, lc in zip(ranks, local):
        mine, f = s.place(lc, total)
        rows.append(mine.cpu().numpy()); filled.add(f)
    assert filled == {ref_filled}
    got = np.concatenate(rows)
    assert got.shape == (n, 6)
    key = lambda a: sorted(map(bytes, np.ascontiguousarray(a)))
    assert key(got) == key(ref)                                  # the same multiset of rows, bit for bit
    assert int((~got.any(1)).sum()) == n - ref_filled
    if rscale > 2:
        assert ref_filled < n                                    # the premise of the third case: rows stay unfilled
    # and the loss of the sharded line set equals the loss of the reference set (order independence, fixed-point sums)
    t1 = torch.from_numpy(p["tri1"]).cuda()[None]; t2 = torch.from_numpy(p["tri2"]).cuda()[None]
    a = rrl.intersected_line_loss(t1, t2, torch.from_numpy(ref).cuda()[None])
    b = rrl.intersected_line_loss(t1, t2, torch.from_numpy(got).cuda()[None])
    assert torch.equal(a, b)
