"""Parity of the sm_100a path against the oracles and the reference-minted golden vectors.  Needs a GPU.

Bars (north star): per-line selected triplet indices bit-exact; tests within 1 ulp of the threshold are counted
and reported (`band`); loss and gradients within 1e-5 relative (Frobenius for tensors)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import golden, hits_to_lists
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


def _run(rrl, tri1, tri2, lines, window=(1, 1, 5, 5)):
    t1 = torch.from_numpy(tri1).cuda().reshape(1, -1, 9).requires_grad_(True)
    t2 = torch.from_numpy(tri2).cuda().reshape(1, -1, 9).requires_grad_(True)
    ln = torch.from_numpy(lines).cuda().reshape(1, -1, 6)
    loss, info = rrl.intersected_line_loss(t1, t2, ln, window, return_info=True)
    loss.sum().backward()
    c1, h1 = info.hits(1)
    c2, h2 = info.hits(2)
    return dict(loss=loss.item(), status=int(info.status[0]), median=float(info.median[0]), stats=info.stats[0].cpu().numpy(),
                counts1=c1[0].cpu().numpy(), hits1=h1[0].cpu().numpy(), counts2=c2[0].cpu().numpy(), hits2=h2[0].cpu().numpy(),
                grad1=t1.grad[0].cpu().numpy(), grad2=t2.grad[0].cpu().numpy())


def _check_against_oracle(out, orc, check_grad=True):
    assert np.array_equal(out["counts1"], orc.counts1)
    assert np.array_equal(out["counts2"], orc.counts2)
    for cnt, mine, ref in ((orc.counts1, out["hits1"], orc.hits1), (orc.counts2, out["hits2"], orc.hits2)):
        keep = cnt <= co.CAP                      # beyond the cap only the count is defined
        assert np.array_equal(mine[keep], ref[keep])
    assert int(out["stats"][0]) == orc.n_selected and int(out["stats"][1]) == orc.n_entries
    assert int(out["stats"][2]) == orc.n_combos
    assert int(out["stats"][5]) == orc.band        # candidates cover every test within 1 ulp of its threshold
    if orc.status == 1:
        assert out["status"] & 1 and out["loss"] == 0.0
        return
    assert out["median"] == orc.median             # exact selection
    assert abs(out["loss"] - orc.loss) <= REL_TOL * abs(orc.loss)
    if check_grad:
        assert _rel(out["grad1"], orc.grad1) <= REL_TOL
        assert _rel(out["grad2"], orc.grad2) <= REL_TOL


@pytest.mark.parametrize("name", ["demo_step0", "synth_sphere", "synth_ragged", "synth_rpm_like", "synth_window", "synth_short_dirs"])
def test_golden_cases(rrl, name):
    g = golden(name)
    w = tuple(int(v) for v in g["krange"])
    out = _run(rrl, g["tri1"], g["tri2"], g["lines"], w)
    orc = co.loss(g["tri1"], g["tri2"], g["lines"], *w)
    _check_against_oracle(out, orc)
    # ... and directly against what the unmodified reference produced
    if out["stats"][5] == 0:
        assert np.array_equal(out["counts1"], g["ref_counts1"]) and np.array_equal(out["counts2"], g["ref_counts2"])
        for counts, nz, hits in ((g["ref_counts1"], g["ref_hits1"], out["hits1"]), (g["ref_counts2"], g["ref_hits2"], out["hits2"])):
            for l, lst in enumerate(hits_to_lists(counts, nz)):
                if len(lst) <= co.CAP:
                    assert list(hits[l][:len(lst)]) == lst
    assert abs(out["loss"] - float(g["ref_loss"][0])) <= REL_TOL * float(g["ref_loss"][0])
    assert _rel(out["grad1"], g["ref_grad1"]) <= REL_TOL
    assert _rel(out["grad2"], g["ref_grad2"]) <= REL_TOL


def test_empty_selection(rrl):
    g = golden("synth_empty")
    out = _run(rrl, g["tri1"], g["tri2"], g["lines"])
    assert out["status"] & 1 and out["loss"] == 0.0
    assert not out["grad1"].any() and not out["grad2"].any()
    args = (1, 1, 5, 5, torch.from_numpy(g["tri1"]).cuda()[None], torch.from_numpy(g["tri2"]).cuda()[None],
            torch.from_numpy(g["lines"]).cuda()[None], "cuda")
    # default: asynchronous, an empty pair is a zero loss with the status bit in the side channel
    res = rrl.loss.cal_loss_intersection_batch_whole_median_pts_lines(*args)
    assert res.shape == (1,) and res.item() == 0.0 and int(rrl.loss.last_info.status[0]) & 1
    rrl.loss.STRICT_EMPTY_RETURN = True             # opt-in: the reference's return value (one host read per call)
    try:
        res = rrl.loss.cal_loss_intersection_batch_whole_median_pts_lines(*args)
    finally:
        rrl.loss.STRICT_EMPTY_RETURN = False
    assert res == (None, None, None)                # loss.py:232


@pytest.mark.parametrize("seed,nf,nl,kw", [
    (101, 64, 257, {}),                                        # smaller than one point tile / one line tile
    (102, 1, 40, {}),                                          # a single triplet
    (103, 1024, 3000, {"shape": "torus"}),
    (104, 1500, 2049, {"nf2": 700, "zero_frac": 0.3}),         # ragged clouds, many all-zero (unfilled) lines
    (105, 2048, 1500, {"noise": 0.01, "outlier_frac": 0.1, "keep_frac": 0.7, "radius_scale": 1.0}),
    (106, 5000, 1024, {"shape": "box"}),                       # several point tiles per CTA
])
def test_seeded_synthetic_vs_oracle(rrl, seed, nf, nl, kw):
    if nf == 1:
        rng = np.random.default_rng(seed)
        tri = rng.normal(size=(1, 9)).astype(np.float32) * 0.05
        p = dict(tri1=tri, tri2=tri + 0.01, lines=synth.chord_lines(rng, nl, 0.5, np.zeros(3), -np.ones(3) * .1, np.ones(3) * .1, -np.ones(3) * .1, np.ones(3) * .1))
    else:
        p = synth.make_pair(seed, nf, nl, **kw)
    out = _run(rrl, p["tri1"], p["tri2"], p["lines"])
    orc = co.loss(p["tri1"], p["tri2"], p["lines"])
    _check_against_oracle(out, orc)


def test_unnormalised_and_degenerate_lines(rrl):
    """the culling bounds must hold for |u| != 1 (the reference never checks), for all-zero rows and for duplicates"""
    p = synth.make_pair(109, 900, 1600, zero_frac=0.2)
    lines = p["lines"].copy()
    rng = np.random.default_rng(5)
    lines[:300, :3] *= rng.uniform(1.0, 1.01, size=(300, 1)).astype(np.float32)      # slightly too long
    lines[300:500, :3] *= rng.uniform(0.3, 1.0, size=(200, 1)).astype(np.float32)    # too short
    lines[500:520, :3] *= 3.0                                                         # far too long
    lines[520:560] = lines[0]                                                         # duplicates
    out = _run(rrl, p["tri1"], p["tri2"], lines)
    _check_against_oracle(out, co.loss(p["tri1"], p["tri2"], lines))


def test_nan_points_and_lines_are_flagged_not_fatal(rrl):
    """the reference prints "Exit the systerm" and exit(0)s on a NaN distance (loss.py:89-91); here NaN triplets and NaN
    lines simply never hit (every comparison with NaN is false, as in the oracle), the status carries RRL_STATUS_NAN,
    and the rest of the pair is evaluated as usual"""
    p = synth.make_pair(113, 400, 700)
    tri1, lines = p["tri1"].copy(), p["lines"].copy()
    tri1[5, 0] = np.nan                      # point 0 of a triplet
    tri1[77, 4] = np.nan                     # point 1 of another
    lines[10, 3] = np.nan                    # a line through nowhere
    lines[11, 1] = np.nan
    out = _run(rrl, tri1, p["tri2"], lines)
    orc = co.loss(tri1, p["tri2"], lines)
    assert np.array_equal(out["counts1"], orc.counts1) and np.array_equal(out["counts2"], orc.counts2)
    keep = orc.counts1 <= co.CAP
    assert np.array_equal(out["hits1"][keep], orc.hits1[keep])
    assert out["counts1"][10] == 0 and out["counts1"][11] == 0 and not np.isin([5, 77], out["hits1"]).any()
    assert out["status"] & 2                                           # RRL_STATUS_NAN
    assert np.isfinite(out["loss"]) and abs(out["loss"] - orc.loss) <= REL_TOL * abs(orc.loss)
    assert np.isfinite(out["grad1"]).all()


def test_clustered_cloud_with_duplicate_points(rrl):
    """many coincident triplets (zero-radius nodes, zero thresholds) and a tight cluster next to a sparse shell"""
    rng = np.random.default_rng(11)
    a = synth.surface_points(rng, 500, "sphere")
    bpts = (rng.normal(size=(300, 3)) * 0.02 + np.array([1.0, 2.0, 0.5])).astype(np.float32)
    dup = np.repeat(a[:20], 5, axis=0)
    src = np.concatenate([a, bpts, dup]).astype(np.float32)
    tri1 = synth.knn_triplets(src)
    tri2 = synth.knn_triplets((src[::-1] * np.float32(1.01)).copy())
    lo, hi = src.min(0), src.max(0)
    lines = synth.chord_lines(rng, 2500, 0.7 * float(np.linalg.norm(hi - lo)), src.mean(0), lo, hi, lo, hi)
    out = _run(rrl, tri1, tri2, lines)
    _check_against_oracle(out, co.loss(tri1, tri2, lines))


def test_large_coordinates_keep_exactness(rrl):
    """the filter's guard band scales with (|p| + |x0|)^2: translate everything far from the origin"""
    p = synth.make_pair(107, 600, 1500)
    shift = np.array([7.0, -9.0, 5.0], np.float32)
    tri1 = (p["tri1"].reshape(-1, 3) + shift).reshape(-1, 9)
    tri2 = (p["tri2"].reshape(-1, 3) + shift).reshape(-1, 9)
    lines = p["lines"].copy()
    lines[:, 3:] += shift
    out = _run(rrl, tri1, tri2, lines)
    _check_against_oracle(out, co.loss(tri1, tri2, lines))


def test_sorted_and_unsorted_node_variants_agree(rrl):
    """Morton-sorted bounding-sphere nodes vs nodes in input order: different candidate sets, identical results"""
    p = synth.make_pair(108, 1024, 4000)
    L = rrl._native.lib()
    try:
        L.rrl_debug_set_dense_variant(0)
        a = _run(rrl, p["tri1"], p["tri2"], p["lines"])
    finally:
        L.rrl_debug_set_dense_variant(1)
    b = _run(rrl, p["tri1"], p["tri2"], p["lines"])
    for c, h in (("counts1", "hits1"), ("counts2", "hits2")):
        assert np.array_equal(a[c], b[c])
        keep = a[c] <= co.CAP                      # beyond the cap the kept subset is arbitrary
        assert np.array_equal(a[h][keep], b[h][keep])
    assert a["loss"] == b["loss"] and a["median"] == b["median"]


def test_bruteforce_kernel_agrees_with_filtered_path(rrl):
    """every (line, triplet) tested exactly on the device (the reference's own formulation) vs the filtered pipeline"""
    p = synth.make_pair(131, 1500, 5000, zero_frac=0.1)
    L = rrl._native.lib()
    try:
        L.rrl_debug_set_param(5, 1)
        a = _run(rrl, p["tri1"], p["tri2"], p["lines"])
    finally:
        L.rrl_debug_set_param(5, 0)
    b = _run(rrl, p["tri1"], p["tri2"], p["lines"])
    for c, h in (("counts1", "hits1"), ("counts2", "hits2")):
        assert np.array_equal(a[c], b[c])
        keep = a[c] <= co.CAP
        assert np.array_equal(a[h][keep], b[h][keep])
    assert a["loss"] == b["loss"] and a["median"] == b["median"]
    _check_against_oracle(a, co.loss(p["tri1"], p["tri2"], p["lines"]))


def test_candidate_queue_overflow_runs_the_exact_test_in_place(rrl):
    """hundreds of triplets on every line: far more exact candidates than the queue holds
    (12 per line), so the warps that find it full run the exact test themselves; results stay exact"""
    rng = np.random.default_rng(17)
    shell = synth.surface_points(rng, 400, "sphere")
    # 600 points 0.02 apart on a straight segment (thr ~ 0.023 > sqrt(2e-4)) and lines running along it
    seg = np.stack([np.arange(600) * 0.02 - 6.0, np.zeros(600), np.zeros(600)], 1) + rng.normal(size=(600, 3)) * 1e-4
    src = np.concatenate([shell, seg]).astype(np.float32)
    tri1 = synth.knn_triplets(src)
    tri2 = synth.knn_triplets((src * np.float32(0.999)).copy())
    d = np.array([1.0, 0, 0]) + rng.normal(size=(1200, 3)) * 2e-4
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x0 = rng.normal(size=(1200, 3)) * 1e-3 + d * rng.uniform(-1, 1, size=(1200, 1))
    lines = np.concatenate([d, x0], 1).astype(np.float32)
    lines[::7] = synth.make_pair(5, 300, 1200)["lines"][::7]                          # and some ordinary lines
    out = _run(rrl, tri1, tri2, lines)
    orc = co.loss(tri1, tri2, lines)
    assert orc.counts1.mean() > 24                                                    # the premise: the queue must overflow
    _check_against_oracle(out, orc)
    # together with an ordinary pair in one batch
    q = synth.make_pair(132, 1000, 1200)
    t1 = torch.from_numpy(np.stack([tri1, q["tri1"]])).cuda()
    t2 = torch.from_numpy(np.stack([tri2, q["tri2"]])).cuda()
    ln = torch.from_numpy(np.stack([lines, q["lines"]])).cuda()
    loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
    oq = co.loss(q["tri1"], q["tri2"], q["lines"])
    c1, _ = info.hits(1)
    assert np.array_equal(c1[0].cpu().numpy(), orc.counts1) and np.array_equal(c1[1].cpu().numpy(), oq.counts1)
    assert abs(float(loss[1]) - oq.loss) <= REL_TOL * oq.loss


def test_batched_equals_per_pair_and_is_permutation_invariant(rrl):
    pairs = [synth.make_pair(200 + i, 512, 2000) for i in range(5)]
    t1 = torch.from_numpy(np.stack([p["tri1"] for p in pairs])).cuda().requires_grad_(True)
    t2 = torch.from_numpy(np.stack([p["tri2"] for p in pairs])).cuda()
    ln = torch.from_numpy(np.stack([p["lines"] for p in pairs])).cuda()
    loss = rrl.intersected_line_loss(t1, t2, ln)
    (loss * torch.arange(1, 6, device="cuda")).sum().backward()
    for i, p in enumerate(pairs):
        orc = co.loss(p["tri1"], p["tri2"], p["lines"])
        assert abs(loss[i].item() - orc.loss) <= REL_TOL * orc.loss
        assert _rel(t1.grad[i].cpu().numpy(), (i + 1) * orc.grad1) <= REL_TOL
    # shuffling the lines of a pair must not change its loss at all: sums are order-independent fixed point
    perm = torch.randperm(ln.shape[1], generator=torch.Generator().manual_seed(0)).cuda()
    loss2 = rrl.intersected_line_loss(t1.detach(), t2, ln[:, perm])
    assert torch.equal(loss2, loss.detach())


def test_raw_c_abi_calls(rrl):
    """the same path through bare ctypes: raw device pointers, explicit workspace, explicit stream"""
    L = rrl._native.lib()
    p = synth.make_pair(300, 700, 1800)
    t1 = torch.from_numpy(p["tri1"]).cuda(); t2 = torch.from_numpy(p["tri2"]).cuda(); ln = torch.from_numpy(p["lines"]).cuda()
    nf1, nf2, nl = t1.shape[0], t2.shape[0], ln.shape[0]
    wsb = L.rrl_workspace_bytes(1, nf1, nf2, nl)
    assert wsb > 0
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    loss = torch.zeros(1, device="cuda"); status = torch.zeros(1, dtype=torch.int32, device="cuda")
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    assert L.rrl_loss_forward(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), 1, nf1, nf2, nl, 1, 1, 5, 5, ws.data_ptr(), wsb,
                              loss.data_ptr(), status.data_ptr(), None, None, s.cuda_stream) == 0
    g1 = torch.empty(nf1, 9, device="cuda"); go = torch.ones(1, device="cuda")
    assert L.rrl_loss_backward(ws.data_ptr(), wsb, go.data_ptr(), 1, nf1, nf2, nl, g1.data_ptr(), None, s.cuda_stream) == 0
    s.synchronize()
    orc = co.loss(p["tri1"], p["tri2"], p["lines"])
    assert abs(loss.item() - orc.loss) <= REL_TOL * orc.loss
    assert _rel(g1.cpu().numpy(), orc.grad1) <= REL_TOL
    # error conventions: status codes, never exit
    assert L.rrl_loss_forward(None, t2.data_ptr(), ln.data_ptr(), 1, nf1, nf2, nl, 1, 1, 5, 5, ws.data_ptr(), wsb,
                              loss.data_ptr(), None, None, None, None) == -1
    assert L.rrl_loss_forward(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), 1, nf1, nf2, nl, 0, 1, 5, 5, ws.data_ptr(), wsb,
                              loss.data_ptr(), None, None, None, None) == -1
    assert L.rrl_loss_forward(t1.data_ptr(), t2.data_ptr(), ln.data_ptr(), 1, nf1, nf2, nl, 1, 1, 5, 5, ws.data_ptr(), 1024,
                              loss.data_ptr(), None, None, None, None) == -2
    with pytest.raises(ValueError):
        rrl.intersected_line_loss(t1, t2[None], ln[None])
    with pytest.raises(rrl.NativeError):
        rrl.intersected_line_loss(t1.cpu()[None], t2.cpu()[None], ln.cpu()[None])      # no CPU fallback


@pytest.mark.parametrize("subbatches", [1, 2, 5])
def test_host_buffer_entry(rrl, subbatches, monkeypatch):
    """host pointers in, host pointers out; the call is pipelined over `subbatches` streams (uneven split for 2)"""
    monkeypatch.setenv("RRL_HOST_SUBBATCHES", str(subbatches))
    L = rrl._native.lib()
    pairs = [synth.make_pair(400 + i, 300, 1000) for i in range(5)]
    tri1 = np.ascontiguousarray(np.stack([p["tri1"] for p in pairs])); tri2 = np.ascontiguousarray(np.stack([p["tri2"] for p in pairs]))
    lines = np.ascontiguousarray(np.stack([p["lines"] for p in pairs]))
    ctx = C.c_void_p()
    assert L.rrl_host_create(5, 300, 300, 1000, 0, C.byref(ctx)) == 0
    try:
        assert L.rrl_host_subbatches(ctx) == subbatches
        loss = np.zeros(5, np.float32); status = np.zeros(5, np.int32); g1 = np.zeros_like(tri1)
        assert L.rrl_host_loss_fwd_bwd(ctx, tri1.ctypes.data, tri2.ctypes.data, lines.ctypes.data, 1, 1, 5, 5,
                                       loss.ctypes.data, status.ctypes.data, g1.ctypes.data) == 0
    finally:
        L.rrl_host_destroy(ctx)
    for i, p in enumerate(pairs):
        orc = co.loss(p["tri1"], p["tri2"], p["lines"])
        assert abs(loss[i] - orc.loss) <= REL_TOL * orc.loss
        assert _rel(g1[i], orc.grad1) <= REL_TOL


def test_host_buffer_pipelined_submit_wait(rrl):
    """double-buffered form: two evaluations in flight, results come back per ticket; a third submit is refused"""
    L = rrl._native.lib()
    sets = []
    for s in range(3):
        pairs = [synth.make_pair(900 + 10 * s + i, 256, 800) for i in range(4)]
        sets.append((pairs,) + tuple(np.ascontiguousarray(np.stack([p[k] for p in pairs])) for k in ("tri1", "tri2", "lines")))
    ctx = C.c_void_p()
    assert L.rrl_host_create(4, 256, 256, 800, 0, C.byref(ctx)) == 0
    try:
        assert L.rrl_host_slots(ctx) == 2
        tk = [C.c_int(-1) for _ in sets]
        out = [(np.zeros(4, np.float32), np.zeros(4, np.int32), np.zeros_like(s[1])) for s in sets]
        sub = lambda i, grad: L.rrl_host_submit(ctx, sets[i][1].ctypes.data, sets[i][2].ctypes.data, sets[i][3].ctypes.data,
                                                1, 1, 5, 5, grad, C.byref(tk[i]))
        wait = lambda i, grad: L.rrl_host_wait(ctx, tk[i].value, out[i][0].ctypes.data, out[i][1].ctypes.data,
                                               out[i][2].ctypes.data if grad else None)
        assert sub(0, 1) == 0 and sub(1, 1) == 0
        assert sub(2, 1) == -4                                   # RRL_ERR_STATE: both slots in flight
        assert wait(0, True) == 0
        assert L.rrl_host_wait(ctx, tk[0].value, out[0][0].ctypes.data, None, None) == -4      # already drained
        assert sub(2, 0) == 0                                    # reuses slot 0 while evaluation 1 is still in flight
        assert wait(1, True) == 0
        assert L.rrl_host_wait(ctx, tk[2].value, out[2][0].ctypes.data, None, out[2][2].ctypes.data) == -1   # no gradient requested
        assert wait(2, False) == 0
    finally:
        L.rrl_host_destroy(ctx)
    for i, s in enumerate(sets):
        for b, p in enumerate(s[0]):
            orc = co.loss(p["tri1"], p["tri2"], p["lines"])
            assert abs(out[i][0][b] - orc.loss) <= REL_TOL * orc.loss
            if i < 2:
                assert _rel(out[i][2][b], orc.grad1) <= REL_TOL


def test_full_size_dcp_batch_properties(rrl):
    """BASELINE config 2 at full size (32 x 1024 triplets x 15000 lines): the oracle checks EVERY pair completely (hit
    lists of both clouds, median, loss), three of them with gradients; every pair also through size-independent
    properties."""
    B, nf, nl = 32, 1024, 15000
    pairs = [synth.make_pair(2000 + i, nf, nl) for i in range(B)]
    t1 = torch.from_numpy(np.stack([p["tri1"] for p in pairs])).cuda()
    t2 = torch.from_numpy(np.stack([p["tri2"] for p in pairs])).cuda()
    ln = torch.from_numpy(np.stack([p["lines"] for p in pairs])).cuda()
    t1.requires_grad_(True)
    loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
    loss.sum().backward()
    _g1 = t1.grad.cpu().numpy()
    t1 = t1.detach()
    loss = loss.detach()
    c1, h1 = info.hits(1)
    c2, h2 = info.hits(2)
    c1n, h1n, c2n, h2n = (x.cpu().numpy() for x in (c1, h1, c2, h2))
    for i in range(B):
        orc = co.loss(pairs[i]["tri1"], pairs[i]["tri2"], pairs[i]["lines"], want_grad=i in (0, 13, 31))
        assert np.array_equal(c1n[i], orc.counts1) and np.array_equal(c2n[i], orc.counts2)
        keep = orc.counts1 <= co.CAP
        assert np.array_equal(h1n[i][keep], orc.hits1[keep])
        keep2 = orc.counts2 <= co.CAP
        assert np.array_equal(h2n[i][keep2], orc.hits2[keep2])
        assert float(info.median[i]) == orc.median
        assert abs(loss[i].item() - orc.loss) <= REL_TOL * orc.loss
        assert int(info.stats[i, 5]) == orc.band
        if i in (0, 13, 31):
            assert _rel(_g1[i], orc.grad1) <= REL_TOL
    # swapping the clouds of every pair transposes every D matrix: same medians, same losses (symmetry of loss.py:227-229)
    loss_sw, info_sw = rrl.intersected_line_loss(t2, t1, ln, return_info=True)
    assert torch.equal(info_sw.median, info.median)
    assert torch.allclose(loss_sw, loss, rtol=1e-6, atol=0)
    # hit lists are sorted and inside range
    h = h1.cpu().numpy(); c = c1.cpu().numpy()
    for s in range(co.CAP - 1):
        ok = (c > s + 1)
        assert np.all(h[..., s][ok] < h[..., s + 1][ok])
    assert h.max() < nf and int(info.stats[:, 6].sum()) == 0


def _bruteforce_all(rrl, t1, t2, ln):
    L = rrl._native.lib()
    try:
        L.rrl_debug_set_param(5, 1)
        loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
        out = [x.cpu().numpy() for x in info.hits(1) + info.hits(2)] + [loss.cpu().numpy(), info.median.cpu().numpy()]
    finally:
        L.rrl_debug_set_param(5, 0)
    return out


@pytest.mark.parametrize("name,B,nf,nl,kw,n_oracle", [
    ("rpm", 64, 2048, 10000, dict(radius_scale=1.0, noise=0.01, outlier_frac=0.1, keep_frac=0.7), 6),
    ("fmr", 128, 1024, 15000, dict(radius_scale=0.5), 6),
])
def test_full_size_rpm_and_fmr_batches(rrl, name, B, nf, nl, kw, n_oracle):
    """BASELINE configs 3 and 4 at full size (64 x 2048 x 10000 with noise / outliers / partial overlap; 128 x 1024 x 15000):
    distinct pairs, `n_oracle` of them checked completely against the C oracle (hit lists, median, loss, gradient), ALL of
    them against the on-device brute-force kernel (every (line, triplet) tested with the reference's own arithmetic)."""
    base = [synth.make_pair(3000 + 131 * B + i, nf, nl, **kw) for i in range(16)]
    rng = np.random.default_rng(B)
    t1l, t2l, lnl = [], [], []
    for i in range(B):
        p = base[i % 16]
        if i < 16:
            Rm, t = np.eye(3), np.zeros(3)
        else:                                              # rigidly moved copies: distinct bits, same statistics
            Rm, t = synth.random_rotation(rng, 180.0), rng.uniform(-0.5, 0.5, 3)
        mv = lambda x: (x.reshape(-1, 3).astype(np.float64) @ Rm.T + t).astype(np.float32)
        t1l.append(mv(p["tri1"]).reshape(-1, 9)); t2l.append(mv(p["tri2"]).reshape(-1, 9))
        l64 = p["lines"].astype(np.float64)
        lnl.append(np.concatenate([l64[:, :3] @ Rm.T, l64[:, 3:] @ Rm.T + t], 1).astype(np.float32))
    tri1, tri2, lines = np.stack(t1l), np.stack(t2l), np.stack(lnl)
    t1 = torch.from_numpy(tri1).cuda().requires_grad_(True)
    t2 = torch.from_numpy(tri2).cuda(); ln = torch.from_numpy(lines).cuda()
    loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
    loss.sum().backward()
    c1, h1 = (x.cpu().numpy() for x in info.hits(1))
    c2, h2 = (x.cpu().numpy() for x in info.hits(2))
    g1 = t1.grad.cpu().numpy()
    for i in list(range(n_oracle - 2)) + [B // 2, B - 1]:
        orc = co.loss(tri1[i], tri2[i], lines[i])
        assert np.array_equal(c1[i], orc.counts1) and np.array_equal(c2[i], orc.counts2)
        assert np.array_equal(h1[i][orc.counts1 <= co.CAP], orc.hits1[orc.counts1 <= co.CAP])
        assert np.array_equal(h2[i][orc.counts2 <= co.CAP], orc.hits2[orc.counts2 <= co.CAP])
        assert float(info.median[i]) == orc.median and int(info.stats[i, 5]) == orc.band
        assert abs(loss[i].item() - orc.loss) <= REL_TOL * orc.loss
        assert _rel(g1[i], orc.grad1) <= REL_TOL
    b1, bh1, b2, bh2, bloss, bmed = _bruteforce_all(rrl, t1.detach(), t2, ln)
    assert np.array_equal(c1, b1) and np.array_equal(c2, b2)
    assert np.array_equal(h1[c1 <= co.CAP], bh1[c1 <= co.CAP]) and np.array_equal(h2[c2 <= co.CAP], bh2[c2 <= co.CAP])
    assert np.array_equal(loss.detach().cpu().numpy(), bloss) and np.array_equal(info.median.cpu().numpy(), bmed)
    assert int(info.stats[:, 6].sum()) == 0 and not bool((info.status != 0).any())


def test_more_d_entries_than_the_median_cache(rrl):
    """a pair whose selected lines hold more D entries than the tail kernel caches in shared memory (49152): the median
    is then selected by sweeps over the records in global memory"""
    nf, nl = 1024, 160000
    p = synth.make_pair(4100, nf, nl)
    out = _run(rrl, p["tri1"], p["tri2"], p["lines"])
    orc = co.loss(p["tri1"], p["tri2"], p["lines"])
    assert orc.n_entries > 49152                               # the premise
    _check_against_oracle(out, orc)


def test_large_cloud_path_vs_oracle(rrl):
    """clouds above 16384 triplets take the other launch plan (radix sort, 16-triplet nodes, group -> node -> triplet
    queues): complete oracle comparison at a size the oracle still finishes in seconds"""
    p = synth.make_pair(141, 40000, 3000, nf2=33000, zero_frac=0.05)
    out = _run(rrl, p["tri1"], p["tri2"], p["lines"])
    _check_against_oracle(out, co.loss(p["tri1"], p["tri2"], p["lines"]))


@pytest.mark.parametrize("scale,shift", [(1.0, (7.0, -9.0, 5.0)), (60.0, (0.0, 0.0, 0.0)), (400.0, (900.0, -300.0, 100.0)), (0.02, (0.0, 0.0, 0.0))])
def test_compressed_records_at_extreme_scales(rrl, scale, shift):
    """the large-cloud modes read 16-bit offsets + half-precision cuts / radii: clouds far from the origin (offsets
    tiny against |p|), scaled up until a cut overflows half precision (-> +inf: always a candidate), scaled down into the regime
    where thr^2 < 2e-4 and nothing can hit; the error of every record is measured and folded in, so the indices stay exact"""
    p = synth.make_pair(143, 20000, 2500, nf2=17000, zero_frac=0.05)
    sh = np.array(shift, np.float32)
    tri1 = (p["tri1"].reshape(-1, 3) * np.float32(scale) + sh).reshape(-1, 9)
    tri2 = (p["tri2"].reshape(-1, 3) * np.float32(scale) + sh).reshape(-1, 9)
    lines = p["lines"].copy()
    lines[:, 3:] = lines[:, 3:] * np.float32(scale) + sh * (np.abs(lines[:, 3:]).sum(1, keepdims=True) > 0)     # all-zero rows stay all-zero
    out = _run(rrl, tri1, tri2, lines)
    _check_against_oracle(out, co.loss(tri1, tri2, lines), check_grad=scale <= 60.0)


@pytest.mark.parametrize("seed", [201, 202, 203, 204, 205, 206])
def test_randomised_large_cloud_sweep_against_the_bruteforce_kernel(rrl, seed):
    """random sizes around the super-node threshold, shapes, scales, offsets, unnormalised / zero / duplicated lines: the filtered
    pipeline (compressed records, per-node slack, super-node level) must report exactly the indices of the device's brute-force
    formulation (every (line, triplet) tested with the reference-order arithmetic)"""
    rng = np.random.default_rng(seed)
    nf1, nf2 = int(rng.integers(16384, 60000)), int(rng.integers(9000, 60000))
    nl = int(rng.integers(1500, 6000))
    p = synth.make_pair(seed, nf1, nl, nf2=nf2, shape=["sphere", "torus", "box"][seed % 3],
                        zero_frac=float(rng.uniform(0, 0.1)), noise=float(rng.choice([0.0, 0.002])), keep_frac=float(rng.choice([1.0, 0.7])))
    scale = np.float32(10.0 ** rng.uniform(-0.7, 1.7))
    shift = (rng.uniform(-3, 3, 3) * scale).astype(np.float32)
    tri1 = (p["tri1"].reshape(-1, 3) * scale + shift).reshape(-1, 9)
    tri2 = (p["tri2"].reshape(-1, 3) * scale + shift).reshape(-1, 9)
    lines = p["lines"].copy()
    nz = np.abs(lines).sum(1, keepdims=True) > 0
    lines[:, 3:] = lines[:, 3:] * scale + shift * nz
    k = nl // 10
    lines[:k, :3] *= rng.uniform(0.5, 1.02, size=(k, 1)).astype(np.float32)          # |u| != 1, both sides
    lines[k:k + 20] = lines[k + 20]                                                    # duplicates
    L = rrl._native.lib()
    try:
        L.rrl_debug_set_param(5, 1)
        a = _run(rrl, tri1, tri2, lines)
    finally:
        L.rrl_debug_set_param(5, 0)
    b = _run(rrl, tri1, tri2, lines)
    for c, h in (("counts1", "hits1"), ("counts2", "hits2")):
        assert np.array_equal(a[c], b[c])
        keep = a[c] <= co.CAP
        assert np.array_equal(a[h][keep], b[h][keep])
    assert a["median"] == b["median"] and a["loss"] == b["loss"]
    assert int(a["stats"][5]) == int(b["stats"][5])                                     # the same 1-ulp band tests were seen


def test_batched_mid_and_large_clouds_share_one_sort(rrl):
    """clouds above 4096 triplets are Hilbert-sorted by ONE radix sort over every cloud of every pair (segment bits above
    a shortened curve index): three ragged pairs per batch, each compared completely with the oracle -- once below
    16384 triplets (per-node kernel) and once above (super-node level)"""
    for nf1, nf2, nl in ((6000, 4500, 1500), (20000, 17000, 1200)):
        pairs = [synth.make_pair(300 + i, nf1, nl, nf2=nf2, zero_frac=0.03) for i in range(3)]
        t1 = torch.from_numpy(np.stack([p["tri1"] for p in pairs])).cuda().requires_grad_(True)
        t2 = torch.from_numpy(np.stack([p["tri2"] for p in pairs])).cuda()
        ln = torch.from_numpy(np.stack([p["lines"] for p in pairs])).cuda()
        loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
        loss.sum().backward()
        c1, h1 = info.hits(1)
        c2, h2 = info.hits(2)
        for i, p in enumerate(pairs):
            orc = co.loss(p["tri1"], p["tri2"], p["lines"])
            assert np.array_equal(c1[i].cpu().numpy(), orc.counts1) and np.array_equal(c2[i].cpu().numpy(), orc.counts2)
            for cnt, mine, ref in ((orc.counts1, h1[i], orc.hits1), (orc.counts2, h2[i], orc.hits2)):
                keep = cnt <= co.CAP
                assert np.array_equal(mine.cpu().numpy()[keep], ref[keep])
            assert float(info.median[i]) == orc.median
            assert abs(loss[i].item() - orc.loss) <= REL_TOL * orc.loss
            assert _rel(t1.grad[i].cpu().numpy(), orc.grad1) <= REL_TOL


def test_super_node_level_with_degenerate_lines_and_without(rrl):
    """the super-node level (clouds from 16384 triplets) must stay a superset filter for |u| != 1, all-zero rows and
    duplicates, and switching it off must not change a single index"""
    p = synth.make_pair(151, 18000, 2500, nf2=16500, zero_frac=0.15)
    lines = p["lines"].copy()
    rng = np.random.default_rng(6)
    lines[:400, :3] *= rng.uniform(1.0, 1.01, size=(400, 1)).astype(np.float32)
    lines[400:700, :3] *= rng.uniform(0.3, 1.0, size=(300, 1)).astype(np.float32)
    lines[700:730, :3] *= 3.0
    lines[730:780] = lines[1]
    out = _run(rrl, p["tri1"], p["tri2"], lines)
    _check_against_oracle(out, co.loss(p["tri1"], p["tri2"], lines))
    L = rrl._native.lib()
    try:
        L.rrl_debug_set_param(7, 1)
        flat = _run(rrl, p["tri1"], p["tri2"], lines)
    finally:
        L.rrl_debug_set_param(7, 0)
    for c, h in (("counts1", "hits1"), ("counts2", "hits2")):
        assert np.array_equal(out[c], flat[c])
        keep = out[c] <= co.CAP                   # beyond the cap only the count is defined
        assert np.array_equal(out[h][keep], flat[h][keep])
    assert out["loss"] == flat["loss"] and out["median"] == flat["median"]


@pytest.mark.parametrize("nf,nl", [(900, 1500), (20000, 1200)])
def test_reused_order_gives_identical_results(rrl, nf, nl):
    """RRL_REUSE_ORDER (LossSession): the second call keeps the spatial order of the first.  Any order must give the
    same indices, median and loss -- for the same clouds, for a rigidly moved source, and even for a different pair of
    the same geometry (the order is then merely a bad one); small-cloud and large-cloud launch plans."""
    p = synth.make_pair(171, nf, nl)
    q = synth.make_pair(172, nf, nl)
    rng = np.random.default_rng(7)
    Rm = synth.random_rotation(rng, 5.0)
    moved = (p["tri1"].reshape(-1, 3).astype(np.float64) @ Rm.T + 0.02).astype(np.float32).reshape(-1, 9)
    sess = rrl.LossSession()
    for tri1, tri2, lines in ((p["tri1"], p["tri2"], p["lines"]), (p["tri1"], p["tri2"], p["lines"]),
                              (moved, p["tri2"], p["lines"]), (q["tri1"], q["tri2"], q["lines"])):
        t1 = torch.from_numpy(tri1).cuda()[None].requires_grad_(True)
        t2 = torch.from_numpy(tri2).cuda()[None]
        ln = torch.from_numpy(lines).cuda()[None]
        loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True, session=sess)
        loss.sum().backward()
        c1, h1 = (x[0].cpu().numpy() for x in info.hits(1))
        c2, h2 = (x[0].cpu().numpy() for x in info.hits(2))
        orc = co.loss(tri1, tri2, lines)
        assert np.array_equal(c1, orc.counts1) and np.array_equal(c2, orc.counts2)
        assert np.array_equal(h1[orc.counts1 <= co.CAP], orc.hits1[orc.counts1 <= co.CAP])
        assert float(info.median[0]) == orc.median
        assert abs(loss.item() - orc.loss) <= REL_TOL * orc.loss
        assert _rel(t1.grad[0].cpu().numpy(), orc.grad1) <= REL_TOL


def test_static_target_session_keeps_the_targets_records(rrl):
    """RRL_REUSE_TARGET (LossSession(static_target=True)): the target's thresholds, records and bounding spheres survive from
    call to call while the source moves and the lines change.  Every call is compared completely with the oracle: same
    lines again; a rigidly moved source; lines that reach farther out than anything before (the device-side extent check
    must rebuild the target's spheres with a larger slack); the first lines again (reuse of the rebuilt records); lines
    pulled towards the origin."""
    nf, nl = 20000, 1500
    p = synth.make_pair(181, nf, nl, nf2=17000)
    rng = np.random.default_rng(8)
    Rm = synth.random_rotation(rng, 4.0)
    moved = (p["tri1"].reshape(-1, 3).astype(np.float64) @ Rm.T + 0.015).astype(np.float32).reshape(-1, 9)
    far = p["lines"].copy()
    far[:, 3:] = far[:, 3:] + far[:, :3] * 3.0 * rng.uniform(0.5, 1.0, size=(nl, 1)).astype(np.float32)   # slide x0 along the line: same lines,
    near = p["lines"].copy()                                                                                  # larger |x0| (same hits)
    d = near[:, :3].astype(np.float64); x0 = near[:, 3:].astype(np.float64)
    near[:, 3:] = (x0 - (x0 * d).sum(1, keepdims=True) * d).astype(np.float32)                              # foot point: smallest |x0|
    sess = rrl.LossSession(static_target=True)
    t2 = torch.from_numpy(p["tri2"]).cuda()[None]
    for tri1, lines in ((p["tri1"], p["lines"]), (p["tri1"], p["lines"]), (moved, p["lines"]), (moved, far), (p["tri1"], p["lines"]),
                        (moved, near)):
        t1 = torch.from_numpy(tri1).cuda()[None].requires_grad_(True)
        ln = torch.from_numpy(lines).cuda()[None]
        loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True, session=sess)
        loss.sum().backward()
        c1, h1 = (x[0].cpu().numpy() for x in info.hits(1))
        c2, h2 = (x[0].cpu().numpy() for x in info.hits(2))
        orc = co.loss(tri1, p["tri2"], lines)
        assert np.array_equal(c1, orc.counts1) and np.array_equal(c2, orc.counts2)
        assert np.array_equal(h1[orc.counts1 <= co.CAP], orc.hits1[orc.counts1 <= co.CAP])
        assert np.array_equal(h2[orc.counts2 <= co.CAP], orc.hits2[orc.counts2 <= co.CAP])
        assert float(info.median[0]) == orc.median and abs(loss.item() - orc.loss) <= REL_TOL * orc.loss
        assert _rel(t1.grad[0].cpu().numpy(), orc.grad1) <= REL_TOL


def test_full_size_large_pair_properties(rrl):
    """BASELINE config 5 at full size (500k triplets x 100k lines): the oracle checks a sample of the lines completely,
    the on-device brute-force kernel (every (line, triplet) tested exactly, the reference's formulation) checks all of
    them, and the loss must not depend on the order of the lines"""
    nf, nl = 500000, 100000
    p = synth.make_pair(5000, nf, nl)
    t1 = torch.from_numpy(p["tri1"]).cuda()[None]
    t2 = torch.from_numpy(p["tri2"]).cuda()[None]
    ln = torch.from_numpy(p["lines"]).cuda()[None]
    loss, info = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
    c1, h1 = (x[0].cpu().numpy() for x in info.hits(1))
    c2, h2 = (x[0].cpu().numpy() for x in info.hits(2))
    assert int(info.stats[0, 6]) == 0 and np.isfinite(loss.item()) and loss.item() > 0
    # (a) a sample of the lines against the C oracle (hit lists only: the loss couples all lines through the median)
    sample = np.random.default_rng(1).choice(nl, 1024, replace=False)
    orc = co.loss(p["tri1"], p["tri2"], p["lines"][sample], want_grad=False)
    assert np.array_equal(c1[sample], orc.counts1) and np.array_equal(c2[sample], orc.counts2)
    for cnt, mine, ref in ((orc.counts1, h1[sample], orc.hits1), (orc.counts2, h2[sample], orc.hits2)):
        keep = cnt <= co.CAP
        assert np.array_equal(mine[keep], ref[keep])
    # (b) every line against the brute-force kernel
    L = rrl._native.lib()
    try:
        L.rrl_debug_set_param(5, 1)
        loss_bf, info_bf = rrl.intersected_line_loss(t1, t2, ln, return_info=True)
        b1, g1 = (x[0].cpu().numpy() for x in info_bf.hits(1))
        b2, g2 = (x[0].cpu().numpy() for x in info_bf.hits(2))
    finally:
        L.rrl_debug_set_param(5, 0)
    assert np.array_equal(c1, b1) and np.array_equal(c2, b2)
    assert np.array_equal(h1[c1 <= co.CAP], g1[c1 <= co.CAP]) and np.array_equal(h2[c2 <= co.CAP], g2[c2 <= co.CAP])
    assert loss_bf.item() == loss.item() and float(info_bf.median[0]) == float(info.median[0])
    # (c) permuting the lines permutes the hit lists and leaves median and loss bit-identical (fixed-point sums)
    perm = torch.randperm(nl, generator=torch.Generator().manual_seed(3)).cuda()
    loss_p, info_p = rrl.intersected_line_loss(t1, t2, ln[:, perm], return_info=True)
    assert float(info_p.median[0]) == float(info.median[0]) and loss_p.item() == loss.item()
    assert np.array_equal(info_p.hits(1)[0][0].cpu().numpy(), c1[perm.cpu().numpy()])
