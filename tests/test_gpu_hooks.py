"""The training hooks of the reference on the GPU: the unchanged per-pair Python loops through the drop-in module
(slice batching), the batched hook helpers (rrl_b200.hooks), FMR's Exp with the reference's ExpMap gradient, and the
differentiable Chamfer distance.  Every loss is compared with the per-pair C oracle sum; gradients with respect to the
predicted transforms are compared with the oracle's point gradient pushed through the transform in float64."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import aux_oracle as ao
from oracle import c_oracle as co
from oracle import synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


@pytest.fixture(scope="module")
def rrl():
    import rrl_b200
    return rrl_b200


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _batch(B, nf, nl, seed):
    pairs = [synth.make_pair(seed + i, nf, nl) for i in range(B)]
    rows1 = np.stack([p["tri1"].reshape(-1, 3) for p in pairs])            # (B, 3nf, 3): the loaders' points_based_neighs_*
    rows2 = np.stack([p["tri2"].reshape(-1, 3) for p in pairs])
    lines = np.stack([p["lines"] for p in pairs])
    rng = np.random.default_rng(seed)
    R = np.stack([synth.random_rotation(rng, 3.0) for _ in range(B)]).astype(np.float32)
    t = rng.uniform(-0.02, 0.02, size=(B, 3)).astype(np.float32)
    return rows1, rows2, lines, R, t


def _oracle_sum(moved_faces, rows2, lines, scale):
    """sum_j oracle loss_j * scale and the per-pair point gradients (already scaled)"""
    total, grads = 0.0, []
    for j in range(moved_faces.shape[0]):
        o = co.loss(moved_faces[j], rows2[j].reshape(-1, 9), lines[j])
        total += o.loss * scale
        grads.append(o.grad1 * scale)
    return total, np.stack(grads)


def _transform_grads(rows1, gpts, R):
    """d/dR, d/dt of sum <R p + t, g> in float64: gR[i][j] = sum g_i p_j, gt = sum g"""
    g = gpts.reshape(gpts.shape[0], -1, 3).astype(np.float64)
    p = rows1.astype(np.float64)
    return np.einsum("bni,bnj->bij", g, p), g.sum(1)


def test_dcp_hook_loop_unchanged_and_batched(rrl):
    """Train_DCP.py:252-297 literally -- torch transform, transpose/reshape, the python loop over B = 1 slices, / 5.0,
    / batch_size -- through rrl_b200.loss; the same with slice batching off; and hooks.dcp_loss"""
    B, nf, nl = 6, 400, 1500
    rows1, rows2, lines, R, t = _batch(B, nf, nl, 700)
    M = rrl.loss
    src_cf = torch.from_numpy(rows1).cuda().transpose(2, 1).contiguous()       # (B, 3, 3nf) channels first, as DCP stores it
    tar_cf = torch.from_numpy(rows2).cuda().transpose(2, 1).contiguous()
    ln = torch.from_numpy(lines).cuda()
    results = {}
    for mode in ("batched-slices", "per-pair", "hook"):
        Rp = torch.from_numpy(R).cuda().requires_grad_(True)
        tp = torch.from_numpy(t).cuda().requires_grad_(True)
        if mode == "hook":
            out = rrl.hooks.dcp_loss(src_cf, Rp, tp, tar_cf, ln)
        else:
            M.BATCH_SLICES = mode == "batched-slices"
            n0 = rrl.launch_count()
            tar_faces = tar_cf.transpose(2, 1).reshape(B, -1, 9)
            pred = (torch.matmul(Rp, src_cf) + tp.unsqueeze(2)).transpose(2, 1).reshape(B, -1, 9)    # utils.py:32-37
            acc = torch.zeros(1, device="cuda")
            for j in range(B):
                acc = acc + M.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, pred[j:j + 1, :, :], tar_faces[j:j + 1, :, :],
                                                                                 ln[j:j + 1, :, :], "cuda") / 5.0
            out = acc / B
            results[mode + "-launches"] = rrl.launch_count() - n0
            M.BATCH_SLICES = True
        assert out.shape == (1,)
        out.backward()
        results[mode] = (out.item(), Rp.grad.cpu().numpy(), tp.grad.cpu().numpy())
    # the loop with slice batching is ONE forward for the whole batch
    assert results["batched-slices-launches"] * 3 < results["per-pair-launches"]
    moved = (np.einsum("bij,bnj->bni", R.astype(np.float64), rows1.astype(np.float64)) + t[:, None].astype(np.float64))
    # oracle on the points the torch transform produced (float32 rounding of the transform is the caller's)
    moved32 = (torch.matmul(torch.from_numpy(R).cuda(), src_cf) + torch.from_numpy(t).cuda().unsqueeze(2)).transpose(2, 1).reshape(B, -1, 9).cpu().numpy()
    want, gpts = _oracle_sum(moved32, rows2, lines, 1.0 / 5.0 / B)
    gR, gt = _transform_grads(rows1, gpts, R)
    for mode in ("batched-slices", "per-pair"):
        v, a, b = results[mode]
        assert abs(v - want) <= REL_TOL * want, mode
        assert _rel(a, gR) <= 2e-5 and _rel(b, gt) <= 2e-5, mode
    assert results["batched-slices"][0] == results["per-pair"][0]             # same per-pair kernels, same bits
    # the fused hook transforms with its own kernel (different rounding of the moved points): compare on ITS points
    mv = rrl.rigid_apply(torch.from_numpy(R).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(rows1).cuda()).reshape(B, -1, 9).cpu().numpy()
    want_h, gpts_h = _oracle_sum(mv, rows2, lines, 1.0 / 5.0 / B)
    gR_h, gt_h = _transform_grads(rows1, gpts_h, R)
    v, a, b = results["hook"]
    assert abs(v - want_h) <= REL_TOL * want_h and _rel(a, gR_h) <= REL_TOL and _rel(b, gt_h) <= REL_TOL
    assert abs(v - want) <= 2e-3 * want and np.abs(moved - mv.reshape(B, -1, 3)).max() < 1e-5


def test_slice_batching_falls_back_when_the_pattern_breaks(rrl):
    """slices visited out of order, a base modified in place between two calls, B = 1 tensors that are no slices: every
    call still returns the B = 1 result of ITS arguments"""
    B, nf, nl = 4, 300, 900
    rows1, rows2, lines, _, _ = _batch(B, nf, nl, 720)
    M = rrl.loss
    f1 = torch.from_numpy(rows1).cuda().reshape(B, -1, 9); f2 = torch.from_numpy(rows2).cuda().reshape(B, -1, 9)
    ln = torch.from_numpy(lines).cuda()
    want = [co.loss(rows1[j].reshape(-1, 9), rows2[j].reshape(-1, 9), lines[j]).loss for j in range(B)]
    call = lambda a, b, c: M.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, a, b, c, "cuda").item()
    for j in (2, 0, 3, 1):                                                       # out of order
        assert abs(call(f1[j:j + 1], f2[j:j + 1], ln[j:j + 1]) - want[j]) <= REL_TOL * want[j]
    assert abs(call(f1[0:1], f2[0:1], ln[0:1]) - want[0]) <= REL_TOL * want[0]   # starts a batch ...
    f1[1] = f1[0]                                                                # ... whose base then changes in place
    w = co.loss(rows1[0].reshape(-1, 9), rows2[1].reshape(-1, 9), lines[1]).loss
    assert abs(call(f1[1:2], f2[1:2], ln[1:2]) - w) <= REL_TOL * w
    single = f1[2].clone()[None]                                                 # not a view of a batch
    assert abs(call(single, f2[2:3].clone(), ln[2:3].clone()) - want[2]) <= REL_TOL * want[2]
    M.STRICT_EMPTY_RETURN = True                                                 # opt-in: the reference's tuple of Nones
    try:
        g = golden("synth_empty")
        res = M.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, torch.from_numpy(g["tri1"]).cuda()[None],
                                                                   torch.from_numpy(g["tri2"]).cuda()[None],
                                                                   torch.from_numpy(g["lines"]).cuda()[None], "cuda")
        assert res == (None, None, None)
    finally:
        M.STRICT_EMPTY_RETURN = False


def test_rpm_and_fmr_hooks(rrl):
    """Train_RPM.py:218-258 (two predicted transforms, / num_iter, discount 0.5^(n-i-1)) and fmr/model.py:285-313 (last 3
    iterates of g_series, / 5.0, discount 0.5^(maxiter-i-1), / batch_size) as the batched helpers, against the same
    arithmetic on per-pair oracle losses"""
    B, nf, nl = 3, 350, 1200
    rows1, rows2, lines, R, t = _batch(B, nf, nl, 740)
    s1 = torch.from_numpy(rows1).cuda(); s2 = torch.from_numpy(rows2).cuda(); ln = torch.from_numpy(lines).cuda()
    rng = np.random.default_rng(9)
    trs = []
    for it in range(2):
        Ri = np.stack([synth.random_rotation(rng, 2.0 + it) for _ in range(B)]).astype(np.float32)
        ti = rng.uniform(-0.02, 0.02, size=(B, 3)).astype(np.float32)
        trs.append(torch.from_numpy(np.concatenate([Ri, ti[:, :, None]], 2)).cuda().requires_grad_(True))
    out = rrl.hooks.rpm_loss(trs, s1, s2, ln)
    out.backward()
    want = 0.0
    for i, g in enumerate(trs):
        mv = rrl.rigid_apply(g.detach()[:, :3, :3], g.detach()[:, :3, 3], s1).reshape(B, -1, 9).cpu().numpy()
        tot, gpts = _oracle_sum(mv, rows2, lines, 1.0 / 2 * 0.5 ** (2 - i - 1))
        want += tot
        gR, gt = _transform_grads(rows1, gpts, None)
        assert _rel(g.grad[:, :3, :3].cpu().numpy(), gR) <= REL_TOL and _rel(g.grad[:, :3, 3].cpu().numpy(), gt) <= REL_TOL
    assert out.shape == (1,) and abs(out.item() - want) <= REL_TOL * want
    # FMR: g_series of maxiter + 1 iterates built from twists through Exp; gradient reaches the twists
    maxiter = 4
    x = torch.from_numpy(rng.normal(scale=0.02, size=(maxiter + 1, B, 6)).astype(np.float32)).cuda().requires_grad_(True)
    gs = rrl.hooks.Exp(x)
    assert gs.shape == (maxiter + 1, B, 4, 4)
    out = rrl.hooks.fmr_loss(gs, s1, s2, ln, maxiter=maxiter)
    out.backward()
    want = 0.0
    gx_want = np.zeros((maxiter + 1, B, 6))
    for i in range(maxiter - 3, maxiter):
        g = gs.detach()[i]
        mv = rrl.rigid_apply(g[:, :3, :3], g[:, :3, 3], s1).reshape(B, -1, 9).cpu().numpy()
        tot, gpts = _oracle_sum(mv, rows2, lines, 1.0 / 5.0 / B * 0.5 ** (maxiter - i - 1))
        want += tot
        gR, gt = _transform_grads(rows1, gpts, None)
        gg = np.zeros((B, 4, 4)); gg[:, :3, :3] = gR; gg[:, :3, 3] = gt
        gx_want[i] = ao.expmap_backward(x.detach()[i].cpu().numpy(), gg)
    assert abs(out.item() - want) <= REL_TOL * want
    assert not x.grad[maxiter].any() and not x.grad[0].any()                     # only the last 3 iterates before maxiter
    assert _rel(x.grad.cpu().numpy(), gx_want) <= 2e-5
    # configs[3] as SURVEY 8(d) states it: per-pair losses with the gradient to the twist (B, 6)
    tw = torch.from_numpy(rng.normal(scale=0.02, size=(B, 6)).astype(np.float32)).cuda().requires_grad_(True)
    per = rrl.hooks.fmr_twist_loss(tw, s1, s2, ln)
    per.sum().backward()
    assert per.shape == (B,) and tw.grad.shape == (B, 6) and bool(tw.grad.abs().sum() > 0)


@pytest.mark.parametrize("B,nf,nl", [(1, 1024, 4000), (3, 700, 1500), (1, 20000, 1200)])
def test_twist_loss_backward_in_pose_space(rrl, B, nf, nl):
    """rrl_b200.twist_loss (exp + transform + loss forward, pose-space contraction backward: no dense point gradient) against
    the two-op chain se3_apply -> intersected_line_loss and against the oracle's point gradient pushed through exp3"""
    pairs = [synth.make_pair(780 + i, nf, nl) for i in range(B)]
    raw = torch.from_numpy(np.stack([p["tri1"] for p in pairs])).cuda()
    t2 = torch.from_numpy(np.stack([p["tri2"] for p in pairs])).cuda()
    ln = torch.from_numpy(np.stack([p["lines"] for p in pairs])).cuda()
    rng = np.random.default_rng(B)
    tw0 = torch.from_numpy(rng.normal(scale=0.02, size=(B, 6)).astype(np.float32)).cuda()
    up = torch.arange(1, B + 1, device="cuda", dtype=torch.float32)
    a = tw0.clone().requires_grad_(True)
    la, info, moved = rrl.twist_loss(a, raw, t2, ln, return_info=True)
    (la * up).sum().backward()
    b = tw0.clone().requires_grad_(True)
    tri1 = rrl.se3_apply(b, raw.reshape(B, -1, 3)).reshape(B, -1, 9)
    lb = rrl.intersected_line_loss(tri1, t2, ln)
    (lb * up).sum().backward()
    assert torch.equal(la, lb) and torch.equal(moved, tri1.detach())
    assert _rel(a.grad.cpu().numpy(), b.grad.cpu().numpy()) <= 1e-6
    for i, p in enumerate(pairs):
        o = co.loss(moved[i].cpu().numpy(), p["tri2"], p["lines"])
        gt = co.se3_backward(tw0[i].cpu().numpy(), p["tri1"].reshape(-1, 3), (i + 1) * o.grad1.reshape(-1, 3))
        assert abs(la[i].item() - o.loss) <= REL_TOL * o.loss and _rel(a.grad[i].cpu().numpy(), gt) <= REL_TOL
        assert float(info.median[i]) == o.median
    # a session keeps the clouds' order from call to call; results do not change
    sess = rrl.LossSession()
    for _ in range(2):
        c = tw0.clone().requires_grad_(True)
        lc = rrl.twist_loss(c, raw, t2, ln, session=sess)
        (lc * up).sum().backward()
        assert torch.equal(lc, la) and _rel(c.grad.cpu().numpy(), a.grad.cpu().numpy()) <= 1e-6


def test_expmap_against_the_reference(rrl):
    """fmr/se_math/se3.py Exp forward and ExpMap.backward (tests/golden/expmap.npz, minted from the reference)"""
    g = golden("expmap")
    x = torch.from_numpy(g["twist"]).cuda().requires_grad_(True)
    out = rrl.hooks.Exp(x)
    out.backward(torch.from_numpy(g["grad_g"]).cuda())
    assert np.abs(out.detach().cpu().numpy() - g["ref_g"]).max() <= 2e-6
    assert _rel(x.grad.cpu().numpy(), g["ref_grad_twist"]) <= REL_TOL
    assert _rel(x.grad.cpu().numpy(), ao.expmap_backward(g["twist"], g["grad_g"])) <= REL_TOL
    lead = rrl.hooks.Exp(x.detach().reshape(4, 6, 6))
    assert lead.shape == (4, 6, 4, 4)


def test_chamfer_is_differentiable_like_the_reference(rrl):
    """loss.py:236-252 with autograd (tests/golden/chamfer_grad.npz)"""
    g = golden("chamfer_grad")
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    y = torch.from_numpy(g["y"]).cuda().requires_grad_(True)
    out = rrl.loss.chamfer_dist(x, y)
    (out * float(g["upstream"])).backward()
    assert abs(out.item() - float(g["ref"])) <= REL_TOL * float(g["ref"])
    assert _rel(x.grad.cpu().numpy(), g["ref_grad_x"]) <= REL_TOL and _rel(y.grad.cpu().numpy(), g["ref_grad_y"]) <= REL_TOL
    # only one side needs a gradient (the hooks: predicted source against a fixed target)
    x2 = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    rrl.chamfer(x2, y.detach()).backward()
    assert _rel(x2.grad.cpu().numpy() * float(g["upstream"]), g["ref_grad_x"]) <= REL_TOL
    assert not rrl.chamfer(x.detach(), y.detach()).requires_grad


def test_tensors_on_a_non_current_device_or_mixed_devices(rrl):
    """the wrappers make the tensors' device current for the native call; mixing devices raises"""
    p = synth.make_pair(760, 200, 500)
    t1 = torch.from_numpy(p["tri1"]).cuda()[None]; t2 = torch.from_numpy(p["tri2"]).cuda()[None]; ln = torch.from_numpy(p["lines"]).cuda()[None]
    if torch.cuda.device_count() > 1:
        want = rrl.intersected_line_loss(t1, t2, ln).item()
        a, b, c = (x.to("cuda:1") for x in (t1, t2, ln))
        assert torch.cuda.current_device() == 0
        assert rrl.intersected_line_loss(a, b, c).item() == want
        with pytest.raises(ValueError):
            rrl.intersected_line_loss(a, t2, ln)
    with pytest.raises(rrl.NativeError):
        rrl.intersected_line_loss(t1.cpu(), t2, ln)
