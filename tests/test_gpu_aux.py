"""se(3), sampler, chamfer and the reference-compatible module on the GPU."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rrl():
    import rrl_b200
    return rrl_b200


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_se3_against_reference_vectors(rrl):
    g = golden("se3")
    tw = torch.from_numpy(g["twists"]).cuda().requires_grad_(True)
    B = tw.shape[0]
    pts = torch.from_numpy(g["pts"]).cuda()[None].expand(B, -1, -1).contiguous()
    out = rrl.se3_apply(tw, pts)
    (out * torch.from_numpy(g["cot"]).cuda()[None]).sum().backward()
    R, T = rrl.se3_exp(tw)
    assert np.allclose(R.cpu().numpy(), g["ref_R"], atol=2e-6)
    assert np.allclose(T.cpu().numpy(), g["ref_T"], atol=2e-6, rtol=2e-6)
    assert np.allclose(out.detach().cpu().numpy(), g["ref_out"], atol=1e-5, rtol=1e-5)
    for i in range(B):
        assert _rel(tw.grad[i].cpu().numpy(), g["ref_twist_grad"][i]) <= 2e-5
        assert _rel(tw.grad[i].cpu().numpy(), co.se3_backward(g["twists"][i], g["pts"], g["cot"])) <= 1e-5


def test_rigid_apply_matches_torch(rrl):
    gen = torch.Generator().manual_seed(3)
    R = torch.linalg.qr(torch.randn(4, 3, 3, generator=gen))[0].cuda().requires_grad_(True)
    t = torch.randn(4, 3, generator=gen).cuda().requires_grad_(True)
    p = torch.randn(4, 100, 3, generator=gen).cuda().requires_grad_(True)
    cot = torch.randn(4, 100, 3, generator=gen).cuda()
    out = rrl.rigid_apply(R, t, p)
    (out * cot).sum().backward()
    R2, t2, p2 = (x.detach().clone().requires_grad_(True) for x in (R, t, p))
    ref = p2 @ R2.transpose(1, 2) + t2[:, None]
    (ref * cot).sum().backward()
    assert torch.allclose(out, ref, atol=1e-5)
    for a, b in ((R, R2), (t, t2), (p, p2)):
        assert torch.allclose(a.grad, b.grad, atol=1e-4, rtol=1e-4)


def test_sampler_replays_recorded_uniforms(rrl):
    g = golden("sampler")
    n = g["uniforms"].shape[2]
    lo1, hi1, lo2, hi2 = (g[k] for k in ("lo1", "hi1", "lo2", "hi2"))
    v1 = torch.from_numpy(np.stack([lo1, hi1]))[None].cuda()
    v2 = torch.from_numpy(np.stack([lo2, hi2]))[None].cuda()
    lines, filled = rrl.sample_lines(torch.tensor([float(g["radius"])]), torch.from_numpy(g["center"])[None], n, v1, v2,
                                     uniforms=torch.from_numpy(g["uniforms"])[None])
    lines = lines[0].cpu().numpy(); filled = int(filled[0])
    ref_filled = int(g["ref_filled"])
    # the area test is a rounding-noise knife edge: same statistics, not the same coin flips
    assert abs(filled - ref_filled) <= 0.05 * ref_filled + 8
    assert np.all(lines[filled:] == 0)
    assert np.all(np.abs(np.linalg.norm(lines[:filled, :3], axis=1) - 1) < 1e-5)
    # every accepted line is one of the reference's candidates (same draws -> same chords to float tolerance),
    # in the same order, and genuinely crosses both boxes
    cand0 = g["ref_cand0"]
    j = 0
    for row in lines[:min(filled, 50)]:
        while j < n and not np.allclose(cand0[j], row, atol=3e-5):
            j += 1
        assert j < n
    d, x0 = lines[:filled, :3].astype(np.float64), lines[:filled, 3:].astype(np.float64)
    for lo, hi in ((lo1, hi1), (lo2, hi2)):
        with np.errstate(divide="ignore", invalid="ignore"):
            ta, tb = (lo - x0) / d, (hi - x0) / d
        assert np.all(np.minimum(ta, tb).max(1) <= np.maximum(ta, tb).min(1) + 1e-4)
    # the C oracle fed the same draws accepts a statistically identical number
    _, cf = co.sample_lines(float(g["radius"]), g["center"], n, lo1, hi1, lo2, hi2, g["uniforms"])
    assert abs(filled - cf) <= 0.05 * cf + 8


@pytest.mark.parametrize("seed,ext,rscale", [(1, (3.6, 9.6, 5.2), 0.5), (2, (2.0, 2.2, 2.1), 1.0), (3, (9.8, 7.8, 0.45), 0.7)])
def test_sampler_statistics_over_a_million_candidates(rrl, seed, ext, rscale):
    """10 rounds x 100 000 recorded uniforms = 10^6 candidates against the bit-exact torch restatement of the reference
    sampler (oracle/torch_port.py, itself pinned by tests/golden/sampler.npz).  The area test `A + B + C <= S` is a
    rounding coin flip for every true hit, so the two implementations cannot take the same decisions; what
    must hold: the per-candidate ACCEPTANCE RATE agrees within 2 % (binomial noise at these counts is ~0.3 %), every row the
    GPU fills is one of the candidates, the rows come in candidate order (checked over ALL rows), and nothing is accepted
    that geometrically misses a box."""
    from oracle import torch_port as tp
    rng = np.random.default_rng(seed)
    n, rounds = 100000, 10
    v1 = (rng.uniform(-0.5, 0.5, (400, 3)) * np.asarray(ext)).astype(np.float32)
    v2 = ((rng.uniform(-0.5, 0.5, (400, 3)) * np.asarray(ext)) @ synth_rot(rng).T + rng.uniform(-0.3, 0.3, 3)).astype(np.float32)
    radius = float(rscale * np.linalg.norm(v2.max(0) - v2.min(0)))
    center = v2.mean(0).astype(np.float32)
    U = torch.rand(rounds, 4, n, generator=torch.Generator().manual_seed(seed))
    # reference acceptance per candidate, all 10 rounds evaluated (no early stop: the RATE is what is compared)
    t1, t2 = tp.box_triangles(torch.from_numpy(v1)), tp.box_triangles(torch.from_numpy(v2))
    ref_acc, cands = 0, []
    for rd in range(rounds):
        cand = tp.lines_from_uniforms(radius, torch.from_numpy(center), *[U[rd, q] for q in range(4)])
        ref_acc += int(((tp.triangle_hits(t1, cand) * tp.triangle_hits(t2, cand)) > 0).sum())
        cands.append(cand.numpy())
    # GPU: N = all candidates, so nothing is cut off and `filled` counts every accepted candidate of every round
    big_n = rounds * n
    Ubig = torch.zeros(1, rounds, 4, big_n)
    Ubig[0, :, :, :n] = U                                        # candidates beyond n: u = 0 -> a degenerate chord (q1 == q2)
    lines, filled = rrl.sample_lines(torch.tensor([radius]), torch.from_numpy(center)[None], big_n, torch.from_numpy(v1)[None].cuda(),
                                     torch.from_numpy(v2)[None].cuda(), uniforms=Ubig.cuda(), rounds=rounds)
    lines = lines[0].cpu().numpy(); filled = int(filled[0])
    assert ref_acc > 20000
    assert abs(filled - ref_acc) <= 0.02 * ref_acc, (filled, ref_acc)
    assert np.all(lines[filled:] == 0)
    # full-order check: walk the candidates of all rounds once; every filled row must be matched, in order
    from scipy.spatial import cKDTree
    flat = np.concatenate(cands).astype(np.float64)               # (rounds * n, 6) in (round, index) order
    dist, idx = cKDTree(flat).query(lines[:filled].astype(np.float64), distance_upper_bound=1e-4)
    assert np.all(np.isfinite(dist))                              # every filled row IS one of the candidates (same draws)
    assert np.all(np.diff(idx) > 0)                               # strictly increasing candidate indices = the reference's order
    d, x0 = lines[:filled, :3].astype(np.float64), lines[:filled, 3:].astype(np.float64)
    for v in (v1, v2):
        with np.errstate(divide="ignore", invalid="ignore"):
            ta, tb = (v.min(0) - x0) / d, (v.max(0) - x0) / d
        assert np.all(np.minimum(ta, tb).max(1) <= np.maximum(ta, tb).min(1) + 1e-4)


def synth_rot(rng):
    from oracle import synth
    return synth.random_rotation(rng, 30.0)


def test_sampler_philox_is_reproducible(rrl):
    g = golden("demo_trajectory")
    v1 = torch.from_numpy(g["src"])[None].cuda(); v2 = torch.from_numpy(g["tgt"])[None].cuda()
    r = torch.tensor([float(g["radius"])]); c = torch.from_numpy(g["center"])[None]
    a, fa = rrl.sample_lines(r, c, 20000, v1, v2, seed=7, offset=3)
    b, fb = rrl.sample_lines(r, c, 20000, v1, v2, seed=7, offset=3)
    d, fd = rrl.sample_lines(r, c, 20000, v1, v2, seed=7, offset=4)
    assert torch.equal(a, b) and int(fa) == int(fb)
    assert not torch.equal(a, d)
    # acceptance statistics of the reference on this pair: 12967..13767 of 20000 rows filled (golden trajectory)
    assert 0.85 * g["ref_filled"].min() <= int(fa) <= 1.15 * g["ref_filled"].max()
    # batched call == per-pair calls (counter-based stream is independent of launch geometry)
    v1b = torch.cat([v1, v1 * 0.9]); v2b = torch.cat([v2, v2 * 1.1])
    ab, fab = rrl.sample_lines(torch.cat([r, r]), torch.cat([c, c]), 5000, v1b, v2b, seed=11)
    a0, _ = rrl.sample_lines(r, c, 5000, v1b[:1], v2b[:1], seed=11)
    assert torch.equal(ab[0], a0[0])


def test_chamfer(rrl):
    g = golden("chamfer")
    x = torch.from_numpy(g["x"])[None].cuda(); y = torch.from_numpy(g["y"])[None].cuda()
    v = rrl.chamfer(x, y).item()
    assert abs(v - float(g["ref_chamfer"])) <= 1e-5 * float(g["ref_chamfer"])
    xb = torch.cat([x, x * 1.5]); yb = torch.cat([y, y * 0.5])
    vb = rrl.chamfer(xb, yb).item()
    want = (co.chamfer(g["x"], g["y"]) + co.chamfer(g["x"] * 1.5, g["y"] * 0.5)) / 2
    assert abs(vb - want) <= 1e-5 * want


def test_reference_named_module_runs_the_demo_step(rrl):
    """test_demo_optimized_Lie_Algebra.py:41-66 with the drop-in names, step 0 of the golden trajectory (lines
    supplied from the reference sampler run are not stored for 20000 lines, so the loss is checked on the stored
    4000-line subset and the twist gradient against the oracle)."""
    g0 = golden("demo_trajectory"); g = golden("demo_step0")
    M = rrl.loss
    rec = M.Reconstruction_point().cuda()
    assert list(rec.state_dict().keys()) == ["parameters_"] and rec.parameters_.shape == (6,)
    rec.parameters_.data = torch.from_numpy(g0["twist0"]).cuda()
    src = torch.from_numpy(g0["src"]).cuda()
    tri_raw = torch.from_numpy(g0["tri1_raw"]).cuda().reshape(1, -1, 3)
    moved, tri1 = rec(src, tri_raw)
    assert moved.shape == src.shape and tri1.shape == (g0["tri1_raw"].shape[0], 9)
    assert np.allclose(tri1.detach().cpu().numpy(), g["tri1"], atol=2e-6)
    loss = M.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, tri1.reshape(1, -1, 9),
                                                                torch.from_numpy(g["tri2"]).cuda().reshape(1, -1, 9),
                                                                torch.from_numpy(g["lines"]).cuda().reshape(1, -1, 6), "cuda")
    assert loss.shape == (1,)
    # the fused transform rounds differently from MKL's sgemm: compare with the oracle on OUR points
    orc = co.loss(tri1.detach().cpu().numpy(), g["tri2"], g["lines"])
    assert abs(loss.item() - orc.loss) <= 1e-5 * orc.loss
    assert abs(loss.item() - float(g["ref_loss"][0])) <= 2e-3 * float(g["ref_loss"][0])
    loss.backward()
    both = np.concatenate([g0["src"], g0["tri1_raw"].reshape(-1, 3)])
    gp = np.concatenate([np.zeros_like(g0["src"]), orc.grad1.reshape(-1, 3)])
    assert _rel(rec.parameters_.grad.cpu().numpy(), co.se3_backward(g0["twist0"], both, gp)) <= 1e-5
    lines = M.Random_uniform_distribution_lines_batch_efficient_resample(
        torch.tensor([[float(g0["radius"])]]).cuda(), torch.from_numpy(g0["center"]).reshape(1, -1).cuda(), 20000,
        moved.detach().view(1, -1, 3), torch.from_numpy(g0["tgt"]).cuda().view(1, -1, 3), "cuda")
    assert lines.shape == (1, 20000, 6)
    cd = M.chamfer_dist(moved.detach().reshape(1, -1, 3), torch.from_numpy(g0["tgt"]).cuda().reshape(1, -1, 3))
    assert abs(cd.item() - float(g0["ref_chamfer"][0])) <= 1e-4 * float(g0["ref_chamfer"][0])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sample_neighs_against_reference_vectors(rrl, tag):
    """Sample_neighs through the CUDA FPS + kNN kernels, bit-exact against the unmodified reference's output; the
    start index comes from torch's CPU generator exactly like the reference's torch.randint (utils.py:288)"""
    g = golden("sample_neighs")
    pts, ns = g[tag + "_points"], int(g[tag + "_num_sample"])
    torch.manual_seed(int(g[tag + "_seed"]))
    out = rrl.loss.Sample_neighs(pts, num_sample=ns, num_neigh=3)
    assert out.dtype == pts.dtype and np.array_equal(out, g[tag + "_ref_neighs"])
    idx = rrl.prep.farthest_point_sample(torch.from_numpy(pts).cuda(), ns, start=int(g[tag + "_ref_fps_idx"][0]))
    assert np.array_equal(idx.cpu().numpy(), g[tag + "_ref_fps_idx"])


@pytest.mark.parametrize("n,npoint,dtype", [(5, 5, np.float32), (1000, 37, np.float64), (70001, 300, np.float32),
                                            (300000, 64, np.float32)])
def test_fps_and_knn_against_oracle(rrl, n, npoint, dtype):
    """single- and multi-block grids (the cooperative kernel adds a block per 4096 points), float64 input, npoint = N"""
    from oracle import neigh_oracle as no
    rng = np.random.default_rng(n)
    pts = (rng.standard_normal((n, 3)) * np.array([2.0, 1.0, 0.5])).astype(dtype)
    pts[n // 2] = pts[0]                                          # a duplicated point: equal distances -> first index wins
    start = int(rng.integers(0, n))
    want = no.fps(pts, npoint, start)
    p = torch.from_numpy(pts).cuda()
    got = rrl.prep.farthest_point_sample(p, npoint, start=start)
    assert np.array_equal(got.cpu().numpy(), want)
    k = min(3, n)
    q = got[: min(npoint, 50)]
    nn = rrl.prep.knn(p, q, k).cpu().numpy()
    assert np.array_equal(nn, no.knn(pts, pts[q.cpu().numpy()], k))


def test_fps_rejects_bad_arguments(rrl):
    p = torch.zeros(10, 3, device="cuda")
    with pytest.raises(ValueError):
        rrl.prep.farthest_point_sample(p, 11)
    with pytest.raises(rrl.NativeError):
        rrl.prep.farthest_point_sample(p.cpu(), 3)
    L = rrl._native.lib()
    assert L.rrl_fps(p.data_ptr(), 0, 10, 3, 10, p.data_ptr(), p.data_ptr(), 1 << 20, None) == -1     # start out of range
    assert L.rrl_knn(p.data_ptr(), 0, 10, p.data_ptr(), 2, 9, p.data_ptr(), None) == -1               # k > 8


def test_demo_flow_with_the_reference_names(rrl):
    """examples/demo_lie_algebra.py = test_demo_optimized_Lie_Algebra.py with `loss` shadowed by the B200 module:
    Sample_neighs -> Reconstruction_point -> per epoch sampler + loss + Adam + chamfer.  A 25-degree misalignment of a
    synthetic ellipsoid must be recovered."""
    import argparse
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_demo", os.path.join(root, "examples", "demo_lie_algebra.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    torch.manual_seed(5); np.random.seed(5); rrl.loss.manual_seed(5)
    args = argparse.Namespace(synthetic=1500, seed=5, angle=25.0, data_path="", label1="0")
    data = demo.load_case(args, "cuda")
    assert data["vertics1_faces_tensor"].shape == (1, 3 * 1500, 3)
    model, hist = demo.test_one_case(data, n_epoch=80, n_sample_line=20000, device="cuda", log=None)     # the demo's own line count
    assert len(hist) >= 70
    # Adam at the demo's learning rate takes normalised steps of ~0.6 degrees per epoch even at the optimum, every epoch draws
    # fresh lines, and the scatter of the point gradient uses float atomics, so repeated runs follow different trajectories and
    # single late epochs wander.  Spread over 8 repeated runs of exactly this case on a B200 (tools/demo_flaky.py 8 80 20000 25,
    # profiles/r02_demo_spread.log): rotation error of the final transform 0.2 .. 3.4 degrees (start: 25), median Chamfer of the
    # last 15 epochs 0.002 .. 0.25 of the start, best epoch below 0.01 of the start.  (At 12 degrees and 8000 lines -- round 1's
    # case -- the same wander is as large as the initial misalignment: final errors up to 4.9 degrees, late medians up to 1.4x the
    # start, which is why that version could only assert "not worse than the start".)
    cf = [h[0] for h in hist]
    first = np.mean(cf[:3])
    assert min(cf) < 0.02 * first, (first, min(cf))
    assert np.median(cf[-15:]) < 0.5 * first, (first, np.median(cf[-15:]))
    Rn = model.Transform()[0][0].cpu().numpy().astype(np.float64)
    ang = np.deg2rad(25.0)
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    # row-vector convention (loss.py:460-461): p' = p @ R and the target is base @ Rz.T, so R must approach Rz.T
    rot_err = np.rad2deg(np.arccos(np.clip((np.trace(Rn @ Rz) - 1) / 2, -1, 1)))
    assert rot_err < 7.0, rot_err
    R, T = model.Transform()
    assert R.shape == (1, 3, 3) and T.shape == (1, 3) and "parameters_" in model.state_dict()
