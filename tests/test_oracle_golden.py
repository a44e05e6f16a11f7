"""The two oracles (plain C and eager-torch restatements) against vectors minted from the unmodified
reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden, hits_to_lists
from oracle import c_oracle as co
from oracle import torch_port as tp

LOSS_CASES = ["demo_step0", "synth_sphere", "synth_ragged", "synth_rpm_like", "synth_window", "synth_short_dirs"]
REL_TOL = 1e-5          # north-star tolerance for loss and gradients (relative, Frobenius for tensors)


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) /
                 max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))


@pytest.mark.parametrize("name", LOSS_CASES)
def test_c_oracle_matches_reference(name):
    g = golden(name)
    k = g["krange"]
    r = co.loss(g["tri1"], g["tri2"], g["lines"], int(k[0]), int(k[1]), int(k[2]), int(k[3]))
    # index-exact hit sets; the oracle's IEEE sqrt may differ from torch.sqrt only inside the 1-ulp band
    assert r.nan == 0
    if r.band == 0:
        assert np.array_equal(r.counts1, g["ref_counts1"])
        assert np.array_equal(r.counts2, g["ref_counts2"])
        for counts, nz, hits in ((g["ref_counts1"], g["ref_hits1"], r.hits1), (g["ref_counts2"], g["ref_hits2"], r.hits2)):
            ref_lists = hits_to_lists(counts, nz)
            for l, lst in enumerate(ref_lists):
                assert list(hits[l][:min(len(lst), co.CAP)]) == lst[:co.CAP]
    else:  # pragma: no cover - not observed on the fixtures; report rather than fail
        mism = int((r.counts1 != g["ref_counts1"]).sum() + (r.counts2 != g["ref_counts2"]).sum())
        assert mism <= r.band
    assert r.status == 0
    assert abs(r.loss - float(g["ref_loss"][0])) <= REL_TOL * abs(float(g["ref_loss"][0]))
    assert _rel(r.grad1, g["ref_grad1"]) <= REL_TOL
    assert _rel(r.grad2, g["ref_grad2"]) <= REL_TOL


@pytest.mark.parametrize("name", LOSS_CASES)
def test_torch_port_matches_reference_bitwise(name):
    g = golden(name)
    k = [int(v) for v in g["krange"]]
    t1 = torch.from_numpy(g["tri1"]).clone().requires_grad_(True)
    t2 = torch.from_numpy(g["tri2"]).clone().requires_grad_(True)
    loss, tr = tp.loss_pair(t1, t2, torch.from_numpy(g["lines"]), k[0], k[1], k[2], k[3], trace=True)
    assert np.array_equal(tr.dense1.counts.numpy(), g["ref_counts1"])
    assert np.array_equal(tr.dense2.counts.numpy(), g["ref_counts2"])
    assert np.array_equal(tr.dense1.hit_tris.numpy(), g["ref_hits1"][:, 1])
    assert np.array_equal(tr.dense2.hit_tris.numpy(), g["ref_hits2"][:, 1])
    if str(torch.__version__) == "2.11.0+cu128":
        assert loss.item() == float(g["ref_loss"][0])        # same ATen ops in the same order
    assert abs(loss.item() - float(g["ref_loss"][0])) <= 1e-6
    loss.backward()
    assert _rel(t1.grad.numpy(), g["ref_grad1"]) <= 1e-6
    assert _rel(t2.grad.numpy(), g["ref_grad2"]) <= 1e-6


def test_empty_selection_is_reported():
    g = golden("synth_empty")
    assert int(g["ref_none"]) == 1
    r = co.loss(g["tri1"], g["tri2"], g["lines"])
    assert r.status == 1 and r.n_combos == 0 and r.loss == 0.0
    assert tp.loss_pair(torch.from_numpy(g["tri1"]), torch.from_numpy(g["tri2"]), torch.from_numpy(g["lines"])) is None


def test_se3_exp_and_backward():
    g = golden("se3")
    for i, tw in enumerate(g["twists"]):
        R, T = co.se3_exp(tw)
        assert np.allclose(R, g["ref_R"][i], rtol=0, atol=2e-6)
        assert np.allclose(T, g["ref_T"][i], rtol=2e-6, atol=2e-6)
        out = co.rigid_apply(R, T, g["pts"])
        assert np.allclose(out, g["ref_out"][i], rtol=1e-5, atol=1e-5)
        gt = co.se3_backward(tw, g["pts"], g["cot"])
        assert _rel(gt, g["ref_twist_grad"][i]) <= 2e-5
        Rt, Tt = tp.se3_exp3(torch.from_numpy(tw))
        assert np.array_equal(Rt[0].numpy(), g["ref_R"][i]) or np.allclose(Rt[0].numpy(), g["ref_R"][i], atol=1e-7)


def test_sampler_from_recorded_uniforms():
    g = golden("sampler")
    n = g["uniforms"].shape[2]
    # torch port: bit-identical to the reference when fed the same draws
    lines, filled = tp.sample_lines(float(g["radius"]), torch.from_numpy(g["center"]), n,
                                    torch.from_numpy(np.stack([g["lo1"], g["hi1"]])),
                                    torch.from_numpy(np.stack([g["lo2"], g["hi2"]])),
                                    uniforms=torch.from_numpy(g["uniforms"]))
    assert filled == int(g["ref_filled"])
    assert np.allclose(lines.numpy(), g["ref_lines"], atol=1e-6)
    # C restatement: libm sin/cos differ from ATen's in the last ulp and the area test is a knife edge
    # (SURVEY 8(a) a7), so the accepted set is compared statistically and geometrically
    cl, cf = co.sample_lines(float(g["radius"]), g["center"], n, g["lo1"], g["hi1"], g["lo2"], g["hi2"], g["uniforms"])
    assert abs(cf - int(g["ref_filled"])) <= 0.15 * int(g["ref_filled"]) + 8
    assert np.all(np.abs(np.linalg.norm(cl[:cf, :3], axis=1) - 1) < 1e-5)
    assert np.all(cl[cf:] == 0)
    # candidate geometry of round 0 agrees with the reference to float tolerance
    import ctypes as C
    cand = np.zeros((n, 6), np.float32)
    u = g["uniforms"][0]
    for i in range(0, n, 97):
        co.lib().rrl_oracle_line_from_uniforms(C.c_float(float(g["radius"])), co._p(co._f32(g["center"])),
                                              C.c_float(float(u[0, i])), C.c_float(float(u[1, i])),
                                              C.c_float(float(u[2, i])), C.c_float(float(u[3, i])), co._p(cand[i]))
        assert np.allclose(cand[i], g["ref_cand0"][i], atol=2e-5)
    # box triangles are exact
    tris = np.zeros(108, np.float32)
    co.lib().rrl_oracle_box_triangles(co._p(co._f32(g["lo1"])), co._p(co._f32(g["hi1"])), co._p(tris))
    assert np.array_equal(tris.reshape(12, 9), g["ref_tris1"])


def test_chamfer():
    g = golden("chamfer")
    assert abs(co.chamfer(g["x"], g["y"]) - float(g["ref_chamfer"])) <= 1e-5 * float(g["ref_chamfer"])
    t = tp.chamfer(torch.from_numpy(g["x"])[None], torch.from_numpy(g["y"])[None]).item()
    assert abs(t - float(g["ref_chamfer"])) <= 1e-6 * float(g["ref_chamfer"])


def test_demo_trajectory_with_port_sampler():
    """Re-runs the first demo steps (test_demo_optimized_Lie_Algebra.py:46-66) with the restated sampler, transform
    and loss on the torch RNG stream of seed 123; only meaningful on the torch build that minted the fixture."""
    g = golden("demo_trajectory")
    if str(g["torch_version"]) != str(torch.__version__):
        pytest.skip("RNG/trig stream differs across torch builds")
    # reproduce the RNG state after Sample_neighs x2 + Reconstruction_point(): re-derive by consuming the same draws
    torch.manual_seed(int(g["seed"]))
    n_src, n_tgt = g["src"].shape[0], g["tgt"].shape[0]
    torch.randint(0, n_src, (1,), dtype=torch.long)            # FPS start index, utils.py:286 (source)
    torch.randint(0, n_tgt, (1,), dtype=torch.long)            # (target)
    V1, V2 = torch.from_numpy(g["src"]), torch.from_numpy(g["tgt"])
    tri2 = torch.from_numpy(g["tri2"])
    tw = torch.from_numpy(g["twist0"]).clone().requires_grad_(True)
    lines, filled = tp.sample_lines(float(g["radius"]), torch.from_numpy(g["center"]), int(g["n_lines"]), V1, V2)
    assert filled == int(g["ref_filled"][0])
    tri1 = tp.rigid_apply(tw, torch.from_numpy(g["tri1_raw"]).reshape(1, -1, 3)).reshape(-1, 9)
    loss = tp.loss_pair(tri1, tri2, lines)
    assert abs(loss.item() - float(g["ref_loss"][0])) <= 1e-6
    loss.backward()
    assert _rel(tw.grad.numpy(), g["ref_twist_grad"][0]) <= 1e-5
    r = co.loss(tri1.detach().numpy(), g["tri2"], lines.numpy())
    assert abs(r.loss - float(g["ref_loss"][0])) <= REL_TOL * float(g["ref_loss"][0])
    gt = co.se3_backward(g["twist0"], g["tri1_raw"].reshape(-1, 3), r.grad1.reshape(-1, 3))
    assert _rel(gt, g["ref_twist_grad"][0]) <= 2e-5


@pytest.mark.parametrize("tag", ["a", "b"])
def test_neigh_oracle_matches_reference_sample_neighs(tag):
    """FPS + 3-NN restatement (oracle/neigh_oracle.py) against the unmodified reference's Sample_neighs and
    utils.farthest_point_sample outputs (oracle/make_golden_neigh.py): index- and bit-exact"""
    from oracle import neigh_oracle as no
    g = golden("sample_neighs")
    pts, ns, ref_idx = g[tag + "_points"], int(g[tag + "_num_sample"]), g[tag + "_ref_fps_idx"]
    assert np.array_equal(no.fps(pts, ns, int(ref_idx[0])), ref_idx)
    assert np.array_equal(no.sample_neighs(pts, ns, 3, int(ref_idx[0])), g[tag + "_ref_neighs"])


def test_expmap_oracle_matches_the_reference():
    """fmr/se_math/se3.py Exp + ExpMap.backward (golden minted by oracle/make_golden_r2.py)"""
    from oracle import aux_oracle as ao
    g = golden("expmap")
    assert np.abs(ao.se3_exp4(g["twist"]) - g["ref_g"]).max() <= 2e-6
    gx = ao.expmap_backward(g["twist"], g["grad_g"])
    assert np.linalg.norm(gx - g["ref_grad_twist"]) <= 1e-5 * np.linalg.norm(g["ref_grad_twist"])


def test_chamfer_gradient_oracle_matches_the_reference():
    from oracle import aux_oracle as ao
    g = golden("chamfer_grad")
    val, gx, gy = ao.chamfer_with_grad(g["x"], g["y"], float(g["upstream"]))
    assert abs(val - float(g["ref"])) <= 1e-5 * float(g["ref"])
    assert np.linalg.norm(gx - g["ref_grad_x"]) <= 1e-5 * np.linalg.norm(g["ref_grad_x"])
    assert np.linalg.norm(gy - g["ref_grad_y"]) <= 1e-5 * np.linalg.norm(g["ref_grad_y"])


def test_se3_log_of_the_shim_matches_the_reference():
    """loss.se3_log (host-side float64, runs once when Reconstruction_point is given an initial (R, T)) against
    LieAlgebra/se3.py:124-134 on exp3 outputs, incl. a near-identity rotation"""
    import torch
    import rrl_b200
    g = golden("se3_log")
    for i in range(g["twist"].shape[0]):
        out = rrl_b200.loss.se3_log(torch.from_numpy(g["R"][i]), torch.from_numpy(g["T"][i])).numpy()
        assert np.abs(out - g["ref_log"][i]).max() <= 2e-5 * max(1.0, np.abs(g["ref_log"][i]).max()), i
        assert np.abs(out - g["twist"][i]).max() <= 5e-5
