"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which never travels to the GPU box):

    python -m oracle.make_golden

Every array under the key prefix `ref_` is an output of the reference's own functions
(/root/reference/code/loss.py, LieAlgebra/se3.py) on torch CPU; the other arrays are the
inputs they were given.  The reference has no golden vectors of its own, so these
files are what pins the two oracles (oracle/torch_port.py, oracle/rrl_oracle.c) and, through
them, the CUDA path.
"""
import os
import sys

import numpy as np
import torch

from . import ref_loader, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SAMPLE = os.path.join(ref_loader.REFERENCE_CODE, "sample_data")


def read_obj_vertices(path):
    return np.array([[float(t) for t in ln.split()[1:4]] for ln in open(path) if ln.startswith("v ")], np.float32)


def ref_labels(L, tri, lines):
    """hit lists straight from the reference's label tensor (loss.py:107-112)."""
    _, _, label = L.cal_intersection_batch2_points_with_line(tri.reshape(1, -1, 9), lines.reshape(1, -1, 6))
    label = label[0]
    counts = label.sum(-1).to(torch.int32).numpy()
    nz = label.nonzero().to(torch.int32).numpy()
    return counts, nz


def ref_loss_case(L, tri1, tri2, lines, rng=(1, 1, 5, 5)):
    t1 = torch.from_numpy(tri1).reshape(1, -1, 9).clone().requires_grad_(True)
    t2 = torch.from_numpy(tri2).reshape(1, -1, 9).clone().requires_grad_(True)
    ln = torch.from_numpy(lines).reshape(1, -1, 6)
    out = L.cal_loss_intersection_batch_whole_median_pts_lines(rng[0], rng[1], rng[2], rng[3], t1, t2, ln, "cpu")
    c1, nz1 = ref_labels(L, t1.detach(), ln)
    c2, nz2 = ref_labels(L, t2.detach(), ln)
    d = dict(tri1=tri1, tri2=tri2, lines=lines, krange=np.array(rng, np.int32),
             ref_counts1=c1, ref_hits1=nz1, ref_counts2=c2, ref_hits2=nz2)
    if isinstance(out, tuple):
        d["ref_none"] = np.int32(1)
        return d
    out.backward()
    d.update(ref_none=np.int32(0), ref_loss=out.detach().numpy().astype(np.float32),
             ref_grad1=t1.grad[0].numpy().copy(), ref_grad2=t2.grad[0].numpy().copy())
    return d


def demo_inputs(L, name="challenge_data", label="0", seed=123):
    """Pre-processing of test_demo_optimized_Lie_Algebra.py:103-134 with a plain OBJ reader."""
    torch.manual_seed(seed)
    np.random.seed(seed)
    v1 = read_obj_vertices(os.path.join(SAMPLE, name, label + "_src_sample.obj"))
    v2 = read_obj_vertices(os.path.join(SAMPLE, name, label + "_tar_sample.obj"))
    n1 = L.Sample_neighs(v1)
    n2 = L.Sample_neighs(v2)
    c1, c2 = v1.mean(0)[None], v2.mean(0)[None]
    v1, v2, n1, n2 = v1 - c1, v2 - c2, n1 - c1, n2 - c2
    return v1.astype(np.float32), v2.astype(np.float32), n1.astype(np.float32), n2.astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    L = ref_loader.load()
    torch.set_num_threads(os.cpu_count() or 1)

    # ---- 1. demo pair (shipped sample data), reference sampler, first Adam steps -------------
    v1, v2, n1, n2 = demo_inputs(L)
    V1, V2 = torch.from_numpy(v1), torch.from_numpy(v2)
    tri1_raw = torch.from_numpy(n1).reshape(1, -1, 3)
    tri2 = torch.from_numpy(n2).reshape(1, -1, 9)
    radius = (V2.max(0)[0] - V2.min(0)[0]).norm(p=2)
    center = V2.mean(0)
    rec = L.Reconstruction_point()
    twist0 = rec.parameters_.detach().numpy().copy()
    opt = torch.optim.Adam(rec.parameters(), lr=1e-2)      # effective lr of the demo
    cur = V1
    traj_loss, traj_filled, traj_twist, traj_grad, traj_cd = [], [], [], [], []
    step0 = None
    n_lines = 20000
    for epoch in range(4):
        lines = L.Random_uniform_distribution_lines_batch_efficient_resample(
            torch.FloatTensor([radius]).reshape(1, 1), center.reshape(1, -1), n_lines, cur.view(1, -1, 3),
            V2.view(1, -1, 3), "cpu").detach()
        cur, tri1 = rec(V1, tri1_raw)
        loss = L.cal_loss_intersection_batch_whole_median_pts_lines(1, 1, 5, 5, tri1.reshape(1, -1, 9), tri2, lines,
                                                                    "cpu")
        opt.zero_grad()
        loss.backward()
        traj_twist.append(rec.parameters_.detach().numpy().copy())
        traj_grad.append(rec.parameters_.grad.numpy().copy())
        traj_loss.append(loss.item())
        traj_filled.append(int((lines[0].abs().sum(-1) > 0).sum()))
        traj_cd.append(L.chamfer_dist(cur.detach().reshape(1, -1, 3), V2.reshape(1, -1, 3)).item())
        if epoch == 0:
            step0 = (tri1.detach().numpy().reshape(-1, 9).copy(), lines[0].numpy().copy())
        opt.step()
        cur = cur.detach()
    print("demo trajectory", traj_loss, traj_filled)
    np.savez_compressed(os.path.join(OUT, "demo_trajectory.npz"), twist0=twist0, ref_twist=np.array(traj_twist),
                        ref_twist_grad=np.array(traj_grad), ref_loss=np.array(traj_loss, np.float32),
                        ref_filled=np.array(traj_filled), ref_chamfer=np.array(traj_cd, np.float32),
                        src=v1, tgt=v2, tri1_raw=n1.reshape(-1, 9), tri2=n2.reshape(-1, 9),
                        radius=np.float32(radius.item()), center=center.numpy(), seed=np.int32(123),
                        n_lines=np.int32(n_lines), torch_version=np.array(torch.__version__))

    # ---- 2. demo step 0 on a 4000-line subset: full internals -----------------------------
    tri1_0, lines0 = step0
    sub = np.concatenate([lines0[:3400], lines0[-600:]])            # keeps 600 all-zero (unfilled) rows
    case = ref_loss_case(L, tri1_0, n2.reshape(-1, 9), sub)
    print("demo_step0", case["ref_loss"], int(case["ref_counts1"].sum()), int(case["ref_counts2"].sum()))
    np.savez_compressed(os.path.join(OUT, "demo_step0.npz"), **case)

    # ---- 3. synthetic cases incl. edge cases -----------------------------------------------
    cases = {
        "synth_sphere": dict(seed=11, nf=300, nl=1500),
        "synth_ragged": dict(seed=12, nf=257, nl=1203, nf2=391, shape="torus", zero_frac=0.1),
        "synth_rpm_like": dict(seed=13, nf=400, nl=1000, shape="box", noise=0.01, outlier_frac=0.1, keep_frac=0.7,
                               radius_scale=1.0),
    }
    for name, kw in cases.items():
        p = synth.make_pair(**kw)
        case = ref_loss_case(L, p["tri1"], p["tri2"], p["lines"])
        print(name, case.get("ref_loss"), int(case["ref_counts1"].sum()))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **case)
    # restricted (k,j) window and a window with no populated combo
    p = synth.make_pair(seed=14, nf=300, nl=1200)
    case = ref_loss_case(L, p["tri1"], p["tri2"], p["lines"], rng=(2, 1, 4, 3))
    print("synth_window", case.get("ref_loss"))
    np.savez_compressed(os.path.join(OUT, "synth_window.npz"), **case)
    far = p["lines"].copy()
    far[:, 3:] += 100.0                                            # lines far away from both clouds: no hits at all
    case = ref_loss_case(L, p["tri1"], p["tri2"], far[:64])
    assert case["ref_none"] == 1
    np.savez_compressed(os.path.join(OUT, "synth_empty.npz"), **case)

    # ---- 4. se(3) exponential + transform (Reconstruction_point) ----------------------------
    g = torch.Generator().manual_seed(5)
    twists = torch.cat([torch.randn(6, 6, generator=g) * 1e-3,          # Taylor branch (|w| < 0.01)
                        torch.randn(6, 6, generator=g) * 0.3,
                        torch.randn(4, 6, generator=g) * 1.5], 0)
    twists[0, :3] = 0.0                                                  # theta == 0 exactly
    pts = torch.randn(64, 3, generator=g) * 3
    cot = torch.randn(64, 3, generator=g)
    Rs, Ts, outs, grads = [], [], [], []
    for tw in twists:
        rec = L.Reconstruction_point()
        rec.parameters_.data = tw.clone()
        R, T = rec.Transform()
        o, _ = rec(pts, pts.reshape(1, -1, 3)[:, :63])
        (o * cot).sum().backward()
        Rs.append(R.detach().numpy()[0])
        Ts.append(T.detach().numpy()[0])
        outs.append(o.detach().numpy())
        grads.append(rec.parameters_.grad.numpy().copy())
    np.savez_compressed(os.path.join(OUT, "se3.npz"), twists=twists.numpy(), pts=pts.numpy(), cot=cot.numpy(),
                        ref_R=np.array(Rs), ref_T=np.array(Ts), ref_out=np.array(outs), ref_twist_grad=np.array(grads))

    # ---- 5. sampler from supplied uniforms ------------------------------------------------
    n = 2048
    g = torch.Generator().manual_seed(9)
    uni = torch.rand(10, 4, n, generator=g)
    it = iter(uni.reshape(40, n))
    orig_rand = torch.rand
    torch.rand = lambda *a, **k: next(it).reshape(1, n).clone()          # feed the recorded draws
    try:
        lines = L.Random_uniform_distribution_lines_batch_efficient_resample(
            torch.FloatTensor([radius]).reshape(1, 1), center.reshape(1, -1), n, V1.view(1, -1, 3), V2.view(1, -1, 3),
            "cpu")
        it = iter(uni.reshape(40, n))
        cand0 = L.Random_uniform_distribution_lines_batch_efficient(torch.FloatTensor([radius]).reshape(1, 1),
                                                                    center.reshape(1, -1), n, "cpu")
    finally:
        torch.rand = orig_rand
    fv1 = L.generate_mesh_by_bbox(L.generate_bbox(V1.view(1, -1, 3)))
    fv2 = L.generate_mesh_by_bbox(L.generate_bbox(V2.view(1, -1, 3)))
    lab1 = L.cal_intersection_batch2_rand_lines(fv1, cand0)[0].numpy()
    lab2 = L.cal_intersection_batch2_rand_lines(fv2, cand0)[0].numpy()
    filled = int((lines[0].abs().sum(-1) > 0).sum())
    print("sampler filled", filled, "round-0 accepted", int((lab1 * lab2 > 0).sum()))
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), uniforms=uni.numpy(), radius=np.float32(radius.item()),
                        center=center.numpy(), lo1=V1.min(0)[0].numpy(), hi1=V1.max(0)[0].numpy(),
                        lo2=V2.min(0)[0].numpy(), hi2=V2.max(0)[0].numpy(), ref_lines=lines[0].numpy(),
                        ref_filled=np.int32(filled), ref_cand0=cand0[0].numpy(), ref_hits1_round0=lab1,
                        ref_hits2_round0=lab2, ref_tris1=fv1[0].numpy(), ref_tris2=fv2[0].numpy())

    # ---- 6. chamfer ------------------------------------------------------------------------
    cd = L.chamfer_dist(V1.reshape(1, -1, 3), V2.reshape(1, -1, 3)).item()
    np.savez_compressed(os.path.join(OUT, "chamfer.npz"), x=v1, y=v2, ref_chamfer=np.float32(cd))
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))}
    print(sizes, sum(sizes.values()))


if __name__ == "__main__":
    sys.exit(main())
