"""Dataset file formats of the reference's DL loaders (pre_dataloader.py:95-123): round trips and the centred-pair
bookkeeping of load_pair.  CPU only."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _io():
    # imported by path: the package __init__ pulls in torch + the CUDA library, which this module does not need
    spec = importlib.util.spec_from_file_location("_rrl_io", os.path.join(ROOT, "a-robust-registration-loss_b200", "io.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_round_trips_and_byte_layout(tmp_path):
    io = _io()
    rng = np.random.default_rng(0)
    v = rng.standard_normal((17, 3))
    io.write_obj_vertices(tmp_path / "a.obj", v)
    assert np.allclose(io.read_obj_vertices(tmp_path / "a.obj"), v, rtol=1e-8)
    tri = rng.standard_normal((12, 3)).astype(np.float32)
    io.write_neigh_bin(tmp_path / "n.bin", tri)
    assert os.path.getsize(tmp_path / "n.bin") == 12 * 3 * 4                      # raw float32, no header
    assert np.array_equal(io.read_neigh_bin(tmp_path / "n.bin"), tri)
    assert np.array_equal(np.fromfile(tmp_path / "n.bin", np.float32), tri.reshape(-1))
    rt = rng.standard_normal((3, 4))
    io.write_transform_bin(tmp_path / "t.bin", rt)
    assert os.path.getsize(tmp_path / "t.bin") == 12 * 8                          # raw float64
    assert np.array_equal(io.read_transform_bin(tmp_path / "t.bin"), rt)


def test_load_pair_recentres_the_ground_truth(tmp_path):
    """src @ R + T must map the centred source onto the centred target when the files hold tar = R_gt src + t_gt"""
    io = _io()
    rng = np.random.default_rng(1)
    src = rng.standard_normal((40, 3)) + np.array([3.0, -1.0, 0.5])
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    q *= np.sign(np.linalg.det(q))
    t = np.array([0.3, -0.2, 0.7])
    tar = src @ q.T + t
    p = io.write_pair(str(tmp_path), 5, src, tar, np.repeat(src, 3, 0), np.repeat(tar, 3, 0), np.concatenate([q, t[:, None]], 1))
    assert os.path.basename(p["transform"]) == "5_transform.bin" and os.path.basename(p["src_neigh"]) == "5_src_sample_neigh.bin"
    d = io.load_pair(p["src"], p["tar"])
    assert abs(d["points_src_sample"].mean(0)).max() < 1e-6 and abs(d["points_tar_sample"].mean(0)).max() < 1e-6
    # the reference stores R_inv = R_gt^T ("rotation") and applies clouds as row vectors: src @ rotation + T = tar
    assert np.allclose(d["points_src_sample"] @ d["R_inv"] + d["T"], d["points_tar_sample"], atol=1e-5)
    assert np.allclose(d["R"], d["R_inv"].T) and d["tar_box"].shape == (8, 3)
    assert np.allclose(d["points_based_neighs_src"].reshape(-1, 3, 3)[:, 0], d["points_src_sample"], atol=1e-6)
    dd = io.load_pair(p["src"], p["tar"], dcp=True, fmr=True)
    assert dd["points_src_sample"].shape[0] == 3 and np.allclose(dd["R"], d["R"].T)
