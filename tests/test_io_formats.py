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


def test_load_pair_matches_the_reference_loader(tmp_path):
    """tests/golden/loader.npz = Dataset_2021_8_29.__getitem__ of the unmodified reference (oracle/make_golden_r2.py) on
    one pair, for the plain / DCP / FMR layout switches; load_pair must reproduce every key bit for bit (tar_box: the two
    corners the hooks read, pre_dataloader.py:111 + Train_DCP.py:234-236)"""
    io = _io()
    g = np.load(os.path.join(ROOT, "tests", "golden", "loader.npz"))
    d = str(tmp_path)
    def wobj(path, v):
        with open(path, "w") as f:
            for p in v:
                f.write("v %.17g %.17g %.17g\n" % tuple(p))
    p = io.pair_paths(d, 7)
    wobj(p["src"], g["src"]); wobj(p["tar"], g["tar"])
    io.write_neigh_bin(p["src_neigh"], g["n_src"]); io.write_neigh_bin(p["tar_neigh"], g["n_tar"])
    io.write_transform_bin(p["transform"], g["rt"])
    for tag, kw in (("plain", {}), ("dcp", dict(dcp=True)), ("fmr", dict(fmr=True))):
        out = io.load_pair(p["src"], p["tar"], **kw)
        keys = [k[len("ref_%s_" % tag):] for k in g.files if k.startswith("ref_%s_" % tag)]
        assert sorted(keys) == sorted(out.keys())
        for k in keys:
            ref = g["ref_%s_%s" % (tag, k)]
            assert out[k].shape == ref.shape and out[k].dtype == ref.dtype, (tag, k)
            if k == "tar_box":
                assert np.array_equal(out[k][[0, -1]], ref[[0, -1]])
                assert sorted(map(tuple, out[k])) == sorted(map(tuple, ref))
            else:
                assert np.array_equal(out[k], ref), (tag, k)
    plain = io.load_pair(p["src"], p["tar"])
    # the reference's numpy `.transpose(0, 1)` is the identity: R == R_inv == gt[:3, :3] outside the DCP switch
    assert np.array_equal(plain["R"], plain["R_inv"]) and np.array_equal(plain["R"], g["rt"][:, :3].astype(np.float32))
