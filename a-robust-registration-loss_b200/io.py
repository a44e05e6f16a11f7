"""On-disk formats of the reference's deep-learning datasets (SURVEY 8(f) row 4;
code/exps_deep_learning/pre_dataloader.py:80-166), so that the op can be fed the same files:

    <i>_src_sample.obj / <i>_tar_sample.obj   vertex lists ("v x y z" lines; the reference reads them with igl)
    <i>_*_sample_neigh.bin                    float32, (3 * nf, 3): the Sample_neighs triplets of that cloud
    <i>_transform.bin                         float64, (3, 4): ground-truth [R | t]

Host-side plumbing only (numpy); nothing here is on the hot path.  `load_pair` restates Dataset_2021_8_29.__getitem__
(centring, the R / T bookkeeping, the DCP / FMR layout switches) and is pinned by tests/golden/loader.npz, minted from the
unmodified reference class (by the test infrastructure, make_golden_r2.py; `igl` stubbed there, so `tar_box` is pinned through its first and
last corner only -- the two the hooks read).
"""
import os

import numpy as np


def read_obj_vertices(path: str) -> np.ndarray:
    """(V, 3) float64, like igl.read_triangle_mesh(path)[0]"""
    rows = [ln.split()[1:4] for ln in open(path) if ln.startswith("v ")]
    return np.array(rows, dtype=np.float64).reshape(-1, 3)


def write_obj_vertices(path: str, v: np.ndarray) -> None:
    with open(path, "w") as f:
        for x, y, z in np.asarray(v, np.float64).reshape(-1, 3):
            f.write("v %.9g %.9g %.9g\n" % (x, y, z))


def read_neigh_bin(path: str) -> np.ndarray:
    """(3 * nf, 3) float32 -- pre_dataloader.py:101-106 (the loader widens it to float64 afterwards)"""
    return np.fromfile(path, np.float32).reshape(-1, 3)


def write_neigh_bin(path: str, triplets: np.ndarray) -> None:
    np.ascontiguousarray(triplets, np.float32).reshape(-1, 3).tofile(path)


def read_transform_bin(path: str) -> np.ndarray:
    """(3, 4) float64 -- pre_dataloader.py:119-121"""
    return np.fromfile(path, np.float64).reshape(3, 4)


def write_transform_bin(path: str, rt: np.ndarray) -> None:
    np.ascontiguousarray(rt, np.float64).reshape(3, 4).tofile(path)


def pair_paths(directory: str, index) -> dict:
    """file names of pair `index` under the reference's naming scheme (pre_dataloader.py:95-118)"""
    tar = os.path.join(directory, "%s_tar_sample.obj" % index)
    src = os.path.join(directory, "%s_src_sample.obj" % index)
    return {"src": src, "tar": tar, "src_neigh": src.replace(".obj", "_neigh.bin", 1), "tar_neigh": tar.replace(".obj", "_neigh.bin", 1),
            "transform": tar.replace("tar_sample", "transform", 1).replace(".obj", ".bin", 1)}


def write_pair(directory: str, index, src: np.ndarray, tar: np.ndarray, src_neigh: np.ndarray, tar_neigh: np.ndarray,
               rt: np.ndarray) -> dict:
    p = pair_paths(directory, index)
    write_obj_vertices(p["src"], src)
    write_obj_vertices(p["tar"], tar)
    write_neigh_bin(p["src_neigh"], src_neigh)
    write_neigh_bin(p["tar_neigh"], tar_neigh)
    write_transform_bin(p["transform"], rt)
    return p


def load_pair(src_obj: str, tar_obj: str, dcp: bool = False, fmr: bool = False) -> dict:
    """Dataset_2021_8_29.__getitem__ (pre_dataloader.py:78-181) without the normals files.  Both clouds and their triplets
    are centred on their own means.  NB the reference's `.transpose(0, 1)` calls on numpy arrays are identity permutations
    (pre_dataloader.py:123,136-160): `rotation` is gt[:3, :3] as stored, R == R_inv == rotation, clouds stay (N, 3) --
    only the DCP switch transposes (pre_dataloader.py:162-173).  tar_box = the 8 corners of the centred target's bounding
    box in libigl's order (corner 0 = all maxima, corner 7 = all minima; the hooks read only these two)."""
    v_src, v_tar = read_obj_vertices(src_obj), read_obj_vertices(tar_obj)
    n_src = read_neigh_bin(src_obj.replace(".obj", "_neigh.bin", 1)).astype(np.float64)
    n_tar = read_neigh_bin(tar_obj.replace(".obj", "_neigh.bin", 1)).astype(np.float64)
    c_tar, c_src = v_tar.mean(0), v_src.mean(0)
    v_tar = v_tar - c_tar
    lo, hi = v_tar.min(0), v_tar.max(0)
    tar_box = np.array([[(lo if (q >> (2 - a)) & 1 else hi)[a] for a in range(3)] for q in range(8)], np.float32)
    v_src = v_src - c_src
    n_src, n_tar = n_src - c_src, n_tar - c_tar
    gt = read_transform_bin(tar_obj.replace("tar_sample", "transform", 1).replace(".obj", ".bin", 1))
    rotation = gt[:3, :3].copy()                                             # pre_dataloader.py:123 (identity transpose)
    translation = gt[:3, 3] + (-c_tar + c_src @ rotation)                    # :124-125
    igt = np.eye(4)
    igt[:3, :3] = rotation
    igt[:3, 3] = -rotation @ translation                                     # :133
    r32, t32 = rotation.astype(np.float32), translation.astype(np.float32)
    data = {"points_tar_sample": v_tar.astype(np.float32), "points_src_sample": v_src.astype(np.float32), "tar_box": tar_box,
            "centers": v_tar.mean(0).astype(np.float32), "R": r32.copy(), "T": t32, "R_inv": r32.copy(), "T_inv": -r32 @ t32,
            "points_based_neighs_src": n_src.astype(np.float32), "points_based_neighs_tar": n_tar.astype(np.float32),
            "igt": igt.astype(np.float32)}
    if dcp:                                               # DCP wants (3, N) clouds and transposed rotations (:162-173)
        for k in ("points_tar_sample", "points_src_sample", "points_based_neighs_src", "points_based_neighs_tar", "R", "R_inv"):
            data[k] = data[k].T
        data["igt"][:3, :3] = data["igt"][:3, :3].T.copy()
    if fmr:                                               # FMR wants equally long clouds (:174-180)
        n = min(data["points_src_sample"].shape[0], data["points_tar_sample"].shape[0])
        data["points_tar_sample"], data["points_src_sample"] = data["points_tar_sample"][:n], data["points_src_sample"][:n]
    return data
