"""TEST INFRASTRUCTURE ONLY -- round-2 golden vectors minted by running the UNMODIFIED reference (build container only):

    python -m oracle.make_golden_r2

  expmap.npz        fmr/se_math/se3.py: Exp forward (B,4,4) and ExpMap.backward for random upstream gradients, 24 twists
                    incl. theta = 0 and the Taylor branch
  chamfer_grad.npz  loss.py:236-252 chamfer_dist with autograd gradients w.r.t. both clouds
  se3_log.npz       LieAlgebra/se3.py:124-134 log of exp3 outputs (what Reconstruction_point's (R, T) initialisation calls)
  loader.npz        exps_deep_learning/pre_dataloader.py:78-181 Dataset_2021_8_29.__getitem__ on files written here, for the
                    three layout switches; igl and h5py are not installed, so `igl` is a stub whose read_triangle_mesh parses
                    "v x y z" lines and whose bounding_box follows libigl's corner order (index = 4 X0 + 2 X1 + X2, X = 1
                    picks the minimum) -- `tar_box` is therefore pinned only through its first and last corner, which is all
                    the hooks read (Train_DCP.py:234-236)
"""
import importlib
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
EXPS = os.path.join(ref_loader.REFERENCE_CODE, "exps_deep_learning")


def _import_from(path, name, pkg_dir=None):
    """import `name` with `pkg_dir` first on sys.path; modules the reference drags in but the path never uses
    (matplotlib's 3-D toolkit in fmr/se_math/mesh.py:5) are stubbed for the duration of the import"""
    saved = list(sys.path)
    stubs = []
    for mod in ("mpl_toolkits", "mpl_toolkits.mplot3d", "matplotlib", "matplotlib.pyplot"):
        if mod not in sys.modules:
            try:
                importlib.import_module(mod)
            except Exception:
                m = types.ModuleType(mod)
                m.Axes3D = object
                sys.modules[mod] = m
                stubs.append(mod)
    sys.path.insert(0, pkg_dir or os.path.dirname(path))
    try:
        return importlib.import_module(name)
    finally:
        sys.path[:] = saved
        for mod in stubs:
            sys.modules.pop(mod, None)


def mint_expmap():
    # fmr/se_math/__init__.py also imports mesh.py (matplotlib): register the package by path and import only se3
    pkg = types.ModuleType("se_math")
    pkg.__path__ = [os.path.join(EXPS, "fmr", "se_math")]
    sys.modules["se_math"] = pkg
    try:
        se3 = importlib.import_module("se_math.se3")
    finally:
        for k in [k for k in sys.modules if k == "se_math" or k.startswith("se_math.")]:
            sys.modules.pop(k)
    rng = np.random.default_rng(41)
    x = rng.normal(size=(24, 6)).astype(np.float32)
    x[0] = 0.0                                               # theta = 0
    x[1, :3] *= 1e-3                                         # Taylor branch
    x[2, :3] *= 3e-3
    x[3, :3] *= 2.5                                          # large rotation
    xt = torch.from_numpy(x).clone().requires_grad_(True)
    g = se3.Exp(xt)
    go = torch.from_numpy(rng.normal(size=(24, 4, 4)).astype(np.float32))
    g.backward(go)
    np.savez(os.path.join(OUT, "expmap.npz"), twist=x, grad_g=go.numpy(), ref_g=g.detach().numpy(), ref_grad_twist=xt.grad.numpy())


def mint_chamfer_grad():
    L = ref_loader.load()
    rng = np.random.default_rng(43)
    x = rng.normal(size=(3, 257, 3)).astype(np.float32)
    y = (rng.normal(size=(3, 300, 3)) * 1.1 + 0.1).astype(np.float32)
    xt, yt = torch.from_numpy(x).clone().requires_grad_(True), torch.from_numpy(y).clone().requires_grad_(True)
    out = L.chamfer_dist(xt, yt)
    (out * 1.7).backward()
    np.savez(os.path.join(OUT, "chamfer_grad.npz"), x=x, y=y, ref=np.float32(out.item()), upstream=np.float32(1.7),
             ref_grad_x=xt.grad.numpy(), ref_grad_y=yt.grad.numpy())


def mint_se3_log():
    code = ref_loader.REFERENCE_CODE
    se3 = _import_from(None, "LieAlgebra.se3", pkg_dir=code)
    rng = np.random.default_rng(47)
    x = rng.normal(size=(12, 6)).astype(np.float32) * 0.7
    x[0, :3] *= 1e-4
    R, T = se3.exp3(torch.from_numpy(x))
    g = torch.zeros(12, 4, 4)
    g[:, :3, :3] = R
    g[:, :3, 3] = T
    g[:, 3, 3] = 1
    lg = se3.log(g)
    np.savez(os.path.join(OUT, "se3_log.npz"), twist=x, R=R.numpy(), T=T.numpy(), ref_log=lg.numpy())


def mint_loader():
    igl = types.ModuleType("igl")

    def read_triangle_mesh(path):
        v = np.array([[float(t) for t in ln.split()[1:4]] for ln in open(path) if ln.startswith("v ")], np.float64)
        return v, np.zeros((0, 3), np.int64)

    def bounding_box(V):
        lo, hi = V.min(0), V.max(0)
        BV = np.array([[(lo if (q >> (2 - a)) & 1 else hi)[a] for a in range(3)] for q in range(8)], np.float64)
        return BV, np.zeros((12, 3), np.int64)
    igl.read_triangle_mesh, igl.bounding_box = read_triangle_mesh, bounding_box
    saved = {k: sys.modules.get(k) for k in ("igl", "h5py")}
    sys.modules["igl"] = igl
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    try:
        mod = _import_from(None, "pre_dataloader", pkg_dir=EXPS)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    rng = np.random.default_rng(53)
    d = tempfile.mkdtemp(prefix="rrl_loader_")
    src = rng.normal(size=(40, 3)) + np.array([0.5, -1.0, 2.0])
    tar = rng.normal(size=(55, 3)) * 1.3 + np.array([-2.0, 0.3, 0.7])
    n_src = rng.normal(size=(120, 3)).astype(np.float32)
    n_tar = rng.normal(size=(165, 3)).astype(np.float32)
    A = rng.normal(size=(3, 3))
    Q, _ = np.linalg.qr(A)
    rt = np.concatenate([Q, rng.normal(size=(3, 1))], 1)

    def wobj(path, v):
        with open(path, "w") as f:
            for p in v:
                f.write("v %.17g %.17g %.17g\n" % tuple(p))
    p_src, p_tar = os.path.join(d, "7_src_sample.obj"), os.path.join(d, "7_tar_sample.obj")
    wobj(p_src, src); wobj(p_tar, tar)
    wobj(p_src.replace("sample", "sample_normals", 1), src); wobj(p_tar.replace("sample", "sample_normals", 1), tar)
    n_src.tofile(p_src.replace(".obj", "_neigh.bin", 1)); n_tar.tofile(p_tar.replace(".obj", "_neigh.bin", 1))
    rt.astype(np.float64).tofile(p_tar.replace("tar_sample", "transform", 1).replace(".obj", ".bin", 1))
    out = dict(src=src, tar=tar, n_src=n_src, n_tar=n_tar, rt=rt)
    for tag, kw in (("plain", {}), ("dcp", dict(DCP_True=True)), ("fmr", dict(FMR_True=True))):
        ds = mod.Dataset_2021_8_29([p_src], [p_tar], **kw)
        item = ds[0]
        for k, v in item.items():
            if k.startswith("normals"):
                continue
            out["ref_%s_%s" % (tag, k)] = np.asarray(v)
    np.savez(os.path.join(OUT, "loader.npz"), **out)


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference sources not present")
    mint_expmap()
    mint_chamfer_grad()
    mint_se3_log()
    mint_loader()
    print("written to", OUT)
