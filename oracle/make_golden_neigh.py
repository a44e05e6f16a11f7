"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/sample_neighs.npz by running the UNMODIFIED reference's Sample_neighs
(code/loss.py:473-485) and utils.farthest_point_sample (code/utils.py:275-296) in the build container:

    python -m oracle.make_golden_neigh

Two float32 clouds (the demo casts igl's vertices to float32 before the call, test_demo_optimized_Lie_Algebra.py:112-118;
a float64 cloud makes the reference raise in utils.py:294 on current torch): the first 3000 vertices of
sample_data/challenge_data/0_src_sample.obj and a seeded random cloud.  `ref_*` arrays are reference outputs.
"""
import os
import sys

import numpy as np
import torch

from . import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    L = ref_loader.load()
    import utils as ref_utils  # the reference's utils, imported by its loss module
    obj = os.path.join(ref_loader.REFERENCE_CODE, "sample_data", "challenge_data", "0_src_sample.obj")
    va = np.array([[float(t) for t in ln.split()[1:4]] for ln in open(obj) if ln.startswith("v ")], np.float32)[:3000]
    rng = np.random.default_rng(11)
    v32 = (rng.standard_normal((2000, 3)) * np.array([1.0, 0.6, 0.3])).astype(np.float32)
    out = {}
    for tag, pts, ns, seed in (("a", va, 400, 7), ("b", v32, 256, 8)):
        torch.manual_seed(seed)
        fps_idx = ref_utils.farthest_point_sample(torch.from_numpy(pts).unsqueeze(0), ns)[0].numpy()
        torch.manual_seed(seed)
        ref = L.Sample_neighs(pts, num_sample=ns, num_neigh=3)
        assert np.array_equal(ref.reshape(-1, 3, 3)[:, 0], pts[fps_idx]), "FPS replay differs from Sample_neighs"
        out.update({tag + "_points": pts, tag + "_num_sample": np.int32(ns), tag + "_seed": np.int32(seed),
                    tag + "_ref_fps_idx": fps_idx.astype(np.int32), tag + "_ref_neighs": ref})
        print(tag, pts.dtype, pts.shape, "->", ref.dtype, ref.shape, "start", fps_idx[0])
    np.savez_compressed(os.path.join(OUT, "sample_neighs.npz"), **out)
    print(os.path.getsize(os.path.join(OUT, "sample_neighs.npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
