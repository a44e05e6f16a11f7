"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/synth_short_dirs.npz by running the UNMODIFIED reference
(build container only; see oracle/make_golden.py).

    python -m oracle.make_golden_edges

Edge case the reference never checks (`line` is used as given): directions that are NOT unit vectors.
Only |u| <= 1 can be minted from the reference -- for |u| > 1 its (AC.u)^2 can exceed |AC|^2, the sqrt argument goes
negative and loss.py:89-91 exits the process -- so the case holds too-short directions (|u| in 0.3 .. 1), all-zero rows
(what the sampler leaves in unfilled rows), duplicated lines and a cloud with duplicated triplets.
"""
import os

import numpy as np

from . import make_golden as mg
from . import ref_loader, synth


def main():
    L = ref_loader.load()
    p = synth.make_pair(seed=21, nf=320, nl=1400, zero_frac=0.15)
    lines = p["lines"].copy()
    rng = np.random.default_rng(22)
    lines[:500, :3] *= rng.uniform(0.3, 1.0, size=(500, 1)).astype(np.float32)      # too short
    lines[500:560] = lines[3]                                                        # duplicates
    tri1 = p["tri1"].copy()
    tri1[100:140] = tri1[60:100]                                                     # duplicated triplets
    case = mg.ref_loss_case(L, tri1, p["tri2"], lines)
    assert case["ref_none"] == 0
    print("synth_short_dirs", case["ref_loss"], int(case["ref_counts1"].sum()), int(case["ref_counts2"].sum()))
    np.savez_compressed(os.path.join(mg.OUT, "synth_short_dirs.npz"), **case)


if __name__ == "__main__":
    main()
