"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's Sample_neighs (code/loss.py:473-485):
farthest-point sampling (code/utils.py:275-296, via Sample_points utils.py:380-385) followed by a k-nearest-neighbour
query (sklearn KDTree on the float64-converted cloud, loss.py:479-480).  Pinned by tests/golden/sample_neighs.npz,
which oracle/make_golden_neigh.py minted by running the unmodified reference.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this module; the product path never does.
"""
import numpy as np


def fps(xyz: np.ndarray, npoint: int, start: int) -> np.ndarray:
    """utils.py:275-296 for B = 1.  The running minimum is float32 (torch.ones(B, N) * 1e10, utils.py:287); the next
    centroid is the first index of the maximum (torch.max, utils.py:295).  The reference is only ever fed float32 (the
    demo casts, test_demo_optimized_Lie_Algebra.py:114-115; float64 raises in utils.py:294 on current torch); for
    float64 input this restatement compares in float64 and rounds on assignment, which is what an implicit cast does."""
    n = xyz.shape[0]
    distance = np.full(n, 1e10, np.float32)
    idx = np.zeros(npoint, np.int64)
    far = int(start)
    for i in range(npoint):
        idx[i] = far
        d = xyz - xyz[far]
        dist = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]          # torch.sum over a size-3 dim: ((a+b)+c)
        mask = dist < distance
        distance[mask] = dist[mask]
        far = int(np.argmax(distance))
    return idx


def knn(points: np.ndarray, queries: np.ndarray, k: int) -> np.ndarray:
    """exact k nearest neighbours, nearest first, squared distances in float64 (loss.py:479-480); ties: smaller index"""
    p = points.astype(np.float64)
    out = np.zeros((queries.shape[0], k), np.int64)
    for s in range(0, queries.shape[0], 256):
        q = queries[s:s + 256].astype(np.float64)
        d = q[:, None, :] - p[None, :, :]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        out[s:s + 256] = np.argsort(d2, axis=1, kind="stable")[:, :k]
    return out


def sample_neighs(points: np.ndarray, num_sample: int, num_neigh: int, start: int) -> np.ndarray:
    """loss.py:473-485 with the torch.randint start index supplied: (num_neigh * min(N, num_sample), 3), rows
    [self, nn1, nn2, ...] per sampled point, in the input's dtype"""
    num_sample = min(num_sample, points.shape[0])
    sel = fps(points, num_sample, start)
    nn = knn(points, points[sel], num_neigh)
    return np.concatenate([points[nn[:, i]].reshape(-1, 3) for i in range(num_neigh)], -1).reshape(-1, 3)
