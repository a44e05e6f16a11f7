"""Pre-processing that feeds the loss: farthest-point sampling + k nearest neighbours (SURVEY 8(f) row 2;
reference: loss.py:473-485, utils.py:275-296,380-385).  Runs once per cloud, outside the hot path; expressed with
device-side torch ops (no per-iteration host sync except the FPS start index)."""
import numpy as np
import torch


def farthest_point_sample(xyz: torch.Tensor, npoint: int, start=None) -> torch.Tensor:
    """xyz (N,3) -> (npoint,) indices; same recurrence as utils.farthest_point_sample (utils.py:275-296)."""
    n = xyz.shape[0]
    idx = torch.zeros(npoint, dtype=torch.long, device=xyz.device)
    dist = torch.full((n,), 1e10, device=xyz.device)
    far = torch.randint(0, n, (1,), dtype=torch.long).to(xyz.device)[0] if start is None else torch.tensor(start, device=xyz.device)
    for i in range(npoint):
        idx[i] = far
        d = torch.sum((xyz - xyz[far]) ** 2, -1)
        dist = torch.minimum(dist, d)
        far = torch.argmax(dist)
    return idx


def knn(points: torch.Tensor, queries: torch.Tensor, k: int, chunk: int = 4096) -> torch.Tensor:
    """brute-force exact kNN: (N,3), (Q,3) -> (Q,k) indices sorted by distance"""
    out = []
    for s in range(0, queries.shape[0], chunk):
        d = torch.cdist(queries[s:s + chunk].double(), points.double())
        out.append(torch.topk(d, k, dim=1, largest=False, sorted=True)[1])
    return torch.cat(out, 0)


def sample_neighs(points: np.ndarray, num_sample=5000, num_neigh=3, device="cuda") -> np.ndarray:
    num_sample = min(num_sample, points.shape[0])
    dev = torch.device(device if torch.cuda.is_available() else "cpu")
    p = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).to(dev)
    sel = farthest_point_sample(p, num_sample)
    nn_idx = knn(p, p[sel], num_neigh)
    out = torch.cat([p[nn_idx[:, i]] for i in range(num_neigh)], -1)
    return out.reshape(-1, 3).cpu().numpy()
