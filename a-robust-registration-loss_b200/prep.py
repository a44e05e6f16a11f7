"""Pre-processing that feeds the loss: farthest-point sampling + k nearest neighbours (reference:
loss.py:473-485, utils.py:275-296,380-385).  Runs once per cloud, outside the hot path, on the hand-written kernels of
csrc/rrl_neigh.cu through the C ABI (rrl_fps, rrl_knn); torch only holds the device memory."""
import numpy as np
import torch

from . import _native as N


def _cloud(xyz: torch.Tensor) -> torch.Tensor:
    if not xyz.is_cuda:
        raise N.NativeError("rrl_b200.prep needs CUDA tensors: there is no CPU path")
    if xyz.dim() != 2 or xyz.shape[1] != 3:
        raise ValueError("expected one cloud of shape (N, 3)")
    if xyz.dtype not in (torch.float32, torch.float64):
        xyz = xyz.float()
    return xyz.contiguous()


def farthest_point_sample(xyz: torch.Tensor, npoint: int, start=None) -> torch.Tensor:
    """xyz (N,3) cuda float32/float64 -> (npoint,) int64 indices; the recurrence of utils.farthest_point_sample
    (utils.py:275-296).  `start` = first centroid; by default drawn like the reference does (torch.randint on the CPU
    generator, utils.py:288), so torch.manual_seed reproduces the reference's selection."""
    xyz = _cloud(xyz)
    n = xyz.shape[0]
    if not 0 < npoint <= n:
        raise ValueError("npoint must lie in 1..N")
    if start is None:
        start = int(torch.randint(0, n, (1,), dtype=torch.long)[0])
    L = N.lib()
    wsb = L.rrl_fps_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=xyz.device)
    out = torch.empty(npoint, dtype=torch.int32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        N.check(L.rrl_fps(xyz.data_ptr(), int(xyz.dtype == torch.float64), n, int(npoint), int(start), out.data_ptr(),
                          ws.data_ptr(), wsb, torch.cuda.current_stream(xyz.device).cuda_stream), "rrl_fps")
    return out.long()


def knn(points: torch.Tensor, query_idx: torch.Tensor, k: int) -> torch.Tensor:
    """exact k nearest neighbours (k <= 8) of points[query_idx] among `points` (N,3): (M,k) int64, nearest first --
    the query point itself comes first (what KDTree.query returns in loss.py:479-480)"""
    points = _cloud(points)
    q = query_idx.to(device=points.device, dtype=torch.int32).contiguous()
    out = torch.empty(q.shape[0], k, dtype=torch.int32, device=points.device)
    with torch.cuda.device(points.device):
        N.check(N.lib().rrl_knn(points.data_ptr(), int(points.dtype == torch.float64), points.shape[0], q.data_ptr(),
                                q.shape[0], int(k), out.data_ptr(), torch.cuda.current_stream(points.device).cuda_stream),
                "rrl_knn")
    return out.long()


def sample_neighs(points: np.ndarray, num_sample=5000, num_neigh=3, device="cuda") -> np.ndarray:
    """Sample_neighs (loss.py:473-485): (N,3) array -> (num_neigh * min(N, num_sample), 3) rows [self, nn1, nn2] per
    sampled point, in the input's dtype."""
    num_sample = min(num_sample, points.shape[0])
    dev = torch.device("cuda" if device in ("cpu", None) else device)        # the reference's default 'cpu' means "wherever"
    p = torch.from_numpy(np.ascontiguousarray(points)).to(dev)
    sel = farthest_point_sample(p, num_sample)
    nn_idx = knn(p, sel, num_neigh)
    out = torch.cat([p[nn_idx[:, i]] for i in range(num_neigh)], -1)
    return out.reshape(-1, 3).cpu().numpy().astype(points.dtype, copy=False)
