"""Bench and test inputs -- deterministic synthetic clouds, triplets and lines.

Follows SURVEY.md 8(d) "Synthetic inputs": points on a closed surface scaled to the extent of
the shipped sample data, triplets = point + its 2 nearest neighbours (the layout
Sample_neighs produces, /root/reference/code/loss.py:473-485), target = rigid transform of an
independent resample, lines = chords of a sphere around the target (loss.py:384-412) that
cross both axis-aligned bounding boxes (geometric slab test; the reference's own rejection
sampler is restated in oracle/torch_port.py and oracle/rrl_oracle.c).

numpy only, no torch: importable everywhere.  This module generates INPUTS; it contains no
part of the loss.
"""
import numpy as np


def surface_points(rng: np.random.Generator, n: int, shape: str = "sphere", extent=(3.6, 9.6, 5.2)) -> np.ndarray:
    if shape == "sphere":
        v = rng.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
    elif shape == "torus":
        a, b = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
        v = np.stack([(1 + 0.35 * np.cos(b)) * np.cos(a), (1 + 0.35 * np.cos(b)) * np.sin(a), 0.35 * np.sin(b)], 1)
    elif shape == "box":
        v = rng.uniform(-1, 1, size=(n, 3))
        ax = rng.integers(0, 3, n)
        v[np.arange(n), ax] = np.sign(v[np.arange(n), ax])
    else:
        raise ValueError(shape)
    return (v * (np.asarray(extent, np.float64) / 2)).astype(np.float32)


def knn_triplets(pts: np.ndarray) -> np.ndarray:
    """(n,3) -> (n,9) rows [p, nn1(p), nn2(p)] (exact kNN; neighbour 0 is the point itself)."""
    n = pts.shape[0]
    if n <= 8192:
        p = pts.astype(np.float64)
        d = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
        idx = np.argpartition(d, 2, axis=1)[:, :3]
        order = np.take_along_axis(d, idx, 1).argsort(1)
        idx = np.take_along_axis(idx, order, 1)
    else:
        from scipy.spatial import cKDTree
        idx = cKDTree(pts).query(pts, 3)[1]
    return np.concatenate([pts[idx[:, 0]], pts[idx[:, 1]], pts[idx[:, 2]]], 1).astype(np.float32)


def random_rotation(rng: np.random.Generator, max_deg: float) -> np.ndarray:
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0, max_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def chord_lines(rng: np.random.Generator, n: int, r: float, center, lo1, hi1, lo2, hi2, zero_frac: float = 0.0):
    """n chords of the sphere (center, r) that cross both AABBs; the last zero_frac*n rows are all-zero
    (the reference leaves unfilled rows zero and still evaluates them, loss.py:423-432)."""
    out = np.zeros((n, 6), np.float32)
    want = n - int(round(zero_frac * n))
    got = 0
    center = np.asarray(center, np.float64)
    while got < want:
        m = max(4 * (want - got), 1024)
        def sph():
            a = rng.uniform(0, 2 * np.pi, m)
            z = rng.uniform(-1, 1, m)
            s = np.sqrt(1 - z * z)
            return r * np.stack([s * np.cos(a), s * np.sin(a), z], 1)
        q1, q2 = sph(), sph()
        d = q2 - q1
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
        x0 = q1 + center
        ok = np.ones(m, bool)
        for lo, hi in ((lo1, hi1), (lo2, hi2)):
            with np.errstate(divide="ignore", invalid="ignore"):
                t0 = (np.asarray(lo) - x0) / d
                t1 = (np.asarray(hi) - x0) / d
            tn = np.minimum(t0, t1).max(1)
            tf = np.maximum(t0, t1).min(1)
            ok &= tn <= tf
        cand = np.concatenate([d, x0], 1)[ok].astype(np.float32)
        take = min(want - got, cand.shape[0])
        out[got:got + take] = cand[:take]
        got += take
    return out


def make_pair(seed: int, nf: int, nl: int, shape: str = "sphere", radius_scale: float = 0.5, noise: float = 0.0,
              outlier_frac: float = 0.0, keep_frac: float = 1.0, zero_frac: float = 0.0, nf2: int = None,
              max_rot_deg: float = 45.0, max_trans: float = 0.5):
    """One synthetic registration pair.  Returns dict(tri1 (nf,9), tri2 (nf2,9), lines (nl,6), src, tgt, R, t)."""
    rng = np.random.default_rng(seed)
    nf2 = nf if nf2 is None else nf2
    src = surface_points(rng, nf, shape)
    tgt0 = surface_points(rng, nf2, shape)
    Rm = random_rotation(rng, max_rot_deg)
    t = rng.uniform(-max_trans, max_trans, 3)
    tgt = (tgt0.astype(np.float64) @ Rm.T + t).astype(np.float32)
    if noise > 0:
        tgt = tgt + np.clip(rng.normal(scale=noise, size=tgt.shape), -5 * noise, 5 * noise).astype(np.float32)
    if keep_frac < 1.0:                      # partial overlap: half-space crop, refill by resampling kept points
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        s = tgt @ nrm
        keep = s <= np.quantile(s, keep_frac)
        kept = tgt[keep]
        tgt = kept[rng.integers(0, kept.shape[0], nf2)] + rng.normal(scale=1e-3, size=(nf2, 3)).astype(np.float32)
        tgt = tgt.astype(np.float32)
    if outlier_frac > 0:
        k = int(outlier_frac * nf2)
        lo, hi = tgt.min(0), tgt.max(0)
        tgt[rng.choice(nf2, k, replace=False)] = rng.uniform(lo, hi, size=(k, 3)).astype(np.float32)
    tri1, tri2 = knn_triplets(src), knn_triplets(tgt)
    lo2, hi2 = tgt.min(0), tgt.max(0)
    lo1, hi1 = src.min(0), src.max(0)
    r = radius_scale * float(np.linalg.norm(hi2 - lo2))
    lines = chord_lines(rng, nl, r, tgt.mean(0), lo1, hi1, lo2, hi2, zero_frac)
    return dict(tri1=tri1, tri2=tri2, lines=lines, src=src, tgt=tgt, R=Rm.astype(np.float32), t=t.astype(np.float32),
                radius=np.float32(r))
