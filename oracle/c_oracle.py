"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/rrl_oracle.c (numpy in, numpy out).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import build as _build  # type: ignore  # noqa: F401  (package-relative when imported as oracle.c_oracle)

CAP = 5
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.build()
        _lib = C.CDLL(path)
        f32p, i32p, i64p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_double))
        _lib.rrl_oracle_threads.restype = C.c_int
        _lib.rrl_oracle_set_threads.argtypes = [C.c_int]
        _lib.rrl_oracle_triplet_thr.restype = C.c_float
        _lib.rrl_oracle_triplet_thr.argtypes = [f32p]
        _lib.rrl_oracle_point_line_d.restype = C.c_float
        _lib.rrl_oracle_point_line_d.argtypes = [f32p, f32p]
        _lib.rrl_oracle_dense.restype = None
        _lib.rrl_oracle_dense.argtypes = [f32p, C.c_int, f32p, C.c_int64, i32p, i32p, f32p, i64p]
        _lib.rrl_oracle_loss.restype = C.c_int
        _lib.rrl_oracle_loss.argtypes = [f32p, C.c_int, f32p, C.c_int, f32p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                         C.c_int, f32p, i64p, i64p, f32p, f32p, i32p, i32p, i32p, i32p, f32p]
        _lib.rrl_oracle_se3_exp.argtypes = [f32p, f32p, f32p]
        _lib.rrl_oracle_rigid_apply.argtypes = [f32p, f32p, f32p, C.c_int64, f32p]
        _lib.rrl_oracle_se3_backward.argtypes = [f32p, f64p, f64p, f64p]
        _lib.rrl_oracle_box_triangles.argtypes = [f32p, f32p, f32p]
        _lib.rrl_oracle_triangle_hits.restype = C.c_int
        _lib.rrl_oracle_triangle_hits.argtypes = [f32p, f32p]
        _lib.rrl_oracle_sample_lines.restype = C.c_int64
        _lib.rrl_oracle_sample_lines.argtypes = [C.c_float, f32p, C.c_int64, f32p, f32p, f32p, f32p, f32p, C.c_int,
                                                 f32p]
        _lib.rrl_oracle_chamfer.restype = C.c_float
        _lib.rrl_oracle_chamfer.argtypes = [f32p, C.c_int64, f32p, C.c_int64]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def threads() -> int:
    return lib().rrl_oracle_threads()


def set_threads(n: int) -> None:
    lib().rrl_oracle_set_threads(int(n))


def triplet_thr(tri: np.ndarray) -> np.ndarray:
    tri = _f32(tri).reshape(-1, 9)
    return np.array([lib().rrl_oracle_triplet_thr(_p(tri[i])) for i in range(tri.shape[0])], np.float32)


@dataclass
class Dense:
    counts: np.ndarray      # (nl,) int32, uncapped
    hits: np.ndarray        # (nl, CAP) int32 ascending, -1 padded
    hit_d: np.ndarray       # (nl, CAP, 3) float32
    nan: int
    band: int


def dense(tri: np.ndarray, lines: np.ndarray) -> Dense:
    tri = _f32(tri).reshape(-1, 9)
    lines = _f32(lines).reshape(-1, 6)
    nl = lines.shape[0]
    counts = np.zeros(nl, np.int32)
    hits = np.zeros((nl, CAP), np.int32)
    hd = np.zeros((nl, CAP, 3), np.float32)
    stats = np.zeros(2, np.int64)
    lib().rrl_oracle_dense(_p(tri), tri.shape[0], _p(lines), nl, _p(counts, C.c_int32), _p(hits, C.c_int32), _p(hd),
                           _p(stats, C.c_int64))
    return Dense(counts, hits, hd, int(stats[0]), int(stats[1]))


@dataclass
class Loss:
    status: int             # 0 ok, 1 = no populated combo (reference returns (None, None, None))
    loss: float
    median: float
    n_combos: int
    n_selected: int
    n_entries: int
    nan: int
    band: int
    n_kj: np.ndarray        # (4,4) int64
    grad1: Optional[np.ndarray]
    grad2: Optional[np.ndarray]
    counts1: np.ndarray
    hits1: np.ndarray
    counts2: np.ndarray
    hits2: np.ndarray
    D: Optional[np.ndarray]  # (nl, 4, 4)


def loss(tri1, tri2, lines, k_lo=1, j_lo=1, k_hi=5, j_hi=5, want_grad=True, want_D=False) -> Loss:
    tri1 = _f32(tri1).reshape(-1, 9)
    tri2 = _f32(tri2).reshape(-1, 9)
    lines = _f32(lines).reshape(-1, 6)
    nl = lines.shape[0]
    sc = np.zeros(2, np.float32)
    cn = np.zeros(5, np.int64)
    nkj = np.zeros(16, np.int64)
    g1 = np.zeros_like(tri1) if want_grad else None
    g2 = np.zeros_like(tri2) if want_grad else None
    c1 = np.zeros(nl, np.int32)
    c2 = np.zeros(nl, np.int32)
    h1 = np.zeros((nl, CAP), np.int32)
    h2 = np.zeros((nl, CAP), np.int32)
    D = np.zeros((nl, 4, 4), np.float32) if want_D else None
    rc = lib().rrl_oracle_loss(_p(tri1), tri1.shape[0], _p(tri2), tri2.shape[0], _p(lines), nl, k_lo, j_lo, k_hi, j_hi,
                               _p(sc), _p(cn, C.c_int64), _p(nkj, C.c_int64), _p(g1), _p(g2), _p(c1, C.c_int32),
                               _p(h1, C.c_int32), _p(c2, C.c_int32), _p(h2, C.c_int32), _p(D))
    return Loss(rc, float(sc[0]), float(sc[1]), int(cn[0]), int(cn[1]), int(cn[2]), int(cn[3]), int(cn[4]),
                nkj.reshape(4, 4), g1, g2, c1, h1, c2, h2, D)


def se3_exp(twist):
    tw = _f32(twist).reshape(6)
    R = np.zeros(9, np.float32)
    T = np.zeros(3, np.float32)
    lib().rrl_oracle_se3_exp(_p(tw), _p(R), _p(T))
    return R.reshape(3, 3), T


def rigid_apply(R, T, pts):
    R = _f32(R).reshape(9)
    T = _f32(T).reshape(3)
    pts = _f32(pts).reshape(-1, 3)
    out = np.zeros_like(pts)
    lib().rrl_oracle_rigid_apply(_p(R), _p(T), _p(pts), pts.shape[0], _p(out))
    return out


def se3_backward(twist, pts, grad_pts):
    """d loss / d twist given d loss / d (transformed points); pts are the untransformed points."""
    tw = _f32(twist).reshape(6)
    p = np.asarray(pts, np.float64).reshape(-1, 3)
    g = np.asarray(grad_pts, np.float64).reshape(-1, 3)
    GR = np.ascontiguousarray(p.T @ g)
    GT = np.ascontiguousarray(g.sum(0))
    out = np.zeros(6, np.float64)
    lib().rrl_oracle_se3_backward(_p(tw), _p(GR, C.c_double), _p(GT, C.c_double), _p(out, C.c_double))
    return out


def sample_lines(r, center, n, lo1, hi1, lo2, hi2, uniforms):
    """uniforms: (rounds, 4, n) float32 in [0,1).  Returns (lines (n,6), filled)."""
    u = _f32(uniforms)
    rounds = u.shape[0]
    out = np.zeros((n, 6), np.float32)
    filled = lib().rrl_oracle_sample_lines(float(r), _p(_f32(center)), n, _p(_f32(lo1)), _p(_f32(hi1)), _p(_f32(lo2)),
                                           _p(_f32(hi2)), _p(u), rounds, _p(out))
    return out, int(filled)


def chamfer(x, y) -> float:
    x = _f32(x).reshape(-1, 3)
    y = _f32(y).reshape(-1, 3)
    return float(lib().rrl_oracle_chamfer(_p(x), x.shape[0], _p(y), y.shape[0]))
