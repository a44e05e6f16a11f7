"""TEST INFRASTRUCTURE ONLY -- eager-PyTorch CPU restatement of the reference hot path.

This is one of the two oracles (the other is the plain-C restatement in
oracle/rrl_oracle.c).  It re-expresses, with the *same ATen operators in the same
order* as the reference (so results are bit-identical to the reference on the same
torch build), the functions of SURVEY.md section 8(a):

    a1  cal_intersection_batch2_points_with_line      /root/reference/code/loss.py:68-112
    a2  cal_loss_intersection_batch_m_n_median_...    /root/reference/code/loss.py:115-167
    a3  cal_loss_intersection_batch_whole_median_...  /root/reference/code/loss.py:170-232
    a5  Reconstruction_point / se3.exp3               loss.py:437-463, LieAlgebra/se3.py:83-106
    a6  Random_uniform_distribution_lines_batch_efficient           loss.py:384-412
    a7  ..._resample + AABB 12-triangle rejection                   loss.py:265-432

Differences from the reference, on purpose:
  * one pair at a time (the reference's B>1 behaviour is broken);
  * the dense phase is chunked over lines so that shapes the reference cannot hold
    in RAM still run (values are unchanged: every op is elementwise in the line axis);
  * the sparse phase gathers the <=4 hit triplets per line straight from the input
    points instead of from an (nl, nf, 9) expanded copy, so autograd's backward is
    sparse.  Same values, much cheaper than the reference -> when this file is used
    as the CPU baseline in bench.py it is a *conservative* (faster-than-reference)
    baseline.  It is reported as kind="port".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path never does.

Parity status: pinned against the unmodified reference run in the build container
(tests/test_oracle_vs_reference.py, fixtures minted by oracle/make_golden.py).  The
reference itself ships no tests or golden vectors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

ADD_EPS = 2e-4          # loss.py:88
THR_SCALE = 1.731       # loss.py:109
MAX_ROUNDS = 10         # loss.py:425


# --------------------------------------------------------------------------------------
# a1: dense point<->line phase
# --------------------------------------------------------------------------------------
def triplet_threshold(tri: torch.Tensor) -> torch.Tensor:
    """thr_f = (delta_f * 1.731) / 2 with delta_f the mean triplet edge (loss.py:94-109)."""
    p0, p1, p2 = tri[:, 0:3], tri[:, 3:6], tri[:, 6:9]
    e01 = torch.sqrt(torch.sum((p1 - p0) ** 2, -1))
    e02 = torch.sqrt(torch.sum((p2 - p0) ** 2, -1))
    e12 = torch.sqrt(torch.sum((p1 - p2) ** 2, -1))
    delta = torch.mean(torch.stack([e01, e02, e12], -1), -1)
    return delta * THR_SCALE / 2


def point_line_distances(tri: torch.Tensor, lines: torch.Tensor) -> torch.Tensor:
    """d[l, f, i] = sqrt(|AC|^2 - (AC.u)^2 + 2e-4)  (loss.py:84-88), shape (nl, nf, 3)."""
    nf = tri.shape[0]
    nl = lines.shape[0]
    pts = tri.view(1, nf, 3, 3)
    x0 = lines[:, 3:6].reshape(nl, 1, 1, 3)
    u = lines[:, 0:3].reshape(nl, 1, 1, 3)
    ac = pts - x0
    proj = torch.sum(ac * u, -1) ** 2
    dac = torch.sum(ac * ac, -1)
    return torch.sqrt(dac - proj + ADD_EPS)


@dataclass
class DenseResult:
    counts: torch.Tensor                    # (nl,) int64: number of hit triplets per line
    hit_lines: torch.Tensor                 # (H,) int64 line index of every hit, line-major
    hit_tris: torch.Tensor                  # (H,) int64 triplet index, ascending inside a line
    hit_d: torch.Tensor                     # (H, 3) float32 distances of the hit triplet
    nan_seen: bool = False
    band: int = 0                           # tests with |d - thr| <= 1 ulp(thr) that can decide the label


def dense_phase(tri: torch.Tensor, lines: torch.Tensor, chunk: int = 1024,
                count_band: bool = False) -> DenseResult:
    """Labels of loss.py:107-110 in sparse form, computed chunk-by-chunk over lines."""
    tri = tri.detach()
    thr = triplet_threshold(tri).view(1, -1, 1)
    nl = lines.shape[0]
    counts = torch.zeros(nl, dtype=torch.int64)
    hl, ht, hd = [], [], []
    nan_seen = False
    band = 0
    for s in range(0, nl, chunk):
        d = point_line_distances(tri, lines[s:s + chunk])
        if torch.isnan(d).any():
            nan_seen = True                      # the reference prints and exit(0)s here (loss.py:89-91)
        label = (d < thr).sum(-1) == 3
        if count_band:
            ulp = torch.nextafter(thr, thr + 1) - thr
            inb = (d - thr).abs() <= ulp                     # within 1 ulp of the threshold ...
            okb = (d < thr) | inb
            others = torch.stack([okb[..., 1] & okb[..., 2], okb[..., 0] & okb[..., 2], okb[..., 0] & okb[..., 1]], -1)
            band += int((inb & others).sum())                # ... and able to decide the label (the other two pass or are in the band)
        counts[s:s + chunk] = label.sum(-1)
        nz = label.nonzero()
        hl.append(nz[:, 0] + s)
        ht.append(nz[:, 1])
        hd.append(d[nz[:, 0], nz[:, 1]])
    return DenseResult(counts, torch.cat(hl), torch.cat(ht), torch.cat(hd), nan_seen, band)


# --------------------------------------------------------------------------------------
# a2 + a3: sparse phase and the loss
# --------------------------------------------------------------------------------------
@dataclass
class LossTrace:
    loss: Optional[torch.Tensor]                       # (1,) or None when no combo is populated
    median: Optional[torch.Tensor] = None
    n_combos: int = 0
    n_kj: Dict[Tuple[int, int], int] = field(default_factory=dict)
    lines_kj: Dict[Tuple[int, int], torch.Tensor] = field(default_factory=dict)   # line ids per combo
    idx1_kj: Dict[Tuple[int, int], torch.Tensor] = field(default_factory=dict)    # (L,k) triplet ids
    idx2_kj: Dict[Tuple[int, int], torch.Tensor] = field(default_factory=dict)    # (L,j)
    D_kj: Dict[Tuple[int, int], torch.Tensor] = field(default_factory=dict)       # (L,k,j)
    dense1: Optional[DenseResult] = None
    dense2: Optional[DenseResult] = None


def _group_hits(dr: DenseResult, lines_sel: torch.Tensor, k: int):
    """For the selected lines (each with exactly k hits) return (L,k) triplet ids and (L,k,3) d."""
    starts = torch.cumsum(dr.counts, 0) - dr.counts
    base = starts[lines_sel].view(-1, 1) + torch.arange(k).view(1, -1)
    return dr.hit_tris[base], dr.hit_d[base]


def _intersection_points(tri: torch.Tensor, idx: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """q = mean_i(w_i * p_i) with w = d / sum(d) detached (loss.py:92,112,155-163)."""
    w = (d / torch.sum(d, -1, keepdim=True)).detach()            # (L,k,3)
    p = tri[idx].view(idx.shape[0], idx.shape[1], 3, 3)          # (L,k,nei,xyz)
    wp = w.unsqueeze(-1) * p                                     # (L,k,nei,xyz)
    return torch.mean(wp.transpose(-1, -2), -1)                  # mean over the 3 neighbours


def sqdist_map(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """(L,k,3),(L,j,3) -> (L,k,j) squared distances (loss.py:38-52)."""
    return torch.sum((x.unsqueeze(2) - y.unsqueeze(1)) ** 2, -1)


def welsch(x: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    return 1 - torch.exp(-((x / c)) / 2.0)                       # loss.py:20-21


def loss_pair(tri1: torch.Tensor, tri2: torch.Tensor, lines: torch.Tensor,
              k_lo: int = 1, j_lo: int = 1, k_hi: int = 5, j_hi: int = 5,
              chunk: int = 1024, trace: bool = False):
    """Restatement of cal_loss_intersection_batch_whole_median_pts_lines for ONE pair.

    tri1 (nf1,9), tri2 (nf2,9), lines (nl,6).  Returns a (1,) tensor (differentiable
    w.r.t. tri1 / tri2) or None if no (k,j) combo is populated (the reference returns the
    tuple (None, None, None), loss.py:232).
    """
    d1 = dense_phase(tri1, lines, chunk)
    d2 = dense_phase(tri2, lines, chunk)
    tr = LossTrace(None, dense1=d1, dense2=d2)
    maps: List[torch.Tensor] = []
    wts: List[float] = []
    for k in range(k_lo, k_hi):
        for j in range(j_lo, j_hi):
            sel = ((d1.counts == k) & (d2.counts == j)).nonzero().reshape(-1)
            if sel.numel() == 0:
                continue
            i1, dd1 = _group_hits(d1, sel, k)
            i2, dd2 = _group_hits(d2, sel, j)
            q1 = _intersection_points(tri1, i1, dd1)
            q2 = _intersection_points(tri2, i2, dd2)
            D = sqdist_map(q1, q2)
            maps.append(D)
            wts.append(k - j)
            tr.n_kj[(k, j)] = int(sel.numel())
            tr.lines_kj[(k, j)] = sel
            tr.idx1_kj[(k, j)] = i1
            tr.idx2_kj[(k, j)] = i2
            tr.D_kj[(k, j)] = D.detach()
    if not maps:
        return (None, tr) if trace else None
    med = torch.median(torch.cat([m.reshape(1, -1) for m in maps], -1)).detach()
    loss = torch.zeros(1)
    for D, kj in zip(maps, wts):
        W = welsch(D, 1 * med)
        w_kj = torch.exp(torch.FloatTensor([-0.5 * abs(kj)]))
        loss = loss + w_kj * (torch.mean(torch.min(W, 2)[0]) + torch.mean(torch.min(W, 1)[0]))
    loss = loss / len(maps)
    tr.loss, tr.median, tr.n_combos = loss, med, len(maps)
    return (loss, tr) if trace else loss


# --------------------------------------------------------------------------------------
# a5: se(3) exponential + rigid transform
# --------------------------------------------------------------------------------------
def _sinc_branch(t: torch.Tensor, small, large) -> torch.Tensor:
    out = torch.zeros_like(t)
    s = torch.abs(t) < 0.01
    c = ~s
    out[s] = small(t[s])
    out[c] = large(t[c])
    return out


def sinc1(t):
    return _sinc_branch(t, lambda x: 1 - x ** 2 / 6 * (1 - x ** 2 / 20 * (1 - x ** 2 / 42)),
                        lambda x: torch.sin(x) / x)


def sinc2(t):
    return _sinc_branch(t, lambda x: 1 / 2 * (1 - x ** 2 / 12 * (1 - x ** 2 / 30 * (1 - x ** 2 / 56))),
                        lambda x: (1 - torch.cos(x)) / x ** 2)


def sinc3(t):
    return _sinc_branch(t, lambda x: 1 / 6 * (1 - x ** 2 / 20 * (1 - x ** 2 / 42 * (1 - x ** 2 / 72))),
                        lambda x: (x - torch.sin(x)) / (x ** 3))


def hat(w: torch.Tensor) -> torch.Tensor:
    w = w.view(-1, 3)
    z = torch.zeros_like(w[:, 0])
    rows = [torch.stack((z, -w[:, 2], w[:, 1]), 1),
            torch.stack((w[:, 2], z, -w[:, 0]), 1),
            torch.stack((-w[:, 1], w[:, 0], z), 1)]
    return torch.stack(rows, 1)


def se3_exp3(x: torch.Tensor):
    """twist (.,6) = [w, v] -> R (n,3,3), T (n,3)   (LieAlgebra/se3.py:83-106)."""
    x = x.view(-1, 6)
    w, v = x[:, 0:3], x[:, 3:6]
    t = w.norm(p=2, dim=1).view(-1, 1, 1)
    W = hat(w)
    S = W.bmm(W)
    eye = torch.eye(3).to(w)
    R = eye + sinc1(t) * W + sinc2(t) * S
    V = eye + sinc2(t) * W + sinc3(t) * S
    T = V.bmm(v.contiguous().view(-1, 3, 1))
    return R, T.reshape(-1, 3)


def rigid_apply(twist: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """Row-vector convention of Reconstruction_point.forward: p' = p @ R + T (loss.py:460-461)."""
    R, T = se3_exp3(twist)
    return points @ R + T.reshape(1, 1, 3)


# --------------------------------------------------------------------------------------
# a6 + a7: line sampler with AABB rejection
# --------------------------------------------------------------------------------------
_BOX_FACES = [[2, 0, 6], [0, 4, 6], [5, 4, 0], [5, 0, 1], [6, 4, 5], [5, 7, 6],
              [3, 0, 2], [1, 0, 3], [3, 2, 6], [6, 7, 3], [5, 1, 3], [3, 7, 5]]   # loss.py:357-358
_PI32 = torch.acos(torch.zeros(1)).item() * 2                                       # loss.py:9


def box_triangles(verts: torch.Tensor) -> torch.Tensor:
    """(n,3) cloud -> (12,9) triangles of its AABB (loss.py:325-362)."""
    lo = torch.min(verts, 0)[0]
    hi = torch.max(verts, 0)[0]
    sel = [(1, 1, 1), (1, 1, 0), (1, 0, 1), (1, 0, 0), (0, 1, 1), (0, 1, 0), (0, 0, 1), (0, 0, 0)]
    corners = torch.stack([torch.stack([hi[a] if s[a] else lo[a] for a in range(3)]) for s in sel])
    f = torch.tensor(_BOX_FACES)
    return torch.cat([corners[f[:, 0]], corners[f[:, 1]], corners[f[:, 2]]], -1)


def lines_from_uniforms(r: float, center: torch.Tensor, a1, u1, a2, u2) -> torch.Tensor:
    """Chords of a sphere from 4 uniform draws in [0,1) (loss.py:394-411)."""
    def on_sphere(a, u):
        alpha = (a * 2 * _PI32).unsqueeze(-1)
        z = (u * 2 - 1.0).unsqueeze(-1)
        return torch.cat([r * torch.sqrt(1 - z * z) * torch.cos(alpha),
                          r * torch.sin(alpha) * torch.sqrt(1 - z * z), r * z], -1)
    q1 = on_sphere(a1, u1)
    q2 = on_sphere(a2, u2)
    direction = torch.nn.functional.normalize(q2 - q1, p=2, dim=-1)
    return torch.cat([direction, q1 + center.view(1, 3)], -1)


def triangle_hits(tris: torch.Tensor, lines: torch.Tensor) -> torch.Tensor:
    """Number of the 12 box triangles each line 'hits' by the area test (loss.py:265-316)."""
    a, b, c = tris[:, 0:3], tris[:, 3:6], tris[:, 6:9]
    nrm = torch.cross(b - a, c - a, dim=-1)
    S = torch.norm(nrm, p=2, dim=-1)
    nrm = torch.nn.functional.normalize(nrm, p=2, dim=-1)
    u = lines[:, None, 0:3]
    x0 = lines[:, None, 3:6]
    t = torch.sum(nrm[None] * (a[None] - x0), -1) / (torch.sum(nrm[None] * u, -1) + 1e-12)
    X = t.unsqueeze(-1) * u + x0
    ca, cb, cc = X - a[None], X - b[None], X - c[None]
    A = torch.norm(torch.cross(cb, cc, dim=-1), p=2, dim=-1)
    B = torch.norm(torch.cross(cc, ca, dim=-1), p=2, dim=-1)
    C = torch.norm(torch.cross(ca, cb, dim=-1), p=2, dim=-1)
    ok = (A > 0) * (B > 0) * (C > 0) * (A + B + C <= S[None])
    return torch.sum(ok, -1)


def sample_lines(r: float, center: torch.Tensor, n: int, verts1: torch.Tensor, verts2: torch.Tensor,
                 uniforms=None, generator: Optional[torch.Generator] = None):
    """Restatement of ..._efficient_resample for one pair (loss.py:415-432).

    `uniforms`, if given, is a (10, 4, n) tensor used instead of torch.rand (draw order
    alpha1, u1, alpha2, u2 per round).  Returns (lines (n,6), filled) -- rows >= filled
    stay all-zero exactly like the reference.
    """
    t1, t2 = box_triangles(verts1), box_triangles(verts2)
    out = torch.zeros(n, 6)
    filled = 0
    for rnd in range(MAX_ROUNDS):
        if uniforms is None:
            draws = [torch.rand(1, n, generator=generator)[0] for _ in range(4)]
        else:
            draws = [uniforms[rnd, q] for q in range(4)]
        cand = lines_from_uniforms(r, center, *draws)
        ok = (triangle_hits(t1, cand) * triangle_hits(t2, cand)).nonzero().reshape(-1)
        take = min(int(ok.numel()), n - filled)
        if take > 0:
            out[filled:filled + take] = cand[ok[:take]]
            filled += take
    return out, filled


# --------------------------------------------------------------------------------------
# monitoring metric ("next" row f1)
# --------------------------------------------------------------------------------------
def chamfer(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """(B,M,3),(B,N,3) -> scalar mean of both directed min-sq-distances (loss.py:236-252)."""
    sq = torch.sum((x.unsqueeze(2) - y.unsqueeze(1)) ** 2, -1)
    return torch.mean(torch.cat([sq.min(2)[0].reshape(-1), sq.min(1)[0].reshape(-1)], 0))
