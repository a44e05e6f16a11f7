"""TEST INFRASTRUCTURE ONLY -- numpy restatements of the small differentiable pieces around the loss.  Pinned by
tests/golden/{expmap,chamfer_grad,se3_log}.npz (minted from the unmodified reference by oracle/make_golden_r2.py).

  se3_exp4 / expmap_backward   exps_deep_learning/fmr/se_math/se3.py:60-84 (exp), :133-165 (ExpMap.backward),
                               :27-55 (mat / genmat), sinc.py (Taylor branch below |t| = 0.01)
  chamfer_with_grad            loss.py:38-52, 236-252 and its autograd
"""
import numpy as np


def _sincs(t):
    t2 = t * t
    if abs(t) < 0.01:                         # sinc.py:7-11, 95-99, 124-128
        a = 1 - t2 / 6 * (1 - t2 / 20 * (1 - t2 / 42))
        b = 0.5 * (1 - t2 / 12 * (1 - t2 / 30 * (1 - t2 / 56)))
        c = 1.0 / 6 * (1 - t2 / 20 * (1 - t2 / 42 * (1 - t2 / 72)))
    else:
        a, b, c = np.sin(t) / t, (1 - np.cos(t)) / t2, (t - np.sin(t)) / (t2 * t)
    return a, b, c


def _hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], np.float64)


def se3_exp4(x):
    """x (B,6) -> g (B,4,4) float64"""
    x = np.asarray(x, np.float64).reshape(-1, 6)
    g = np.zeros((x.shape[0], 4, 4))
    for b, row in enumerate(x):
        w, v = row[:3], row[3:]
        a, bb, c = _sincs(np.linalg.norm(w))
        W = _hat(w)
        S = W @ W
        g[b, :3, :3] = np.eye(3) + a * W + bb * S
        g[b, :3, 3] = (np.eye(3) + bb * W + c * S) @ v
        g[b, 3, 3] = 1
    return g


def expmap_backward(x, grad_g):
    """ExpMap.backward: grad_x[k] = sum_ij grad_g[i,j] (gen_k g)[i,j]"""
    g = se3_exp4(x)
    gen = np.zeros((6, 4, 4))
    for k in range(3):
        e = np.zeros(3); e[k] = 1
        gen[k, :3, :3] = _hat(e)
        gen[3 + k, k, 3] = 1
    dg = np.einsum("kim,bmj->bkij", gen, g)
    return np.einsum("bij,bkij->bk", np.asarray(grad_g, np.float64), dg)


def chamfer_with_grad(x, y, upstream=1.0):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    B, M, _ = x.shape
    N = y.shape[1]
    d = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1)
    a1, a2 = d.argmin(2), d.argmin(1)
    val = (d.min(2).sum() + d.min(1).sum()) / (B * (M + N))
    gx, gy = np.zeros_like(x), np.zeros_like(y)
    s = 2.0 * upstream / (B * (M + N))
    for b in range(B):
        diff = x[b] - y[b, a1[b]]
        gx[b] += s * diff
        np.add.at(gy[b], a1[b], -s * diff)
        diff2 = y[b] - x[b, a2[b]]
        gy[b] += s * diff2
        np.add.at(gx[b], a2[b], -s * diff2)
    return val, gx, gy
