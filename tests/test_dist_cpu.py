"""Host-side logic of the multi-GPU paths, exercised on CPU with the gloo backend (world_size 2).

The per-rank stages are stood in for by an oracle-backed backend (numpy + the C oracle), so what is tested here is
exactly the code the GPU path shares: rrl_b200.dist.shard_range / gather_entries / line_shard_forward /
batch_sharded_loss's reduction -- i.e. the exchange protocol of SURVEY 8(e)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = 2.0 ** 40


class OracleShardBackend:
    """CPU stand-in for NativeShardBackend: same stage semantics, computed with the C oracle."""

    def __init__(self, tri1, tri2, lines_local):
        from oracle import c_oracle as co
        self.r = co.loss(tri1, tri2, lines_local, want_grad=False, want_D=True)
        k, j = self.r.counts1, self.r.counts2
        self.sel = np.where((k >= 1) & (k <= 4) & (j >= 1) & (j <= 4))[0]

    def stage1_counts(self):
        c = np.zeros(18, np.int64)
        c[:16] = self.r.n_kj.reshape(-1)
        c[16] = len(self.sel)
        c[17] = sum(int(self.r.counts1[l]) * int(self.r.counts2[l]) for l in self.sel)
        return torch.from_numpy(c)

    def _entries(self):
        out = []
        for l in self.sel:
            out.append(self.r.D[l][:self.r.counts1[l], :self.r.counts2[l]].reshape(-1))
        return np.concatenate(out) if out else np.zeros(0, np.float32)

    def pack_entries(self, n_local):
        e = self._entries()
        assert e.shape[0] == n_local
        return torch.from_numpy(np.concatenate([e, np.zeros(max(1 - n_local, 0), np.float32)]).astype(np.float32))

    def median(self, entries):
        v = np.sort(entries.numpy())
        return torch.tensor([v[(len(v) - 1) // 2] if len(v) else 0.0], dtype=torch.float32)

    # distributed lower median: the semantics of rrl_shard_select_hist / rrl_shard_select_pick in numpy
    def select_hist(self, rnd, state):
        keys = self._entries().astype(np.float32).view(np.uint32)
        if rnd == 1:
            keys = keys[(keys >> 16) == (np.uint32(int(state[0])) >> 16)] & np.uint32(0xFFFF)
        else:
            keys = keys >> 16
        return torch.from_numpy(np.bincount(keys.astype(np.int64), minlength=65536).astype(np.int32))

    def select_pick(self, rnd, ghist, gcounts, state, med):
        n = int(gcounts[17])
        if n <= 0:
            state[:] = 0
            med[0] = 0.0
            return
        rank = (n - 1) // 2 if rnd == 0 else int(state[1])
        cum = np.cumsum(ghist.numpy().astype(np.int64))
        b = int(np.searchsorted(cum, rank, side="right"))
        state[1] = rank - (int(cum[b - 1]) if b else 0)
        state[0] = (b << 16) if rnd == 0 else ((int(state[0]) & 0xFFFF0000) | b)
        if rnd == 1:
            med[0] = float(np.array([int(state[0])], np.uint32).view(np.float32)[0])

    def stage2_sums(self, gcounts, med):
        m = np.float32(med.item())
        sums = np.zeros(32, np.int64)
        for l in self.sel:
            k, j = int(self.r.counts1[l]), int(self.r.counts2[l])
            D = self.r.D[l][:k, :j]
            W = (np.float32(1) - np.exp(-(D / m).astype(np.float64) / 2).astype(np.float32)).astype(np.float32)
            c = (k - 1) * 4 + (j - 1)
            sums[c] += int(round(float(W.min(1).astype(np.float64).sum()) * FIX))
            sums[16 + c] += int(round(float(W.min(0).astype(np.float64).sum()) * FIX))
        self.gcounts = gcounts.numpy().copy()
        return torch.from_numpy(sums)

    def stage3_loss(self, gsums):
        s = gsums.numpy().astype(np.float64) / FIX
        loss, C = 0.0, 0
        for k in range(1, 5):
            for j in range(1, 5):
                c = (k - 1) * 4 + (j - 1)
                n = float(self.gcounts[c])
                if n > 0:
                    C += 1
                    loss += np.exp(-0.5 * abs(k - j)) * (s[c] / (n * k) + s[16 + c] / (n * j))
        return torch.tensor([loss / C if C else 0.0], dtype=torch.float32), torch.tensor([0 if C else 1], dtype=torch.int32)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rrl_b200
    from oracle import c_oracle as co
    from oracle import synth
    D = rrl_b200.dist
    # ---- line shard: every rank holds both clouds, a contiguous block of the lines ----
    p = synth.make_pair(77, 400, 1501)
    lo, hi = D.shard_range(1501, rank, world)
    loss, status, med = D.line_shard_forward(OracleShardBackend(p["tri1"], p["tri2"], p["lines"][lo:hi]))
    full = co.loss(p["tri1"], p["tri2"], p["lines"], want_grad=False)
    assert float(med) == full.median, (float(med), full.median)
    assert abs(float(loss) - full.loss) <= 1e-6 * full.loss, (float(loss), full.loss)
    assert int(status) == 0
    # a shard with no selected line at all still takes part in every collective
    far = p["lines"].copy()
    far[:, 3:] += 100.0
    mixed = p["lines"][lo:hi] if rank == 0 else far[:50]
    loss2, _, med2 = D.line_shard_forward(OracleShardBackend(p["tri1"], p["tri2"], mixed))
    ref2 = co.loss(p["tri1"], p["tri2"], p["lines"][:D.shard_range(1501, 0, world)[1]], want_grad=False)
    assert float(med2) == ref2.median and abs(float(loss2) - ref2.loss) <= 1e-6 * ref2.loss
    # ---- variable-length gather ----
    mine = torch.arange(3 + 2 * rank, dtype=torch.float32) + 100 * rank
    got = D.gather_entries(mine, mine.numel(), [3 + 2 * r for r in range(world)])
    want = torch.cat([torch.arange(3 + 2 * r, dtype=torch.float32) + 100 * r for r in range(world)])
    assert torch.equal(got, want)
    assert D.gather_entries(torch.zeros(1), 0, [0] * world).numel() == 0
    # ---- batch shard: value = global sum, gradient = local pairs only ----
    local = torch.tensor([1.0 + rank, 2.0 + rank], requires_grad=True)
    total = D.global_sum_with_local_grad(local.sum())
    assert float(total) == sum(3.0 + 2 * r for r in range(world))
    total.backward()
    assert torch.equal(local.grad, torch.ones(2))
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


def test_shard_range_partitions_exactly():
    sys.path.insert(0, ROOT)
    import rrl_b200
    for n in (0, 1, 7, 100000, 15000):
        for world in (1, 2, 3, 8):
            spans = [rrl_b200.dist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_line_and_batch_shard_protocol_world2(tmp_path):
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
