"""e2e (host buffers -> host results) time per call of rrl_host_loss_fwd_bwd for several sub-batch counts."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rrl_b200
from tools import synth
L = rrl_b200._native.lib()
CONFIGS = {"demo": (1, 1024, 20000), "dcp": (32, 1024, 15000), "rpm": (64, 2048, 10000), "fmr": (128, 1024, 15000)}
name = sys.argv[1] if len(sys.argv) > 1 else "dcp"
B, nf, nl = CONFIGS[name]
pairs = [synth.make_pair(1000 + i, nf, nl) for i in range(min(B, 4))]
idx = [i % len(pairs) for i in range(B)]
host = [torch.from_numpy(np.stack([pairs[i][k] for i in idx])).pin_memory() for k in ("tri1", "tri2", "lines")]
loss = np.zeros(B, np.float32); status = np.zeros(B, np.int32)
for S in [int(x) for x in (sys.argv[2:] or ["1", "2", "4", "8"])]:
    os.environ["RRL_HOST_SUBBATCHES"] = str(S)
    ctx = C.c_void_p()
    assert L.rrl_host_create(B, nf, nf, nl, 0, C.byref(ctx)) == 0
    def call():
        rc = L.rrl_host_loss_fwd_bwd(ctx, host[0].data_ptr(), host[1].data_ptr(), host[2].data_ptr(), 1, 1, 5, 5,
                                     loss.ctypes.data, status.ctypes.data, None)
        assert rc == 0, rc
    for _ in range(5): call()
    n = 100
    t0 = time.perf_counter()
    for _ in range(n): call()
    dt = (time.perf_counter() - t0) / n
    print("%s S=%d (actual %d): %.1f us per call, %.1f M pair*lines/s, loss sum %.6f" %
          (name, S, L.rrl_host_subbatches(ctx), dt * 1e6, B * nl / dt / 1e6, float(loss.sum())), flush=True)
    tk = [C.c_int(-1), C.c_int(-1)]
    sub = lambda i: L.rrl_host_submit(ctx, host[0].data_ptr(), host[1].data_ptr(), host[2].data_ptr(), 1, 1, 5, 5, 0, C.byref(tk[i % 2]))
    wt = lambda i: L.rrl_host_wait(ctx, tk[i % 2].value, loss.ctypes.data, status.ctypes.data, None)
    def run(n):
        assert sub(0) == 0
        for i in range(n):
            if i + 1 < n: assert sub(i + 1) == 0
            assert wt(i) == 0
    run(5)
    t0 = time.perf_counter(); run(n); dt = (time.perf_counter() - t0) / n
    print("%s S=%d pipelined submit/wait: %.1f us per step, %.1f M pair*lines/s" % (name, S, dt * 1e6, B * nl / dt / 1e6), flush=True)
    L.rrl_host_destroy(ctx)
