"""which lines of the degenerate-lines test differ from the oracle (debugging aid)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import rrl_b200
from oracle import c_oracle as co
from tools import synth
p = synth.make_pair(151, 18000, 2500, nf2=16500, zero_frac=0.15)
lines = p["lines"].copy()
rng = np.random.default_rng(6)
lines[:400, :3] *= rng.uniform(1.0, 1.01, size=(400, 1)).astype(np.float32)
lines[400:700, :3] *= rng.uniform(0.3, 1.0, size=(300, 1)).astype(np.float32)
lines[700:730, :3] *= 3.0
lines[730:780] = lines[1]
t1 = torch.from_numpy(p["tri1"]).cuda()[None]; t2 = torch.from_numpy(p["tri2"]).cuda()[None]; ln = torch.from_numpy(lines).cuda()[None]
orc = co.loss(p["tri1"], p["tri2"], lines)
for sup in (0, 1):
    rrl_b200._native.lib().rrl_debug_set_param(7, 1 - sup)
    loss, info = rrl_b200.intersected_line_loss(t1, t2, ln, return_info=True)
    for c, oc in ((1, orc.counts1), (2, orc.counts2)):
        cnt = info.hits(c)[0][0].cpu().numpy()
        bad = np.nonzero(cnt != oc)[0]
        print("supers", sup, "cloud", c, "differing lines", len(bad), bad[:20].tolist())
        for l in bad[:8]:
            u = lines[l, :3]; print("   line", l, "|u|^2", float((u.astype(np.float64) ** 2).sum()), "x0", lines[l, 3:].tolist(), "got", int(cnt[l]), "want", int(oc[l]))
