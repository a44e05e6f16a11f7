"""Spread of the demo-flow convergence check over repeated runs (same seeds): python tools/demo_flaky.py [runs] [epochs]"""
import argparse, importlib.util, os, sys
import numpy as np, torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import rrl_b200 as rrl
spec = importlib.util.spec_from_file_location("_demo", os.path.join(root, "examples", "demo_lie_algebra.py"))
demo = importlib.util.module_from_spec(spec); spec.loader.exec_module(demo)
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 60
nlines = int(sys.argv[3]) if len(sys.argv) > 3 else 8000
angle = float(sys.argv[4]) if len(sys.argv) > 4 else 12.0
for r in range(runs):
    torch.manual_seed(5); np.random.seed(5); rrl.loss.manual_seed(5)
    args = argparse.Namespace(synthetic=1500, seed=5, angle=angle, data_path="", label1="0")
    data = demo.load_case(args, "cuda")
    model, hist = demo.test_one_case(data, n_epoch=epochs, n_sample_line=nlines, device="cuda", log=None)
    cf = [h[0] for h in hist]
    R, T = model.Transform()
    Rn = R[0].cpu().numpy().astype(np.float64)
    ang = np.deg2rad(angle)
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    # row-vector convention: p' = p @ R, the target is base @ Rz.T  =>  R should approach Rz.T
    err = np.rad2deg(np.arccos(np.clip((np.trace(Rn @ Rz) - 1) / 2, -1, 1)))
    late = np.array(cf[20:])
    print("rot err deg %.3f  median last15/first %.3f  frac of epochs>=20 below 0.25*first: %.2f  max late/first %.2f" %
          (err, np.median(cf[-15:]) / np.mean(cf[:3]), float((late < 0.25 * np.mean(cf[:3])).mean()), late.max() / np.mean(cf[:3])))
    print(r, len(hist), "chamfer first3 %.5f last3 %.5f min %.5f | at 20/40/60/80/100: %s" % (
        np.mean(cf[:3]), np.mean(cf[-3:]), min(cf), " ".join("%.5f" % cf[i] for i in range(19, len(cf), 20))), flush=True)
