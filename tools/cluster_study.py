"""CPU study of node compactness (numpy only): bounding radius of G consecutive points in Hilbert order on a synthetic
surface cloud, with and without k-d refinement inside windows of W sorted points -- the numbers quoted in DESIGN.md 4.1.

    python tools/cluster_study.py [n_points]
"""
import numpy as np, sys
sys.path.insert(0, '.')
from tools import synth
def hilbert_key(P, bits=10):
    # Skilling transpose, vectorised. P: (n,3) ints in [0,2^bits)
    X = P.astype(np.uint32).T.copy()
    M = 1 << (bits-1)
    Q = M
    while Q > 1:
        Pm = Q-1
        for i in range(3):
            sel = (X[i] & Q) != 0
            X[0] = np.where(sel, X[0]^Pm, X[0])
            t = (X[0]^X[i]) & Pm
            X[0] = np.where(sel, X[0], X[0]^t)
            X[i] = np.where(sel, X[i], X[i]^t)
        Q >>= 1
    X[1]^=X[0]; X[2]^=X[1]
    t = np.zeros_like(X[0]); Q=M
    while Q>1:
        t = np.where((X[2]&Q)!=0, t^(Q-1), t); Q>>=1
    X ^= t
    key = np.zeros(X.shape[1], np.uint64)
    for b in range(bits-1,-1,-1):
        for i in range(3):
            key = (key<<np.uint64(1)) | ((X[i]>>b)&1).astype(np.uint64)
    return key
n=int(sys.argv[1]) if len(sys.argv)>1 else 200000
rng=np.random.default_rng(0)
pts=synth.surface_points(rng,n,"sphere").astype(np.float64)
Pm=np.abs(pts).max()*1.0001
g=np.clip(((pts+Pm)*(512/Pm)).astype(int),0,1023)
key=hilbert_key(g)
order=np.argsort(key,kind='stable')
area=4*np.pi*np.mean(np.linalg.norm(pts,axis=1))**2
delta=np.sqrt(area/n)
for G in (8,16,256):
    s=pts[order][:n//G*G].reshape(-1,G,3)
    c=s.mean(1,keepdims=True)
    r=np.linalg.norm(s-c,axis=2).max(1)/delta
    ideal=np.sqrt(G/np.pi)
    print("G=%d radius/delta: median %.2f mean %.2f p90 %.2f p99 %.2f max %.2f ; ideal disk %.2f; mean r^2 %.2f vs ideal %.2f"%(G,np.median(r),r.mean(),np.percentile(r,90),np.percentile(r,99),r.max(),ideal,(r**2).mean(),ideal**2))

def kd_refine(P, W, G):
    # P: (n,3) sorted points; windows of W consecutive points are re-partitioned into leaves of G by median splits along the longest axis
    n = P.shape[0]//W*W
    Q = P[:n].reshape(-1, W, 3).copy()
    size = W
    while size > G:
        S = Q.reshape(-1, size, 3)
        ext = S.max(1) - S.min(1)
        ax = ext.argmax(1)
        coord = np.take_along_axis(S, ax[:,None,None].repeat(size,1), 2)[:,:,0]
        o = np.argsort(coord, axis=1, kind='stable')
        S = np.take_along_axis(S, o[:,:,None].repeat(3,2), 1)
        Q = S
        size //= 2
        Q = Q.reshape(-1, size, 3)
    return Q.reshape(-1, 3)
thr=0.87
def stat(S, G, name):
    s=S[:S.shape[0]//G*G].reshape(-1,G,3); c=s.mean(1,keepdims=True)
    r=np.linalg.norm(s-c,axis=2).max(1)/delta
    print("%-28s G=%3d mean r %.2f  mean (r+thr)^2 %.2f  (ideal %.2f)"%(name,G,r.mean(),((r+thr)**2).mean(),(np.sqrt(G/np.pi)+thr)**2))
S0=pts[order]
for G in (8,16):
    stat(S0,G,"hilbert")
    for W in (32,64,256,1024):
        if W>G: stat(kd_refine(S0,W,G),G,"kd window %d"%W)
