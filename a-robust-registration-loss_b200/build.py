"""Builds librrl_b200.so (hand-written sm_100a CUDA behind include/rrl_b200.h) in-tree with nvcc.

    python a-robust-registration-loss_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# A/B builds: RRL_VARIANT=name RRL_DEFS="-DFOO=1 ..." writes build/variants/librrl_b200_<name>.so next to the product
# library; tools load it with RRL_LIB_PATH.  The product build never sets either.
VARIANT = os.environ.get("RRL_VARIANT", "")
OUT = os.path.join(HERE, "build", "variants", "librrl_b200_%s.so" % VARIANT) if VARIANT else os.path.join(HERE, "librrl_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
COMMON += os.environ.get("RRL_DEFS", "").split()
if os.environ.get("RRL_MARKS"):          # measurement builds: in-kernel phase timestamps (rrl_debug_read_marks)
    COMMON.append("-DRRL_MARKS")
# rrl_sampler.cu restates a floating-point knife-edge test and must not contract a*b+c into FMAs
SOURCES = {"rrl_api.cu": [], "rrl_dense.cu": [], "rrl_sparse.cu": [], "rrl_se3.cu": [], "rrl_aux.cu": [], "rrl_neigh.cu": [], "rrl_comm.cu": [],
           "rrl_sampler.cu": ["-fmad=false"]}


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return "nvcc"


def build(force=False, verbose=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rrl_b200.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    objdir = os.path.join(HERE, "build", "variants", VARIANT) if VARIANT else os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src, extra in SOURCES.items():
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        if p.returncode:
            raise RuntimeError("nvcc failed on " + src)
    subprocess.run([_nvcc()] + ARCH + ["-shared", "-o", OUT] + objs + ["-lcudart", "-ldl"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
