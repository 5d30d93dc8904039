"""In-tree build of libradnet_b200.so (sm_100a only) with plain nvcc.

`python -m rock_art_radnet_b200.build` or `__graft_entry__.build()`.  The shared
library has a C ABI (include/radnet_b200.h), links the static CUDA runtime and
does not depend on torch; it lands next to the package so the snapshot taken by
`gpurun` carries it to the GPU box.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OUT_DIR = os.path.join(PKG_DIR, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libradnet_b200.so")
SOURCES = ["capi.cu", "decode.cu", "sort_nms.cu", "roipool.cu", "rpn_targets.cu", "targets.cu", "sampling.cu", "detect.cu", "synth.cu", "losses.cu", "evalmap.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # parity: every NumPy/TF operation rounds once; never contract a*b+c into an FMA
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libradnet_b200.so")


def _fingerprint():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    names = sorted(os.listdir(CSRC)) + ["../../include/radnet_b200.h"]
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            with open(path, "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Returns its path."""
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "build.stamp")
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append("== %s\n%s" % (src, out))
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
        objs.append(obj)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("link failed")
    with open(os.path.join(OUT_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    with open(stamp, "w") as f:
        f.write(fp)
    if verbose:
        print("\n".join(log))
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
