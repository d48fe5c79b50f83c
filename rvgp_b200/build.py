"""Build librvgp_b200.so in-tree with nvcc for sm_100a.  No JIT cache: the built rvgp_b200/lib/librvgp_b200.so is git-ignored
(history stays source-only) but is part of the gpurun snapshot, so it travels to the GPU box with the working tree."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "librvgp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "rvgp_b200.h"))
    return max(os.path.getmtime(p) for p in hdrs)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hm = _deps_mtime()
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        p = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        log = p.stdout + p.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
        return src, log

    with ThreadPoolExecutor(max_workers=8) as ex:
        for src, log in ex.map(cc, jobs):
            if verbose:
                print("compiled", os.path.basename(src))
                spills = [l for l in log.splitlines() if "spill" in l and "0 bytes spill stores, 0 bytes spill loads" not in l]
                for l in spills[:6]:
                    print("   ", l.strip())
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(LIB):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
