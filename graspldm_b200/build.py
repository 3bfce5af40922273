"""In-tree build of libgraspldm_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgraspldm_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = sources()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "graspldm_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs, jobs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        p = subprocess.run(["nvcc", "-c", s, "-o", o] + NVCC_FLAGS, capture_output=True, text=True)
        return s, p

    with ThreadPoolExecutor(max_workers=8) as ex:
        for s, p in ex.map(compile_one, jobs):
            if verbose or p.returncode != 0:
                sys.stderr.write(f"--- {os.path.basename(s)}\n{p.stdout}{p.stderr}\n")
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
    if jobs or force or not os.path.exists(LIB):
        subprocess.check_call(["nvcc", "-shared", "-o", LIB] + objs +
                              ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
