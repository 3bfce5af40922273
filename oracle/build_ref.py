"""Build recipe for oracle/_ref/_pvcnn_backend.so  (TEST INFRASTRUCTURE, not product).

Compiles the reference's own operator extension (the 13 sources listed at
/root/reference/grasp_ldm/models/modules/ext/pvcnn/modules/functional/backend.py:12-24)
for sm_100a, from the sources where they lie under /root/reference, into
oracle/_ref/.  Nothing from /root/reference is copied into the repo; only the
built .so lands in oracle/_ref/ (git-ignored, but shipped to the GPU box).

The reference loads this module with torch's JIT `load()` at import time; we do
not run that loader - this recipe calls nvcc / g++ directly.  The module name
must stay `_pvcnn_backend` because the reference's bindings.cpp:10 hard-codes it.

Used only by tests/ (GPU parity of our kernels against the real reference
kernels) and by tests/golden/make_golden_gpu.py.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

REF_SRC = "/root/reference/grasp_ldm/models/modules/ext/pvcnn/modules/functional/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SOURCES = [
    "ball_query/ball_query.cpp", "ball_query/ball_query.cu",
    "grouping/grouping.cpp", "grouping/grouping.cu",
    "interpolate/neighbor_interpolate.cpp", "interpolate/neighbor_interpolate.cu",
    "interpolate/trilinear_devox.cpp", "interpolate/trilinear_devox.cu",
    "sampling/sampling.cpp", "sampling/sampling.cu",
    "voxelization/vox.cpp", "voxelization/vox.cu",
    "bindings.cpp",
]


def ref_so_path():
    return os.path.join(OUT, "_pvcnn_backend.so")


def build(force=False, verbose=False):
    """Returns the path of the built module, or None when /root/reference is absent."""
    so = ref_so_path()
    if os.path.exists(so) and not force:
        return so
    if not os.path.isdir(REF_SRC):
        return None
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    objdir = os.path.join(OUT, "obj")
    os.makedirs(objdir, exist_ok=True)
    inc = []
    for p in ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"]]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = ["-DTORCH_EXTENSION_NAME=_pvcnn_backend", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(REF_SRC, s)
        obj = os.path.join(objdir, s.replace("/", "_") + ".o")
        objs.append(obj)
        if s.endswith(".cu"):
            # reference flags: -O3 -std=c++17 on the host side, nvcc defaults (fmad on) for device
            cmd = ["nvcc", "-c", src, "-o", obj, "-O3", "-std=c++17", "--expt-relaxed-constexpr",
                   "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"] + common + inc
        else:
            cmd = ["g++", "-c", src, "-o", obj, "-O3", "-std=c++17", "-fPIC"] + common + inc
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("reference backend compile failed: " + " ".join(cmd))
        if verbose:
            sys.stdout.write(out.decode())
    libdirs = ce.library_paths(device_type="cuda") if "device_type" in ce.library_paths.__code__.co_varnames else ce.library_paths(cuda=True)
    link = ["g++", "-shared", "-o", so] + objs
    for d in libdirs:
        link += ["-L", d, f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    subprocess.check_call(link)
    shutil.rmtree(objdir, ignore_errors=True)
    return so


def load():
    """Import the built reference module (needs a CUDA device to run anything)."""
    so = ref_so_path()
    if not os.path.exists(so):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("_pvcnn_backend", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
