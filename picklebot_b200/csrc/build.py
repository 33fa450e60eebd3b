"""Build libpicklebot_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python picklebot_b200/csrc/build.py [--force] [--verbose]

The shared library has a plain C ABI (include/picklebot_b200.h) and links only the CUDA runtime
(statically) -- no torch, no libcuda at link time (the TMA descriptor encoder is fetched through
cudaGetDriverEntryPoint).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "libpicklebot_b200.so")
OBJ = os.path.join(HERE, "_build")
SOURCES = ["runtime.cu", "dwconv.cu", "dwconv_tiled.cu", "dwconv_mma.cu", "pwgemm_simt.cu", "pwgemm_tc.cu", "pwwgrad_tc.cu", "norm_act.cu",
           "se_pool.cu", "se_fc.cu", "stem.cu", "stem_tc.cu", "optim.cu", "loss.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v" if "--verbose" in sys.argv else "-O3"]


def _deps():
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "picklebot_b200.h"))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(HERE, src)
    if "--force" in sys.argv or _stale(obj, [path] + _deps()):
        cmd = [NVCC] + FLAGS + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if "--verbose" in sys.argv:
            print(r.stderr)
    return obj


def build(force=False):
    if force and "--force" not in sys.argv:
        sys.argv.append("--force")
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    if "--force" in sys.argv or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


if __name__ == "__main__":
    print(build())
