"""Builds libcgb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).  One translation unit per kernel
family, compiled in parallel, then linked."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, os.environ.get("CGB_LIB_NAME", "libcgb200.so"))      # developer knobs for A/B builds
SOURCES = ["cgb_api.cu", "cgb_tu_fast.cu", "cgb_tu_sfcc_fast.cu", "cgb_tu_newton_tab.cu", "cgb_tu_euler.cu", "cgb_tu_euler_diag.cu", "cgb_tu_lite.cu", "cgb_tu_lite_warp.cu",
           "cgb_tu_salt.cu", "cgb_tu_presolver.cu"]
EXTRA_FLAGS = {"cgb_tu_presolver.cu": ["-fmad=false"]}      # knots must agree with the (uncontracted) host builders to rounding
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cgb200.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    headers = [d for d in deps if not d.endswith(".cu")]

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= _newest(headers + [path]):
            return obj, ""
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("CGB_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + EXTRA_FLAGS.get(src, []) + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    cmd = [nvcc, "-shared", "-o", LIB] + [o for o, _ in results]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
