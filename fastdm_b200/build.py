"""Build libfastdm_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fastdm_b200.build [--force]

Each csrc/*.cu is compiled to an object (in parallel) and linked into
fastdm_b200/libfastdm_b200.so with the static CUDA runtime, so the only run-time dependency is
the CUDA driver. Objects are cached under build/ keyed by source + header mtimes.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfastdm_b200.so")
OBJDIR = os.path.join(ROOT, "build", "obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include")]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _compile(src, obj):
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_time = _newest(hdrs)
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(OBJDIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_time):
            jobs.append((s, o))
    if jobs:
        if verbose:
            print(f"[fastdm_b200.build] compiling {len(jobs)} file(s) for sm_100a", flush=True)
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda a: _compile(*a), jobs))
    if jobs or not os.path.exists(OUT) or os.path.getmtime(OUT) < _newest(objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[fastdm_b200.build] linked {OUT}", flush=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
