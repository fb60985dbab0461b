"""Build libuwr_b200.so (every .cu under csrc/) for sm_100a with nvcc, in-tree.

    python underwater-image-restoration_b200/build.py [--force]

The shared library has a plain C ABI (include/uwr_b200.h); there is no torch/pybind dependency,
so a file compiles in seconds and the .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libuwr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--use_fast_math=false" if False else "-Xptxas=-v"]
FLAGS = [f for f in FLAGS if f]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "uwr_b200.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, hdr_mtime):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(srcp)
            and os.path.getmtime(obj) >= hdr_mtime):
        return src, False, ""
    cmd = [NVCC] + FLAGS + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return src, True, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    hdr_mtime = _deps_mtime()
    rebuilt = False
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, did, log in ex.map(lambda s: _compile(s, force, hdr_mtime), srcs):
            rebuilt |= did
            if verbose and did:
                print(f"[uwr build] compiled {src}")
                spills = [l for l in log.splitlines() if "spill" in l and "0 bytes spill stores" not in l]
                for l in spills:
                    print("   ", l.strip())
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if rebuilt or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[uwr build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
