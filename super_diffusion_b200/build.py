"""In-tree build of the sm_100a CUDA extension (explicit nvcc, no JIT cache).

``python -m super_diffusion_b200.build`` compiles every ``csrc/*.cu`` for
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` into
``super_diffusion_b200/libsuperdiff_b200.so`` (git-ignored, travels to the GPU
box with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(HERE, "libsuperdiff_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-diag-suppress", "128",
          "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha256()
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + \
            [os.path.join(HERE, "..", "include", "superdiff_b200.h")]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH_FLAGS + COMMON).encode())
    return h.hexdigest()


def _compile_one(src):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, False
    cmd = [NVCC] + ARCH_FLAGS + COMMON + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return obj, True


def build(verbose=True, force=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = _sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(_compile_one, srcs))
    objs = [o for o, _ in results]
    rebuilt = any(c for _, c in results) or not os.path.exists(LIB_PATH)
    if rebuilt:
        cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", LIB_PATH] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[super_diffusion_b200.build] {'built' if rebuilt else 'up to date'}: {LIB_PATH}", file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
