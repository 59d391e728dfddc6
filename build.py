"""Build libpolymath_b200.so (CUDA, sm_100a) and the oracle's C++ restatement, in-tree.

Usage: python build.py [--force] [--jobs N].  nvcc cross-compiles without a GPU.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "polymath_b200", "csrc")
OBJDIR = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(ROOT, "polymath_b200", "libpolymath_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"),
]


def _sources():
    out = []
    for dirpath, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cu", ".cpp")):
                out.append(os.path.join(dirpath, f))
    return sorted(out)


def _headers_digest():
    h = hashlib.sha256()
    for dirpath, _, files in os.walk(CSRC):
        for f in sorted(files):
            if f.endswith((".cuh", ".hpp", ".h")):
                with open(os.path.join(dirpath, f), "rb") as fh:
                    h.update(fh.read())
    with open(os.path.join(ROOT, "include", "polymath_b200.h"), "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src, hdr_digest, force):
    rel = os.path.relpath(src, CSRC).replace(os.sep, "_")
    obj = os.path.join(OBJDIR, rel + ".o")
    stamp = obj + ".stamp"
    with open(src, "rb") as fh:
        digest = hashlib.sha256(fh.read() + hdr_digest.encode()).hexdigest()
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, False
    cmd = [NVCC] + NVCC_FLAGS + ["-x", "cu", "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, res.stdout, res.stderr))
    with open(stamp, "w") as fh:
        fh.write(digest)
    return obj, True


def build_cuda(force=False, jobs=None):
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    hd = _headers_digest()
    jobs = jobs or min(len(srcs), os.cpu_count() or 4)
    objs, rebuilt = [], False
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        for obj, did in ex.map(lambda s: _compile_one(s, hd, force), srcs):
            objs.append(obj)
            rebuilt |= did
    if rebuilt or force or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (res.stdout, res.stderr))
    return LIB


def build_oracle(force=False):
    mk = os.path.join(ROOT, "oracle", "Makefile")
    if not os.path.exists(mk):
        return None
    res = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")] + (["-B"] if force else []),
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n%s\n%s" % (res.stdout, res.stderr))
    return os.path.join(ROOT, "oracle", "libpm_oracle.so")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    a = ap.parse_args()
    print(build_cuda(a.force, a.jobs))
    o = build_oracle(a.force)
    if o:
        print(o)


if __name__ == "__main__":
    sys.exit(main())
