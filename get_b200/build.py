"""Build libget_b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

    python -m get_b200.build [--force]

The shared library lands next to the sources (get_b200/csrc/libget_b200.so); it is git-ignored but ships to
the GPU box with the repository snapshot. No torch / pybind dependency: the ABI is plain C (include/get_b200.h).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libget_b200.so")
STAMP = os.path.join(CSRC, ".build_stamp")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("GETB_EXTRA_NVCC_FLAGS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(HERE), "include", "get_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode())      # location-independent: the built library travels with the tree
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def is_current() -> bool:
    """Does the built library match the sources (and nvcc flags) it sits next to?"""
    try:
        return os.path.exists(LIB) and open(STAMP).read().strip() == _digest()
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if stale or forced) under an exclusive file lock: concurrent ranks / test workers build once, the others
    wait and pick up the result; the library is linked under a temporary name and renamed into place."""
    import fcntl
    with open(os.path.join(CSRC, ".build_lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = nvcc_path()
    objs = []
    procs = []
    for src in _sources():
        obj = src[:-3] + ".o"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("[%s]\n%s\n" % (os.path.basename(src), out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building get_b200 kernels")
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs + ["-lcudart"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
