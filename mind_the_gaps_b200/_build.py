"""Build recipe for ``libmtg_b200.so`` (hand-written CUDA, sm_100a only).

The library is compiled IN-TREE (``mind_the_gaps_b200/csrc/libmtg_b200.so``) so that it travels with
a repository snapshot; ``__graft_entry__.build()`` calls :func:`build`.  One translation unit per
``.cu`` file, compiled in parallel, linked with a static cudart and the CUDA toolkit's cuFFT.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libmtg_b200.so")
OBJ = os.path.join(CSRC, "_obj")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
              "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++"]


# the simulator restates NumPy expressions whose index results must match bit for bit: no FMA contraction
EXTRA_FLAGS = {"mtg_sim.cu": ["-fmad=false"]}


def nvcc():
    for cand in (os.environ.get("MTG_NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmtg_b200 cannot be built")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mt for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(CSRC, "..", "..", "include", "mtg_b200.h"))
    if not _stale(obj, [os.path.join(CSRC, src)] + headers):
        return obj, ""
    cmd = [nvcc()] + ARCH + NVCC_FLAGS + EXTRA_FLAGS.get(src, []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, p.stdout))
    with open(obj[:-2] + ".ptxas.log", "w") as fh:
        fh.write(p.stdout)
    return obj, p.stdout if verbose else ""


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` for sm_100a and link ``libmtg_b200.so``.  Returns its path."""
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [r[0] for r in results]
    if verbose:
        for r in results:
            sys.stdout.write(r[1])
    if force or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + [
            "-cudart", "static", "-lcufft", "-Xlinker", "-rpath,/usr/local/cuda/lib64", "-ccbin", "/usr/bin/g++"]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s" % p.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
