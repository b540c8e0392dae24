"""memotr_b200/build.py -- compile the sm_100a kernels into memotr_b200/csrc/libmemotr_b200.so (in-tree).

nvcc cross-compiles without a GPU; the shared library depends only on libcudart (no torch, no libcuda link:
the driver entry points the TMA descriptors need are resolved at run time through cudaGetDriverEntryPoint).
"""
import concurrent.futures
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libmemotr_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(CSRC)), "include", "memotr_b200.h"))
    objs, jobs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run(["nvcc", *NVCC_FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(f"--- nvcc {os.path.basename(s)}\n{r.stdout}{r.stderr}\n")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
        open(o + ".log", "w").write(r.stdout + r.stderr)
        return o

    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or _stale(LIB, objs):
        subprocess.check_call(["nvcc", "-shared", "-o", LIB, *objs, "-cudart", "shared",
                               "-Xlinker", "-rpath,/usr/local/cuda/lib64"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
