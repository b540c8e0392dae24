#!/usr/bin/env python
"""Build the REAL reference CUDA operator into oracle/_ref/ (test infrastructure only).

TEST INFRASTRUCTURE -- nothing under memotr_b200/ may import or load anything from oracle/.

Recipe (no reference source is copied into this repository):
  1. read the reference's own sources where they lie:  /root/reference/models/ops/src/{vision.cpp,
     ms_deform_attn.h, cpu/*, cuda/*}  (SURVEY.md section 2.1);
  2. stage them into a throw-away directory under /tmp and apply the 2-token patch the survey found
     necessary for torch >= 2.x headers: `value.type()` -> `value.scalar_type()` inside the two
     AT_DISPATCH_FLOATING_TYPES calls (src/cuda/ms_deform_attn_cuda.cu:64 and :134) -- the kernels
     themselves (src/cuda/ms_deform_im2col_cuda.cuh) are compiled byte-for-byte unmodified;
  3. compile with nvcc for sm_100a (plus the reference's own -D flags, models/ops/setup.py:40-45)
     and link against the installed torch, writing ONLY  oracle/_ref/MultiScaleDeformableAttention.so
     (git-ignored, travels to the GPU box with the snapshot).

The resulting module is the reference `models/ops` CUDA build: the GPU parity tests use it as the
bit-level checker for the fp32 kernels and bench.py's `gpu_reference` leg times it beside ours.
It needs a GPU to *run*; building is done here by cross-compilation.
"""
import glob
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

REF = "/root/reference/models/ops/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "MultiScaleDeformableAttention.so")


def build(force: bool = False) -> str | None:
    if not os.path.isdir(REF):
        return OUT if os.path.exists(OUT) else None       # GPU box: use the prebuilt file
    if os.path.exists(OUT) and not force:
        return OUT
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="msda_ref_")
    try:
        stage = os.path.join(tmp, "src")
        shutil.copytree(REF, stage)
        cu = os.path.join(stage, "cuda", "ms_deform_attn_cuda.cu")
        text = open(cu).read()
        n = text.count("AT_DISPATCH_FLOATING_TYPES(value.type()")
        assert n == 2, f"unexpected reference source layout ({n} dispatch sites)"
        text = text.replace("AT_DISPATCH_FLOATING_TYPES(value.type()", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type()")
        open(cu, "w").write(text)

        inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{stage}", f"-I{sysconfig.get_paths()['include']}"]
        common = ["-DWITH_CUDA", "-DTORCH_EXTENSION_NAME=MultiScaleDeformableAttention",
                  "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))]
        objs = []
        for src in [os.path.join(stage, "vision.cpp")] + glob.glob(os.path.join(stage, "cpu", "*.cpp")):
            o = os.path.join(tmp, os.path.basename(src) + ".o")
            subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-w", *common, *inc, "-c", src, "-o", o])
            objs.append(o)
        for src in glob.glob(os.path.join(stage, "cuda", "*.cu")):
            o = os.path.join(tmp, os.path.basename(src) + ".o")
            subprocess.check_call(["nvcc", "-O3", "-std=c++17", "-w", "-Xcompiler", "-fPIC",
                                   "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                                   "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                   "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                                   *common, *inc, "-c", src, "-o", o])
            objs.append(o)
        libdirs = ce.library_paths("cuda")
        link = ["g++", "-shared", *objs, "-o", OUT]
        for d in libdirs:
            link += [f"-L{d}", f"-Wl,-rpath,{d}"]
        link += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
        subprocess.check_call(link)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print("oracle/_ref:", p)
