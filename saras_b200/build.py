"""Builds the native libraries in-tree (nvcc cross-compiles sm_100a without a GPU).

    python -m saras_b200.build            # libsarasmg.so (+ libsarashost.so when the host sources exist)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(HERE, "lib")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",              # parity: no FMA contraction, every rounding where the reference source has it
    "-Xcompiler", "-fPIC", "-shared",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda(force=False, verbose=False, variant=None, defines=()):
    """variant: build saras_b200/lib/variants/libsarasmg_<variant>.so with extra -D flags (kernel tuning experiments are
    compiled here, where nvcc costs no GPU time, and selected on the GPU box with SARASMG_LIB=<path>)."""
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libsarasmg.so")
    if variant:
        os.makedirs(os.path.join(LIBDIR, "variants"), exist_ok=True)
        out = os.path.join(LIBDIR, "variants", "libsarasmg_%s.so" % variant)
    srcs = [os.path.join(HERE, "csrc", "sarasmg.cu")]
    deps = srcs + [os.path.join(HERE, "csrc", f) for f in ("mg_kernels.cuh", "gs_skew.cuh", "gs_strip.cuh", "mg_tma.cuh", "mg_fused.cuh", "proj_kernels.cuh", "impl_kernels.cuh")] + [os.path.join(ROOT, "include", "sarasmg.h")]
    if force or _stale(out, deps):
        extra = os.environ.get("MG_NVCC_EXTRA", "").split() + list(defines)     # tuning experiments: -DMG_MARCH_MINB=5 ...
        cmd = ["nvcc"] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + srcs + ["-lnccl"]
        subprocess.run(cmd, check=True)
    return out


def build_host(force=False):
    hostdir = os.path.join(HERE, "host")
    mk = os.path.join(hostdir, "Makefile")
    if os.path.exists(mk):
        subprocess.run(["make", "-s", "-C", hostdir] + (["-B"] if force else []), check=True)


def build_all(force=False, verbose=False):
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    if "--variant" in sys.argv:
        name = sys.argv[sys.argv.index("--variant") + 1]
        print(build_cuda(force=True, verbose="-v" in sys.argv, variant=name, defines=[a for a in sys.argv if a.startswith("-D")]))
    else:
        build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
